"""Default hyper-parameters of the hot path.

Values restate the reference's argparse defaults (config.py:29-59, :78) and the
architecture list built in code (training_3DMatch.py:44-56); only the attributes
the hot path reads are kept ("Config / flags").
"""
from types import SimpleNamespace


def build_architecture(num_layers=5, deformable_from=None):
    """training_3DMatch.py:44-56.  `deformable_from`=l makes the non-strided resnetb blocks of
    layers >= l deformable (BASELINE config 4: 'resnetb_deformable' in layers 3-4)."""
    def rb(layer, strided=False):
        d = deformable_from is not None and layer >= deformable_from
        return "resnetb" + ("_deformable" if d and not strided else "") + ("_strided" if strided else "")

    arch = ["simple", rb(0)]
    for i in range(num_layers - 1):
        arch += [rb(i, strided=True), rb(i + 1), rb(i + 1)]
    for _ in range(num_layers - 2):
        arch += ["nearest_upsample", "unary"]
    arch += ["nearest_upsample", "last_unary"]
    return arch


def default_config(**overrides):
    cfg = SimpleNamespace(
        # network (config.py:29-46)
        num_layers=5, in_points_dim=3, first_features_dim=128, first_subsampling_dl=0.03,
        in_features_dim=1, conv_radius=2.5, deform_radius=5.0, num_kernel_points=15,
        KP_extent=2.0, KP_influence="linear", aggregation_mode="sum",
        fixed_kernel_points="center", use_batch_norm=False, batch_norm_momentum=0.02,
        deformable=False, modulated=False,
        # loss (config.py:49-59)
        dist_type="euclidean", desc_loss="circle", pos_margin=0.1, neg_margin=1.4,
        log_scale=10.0, safe_radius=0.1, desc_loss_weight=1.0, det_loss_weight=1.0,
        # data (config.py:78)
        num_node=128,
    )
    for k, v in overrides.items():
        setattr(cfg, k, v)
    if not hasattr(cfg, "architecture"):
        cfg.architecture = build_architecture(cfg.num_layers)
    return cfg
