"""ctypes binding of libd3feat_b200.so (the C ABI declared in include/d3feat_b200.h).

There is NO CPU fallback: if the shared library is missing or a CUDA device is not
available the ops raise.  ``build()`` compiles the library in-tree with nvcc for sm_100a
(cross-compiles on a GPU-less host); ``__graft_entry__.build()`` calls it.
"""
import ctypes
import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libd3feat_b200.so")
CSRC = os.path.join(_PKG, "csrc")
HEADER = os.path.join(os.path.dirname(os.path.dirname(_PKG)), "include", "d3feat_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
              "-Xcompiler", "-fPIC", "-shared"]

c_p, c_i, c_f, c_d, c_sz, c_i64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double,
                                  ctypes.c_size_t, ctypes.c_int64)

# name -> (restype, argtypes); mirrors include/d3feat_b200.h one to one
SIGNATURES = {
    "d3f_version": (c_i, []),
    "d3f_last_error_string": (ctypes.c_char_p, []),
    "d3f_launch_count": (ctypes.c_ulonglong, []),
    "d3f_radius_neighbors_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "d3f_radius_neighbors": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_i, c_p, c_i, c_i, c_p, c_i, c_p, c_sz, c_p]),
    "d3f_grid_subsample_workspace_bytes": (c_sz, [c_i, c_i]),
    "d3f_grid_subsample": (c_i, [c_p, c_p, c_i, c_i, c_f, c_p, c_i, c_p, c_p, c_sz, c_p]),
    "d3f_kpconv_workspace_bytes": (c_sz, [c_i] * 6),
    "d3f_kpconv_forward": (c_i, [c_p, c_p, c_p, c_i, c_i64, c_p, c_p, c_p, c_i, c_p,
                                 c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i,
                                 c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "d3f_kpconv_backward": (c_i, [c_p, c_p, c_p, c_i, c_i64, c_p, c_p, c_p, c_i, c_p,
                                  c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i,
                                  c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "d3f_kpconv_backward_ex": (c_i, [c_p, c_p, c_p, c_i, c_i64, c_p, c_p, c_p, c_i, c_p,
                                     c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i,
                                     c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "d3f_kpconv_gather_transposed": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_p, c_p]),
    "d3f_kpconv_grads_from_gathered": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_p]),
    "d3f_neighbors_transpose_workspace_bytes": (c_sz, [c_i]),
    "d3f_neighbors_transpose": (c_i, [c_p, c_i, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_sz, c_p]),
    "d3f_set_kpconv_impl": (None, [c_i]),
    "d3f_get_kpconv_impl": (c_i, []),
    "d3f_kpconv_fused_eligible": (c_i, [c_i, c_i, c_i, c_i]),
    "d3f_kpconv_set_gather_events": (None, [c_p, c_p]),
    "d3f_colsum": (c_i, [c_p, c_i, c_i, c_p, c_p]),
    "d3f_leaky_backward_colsum": (c_i, [c_p, c_p, c_f, c_i, c_i, c_p, c_p, c_p]),
    "d3f_max_pool_forward": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "d3f_max_pool_backward": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "d3f_gather_rows_forward": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, c_i, c_p, c_p]),
    "d3f_gather_rows_backward": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, c_i, c_p, c_p]),
    "d3f_detection_scores_forward": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "d3f_detection_scores_backward": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "d3f_det_loss_forward": (c_i, [c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p]),
    "d3f_det_loss_backward": (c_i, [c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p]),
    "d3f_exchange_chunk_bytes": (c_sz, [c_i, c_i]),
    "d3f_exchange_pack": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "d3f_exchange_unpack": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "d3f_set_gemm_impl": (None, [c_i]),
    "d3f_set_gemm_tuning": (None, [c_i, c_i]),
    "d3f_gemm_tcgen05_failed": (c_i, []),
    "d3f_gemm": (c_i, [c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_f, c_p]),
    "d3f_gemm_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "d3f_gemm_ex": (c_i, [c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f,
                          c_p, c_sz, c_p]),
    "d3f_kpconv_forward_ex": (c_i, [c_p, c_p, c_p, c_i, c_i64, c_p, c_p, c_p, c_i, c_p,
                                    c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_p, c_i, c_f,
                                    c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "d3f_gemm_status_snapshot": (c_i, [c_p, c_p]),
    "d3f_sgd_step": (c_i, [c_p, c_p, c_p, c_sz, c_p, c_f, c_f, c_p, c_i, c_i, c_p]),
    "d3f_gemm_prezeroed": (c_i, [c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_f, c_p]),
    "d3f_colsum_prezeroed": (c_i, [c_p, c_i, c_i, c_p, c_p]),
    "d3f_leaky_backward_colsum_prezeroed": (c_i, [c_p, c_p, c_f, c_i, c_i, c_p, c_p, c_p]),
    "d3f_mutual_nn": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "d3f_pair_loss_aux_floats": (c_sz, [c_i]),
    "d3f_pair_dist": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "d3f_pair_loss_forward": (c_i, [c_p, c_p, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_d, c_f, c_f, c_f,
                                    c_p, c_p, c_p, c_p, c_p, c_p]),
    "d3f_pair_loss_backward": (c_i, [c_p, c_p, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_d, c_f, c_f, c_f,
                                     c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
}

_lib = None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> d3feat/pytorch_b200/libd3feat_b200.so"""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + sources()
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), out.stdout + out.stderr))
    if verbose:
        print(" ".join(cmd))
    return LIB_PATH


def load():
    """Load the C-ABI library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "d3feat.pytorch_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class D3FError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().d3f_last_error_string().decode("utf-8", "replace")
        raise D3FError("libd3feat_b200 status %d: %s" % (rc, msg))
