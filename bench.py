#!/usr/bin/env python
"""bench.py -- fragment-pairs/sec of the D3Feat hot path on B200 (BASELINE.json metric).

One STEP = one 20k+20k synthetic fragment pair through the whole hot path (BASELINE config 3):
device pyramid build (5 radius searches + 4 grid subsamplings + 4 pool + 4 upsample searches)
-> KPFCNN forward (14 KPConv) -> circle + detector loss -> backward -> SGD step.

  python bench.py --gpus N --steps K --warmup W        (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                  (the reference's CPU path, timed on host cores)

Prints ONE JSON line (rank 0).  `value` = pairs/s with the raw pair already resident in HBM,
`e2e` = the same through the public API from pinned HOST buffers with the loss read back.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# Keep stdout for the ONE JSON line: everything else (NCCL banners, warnings) goes to stderr.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

N_POINTS = 20000
POOL = 4  # distinct synthetic pairs cycled through the steps (per rank)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fwd-only", action="store_true", help="diagnostic: forward + loss only")
    ap.add_argument("--host-profile", action="store_true", help="diagnostic: CPU enqueue time per phase -> stderr")
    ap.add_argument("--no-graph", action="store_true", help="diagnostic: static pipeline without CUDA-graph capture")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- helpers
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def kpconv_logical_bytes(nq, ns, H, cin, cout, K=15):
    """algorithmic (logical gather) bytes of one KPConv forward."""
    return nq * H * (4 * cin + 12 + 4) + nq * (12 + 4 * cout) + 4 * K * cin * cout


def kpconv_flops(nq, H, cin, cout, K=15):
    return 2 * K * nq * cin * (H + cout) + 12 * nq * H * K


def make_pairs(n, count, seed0):
    from d3feat.pytorch_b200 import synthetic
    return [synthetic.fragment_pair(n, seed=seed0 + i) for i in range(count)]


# ------------------------------------------------------------------------------------------- reference arm
def cpu_reference_setup(n_points):
    import _inputs
    from d3feat.pytorch_b200.config import default_config
    from oracle import cpu
    cfg = default_config()
    sd = _inputs.kpfcnn_state_dict(cfg, seed=0)
    impl = "ref" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libd3feat_ref.so")) else "port"
    cpu.build()
    return cfg, sd, impl


def tune_cpu_threads(cfg, sd, impl):
    """Pick the torch thread count that runs the CPU path fastest on this host (a 128-thread pool is
    far slower than 16-32 threads for these small ops); probed on a 4k+4k pair."""
    from oracle import pipeline
    cores = os.cpu_count() or 1
    probe = make_pairs(4000, 1, 999)[0]
    lim = [40, 43, 44, 43, 26]
    best, best_t = cores, float("inf")
    for th in sorted({min(cores, v) for v in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(th)
        pipeline.cpu_pair_step(probe, dict(sd), cfg, lim, impl=impl, backward=True)
        t0 = time.perf_counter()
        pipeline.cpu_pair_step(probe, dict(sd), cfg, lim, impl=impl, backward=True)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = th, dt
    torch.set_num_threads(best)
    return best


def cpu_limits(pairs, cfg, impl):
    """neighborhood_limits by the reference rule (dataloader.py:191-223) from the CPU oracle."""
    from oracle import pipeline
    hist_n = int(np.ceil(4 / 3 * np.pi * (cfg.deform_radius + 1) ** 3))
    hists = np.zeros((cfg.num_layers, hist_n), np.int64)
    for d in pairs:
        b = pipeline.cpu_collate(d, cfg, [hist_n] * cfg.num_layers, impl=impl)
        for l, m in enumerate(b["neighbors"]):
            c = (m < m.shape[0]).sum(1).numpy()
            hists[l] += np.bincount(c, minlength=hist_n)[:hist_n]
    cs = np.cumsum(hists.T, axis=0)
    return (cs < 0.8 * cs[hist_n - 1]).sum(0).tolist()


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores: reference C++
    (oracle/_ref) for the pyramid + the torch-CPU restatement of KPFCNN / losses (the reference's
    Python cannot travel to the GPU box).  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pipeline
    cfg, sd, impl = cpu_reference_setup(args.points)
    cores = tune_cpu_threads(cfg, sd, impl)
    pairs = make_pairs(args.points, min(POOL, max(1, args.steps)), 0)
    limits = LIMITS_20K if args.points == N_POINTS else cpu_limits(pairs[:1], cfg, impl)
    sgd = {}
    for i in range(args.warmup):
        pipeline.cpu_pair_step(pairs[i % len(pairs)], sd, cfg, limits, impl=impl, backward=not args.fwd_only, sgd=sgd)
    t0 = time.perf_counter()
    stages = {}
    for i in range(args.steps):
        t, _ = pipeline.cpu_pair_step(pairs[i % len(pairs)], sd, cfg, limits, impl=impl, backward=not args.fwd_only, sgd=sgd)
        for k, v in t.items():
            stages[k] = stages.get(k, 0.0) + v
    dt = time.perf_counter() - t0
    val = args.steps / dt
    sample = "%d pairs of %d+%d points, whole path (collate+fwd+loss+bwd), %s" % (
        args.steps, args.points, args.points, "pyramid by the reference C++ (oracle/_ref), model by the torch-CPU oracle"
        if impl == "ref" else "oracle port")
    line = {"impl": "reference", "metric": "fragment-pairs/sec", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, limits),
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "reference" if impl == "ref" else "port",
                             "sample": sample, "stage_seconds_per_pair": {k: v / args.steps for k, v in stages.items()}},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


LIMITS_20K = [35, 42, 42, 45, 47]  # 80th-percentile rule on the synthetic 20k pairs (recomputed on the device below)


def workload_config(args, limits):
    return {"workload": "BASELINE config 3: %d+%d-point synthetic room-shell pair, device pyramid build + KPFCNN fwd "
                        "(default arch, 24.3M params, K=15, first_features_dim=128) + circle & detector loss (P=128) "
                        "+ bwd + SGD step" % (args.points, args.points),
            "points_per_fragment": args.points, "neighborhood_limits": [int(v) for v in limits], "pairs_per_step_per_gpu": 1,
            "parallelism": "dp%d (1 pair per GPU, descriptor all-gather)" % args.gpus,
            "l2_policy": "256 MiB L2 flush write between timed steps (outside the per-step event brackets)"}


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    from d3feat.pytorch_b200 import _lib, ops
    from d3feat.pytorch_b200.architectures import KPFCNN
    from d3feat.pytorch_b200.blocks import gather
    from d3feat.pytorch_b200.config import default_config
    from d3feat.pytorch_b200.dataloader import calibrate_neighbors, collate_fn_descriptor
    from d3feat.pytorch_b200.loss import PairLoss
    from d3feat.pytorch_b200 import parallel
    from d3feat.pytorch_b200.engine import PairStep, plan_capacities

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()  # raises if the sm_100a library has not been built: no fallback
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    cfg = default_config()
    torch.manual_seed(0); np.random.seed(0)
    model = KPFCNN(cfg).to(dev)
    model.train()
    # config.py:62-69; fused=True: one multi-tensor kernel per parameter group instead of three foreach passes
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.98, weight_decay=1e-6, fused=True)
    # multi-GPU: all gradients are views of one buffer (zero() + ONE all-reduce); single GPU: plain per-parameter
    # gradients created by the backward pass (saves one accumulate kernel per parameter, ~130 launches per step)
    flat = parallel.FlatGradients(model) if world > 1 else None
    loss_fn = PairLoss("circle", "euclidean", cfg.log_scale, cfg.safe_radius, cfg.pos_margin, cfg.neg_margin)

    pairs = make_pairs(args.points, POOL, 100 * rank)

    class _DS:
        config = cfg
        def __len__(self): return 2
        def __getitem__(self, i): return pairs[i]
    limits = [int(v) for v in calibrate_neighbors(_DS(), cfg, collate_fn_descriptor, samples_threshold=10 ** 9)]

    host = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in p) for p in pairs]
    devp = [tuple(t.to(dev) for t in p) for p in host]
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    hp = {}

    def tick(name, t0):
        if args.host_profile:
            hp[name] = hp.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def step(data, read_loss):
        t0 = time.perf_counter()
        batch = collate_fn_descriptor([data], cfg, limits)
        t0 = tick("collate", t0)
        feats, scores = model(batch)
        t0 = tick("forward", t0)
        c = batch["corr"].long()
        n0 = data[0].shape[0]
        ia, ip = c[:, 0], c[:, 1] + n0
        a, p = gather(feats, ia), gather(feats, ip)          # trainer.py:91-94 row selects
        sa, sp = gather(scores, ia), gather(scores, ip)
        if world > 1:
            out = parallel.cross_fragment_loss(loss_fn, a, p, batch["dist_keypts"], sa, sp)
        else:
            out = loss_fn(a, p, batch["dist_keypts"], sa, sp)
        loss = out["desc_loss"] * cfg.desc_loss_weight + out["det_loss"] * cfg.det_loss_weight
        t0 = tick("loss", t0)
        if not args.fwd_only:
            if flat is not None:
                flat.zero()
            else:
                opt.zero_grad(set_to_none=True)
            loss.backward()
            t0 = tick("backward", t0)
            if flat is not None:
                flat.allreduce()
            opt.step()
            t0 = tick("optimizer", t0)
        return float(loss.detach()) if read_loss else loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the production path: static capacities, no host sync, one CUDA graph per pair step
    sizes = [[int(t.shape[0]) for t in collate_fn_descriptor([p], cfg, limits)["points"]] for p in pairs]
    caps = plan_capacities(sizes)
    stepper = PairStep(model, cfg, limits, caps, args.points, args.points, loss_fn,
                       None if args.fwd_only else opt, None if args.fwd_only else flat, num_node=cfg.num_node,
                       cross_fragment=parallel.cross_fragment_loss if world > 1 else None)
    l0 = lib.d3f_launch_count()
    stepper(devp[0])
    torch.cuda.synchronize()
    launches_per_step = lib.d3f_launch_count() - l0
    graph_mode = "cuda-graph"
    if args.no_graph:
        graph_mode = "eager (static shapes)"
    else:
        try:
            stepper.capture()
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("CUDA graph capture failed (%s: %s); running the static pipeline eagerly\n" % (type(e).__name__, e))
            stepper.graph = None
            graph_mode = "eager (static shapes; capture failed)"
            torch.cuda.synchronize()

    def timed(kind, steps, profile=False):
        src = devp if kind == "device" else host
        evs = []
        barrier()
        ops.PROFILE = {} if profile else None
        launches0 = lib.d3f_launch_count()
        wall0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)          # L2 flush, outside the event bracket
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if profile:
                step(src[i % POOL], read_loss=False)        # exact-shape eager path with per-op events
            else:
                stepper(src[i % POOL])                      # H2D (e2e) or D2D copies into the static inputs + replay
                if kind == "host":
                    float(stepper.loss)                     # D2H read of the step's result
            e1.record()
            evs.append((e0, e1))
        barrier()
        wall = time.perf_counter() - wall0
        prof, ops.PROFILE = ops.PROFILE, None
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), wall, lib.d3f_launch_count() - launches0, prof

    for i in range(max(args.warmup, 3)):
        stepper(devp[i % POOL])
    for i in range(2):
        stepper(host[i % POOL]); float(stepper.loss)
    stepper.check()
    for i in range(2):
        step(devp[i % POOL], False)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    hp.clear()
    ms_dev, wall_dev, _, _ = timed("device", args.steps)
    ms_e2e, wall_e2e, _, _ = timed("host", args.steps)
    stepper.check()                  # no capacity / candidate-buffer overflow in any timed step
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    # separate pass with per-op CUDA events (same steps, same stream) for the roofline / op breakdown only
    _, _, _, prof = timed("device", args.steps, profile=True)

    def finish():
        """Leave without interpreter / NCCL teardown: with CUDA graphs that captured NCCL kernels,
        destroy_process_group() (and the implicit teardown at exit) hung a 2-GPU run after its JSON line was out
        (round 1).  All ranks meet at a barrier first so nobody exits under a peer that still needs it."""
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        _REAL_STDOUT.flush()
        sys.stderr.flush()
        os._exit(0)

    if rank != 0:
        finish()

    if args.host_profile:
        n_calls = 2 * args.steps + args.steps
        sys.stderr.write("host enqueue ms/step (timed passes only approx): %s\n" % {k: round(1e3 * v / max(n_calls, 1), 3) for k, v in hp.items()})
    pk, pk_src = peaks()
    # roofline of the dominant op: the KPConv forward with the most algorithmic bytes (L0 resnetb 32->32)
    fwd = {k: v for k, v in prof.items() if k[0] == "kpconv_fwd"}
    per_layer, tot_bytes, tot_ms = [], 0, 0.0
    for k, evs in fwd.items():
        _, nq, ns, H, cin, cout, deformed = k
        calls_per_step = len(evs) / args.steps
        avg_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        by = kpconv_logical_bytes(nq, ns, H, cin, cout)
        per_layer.append({"nq": nq, "ns": ns, "H": H, "cin": cin, "cout": cout, "calls_per_step": calls_per_step,
                          "ms": avg_ms, "logical_MB": by / 1e6, "GBps": by / avg_ms / 1e6,
                          "TFLOPs": kpconv_flops(nq, H, cin, cout) / avg_ms / 1e9})
        tot_bytes += by * calls_per_step
        tot_ms += avg_ms * calls_per_step
    per_layer.sort(key=lambda d: -d["logical_MB"])
    dom = per_layer[0]
    # the gather kernel of that layer alone (events recorded inside d3f_kpconv_forward around kp2_correlate): it reads
    # the neighbour rows / positions / indices and writes wf [Nq, K*Cin] for the contraction
    gather = None
    gk = [(k, v) for k, v in prof.items() if k[0] == "kpconv_gather" and k[1:6] == (dom["nq"], dom["ns"], dom["H"], dom["cin"], dom["cout"])]
    if gk:
        evs = gk[0][1]
        g_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        g_bytes = dom["nq"] * dom["H"] * (4 * dom["cin"] + 16) + 12 * dom["nq"] + 4 * dom["nq"] * 15 * dom["cin"]
        gather = {"kernel": "kp2_correlate (gather + kernel-point correlation)", "avg_ms_per_launch": g_ms,
                  "algorithmic_bytes_per_launch": g_bytes, "GBps": g_bytes / g_ms / 1e6,
                  "frac_of_hbm_peak": g_bytes / g_ms / 1e6 / pk["hbm_gbs"],
                  "note": "bytes = Nq*H*(4Cin+12+4) + 12Nq read + 4*Nq*K*Cin written (wf)"}
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
    stage_ms = {}
    for k, evs in prof.items():
        stage_ms[k[0]] = stage_ms.get(k[0], 0.0) + sum(a.elapsed_time(b) for a, b in evs) / args.steps
    op_breakdown = sorted(([list(map(str, k)), len(evs) / args.steps, sum(a.elapsed_time(b) for a, b in evs) / args.steps]
                           for k, evs in prof.items()), key=lambda t: -t[2])[:40]
    roofline = {"bound": "hbm", "achieved": dom["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": dom["GBps"] / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "KPConv forward op (kp_rowpos + kp2_correlate gather + tcgen05 contraction) of the layer with the "
                          "most algorithmic bytes: Nq=%d Ns=%d H=%d Cin=%d Cout=%d; achieved = SURVEY 8(d) logical gather "
                          "bytes / CUDA-event time of the whole op" % (dom["nq"], dom["ns"], dom["H"], dom["cin"], dom["cout"]),
                "gather_kernel": gather, "kpconv_impl": int(lib.d3f_get_kpconv_impl()),
                "algorithmic_bytes_per_launch": dom["logical_MB"] * 1e6, "avg_ms_per_launch": dom["ms"],
                "peak_source": pk_src + ", burst copy bandwidth",
                "all_kpconv_fwd": {"logical_GB_per_step": tot_bytes / 1e9, "ms_per_step": tot_ms,
                                   "GBps": tot_bytes / tot_ms / 1e6, "frac": tot_bytes / tot_ms / 1e6 / pk["hbm_gbs"]},
                "layers": per_layer[:6], "stage_ms_per_step": stage_ms,
                "op_breakdown_ms_per_step": [{"op": " ".join(k), "calls": c, "ms": round(ms, 4)} for k, c, ms in op_breakdown]}

    pairs_per_step = world
    value = pairs_per_step * args.steps / (ms_dev / 1e3)
    e2e_val = pairs_per_step * args.steps / (ms_e2e / 1e3)
    line = {"metric": "fragment-pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, limits),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "api": "engine.PairStep(pinned host tensors): H2D -> [pyramid build -> KPFCNN -> PairLoss -> backward -> SGD] "
                           "as one CUDA graph -> float(loss)"},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
            "execution": {"mode": graph_mode, "capacities": caps, "final_loss": float(stepper.loss)},
            "wall_s": {"device": wall_dev, "e2e": wall_e2e},
            "roofline": roofline}

    if not args.no_cpu_baseline and world == 1:
        from oracle import pipeline
        cfg_c, sd, impl = cpu_reference_setup(args.points)
        cores = tune_cpu_threads(cfg_c, sd, impl)
        sgd = {}
        pipeline.cpu_pair_step(pairs[0], sd, cfg_c, limits, impl=impl, backward=not args.fwd_only, sgd=sgd)
        t0 = time.perf_counter()
        n_cpu, stages = 0, {}
        while n_cpu < 3 or (time.perf_counter() - t0 < 10 and n_cpu < 10):
            t, _ = pipeline.cpu_pair_step(pairs[n_cpu % POOL], sd, cfg_c, limits, impl=impl, backward=not args.fwd_only, sgd=sgd)
            for k, v in t.items():
                stages[k] = stages.get(k, 0.0) + v
            n_cpu += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "pairs/s", "cores": cores,
                                "kind": "reference" if impl == "ref" else "port",
                                "sample": "%d pairs of the same workload (pyramid: reference C++ via oracle/_ref, single thread as "
                                          "in the reference; model+loss+bwd+SGD: torch-CPU oracle on %d threads = fastest of {8,16,32,64,all %d})" % (n_cpu, cores, os.cpu_count() or 1),
                                "stage_seconds_per_pair": {k: v / n_cpu for k, v in stages.items()}}
    emit(line)
    finish()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
