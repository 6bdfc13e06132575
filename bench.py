#!/usr/bin/env python
"""bench.py -- headline benchmark of the octree volume-rendering hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (config C3 of BASELINE.json, the configuration the metric is quoted on): depth-8 ball octree
(1 897 408 leaf rows x 32 channels), 2^20 random rays PER GPU, one step = forward feature render + backward into
the leaf feature table (+ NCCL all-reduce of the leaf gradients when N > 1). Weak scaling: rays shard across
ranks with a fixed per-GPU batch, tree and features replicated, the only exchange is the gradient sum.

One JSON line on stdout (rank 0). `value` = whole-job Mrays/s with inputs resident in HBM; `e2e` = the same step
through the public API (VolumeRenderer + autograd) with rays and targets coming from pinned host memory and the
loss read back every step; `roofline` = the dominant kernel (backward march) against the measured HBM peak;
`cpu_baseline` = the CPU oracle (a port of the reference's CUDA algorithm; the reference has no working CPU path)
timed on this box's cores on a bounded ray sample. `--impl reference` times that CPU port as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line. Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


import numpy as np  # noqa: E402

L_TREE, SHAPE, D_FEAT = 8, "ball", 32
Q_PER_GPU = 1 << 20
CPU_SAMPLE = 32768          # rays per CPU-baseline measurement / per reference-arm step
WORKLOAD = ("C3 training step: depth-8 ball octree (1897408 leaf rows x 32 ch, 281697 nodes), 2^20 random rays per "
            "GPU, fwd feature render + bwd into leaf features")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(cnt, D, M, explicit_rays=True, depth=False):
    """bytes per launch from the oracle's counters (S samples, LV child lookups, V valid rows, H hits)."""
    Q, S, LV, V, H = cnt["Q"], cnt["S"], cnt["LV"], cnt["V"], cnt["H"]
    r_in = 36 if explicit_rays else 0
    d_out = 4 * D + (4 if depth else 0)
    b_fwd = Q * (r_in + d_out) + 4 * LV + 4 * S + 4 * V + 4 * (D - 1) * H
    b_bwd = Q * (r_in + 8 * D) + 4 * LV + 4 * S + 4 * V + 4 * (D - 1) * H + 4 * D * H + 4 * M * D
    return b_fwd, b_bwd


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50", "-i", uuid],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def scene_numpy(seed_rays):
    from svox_t_b200 import synth
    tr = synth.synth_tree(L_TREE, SHAPE)
    f = synth.synth_features(tr["M"], D_FEAT, seed=0)
    o, d = synth.synth_rays(Q_PER_GPU, seed=seed_rays)
    return tr, f, o, d


def host_threads():
    """The CPU legs use every host core; torchrun injects OMP_NUM_THREADS=1, which would silently serialise them."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def cpu_port_step(T, f, o, d, g, orc):
    """One fwd+bwd of the CPU oracle on a ray sample; returns (seconds, counters)."""
    t0 = time.perf_counter()
    _, _, cnt = orc.render_rays(T, f, o, d, want_counters=True)
    orc.render_rays_backward(T, f, o, d, g)
    return time.perf_counter() - t0, cnt


def run_reference_arm(args):
    """The reference's algorithm on the host cores (CPU port = the oracle; the reference itself has no runnable CPU
    path and its CUDA extension is not a CPU baseline). Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = host_threads()
    from oracle import oracle as orc
    tr, f, o, d = scene_numpy(1)
    T = orc.Tree(tr["child"], tr["data"])
    rng = np.random.default_rng(5)
    n = CPU_SAMPLE
    times = []
    for s in range(args.warmup + args.steps):
        lo = (s * n) % (Q_PER_GPU - n)
        g = rng.standard_normal((n, D_FEAT)).astype(np.float32)
        dt, _ = cpu_port_step(T, f, o[lo:lo + n], d[lo:lo + n], g, orc)
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = n / (ms * 1e-3) / 1e6
    sample = f"each step = fwd+bwd of a {n}-ray slice of the 2^20-ray batch (C oracle, OpenMP over rays)"
    emit({
        "impl": "reference", "metric": "Mrays/s fwd+bwd feature render", "value": val, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": n},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--skip-extras", action="store_true", help="skip the C2 image / reference-CUDA side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import svox_t_b200 as sv
    from svox_t_b200 import csrc as C, dist as svd, synth

    rank, world, local_rank = svd.init_from_env()
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    C.load_library()

    tr, f, o, d = scene_numpy(1 + rank)
    Q, D, M = Q_PER_GPU, D_FEAT, tr["M"]
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    feats = torch.from_numpy(f).to(dev)
    o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    g_t = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + rank))
    renderer = sv.VolumeRenderer(tree)
    opt = renderer._get_options()
    ts = tree._spec(feats)
    rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
    accel = tree.accel(feats)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(rec=None):
        e = [ev() for _ in range(4)] if rec is not None else None
        if e: e[0].record()
        # features change every training step: the per-row activation pass and the refresh of the accelerator's
        # hit marks (both derived from the features) are part of the step
        ts._act = C.Activated(feats)
        if accel is not None:
            accel._marks_key = None
            accel.mark_hits(feats)
        out = C.volume_render(ts, rs, opt)
        if e: e[1].record()
        grad = torch.zeros_like(feats)
        if e: e[2].record()
        C._check(C.load_library().svoxb_render_rays_bwd(
            C.ctypes.byref(ts._c()), C._ptr(o_t), C._ptr(d_t), C._ptr(d_t), Q,
            C.ctypes.byref(opt._c(sigma_thresh=0.0, stop_thresh=-1.0)),
            C._ptr(g_t), C._ptr(out), C._ptr(grad), C._stream()))
        if e: e[3].record()
        svd.all_reduce_leaf_grads(grad)
        if rec is not None:
            rec.append(e)
        return out, grad

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    sampler = ClockSampler(uuid) if rank == 0 else None
    time.sleep(0.25)
    svd.barrier(); torch.cuda.synchronize()
    launches0 = C.launch_count()
    t_wall0 = time.time()
    rec = []
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        step(rec)
    e1.record()
    torch.cuda.synchronize(); svd.barrier()
    t_wall1 = time.time()
    launches = C.launch_count() - launches0
    total_ms = svd.max_over_ranks(e0.elapsed_time(e1), dev)
    ms_per_step = total_ms / args.steps
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in rec]))       # includes the activation pass
    bwd_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in rec]))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * Q / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API: pinned host rays + targets in, loss out, every step ----------------------
    # Data-loader style pipeline: two device buffer sets; while step k renders out of one, a copy stream uploads the
    # inputs of step k+1 into the other. Every step's host->device copy and loss read-back happen inside the timed
    # region (K uploads + K renders + K read-backs for K steps; the first upload is exposed, the rest overlap).
    # Per-step host inputs: ray origins + directions (what VolumeRenderer.forward takes) and the supervision a
    # training step consumes -- an RGB target and an opacity target per ray, as in image-supervised training, where
    # the 31 rendered feature channels pass through a fixed decoder (mean of three channel groups) before the loss.
    h_o, h_d = torch.from_numpy(o).pin_memory(), torch.from_numpy(d).pin_memory()
    gen = torch.Generator().manual_seed(7 + rank)
    h_rgb = torch.rand(Q, 3, generator=gen).pin_memory()
    h_alpha = torch.rand(Q, generator=gen).pin_memory()
    fparam = feats.clone().requires_grad_(True)
    bufs = [(torch.empty(Q, 3, device=dev), torch.empty(Q, 3, device=dev), torch.empty(Q, 3, device=dev),
             torch.empty(Q, device=dev)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    free_ev = [None, None]
    G3 = (D - 1) // 3                                  # channels per decoded colour
    w_dec = torch.zeros(D, 4, device=dev)              # fixed linear decoder: RGB = means of three channel groups,
    for c in range(3):                                 # column 3 passes the opacity through
        w_dec[c * G3:(c + 1) * G3, c] = 1.0 / G3
    w_dec[D - 1, 3] = 1.0

    def upload(k):
        bo, bd, brgb, ba = bufs[k & 1]
        with torch.cuda.stream(copy_stream):
            if free_ev[k & 1] is not None:
                copy_stream.wait_event(free_ev[k & 1])          # the step that last used this buffer set is done
            bo.copy_(h_o, non_blocking=True); bd.copy_(h_d, non_blocking=True)
            brgb.copy_(h_rgb, non_blocking=True); ba.copy_(h_alpha, non_blocking=True)
            e = torch.cuda.Event(); e.record(copy_stream)
        return e

    def e2e_run(n_steps):
        losses = []
        ready = upload(0)
        for k in range(n_steps):
            nxt = upload(k + 1) if k + 1 < n_steps else None
            main_stream.wait_event(ready)
            bo, bd, brgb, ba = bufs[k & 1]
            fparam.grad = None
            with torch.no_grad():
                fparam.add_(0.0)                       # stands in for the optimiser update: features change every step
            out = renderer(fparam, sv.Rays(bo, bd, bd))
            dec = out @ w_dec                          # [Q, 4]: decoded RGB + opacity
            loss = 0.5 * ((dec[:, :3] - brgb) ** 2).mean() + 0.5 * ((dec[:, 3] - ba) ** 2).mean()
            loss.backward()
            svd.all_reduce_leaf_grads(fparam.grad)
            free_ev[k & 1] = torch.cuda.Event(); free_ev[k & 1].record(main_stream)
            losses.append(float(loss.item()))          # device -> host read of the step's result
            ready = nxt
        return losses

    e2e_run(3)
    svd.barrier(); torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    torch.cuda.synchronize(); svd.barrier()
    e2e_ms = svd.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    e2e = {"value": world * Q / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(h_o.numel() + h_d.numel() + h_rgb.numel() + h_alpha.numel()) * 4,
           "d2h_bytes_per_step": 4,
           "api": "VolumeRenderer.forward + autograd backward; loss = MSE(decoded RGB, rgb target) + MSE(opacity, alpha "
                  "target), RGB / opacity = a fixed linear decoder (out @ W[32,4]: means of three groups of the rendered feature channels, opacity passed through). Every step uploads its ray "
                  "origins, directions, RGB and opacity targets from pinned host memory and reads the loss back; the "
                  "upload of step k+1 overlaps the render of step k (double buffering)"}

    if world > 1:
        import torch.distributed as tdist
        svd.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    # ---- CPU baseline + counters (rank 0, bounded sample) ---------------------------------------------------------------
    cores = host_threads()
    from oracle import oracle as orc
    T = orc.Tree(tr["child"], tr["data"])
    n = CPU_SAMPLE
    g_np = np.random.default_rng(5).standard_normal((n, D)).astype(np.float32)
    cpu_s, cnt = cpu_port_step(T, f, o[:n], d[:n], g_np, orc)
    cpu_baseline = {"value": n / cpu_s / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                    "sample": f"first {n} of the 2^20 rays, fwd+bwd once, C oracle with OpenMP over rays ({cpu_s:.2f} s)"}
    scale = Q / cnt["Q"]
    cnt_full = {k: (v * scale if k != "Q" else Q) for k, v in cnt.items()}
    b_fwd, b_bwd = algorithmic_bytes(cnt_full, D, M)
    peak, peak_src = measured_peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    roof = lambda b, ms, key: {"bound": "hbm", "achieved": b / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": b / (ms * 1e-3) / 1e9 / peak, "traffic": traffic.get(key),
                               "kernel": key, "ms_per_launch": ms, "algorithmic_bytes_per_launch": b,
                               "peak_source": peak_src}
    out = {
        "metric": "Mrays/s fwd+bwd feature render", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": Q, "global_rays": world * Q, "D": D, "leaf_rows": M,
                   "nodes": tr["n_nodes"], "options": "step_size=1e-3, background=1, sigma_thresh=stop_thresh=0",
                   "parallelism": f"ray-sharded x{world}, tree+features replicated, NCCL all-reduce of grad[M,D]",
                   "l2": "inputs larger than L2: features 243 MB + grad 243 MB + out 134 MB + rays 25 MB vs 126 MB",
                   "accelerator": accel.describe() if accel is not None else None},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roof(b_bwd, bwd_ms, "march_bwd_kernel"),
        "roofline_fwd": roof(b_fwd, fwd_ms, "march_fwd_kernel"),
        "roofline_step": {"achieved": (b_fwd + b_bwd) / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": (b_fwd + b_bwd) / (ms_per_step * 1e-3) / 1e9 / peak,
                          "bytes_per_ray": (b_fwd + b_bwd) / Q},
        "kernel_ms": {"fwd": fwd_ms, "bwd": bwd_ms},
        "fwd_only": {"value": world * Q / (fwd_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                     "note": "forward feature render alone (BASELINE metric (i)), incl. the per-step activation + marking passes"},
        "counters_per_ray": {k: cnt[k] / cnt["Q"] for k in ("S", "LV", "V", "H")},
        "cpu_baseline": cpu_baseline,
    }
    if not args.skip_extras and world == 1:
        out["extras"] = extras(sv, C, synth, tree, feats, renderer, opt, ts, rs, o_t, d_t, g_t, dev, peak)
    emit(out)


def extras(sv, C, synth, tree, feats, renderer, opt, ts, rs, o_t, d_t, g_t, dev, peak):
    """Side measurements (N = 1): config C2 image render, and the reference's own CUDA kernels on the same inputs."""
    import torch
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def best(fn, warm=2, it=5):
        for _ in range(warm):
            fn()
        ts_ = []
        for _ in range(it):
            a, b = ev(), ev()
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts_.append(a.elapsed_time(b))
        return float(np.median(ts_))

    ex = {}
    cam = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev)
    cs = sv.renderer._make_camera_spec(cam, 800, 800, 1111.111, 1111.111)
    ms = best(lambda: C.volume_render_image_with_depth(ts, cs, opt))
    ex["c2_image_800x800_fwd_with_depth"] = {"ms": ms, "Mrays/s": 0.64 / (ms * 1e-3)}
    try:   # secondary forward-only series of SURVEY 8d: fast=True thresholds (sigma_thresh = stop_thresh = 1e-2)
        fast = renderer._get_options(True)
        ms = best(lambda: C.volume_render(ts, rs, fast))
        ex["c3_fwd_fast_thresholds"] = {"ms": ms, "Mrays/s": o_t.shape[0] / (ms * 1e-3) / 1e6,
                                        "sigma_thresh": fast.sigma_thresh, "stop_thresh": fast.stop_thresh}
    except Exception as e:
        ex["c3_fwd_fast_thresholds"] = {"unavailable": str(e)[:200]}
    try:   # motion-feature render on the same tree and rays: J = 24 joints, F = 32, B = 4
        rng = np.random.default_rng(0)
        M = feats.shape[0]
        jf = torch.randn(24, 32, device=dev, requires_grad=True)
        sw = torch.from_numpy(rng.dirichlet(np.ones(4), M).astype(np.float32)).to(dev)
        ji = torch.from_numpy(rng.integers(0, 24, (M, 4)).astype(np.int32)).to(dev)
        rays = sv.Rays(o_t, d_t, d_t)
        gm = torch.randn(o_t.shape[0], 32, device=dev)
        f_ms = best(lambda: renderer.motion_feature_render(feats, jf.detach(), sw, ji, rays), 1, 3)

        def fb():
            jf.grad = None
            renderer.motion_feature_render(feats, jf, sw, ji, rays).backward(gm)
        ex["motion_feature_render_J24_F32_B4"] = {"fwd_ms": f_ms, "fwd_bwd_ms": best(fb, 1, 3)}
        del jf, sw, ji, gm
    except Exception as e:
        ex["motion_feature_render_J24_F32_B4"] = {"unavailable": str(e)[:200]}
    try:   # config C4: animated frame = LBS warp of 2^20 points + p2v splat (256^3) + octree rebuild to depth 8 + accelerator
        P = 1 << 20   #            + 1920x1080 render with opacity and depth; per-frame latency, CUDA events around the frame
        vox = synth._occupied_keys(8, "ball")
        pts = synth.voxel_centers(vox[np.random.default_rng(2).permutation(len(vox))[:P]], 8)
        Tm, w4, j4 = synth.synth_skeleton(P)
        p_t, Tm_t, w_t, j_t = (torch.from_numpy(a).to(dev) for a in (pts, Tm, w4, j4))
        f4 = torch.from_numpy(synth.synth_features(P, 32)).to(dev)
        corner, size = torch.zeros(3, device=dev), torch.ones(3, device=dev)
        tree4 = sv.N3Tree(N=2, data_dim=32, map_location=dev)
        r4 = sv.VolumeRenderer(tree4)

        def frame():
            warped, _ = sv.warp_vertices(Tm_t, p_t, w_t, j_t)
            sv.voxelize(warped, f4, corner, size, 256, 1.5 / 256, 2.0 / 256)
            tree4.build_from_points(warped, 8)
            r4.render_persp_with_depth(f4, cam, width=1920, height=1080, fx=1500.0)
        ex["c4_animated_frame_1080p"] = {"ms_per_frame": best(frame, 2, 5), "points": P, "nodes": int(tree4.filled)}
    except Exception as e:
        ex["c4_animated_frame_1080p"] = {"unavailable": str(e)[:200]}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refdrv
        if refdrv.available():
            m = refdrv.module()
            rts = refdrv.tree_spec(feats, tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
            rrs, ro = refdrv.rays_spec(o_t, d_t), refdrv.options()
            f_ms = best(lambda: m.volume_render(rts, rrs, ro), 1, 3)
            b_ms = best(lambda: m.volume_render_backward(rts, rrs, ro, g_t), 1, 3)
            ex["reference_cuda_same_inputs"] = {"fwd_ms": f_ms, "bwd_ms": b_ms,
                                                "Mrays/s_fwd_bwd": Q_PER_GPU / ((f_ms + b_ms) * 1e-3) / 1e6,
                                                "note": "unmodified svox_t csrc compiled for sm_100a (oracle/_ref)"}
    except Exception as e:  # the checker is optional here
        ex["reference_cuda_same_inputs"] = {"unavailable": str(e)[:200]}
    return ex


if __name__ == "__main__":
    main()
