#!/usr/bin/env python
"""bench.py — simulated+rendered frames/s at 800x800 (BASELINE.json metric), chair config (configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--density-scale S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one GUI frame of the reference (nerf/gui.py:556-645 -> trainer.py:284-329): read IP state, one
Q-GMLS `stepforward` (10 local-global iterations, ~2k IPs), deformed-space render of 800x800 rays.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_SAMPLE_FUSED = 1036        # 12 B in + 16 levels x 8 corners x 8 B gathered; encodings never written
ALGO_BYTES_PER_SAMPLE_GRID = 1164         # stand-alone grid_encode_forward also writes 128 B / sample
MLP_FLOP_PER_SAMPLE = 18688               # 9344 MAC


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"     # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pienerf_b200.frame import DistFrameDriver, build_scene

    model, sim, opt, pose, intr, body, field = build_scene(args.config, device=dev, density_scale=args.density_scale)
    drv = DistFrameDriver(model, sim, opt, overlap_sim=not args.no_overlap_sim)
    W, H = opt.W, opt.H
    N = W * H
    host_pose = torch.from_numpy(pose).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # > 126 MB L2
    calib = None
    if world > 1 and not args.equal_tiles and args.no_overlap_sim:
        # only when the step is serialised with rank 0's render does rank 0 need a smaller share of the tiles; with the
        # step on its concurrent side stream (default) equal shares measured better (2 GPUs: 507 vs 484 fps)
        calib = drv.calibrate(host_pose, intr)
    from pienerf_b200 import _lib
    n_pass = int(_lib.lib.pn_render_pass_count(int(opt.max_steps)))

    def frame(e2e, prof=None):
        """One GUI frame through the public multi-GPU frame API.  e2e=True adds the host<->device traffic: rays are
        regenerated from the host pose and the gathered frame is copied to pinned host memory."""
        out, _ = drv.frame(host_pose, intr, to_host=e2e, regenerate_rays=e2e, profile_events=prof)
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    enqueue_s = [0.0]                                                        # host time spent enqueueing frames (all timed loops)
    enqueue_n = [0]

    def timed(K, e2e, with_prof=False):
        enqueue_n[0] += K
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        # per frame: (start, stop) around all render passes + one event pair per field-kernel launch (one per pass)
        pv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
               [torch.cuda.Event(enable_timing=True) for _ in range(2 * n_pass)]) for _ in range(K)] if with_prof else None
        if pv:
            for a, b, lst in pv:
                a.record(); b.record()                                        # instantiate the handles
                for e in lst:
                    e.record()
        samples0 = []
        sync_all()
        t0 = time.perf_counter()
        for i in range(K):
            flush.fill_(float(i))                                             # evict L2 between timed frames (not timed)
            ev[i][0].record()
            c0 = time.perf_counter()
            out = frame(e2e, pv[i] if pv else None)
            enqueue_s[0] += time.perf_counter() - c0
            ev[i][1].record()
            samples0.append(out["stats"].clone())
        if e2e:
            drv.wait_host()                                                   # every frame has landed in pinned host memory
        sync_all()
        wall = time.perf_counter() - t0
        ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = float(sum(ms))
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        kms = None
        if pv:
            kms = {"render": [a.elapsed_time(b) for a, b, _ in pv],
                   "field": [sum(lst[2 * k].elapsed_time(lst[2 * k + 1]) for k in range(n_pass)) for _, _, lst in pv],
                   "field_launch": [lst[0].elapsed_time(lst[1]) for _, _, lst in pv]}
        samples = [[int(v) for v in s[:4]] for s in samples0]
        return total_ms, wall, kms, samples

    for _ in range(max(args.warmup, 3)):
        frame(False); frame(True)
    drv.wait_host()
    drv.launches = 0
    sampler = ClockSampler(local) if rank == 0 else None
    torch.cuda.profiler.start()                                               # `ncu --profile-from-start off` sees exactly the timed frames
    total_ms, wall, _, samples = timed(args.steps, e2e=False)
    torch.cuda.profiler.stop()
    n_launch = drv.launches
    # the roofline kernel's own duration: a separate, shorter loop with an event pair around every field-kernel launch
    # (kept out of the headline loop so that the instrumentation cannot perturb it)
    _, _, kms, _ = timed(min(args.steps, 10), e2e=False, with_prof=True)
    e2e_ms, e2e_wall, _, _ = timed(args.steps, e2e=True)
    clocks = sampler.stop() if sampler else None

    # ---- stand-alone kernel rooflines on rank 0 (hash microbench = BASELINE.json configs[4], MLP pass)
    extra = {}
    hbm, tf, src = measured_peaks()
    if rank == 0:
        from pienerf_b200.gridencoder import grid_encode
        B = 1 << 22
        g = torch.Generator(device=dev).manual_seed(0)
        pts = torch.rand(B, 3, device=dev, generator=g)
        enc = model.encoder

        def time_kernel(fn, reps=5):
            fn(); torch.cuda.synchronize()
            best = []
            for _ in range(reps):
                flush.fill_(1.0)
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                best.append(a.elapsed_time(b))
            return float(np.mean(best))
        outbuf = torch.empty(16, B, 2, device=dev)
        from pienerf_b200 import _gridencoder
        S = float(np.log2(enc.per_level_scale))
        def grid_ms(x):
            return time_kernel(lambda: _gridencoder.grid_encode_forward(x, enc.embeddings.data, enc.offsets, outbuf, B, 3, 2, 16, S, 16, None, 0, False, 0))
        # "warped samples" (BASELINE.json configs[4]): what the renderer feeds the encoder — 2^15 rays x 128 consecutive samples,
        # 0.0017 apart in table coordinates (= the chair's dt_min 0.0034 in [-1,1]); and the worst case, uniform random points
        nr = B // 128
        o = torch.rand(nr, 1, 3, device=dev, generator=g) * 0.5 + 0.1
        dd = torch.nn.functional.normalize(torch.rand(nr, 1, 3, device=dev, generator=g) + 0.1, dim=-1)
        coh = (o + dd * (torch.arange(128, device=dev).view(1, 128, 1) * 0.0017)).reshape(B, 3).contiguous().clamp(0, 1)
        ms_coh, ms_rand = grid_ms(coh), grid_ms(pts)
        gbs_c, gbs_r = (B * ALGO_BYTES_PER_SAMPLE_GRID / (m * 1e-3) / 1e9 for m in (ms_coh, ms_rand))
        extra["hash_microbench"] = {"samples": B, "ms": ms_coh, "kernel": "grid_forward_d3c2 (stand-alone grid_encode_forward)",
                                    "roofline": {"bound": "hbm", "achieved": gbs_c, "peak": hbm, "unit": "GB/s", "frac": gbs_c / hbm, "traffic": None,
                                                 "peak_source": src, "inputs": "ray-coherent warped samples: 2^15 rays x 128 samples, 0.0017 apart"},
                                    "uniform_random": {"ms": ms_rand, "achieved": gbs_r, "frac": gbs_r / hbm, "inputs": "uniform random in [0,1]^3 (no reuse between lanes)"}}
        M = 1 << 21
        xs = (torch.rand(M, 3, device=dev, generator=g) * 2 - 1); ds = torch.nn.functional.normalize(torch.randn(M, 3, device=dev, generator=g), dim=-1)
        ms_field = time_kernel(lambda: model.forward_fused(xs, ds, mode=0))
        ms_field_tc = time_kernel(lambda: model.forward_fused(xs, ds, mode=1))
        # tensor work actually issued: 3 bf16 MMAs per GEMM on padded tiles (20480 MAC/sample x 3), vs the algorithmic 18688 FLOP
        tfl = M * MLP_FLOP_PER_SAMPLE / (ms_field_tc * 1e-3) / 1e12
        extra["field_pass"] = {"samples": M, "ms_fp32_simt": ms_field, "ms_tcgen05": ms_field_tc,
                               "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf, "unit": "TFLOP/s", "frac": tfl / tf, "traffic": None, "peak_source": src,
                                            "note": "algorithmic 18688 FLOP/sample; kernel = hash encode + 5-layer MLP (bf16x3 split, 3 MMAs per GEMM); gather-bound, not tensor-bound"}}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = args.steps
    fps = K / (total_ms * 1e-3)
    render_ms = float(np.mean(kms["render"])); field_ms = float(np.mean(kms["field"])); field0_ms = float(np.mean(kms["field_launch"]))
    st = np.mean(np.asarray(samples, dtype=np.float64), axis=0)           # composited, rays hit, field evaluations, field rows
    samp, evaluated, rows = float(st[0]), float(st[2]), float(st[3])
    achieved = evaluated * ALGO_BYTES_PER_SAMPLE_FUSED / (field_ms * 1e-3) / 1e9
    line = {
        "metric": f"simulated+rendered frames/s at {W}x{H}", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 render / f64 sim",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": f"{args.config}: {sim.n_ip}-IP Q-GMLS body ({sim.n_k} kernels, sim_iters {sim.iters}) + {W}x{H} deformed render, "
                               f"random-init 16-level hash grid + 64-wide MLP, density_scale {args.density_scale}, num_seek_IP {opt.num_seek_IP}",
                   "rays": N, "n_ip": sim.n_ip, "kept_samples_per_frame": samp, "field_evaluations_per_frame": evaluated,
                   "parallelism": f"16x16 ray tiles over {world} GPU(s), simulator on rank 0" + (f", sim-aware tile weights {[round(x, 4) for x in calib['weights']]}" if calib else ""),
                   "l2": "flushed between timed frames (256 MiB fill)"},
        "e2e": {"value": K / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 64 + 16, "d2h_bytes_per_step": N * 5 * 4,
                "wall_fps": K / e2e_wall, "api": "pienerf_b200.frame.DistFrameDriver.frame(): host pose -> pn_get_rays -> sim step -> (bcast) -> pn_render_deformed -> (gather) -> async copy to pinned host frame (double buffered)"},
        "gpu_launches": n_launch,
        "roofline": {"kernel": "wave_field_ws_kernel (16-level hash-grid gather + tcgen05 MLP over 128-row sample tiles; one launch per wavefront pass)",
                     "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": 96.1e6, "traffic_note": "dram__bytes_read+write of the first-pass launch (2.15 M rows, 2.2 GB algorithmic) in profiles/r1_ncu_summary.txt; the 46.7 MiB table is L2-resident",
                     "peak_source": src, "kernel_ms_per_frame": field_ms, "launches_per_frame": n_pass, "first_pass_launch_ms": field0_ms,
                     "algorithmic_bytes": f"{ALGO_BYTES_PER_SAMPLE_FUSED} B/sample x {evaluated:.0f} field evaluations per frame (summed over the frame's launches; rows incl. slab padding: {rows:.0f})",
                     "share_of_step": field_ms / (total_ms / K), "render_passes_ms_per_frame": render_ms},
        "clocks": clocks, "wall_fps": K / wall, "host_enqueue_ms_per_frame": 1e3 * enqueue_s[0] / max(enqueue_n[0], 1),
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, budget_s=args.cpu_budget)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(args, budget_s=20.0):
    """The oracle restatement timed on the host cores: full Q-GMLS step (C, OpenMP) + the numpy renderer on every
    `stride`-th pixel of the same frame, scaled to whole frames.  A reported baseline, not the target."""
    from oracle import render_oracle as ro
    from oracle.sim_oracle import OracleSimulator
    from pienerf_b200.synthetic import CONFIGS, make_body, make_field, occupancy_bitfield, orbit_intrinsics, orbit_pose
    cfg = CONFIGS[args.config]
    body = make_body(cfg["body"], dx=cfg["sim_dx"], bound=cfg["bound"])
    orc = OracleSimulator(dt=1e-2, iters=10, bbox=[2 * cfg["bound"]] * 3, dx=cfg["sim_dx"], stiff=1e5, base=[-cfg["bound"]] * 3)
    orc.initialize(body["pos"], body["mass"], body["mu"], body["lam"], body["pin"])
    orc.stepforward()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < budget_s * 0.3 and n < 200:
        orc.stepforward(); n += 1
    sim_s = (time.perf_counter() - t0) / max(n, 1)
    field = make_field(bound=cfg["bound"]); bits = occupancy_bitfield(body["pos"], 0.6 * cfg["sim_dx"], bound=cfg["bound"])
    pose = orbit_pose(radius=cfg["radius"]); intr = orbit_intrinsics(cfg["W"], cfg["H"], cfg["fovy"])
    rays_o, rays_d = ro.get_rays(pose, intr, cfg["H"], cfg["W"])
    stride = args.cpu_stride
    sel = (np.arange(cfg["H"])[::stride, None] * cfg["W"] + np.arange(cfg["W"])[None, ::stride]).reshape(-1)
    pos, F, dF = orc.get_IP_info()
    t0 = time.perf_counter()
    ro.rund_cuda(ro.OracleField(field), rays_o[sel], rays_d[sel], pos, orc.IP_pos.astype(np.float32), F, dF, 1.05 * cfg["sim_dx"], bits, cfg["bound"], 1,
                 min_near=cfg["min_near"], density_scale=args.density_scale, dt_gamma=cfg["dt_gamma"], max_steps=cfg["max_steps"], T_thresh=cfg["T_thresh"],
                 max_iter_num=cfg["max_iter_num"], hash_grid_size=1.2 * cfg["sim_dx"], num_seek_IP=cfg["num_seek_IP"])
    render_s = (time.perf_counter() - t0) * (cfg["W"] * cfg["H"] / sel.size)
    return {"value": 1.0 / (sim_s + render_s), "unit": "frames/s", "cores": OracleSimulator.threads(), "kind": "port",
            "sample": f"sim: {n} full stepforward() calls of oracle/sim_oracle.c ({OracleSimulator.threads()} OpenMP threads, {sim_s * 1e3:.1f} ms/step); "
                      f"render: numpy oracle (1 thread) on every {stride}th pixel in x and y ({sel.size} of {cfg['W'] * cfg['H']} rays), time scaled x{cfg['W'] * cfg['H'] / sel.size:.0f} "
                      f"({render_s:.1f} s/frame)",
            "sim_step_ms": sim_s * 1e3, "render_frame_s": render_s}


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own code path for the same frame: its CUDA kernels compiled unmodified for sm_100a
    (oracle/_ref) driven by its rund_cuda loop + fp32 nn.Linear MLP (oracle/ref_renderer.py), and — because Warp
    is not installable here — the fp64 C restatement of its simulator on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.sim_oracle import OracleSimulator
    from pienerf_b200.synthetic import CONFIGS, make_body, make_field, occupancy_bitfield, orbit_intrinsics, orbit_pose
    cfg = CONFIGS[args.config]
    K, Wm = args.steps, max(args.warmup, 3)
    body = make_body(cfg["body"], dx=cfg["sim_dx"], bound=cfg["bound"])
    orc = OracleSimulator(dt=1e-2, iters=10, bbox=[2 * cfg["bound"]] * 3, dx=cfg["sim_dx"], stiff=1e5, base=[-cfg["bound"]] * 3)
    orc.initialize(body["pos"], body["mass"], body["mu"], body["lam"], body["pin"])
    field = make_field(bound=cfg["bound"]); bits = occupancy_bitfield(body["pos"], 0.6 * cfg["sim_dx"], bound=cfg["bound"])
    pose = orbit_pose(radius=cfg["radius"]); intr = orbit_intrinsics(cfg["W"], cfg["H"], cfg["fovy"])
    common = {"metric": f"simulated+rendered frames/s at {cfg['W']}x{cfg['H']}", "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K,
              "warmup": Wm, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 render / f64 sim", "data": "synthetic",
              "impl": "reference"}
    have_gpu = torch.cuda.is_available()
    try:
        from oracle.ref_renderer import ReferenceRenderer
        ref = ReferenceRenderer(field, bits, bound=cfg["bound"], density_scale=args.density_scale, min_near=cfg["min_near"]) if have_gpu else None
    except Exception as e:                                                    # no prebuilt reference kernels: CPU port on a bounded sample
        ref = None
        why = str(e)
    if ref is None:
        cb = cpu_baseline(args, budget_s=30.0)
        line = dict(common, value=cb["value"], ms_per_step=1e3 / cb["value"], cpu_baseline=cb,
                    config={"workload": f"{args.config} (oracle port on a bounded sample: reference kernels unavailable)"},
                    e2e={"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return
    from oracle import render_oracle as ro
    dev = torch.device("cuda")
    rays_o, rays_d = ro.get_rays(pose, intr, cfg["H"], cfg["W"])
    rays_o = torch.from_numpy(rays_o).to(dev); rays_d = torch.from_numpy(rays_d).to(dev)
    p_ori = torch.from_numpy(orc.IP_pos.astype(np.float32)).to(dev)
    kw = dict(dt_gamma=cfg["dt_gamma"], max_steps=cfg["max_steps"], T_thresh=cfg["T_thresh"], max_iter_num=cfg["max_iter_num"],
              hash_grid_size=1.2 * cfg["sim_dx"], cut=cfg["cut"], cut_bounds=tuple(cfg.get("cut_bounds", [0.0] * 6)), num_seek_IP=cfg["num_seek_IP"])
    sim_t = []

    def frame():
        t0 = time.perf_counter()
        pos, F, dF = orc.get_IP_info()                                        # trainer.py:303-308
        orc.stepforward()
        sim_t.append(time.perf_counter() - t0)
        out = ref.rund_cuda(rays_o, rays_d, torch.from_numpy(pos).to(dev), p_ori, torch.from_numpy(F).to(dev), torch.from_numpy(dF).to(dev),
                            1.05 * cfg["sim_dx"], return_stats=True, **kw)
        img = out["image"].cpu().numpy(); out["depth"].cpu().numpy(); out["depth_0"].cpu().numpy()   # trainer.py:589-593
        return out, img
    for _ in range(Wm):
        frame()
    sim_t.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        out, img = frame()
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    fps = K / total
    line = dict(common, value=fps, ms_per_step=total / K * 1e3,
                config={"workload": f"{args.config}: reference CUDA kernels (oracle/_ref, sm_100a build) in the reference rund_cuda loop + fp32 nn.Linear MLP; "
                                    f"simulator = fp64 C restatement on host cores (Warp unavailable); density_scale {args.density_scale}",
                        "kept_samples_per_frame": out["n_samples"], "loop_iterations": out["iters"],
                        "timing": "wall clock between device synchronizes (the CPU sim step is part of the frame)"},
                cpu_baseline={"value": fps, "unit": "frames/s", "cores": OracleSimulator.threads(), "kind": "reference",
                              "sample": f"{K} whole frames; sim step {np.mean(sim_t) * 1e3:.1f} ms on {OracleSimulator.threads()} host threads, rest = reference GPU render loop"},
                e2e={"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner, torch warnings) write to fd 1 too, so
    # everything goes to stderr while the run lasts and the line is printed on the real stdout at the end
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="chair")
    ap.add_argument("--density-scale", type=float, default=1.0, dest="density_scale")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap-sim", action="store_true", help="run the simulator step on the render stream instead of a concurrent side stream")
    ap.add_argument("--equal-tiles", action="store_true", help="N>1: equal tile shares instead of sim-aware weights")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--cpu-stride", type=int, default=16)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
