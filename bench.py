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
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi takes a few hundred ms to
    deliver its first sample and the timed region is ~0.1 s, so the sampler is started when the bench process starts and the samples
    are selected afterwards by their timestamps: those inside [begin(), end()], else the three closest to it (and the line says so)."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.t0 = self.t1 = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        import atexit
        atexit.register(self._kill)                                           # never leave the sampler behind if the bench dies early

    def _kill(self):
        try:
            if self.p is not None and self.p.poll() is None:
                self.p.kill()
        except Exception:
            pass

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    @classmethod
    def parse(cls, text, t0=None, t1=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        rows = []
        for line in text.splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 8:
                continue
            try:
                sm, mx = float(t[1]), float(t[2])
            except ValueError:
                continue
            try:
                import datetime
                ts = datetime.datetime.strptime(t[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()      # local time, like time.time() once converted
            except Exception:
                ts = None
            rows.append((ts, sm, mx, {n for n, v in zip(cls.REASONS, t[4:8]) if v.lower().startswith("active")}))
        if not rows:
            return out
        sel, window = rows, "every sample of the bench process (no usable timestamps)"
        stamped = [r for r in rows if r[0] is not None]
        if stamped and t0 is not None and t1 is not None:
            inside = [r for r in stamped if t0 <= r[0] <= t1]
            if inside:
                sel, window = inside, "timed region"
            else:
                near = sorted(stamped, key=lambda r: max(t0 - r[0], r[0] - t1, 0.0))[:3]
                sel, window = near, "the 3 samples closest to the timed region (none fell inside it)"
        return {"sm_mhz": float(np.median([r[1] for r in sel])), "sm_max_mhz": float(max(r[2] for r in sel)),
                "reasons": sorted(set().union(*[r[3] for r in sel])), "samples": len(sel), "window": window}

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.05)                                                      # let the sample that covers the end of the window land
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        try:
            self.f.flush(); self.f.seek(0)
            out = self.parse(self.f.read(), self.t0, self.t1)
        except Exception:
            pass
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def workload_config(args, n_ip, n_k, iters, W, H, num_seek_IP):
    """`config` of the JSON line: the SAME dict in both arms (what is measured); how each arm runs it goes into `details`."""
    return {"workload": f"{args.config}: {n_ip}-IP Q-GMLS body ({n_k} kernels, sim_iters {iters}) + {W}x{H} deformed render, random-init 16-level hash grid "
                        f"+ 64-wide MLP, density_scale {args.density_scale}, num_seek_IP {num_seek_IP}", "rays": W * H, "n_ip": n_ip}


# ------------------------------------------------------------------------------------------------- ours
def hash_microbench(model, dev, rank, world, flush, hbm, src):
    """BASELINE.json configs[4]: 2^22 samples x 16 levels through the stand-alone grid_encode_forward, samples split evenly
    over the ranks (table replicated).  Returns this rank's (ms_coherent, ms_random, samples)."""
    import torch
    from pienerf_b200 import _gridencoder
    B = (1 << 22) // world
    g = torch.Generator(device=dev).manual_seed(rank)
    enc = model.encoder
    S = float(np.log2(enc.per_level_scale))
    outbuf = torch.empty(16, B, 2, device=dev)
    pts = torch.rand(B, 3, device=dev, generator=g)
    # "warped samples": what the renderer feeds the encoder — rays x 128 consecutive samples, 0.0017 apart in table
    # coordinates (= the chair's dt_min 0.0034 in [-1,1]); and the worst case, uniform random points
    nr = B // 128
    o = torch.rand(nr, 1, 3, device=dev, generator=g) * 0.5 + 0.1
    dd = torch.nn.functional.normalize(torch.rand(nr, 1, 3, device=dev, generator=g) + 0.1, dim=-1)
    coh = (o + dd * (torch.arange(128, device=dev).view(1, 128, 1) * 0.0017)).reshape(B, 3).contiguous().clamp(0, 1)

    def run(x, reps=5):
        fn = lambda: _gridencoder.grid_encode_forward(x, enc.embeddings.data, enc.offsets, outbuf, B, 3, 2, 16, S, 16, None, 0, False, 0)
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1.0)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))
    return run(coh), run(pts), B


def run_ours(args):
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")               # frame slots + simulator + copy streams each get their own queue
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local) if rank == 0 else None                     # started now, samples selected by timestamp later
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pienerf_b200 import _lib
    from pienerf_b200.frame import build_scene
    from pienerf_b200.pipeline import FramePipeline

    model, sim, opt, pose, intr, body, field = build_scene(args.config, device=dev, density_scale=args.density_scale)
    W, H = opt.W, opt.H
    N = W * H
    weights = None
    share = args.rank0_share
    if world > 1 and share is None:                                           # rank 0 also runs the simulator: measured smaller tile share
        share = FramePipeline.calibrate_rank0_share(model, sim, opt, pose, intr, slots=args.slots, sim_sm_reserve=args.sim_sm_reserve)
    if world > 1 and share < 0.999:
        f0 = share / world
        weights = [f0] + [(1.0 - f0) / (world - 1)] * (world - 1)
    pipe = FramePipeline(model, sim, opt, slots=args.slots, weights=weights, sim_sm_reserve=args.sim_sm_reserve)
    pipe.build(pose, intr)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # > 126 MB L2 (stand-alone kernel timings only)
    n_pass = int(_lib.lib.pn_render_pass_count(int(opt.max_steps)))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- proof that the N-rank frame is the 1-rank frame: 12 frames from the rest state with a drag force, checksum of the last
    check_frames = 12
    if rank == 0:
        sim.update_force(sim.n_ip // 3, torch.tensor([3e4, -1e4, 2e4]))
    for k in range(check_frames):
        slot = pipe.frame(pose, intr, to_host=True)
    checksum = None
    if rank == 0:
        import hashlib
        h = pipe.wait_host(slot)
        img = np.nan_to_num(h["image"].numpy(), nan=-1.0)
        checksum = {"frame": check_frames - 1, "sha1_image": hashlib.sha1(img.tobytes()).hexdigest(), "sum_image": float(img.astype(np.float64).sum()),
                    "sha1_depth_0": hashlib.sha1(h["depth_0"].numpy().tobytes()).hexdigest(),
                    "note": "frame after 11 simulator steps with a drag force, as delivered to the host; identical for every --gpus N"}
        sim.clear_force()
    pipe.drain(); sync_all()

    enqueue_s, enqueue_n = [0.0], [0]

    def timed(K, to_host):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            c0 = time.perf_counter()
            pipe.frame(pose, intr, to_host=to_host)
            enqueue_s[0] += time.perf_counter() - c0
        enqueue_n[0] += K
        pipe.drain()
        e1.record()
        sync_all()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    Wm = max(args.warmup, 3)
    for _ in range(Wm):
        pipe.frame(pose, intr, to_host=False)
    for _ in range(Wm):
        pipe.frame(pose, intr, to_host=True)
    pipe.drain(); sync_all()
    if sampler:
        sampler.begin()
    torch.cuda.profiler.start()                                               # `ncu --profile-from-start off` sees exactly the timed frames
    total_ms, wall = timed(args.steps, to_host=False)
    torch.cuda.profiler.stop()
    stats = np.mean(np.asarray([[int(v) for v in sl["stats"].tolist()] for sl in pipe.slots], dtype=np.float64), axis=0)
    e2e_ms, e2e_wall = timed(args.steps, to_host=True)
    if sampler:
        sampler.end()
    clocks = sampler.stop() if sampler else None
    pipe.check()

    # ---- the roofline kernel's own duration: eager (un-pipelined) frames with an event pair around every field-kernel launch
    # (kept out of the headline loop so that the instrumentation cannot perturb it; L2 flushed before each)
    reps = min(args.steps, 10)
    pv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), [torch.cuda.Event(enable_timing=True) for _ in range(2 * n_pass)])
          for _ in range(reps)]
    for a, b, lst in pv:
        a.record(); b.record()
        for e in lst:
            e.record()
    def eager_profile():
        sync_all()
        for i in range(reps):
            flush.fill_(float(i))
            a, b, lst = pv[i]
            _lib.lib.pn_set_profile_events(_lib.vp(a.cuda_event), _lib.vp(b.cuda_event))
            arr = (_lib.vp * len(lst))(*[e.cuda_event for e in lst])
            _lib.lib.pn_set_profile_event_list(arr, len(lst))
            pipe._frame_body(pipe.slots[i % pipe.S], sync=False)
            _lib.lib.pn_set_profile_events(_lib.vp(0), _lib.vp(0)); _lib.lib.pn_set_profile_event_list(None, 0)
        sync_all()
        return (float(np.mean([a.elapsed_time(b) for a, b, _ in pv])),
                float(np.mean([sum(lst[2 * k].elapsed_time(lst[2 * k + 1]) for k in range(n_pass)) for _, _, lst in pv])),
                float(np.mean([lst[0].elapsed_time(lst[1]) for _, _, lst in pv])))
    render_ms, field_ms, field0_ms = eager_profile()                          # as in the timed loop: grids sized for (SMs - reserve)
    field_ms_all = field_ms
    if pipe.sim_sm_reserve:                                                   # and with every SM, for the kernel's own roofline
        _lib.lib.pn_set_render_sm_reserve(0)
        _, field_ms_all, _ = eager_profile()
        _lib.lib.pn_set_render_sm_reserve(pipe.sim_sm_reserve)

    # ---- stand-alone kernel rooflines (hash microbench on EVERY rank = BASELINE.json configs[4]; MLP pass on rank 0)
    hbm, tf, src = measured_peaks()
    ms_coh, ms_rand, Bm = hash_microbench(model, dev, rank, world, flush, hbm, src) if not args.quick else (1.0, 1.0, 1)
    per_rank = torch.tensor([ms_coh, ms_rand, stats[2], field_ms], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, per_rank)
    else:
        allr = [per_rank]
    allr = np.asarray([t.tolist() for t in allr])
    extra = {}
    if rank == 0 and not args.quick:
        gb = Bm * ALGO_BYTES_PER_SAMPLE_GRID / 1e9
        agg_c, agg_r = world * gb / (allr[:, 0].max() * 1e-3), world * gb / (allr[:, 1].max() * 1e-3)
        extra["hash_microbench"] = {
            "samples": Bm * world, "samples_per_gpu": Bm, "n_gpus": world, "kernel": "grid_forward_d3c2 (stand-alone grid_encode_forward)",
            "ms_per_gpu": [round(float(v), 4) for v in allr[:, 0]],
            "roofline": {"bound": "hbm", "achieved": agg_c, "peak": hbm * world, "unit": "GB/s", "frac": agg_c / (hbm * world), "traffic": None, "peak_source": src,
                         "inputs": "ray-coherent warped samples: rays x 128 samples, 0.0017 apart; aggregate over all GPUs = total bytes / slowest GPU"},
            "uniform_random": {"ms_per_gpu": [round(float(v), 4) for v in allr[:, 1]], "achieved": agg_r, "frac": agg_r / (hbm * world),
                               "inputs": "uniform random in [0,1]^3 (no reuse between lanes)"},
            "reference_kernel": "timed on the same inputs by `bench.py --impl reference` (key hash_microbench_reference)"}
        M = 1 << 21
        g = torch.Generator(device=dev).manual_seed(0)
        xs = (torch.rand(M, 3, device=dev, generator=g) * 2 - 1); ds = torch.nn.functional.normalize(torch.randn(M, 3, device=dev, generator=g), dim=-1)

        def time_kernel(fn, reps=5):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush.fill_(1.0)
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.mean(ts))
        ms_field = time_kernel(lambda: model.forward_fused(xs, ds, mode=0))
        ms_field_tc = time_kernel(lambda: model.forward_fused(xs, ds, mode=1))
        tfl = M * MLP_FLOP_PER_SAMPLE / (ms_field_tc * 1e-3) / 1e12
        extra["field_pass"] = {"samples": M, "ms_fp32_simt": ms_field, "ms_tcgen05": ms_field_tc,
                               "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf, "unit": "TFLOP/s", "frac": tfl / tf, "traffic": None, "peak_source": src,
                                            "note": "algorithmic 18688 FLOP/sample; kernel = hash encode + 5-layer MLP (bf16x3 split, 3 MMAs per GEMM); gather-bound, not tensor-bound"}}
        # the MLP pass by itself: features in, sigma / rgb out, through the frame renderer's tcgen05 pipeline; timed with a CUDA
        # event pair recorded by the library right around the tensor-core kernel
        enc = torch.randn(M, 32, device=dev, generator=g)
        model.mlp_only(enc, ds); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.fill_(1.0)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); b.record()
            arr = (_lib.vp * 2)(a.cuda_event, b.cuda_event)
            _lib.lib.pn_set_profile_event_list(arr, 2)
            model.mlp_only(enc, ds)
            _lib.lib.pn_set_profile_event_list(None, 0)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms_mlp = float(np.mean(ts))
        tfm = M * MLP_FLOP_PER_SAMPLE / (ms_mlp * 1e-3) / 1e12
        extra["mlp_pass"] = {"samples": M, "ms": ms_mlp, "kernel": "wave_field_ws_kernel fed with pre-encoded features (pn_mlp_forward): 5 layers on tcgen05, activations in TMEM",
                             "roofline": {"bound": "tensor", "achieved": tfm, "peak": tf, "unit": "TFLOP/s", "frac": tfm / tf, "traffic": None, "peak_source": src,
                                          "issued_tflops": tfm * 3 * 20480 / 18688,
                                          "note": "achieved counts the algorithmic 18688 FLOP/sample; issued = 3 bf16 MMAs per GEMM (hi/lo split for fp32-level accuracy) on 32/64/16-padded tiles"}}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = args.steps
    fps = K / (total_ms * 1e-3)
    evaluated_all = float(allr[:, 2].sum())
    samp, evaluated, rows = float(stats[0]), float(stats[2]), float(stats[3])
    achieved = evaluated * ALGO_BYTES_PER_SAMPLE_FUSED / (field_ms * 1e-3) / 1e9
    traffic = ncu_traffic(args.config, world)
    line = {
        "metric": f"simulated+rendered frames/s at {W}x{H}", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 render / f64 sim",
        "data": "synthetic", "impl": "ours",
        "config": workload_config(args, sim.n_ip, sim.n_k, sim.iters, W, H, opt.num_seek_IP),
        "details": {"kept_samples_per_frame_rank0": samp, "field_evaluations_per_frame": evaluated_all,
                    "parallelism": f"16x16 ray tiles over {world} GPU(s), simulator on rank 0; {pipe.S} frames in flight per GPU (one CUDA graph per rank-frame); "
                                   + ("IP state pushed and pixels returned by peer-memory stores over NVLink (no collective in the frame loop)" if world > 1 else "single GPU")
                                   + (f"; tile shares {[round(x, 4) for x in weights]}" if weights else "")
                                   + (f"; persistent render grids sized for {pipe.sim_sm_reserve} SMs fewer than the GPU has (room for the small kernels of the frames in flight)" if pipe.sim_sm_reserve else ""),
                    "l2": f"no flush in the timed loop: inputs larger than L2 — {pipe.S} frame slots rotate, each with its own 46.7 MiB copy of the hash table, "
                          "its own sample lists and rays (per-frame traffic on rank 0 ~ %.0f MB); stand-alone kernel timings flush with a 256 MiB fill" % (rows * 40 / 1e6 + 46.7)},
        "e2e": {"value": K / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 80, "d2h_bytes_per_step": N * 5 * 4,
                "wall_fps": K / e2e_wall, "api": "pienerf_b200.pipeline.FramePipeline.frame(): pinned host pose -> H2D -> sim state + step -> (peer push) -> rays -> "
                                                 "pn_render_deformed_ex -> (peer pixel stores) -> async D2H into pinned host frames (one per slot)"},
        "gpu_launches": int(pipe.launches_per_frame * K),
        "frame_checksum": checksum,
        "roofline": {"kernel": "wave_field_ws_kernel (16-level hash-grid gather + tcgen05 MLP over 128-row sample tiles; one launch per wavefront pass)",
                     "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": traffic["bytes"], "traffic_note": traffic["note"],
                     "peak_source": src, "kernel_ms_per_frame": field_ms, "launches_per_frame": n_pass, "first_pass_launch_ms": field0_ms,
                     "algorithmic_bytes": f"{ALGO_BYTES_PER_SAMPLE_FUSED} B/sample x {evaluated:.0f} field evaluations per frame on rank 0 (summed over the frame's launches; rows incl. slab padding: {rows:.0f})",
                     "measured": "eager un-pipelined frames, CUDA events around every field-kernel launch, L2 flushed before each frame",
                     "share_of_eager_frame": field_ms / render_ms, "render_passes_ms_per_frame": render_ms,
                     "sms_used": f"{148 - pipe.sim_sm_reserve} of 148 (the frame pipeline keeps {pipe.sim_sm_reserve} SMs free for the small kernels of the frames in flight)",
                     "all_sms": {"kernel_ms_per_frame": field_ms_all, "achieved": evaluated * ALGO_BYTES_PER_SAMPLE_FUSED / (field_ms_all * 1e-3) / 1e9,
                                 "frac": evaluated * ALGO_BYTES_PER_SAMPLE_FUSED / (field_ms_all * 1e-3) / 1e9 / hbm}},
        "clocks": clocks, "wall_fps": K / wall, "host_enqueue_ms_per_frame": 1e3 * enqueue_s[0] / max(enqueue_n[0], 1),
    }
    line.update(extra)
    if world == 1 and not args.quick and args.config == "chair":
        pipe.close()
        line["other_configs"] = other_configs(args, dev)
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        line["cpu_baseline"] = cpu_baseline(args, budget_s=args.cpu_budget)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_configs(args, dev, frames=12):
    """Short pipelined runs (N = 1) of the other BASELINE.json configs and of the variants the headline hides, so that the driver's
    record carries them: the opaque field (density_scale 50: early termination, what a trained model looks like), the chair-like
    body (100-250 kernels: the realistic global solve), trex (configs[2]) and the 1080p / 4k-IP frame (configs[3])."""
    import gc

    import torch
    from pienerf_b200.frame import build_scene
    from pienerf_b200.pipeline import FramePipeline
    out = {}
    for name, config, ds in (("chair_opaque", "chair", 50.0), ("chairlike", "chairlike", 1.0), ("trex", "trex", 1.0), ("synth1080", "synth1080", 1.0)):
        model, sim, opt, pose, intr, body, field = build_scene(config, device=dev, density_scale=ds)
        pipe = FramePipeline(model, sim, opt, slots=args.slots)
        pipe.build(pose, intr)
        for _ in range(4):
            pipe.frame(pose, intr, to_host=True)
        pipe.drain(); torch.cuda.synchronize()
        res = {}
        for key, to_host in (("value", False), ("e2e", True)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(frames):
                pipe.frame(pose, intr, to_host=to_host)
            pipe.drain(); e1.record(); torch.cuda.synchronize()
            res[key] = frames / (e0.elapsed_time(e1) * 1e-3)
        st = pipe.check()[0]
        out[name] = {"frames_per_s": res["value"], "e2e_frames_per_s": res["e2e"], "config": f"{config}: {opt.W}x{opt.H}, {sim.n_ip} IPs, {sim.n_k} kernels, "
                     f"density_scale {ds}, num_seek_IP {opt.num_seek_IP}", "kept_samples_per_frame": st[0], "field_evaluations_per_frame": st[2],
                     "sim_launches_per_step": sim.step_launches, "passes": pipe.max_passes, "frames_timed": frames}
        pipe.close()
        del pipe, model, sim
        gc.collect(); torch.cuda.empty_cache()
    return out


def ncu_traffic(config, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first-pass field-kernel launch, from the committed `ncu --set full`
    capture of this config (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py).  None when there is no capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world == 1 and os.path.exists(p):
        d = json.load(open(p)).get(config)
        if d:
            return {"bytes": d["dram_bytes"], "note": f"first-pass launch ({d['rows']} rows), {d['source']}; the 46.7 MiB table is L2-resident within a frame"}
    return {"bytes": None, "note": "no committed ncu capture for this config / GPU count"}


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
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the reference arm's host-side simulator must still get all cores
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    import torch
    from oracle.sim_oracle import OracleSimulator
    OracleSimulator.set_threads(ncores)
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
    why = "no CUDA device"
    try:
        from oracle.ref_renderer import ReferenceRenderer
        ref = ReferenceRenderer(field, bits, bound=cfg["bound"], density_scale=args.density_scale, min_near=cfg["min_near"]) if have_gpu else None
    except Exception as e:                                                    # no prebuilt reference kernels: CPU port on a bounded sample
        ref = None
        why = str(e)
    if ref is None:
        cb = cpu_baseline(args, budget_s=30.0)
        line = dict(common, value=cb["value"], ms_per_step=1e3 / cb["value"], cpu_baseline=cb,
                    config=workload_config(args, orc.n_ip, orc.n_k, 10, cfg["W"], cfg["H"], cfg["num_seek_IP"]),
                    details={"arm": "oracle port on a bounded sample: reference kernels (oracle/_ref) unavailable: " + why},
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
                            1.05 * cfg["sim_dx"], return_stats=False, **kw)
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
    # sample / iteration counts of this frame: one extra UNTIMED frame (the per-iteration count costs a host sync each)
    pos, F, dF = orc.get_IP_info()
    out = ref.rund_cuda(rays_o, rays_d, torch.from_numpy(pos).to(dev), p_ori, torch.from_numpy(F).to(dev), torch.from_numpy(dF).to(dev),
                        1.05 * cfg["sim_dx"], return_stats=True, **kw)
    # the reference's own kernel_grid (gridencoder.cu:88-245) on the inputs of our hash microbench (BASELINE.json configs[4])
    B = 1 << 22
    g = torch.Generator(device=dev).manual_seed(0)
    pts = torch.rand(B, 3, device=dev, generator=g)
    nr = B // 128
    o_ = torch.rand(nr, 1, 3, device=dev, generator=g) * 0.5 + 0.1
    dd = torch.nn.functional.normalize(torch.rand(nr, 1, 3, device=dev, generator=g) + 0.1, dim=-1)
    coh = (o_ + dd * (torch.arange(128, device=dev).view(1, 128, 1) * 0.0017)).reshape(B, 3).contiguous().clamp(0, 1)
    outbuf = torch.empty(16, B, 2, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def ref_grid_ms(x):
        fn = lambda: ref.ge.grid_encode_forward(x, ref.emb, ref.offsets, outbuf, B, 3, 2, 16, ref.S, ref.base_res, None, 0, False, 0)
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.fill_(1.0)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))
    hbm, _, src = measured_peaks()
    mc, mr = ref_grid_ms(coh), ref_grid_ms(pts)
    hash_ref = {"kernel": "reference kernel_grid<float,3,2> (oracle/_ref/_ref_gridencoder.so, unmodified source, sm_100a build)", "samples": B,
                "ms": mc, "achieved": B * ALGO_BYTES_PER_SAMPLE_GRID / (mc * 1e-3) / 1e9, "frac": B * ALGO_BYTES_PER_SAMPLE_GRID / (mc * 1e-3) / 1e9 / hbm,
                "uniform_random": {"ms": mr, "achieved": B * ALGO_BYTES_PER_SAMPLE_GRID / (mr * 1e-3) / 1e9, "frac": B * ALGO_BYTES_PER_SAMPLE_GRID / (mr * 1e-3) / 1e9 / hbm},
                "peak": hbm, "unit": "GB/s", "peak_source": src, "inputs": "same generators as the `ours` arm's hash_microbench at 1 GPU"}
    line = dict(common, value=fps, ms_per_step=total / K * 1e3, hash_microbench_reference=hash_ref,
                config=workload_config(args, orc.n_ip, orc.n_k, 10, cfg["W"], cfg["H"], cfg["num_seek_IP"]),
                details={"arm": "reference CUDA kernels (oracle/_ref, unmodified source, sm_100a build) in the reference rund_cuda loop + fp32 nn.Linear MLP; "
                                "simulator = fp64 C restatement on host cores (Warp unavailable)",
                         "kept_samples_per_frame": out["n_samples"], "loop_iterations": out["iters"],
                         "timing": "wall clock between device synchronizes (the CPU sim step is part of the frame)"},
                cpu_baseline={"value": fps, "unit": "frames/s", "cores": OracleSimulator.threads(), "kind": "reference",
                              "sample": f"{K} whole frames; sim step {np.mean(sim_t) * 1e3:.1f} ms on {OracleSimulator.threads()} host threads, rest = reference GPU render loop"},
                e2e={"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner, torch warnings) write to fd 1 too, so
    # everything goes to stderr while the run lasts and the line is printed on the real stdout at the end
    # watchdog: a hang (a wedged kernel, a lost rank) must not eat the box: dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("PN_BENCH_WATCHDOG_S", "480")), exit=True)
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
    ap.add_argument("--slots", type=int, default=3, help="frames in flight per GPU (each with its own workspace and hash-table copy)")
    ap.add_argument("--rank0-share", type=float, default=None, dest="rank0_share", help="N>1: rank 0's tile share relative to an equal split (default: measured, FramePipeline.calibrate_rank0_share)")
    ap.add_argument("--sim-sm-reserve", type=int, default=None, dest="sim_sm_reserve", help="SMs every rank keeps out of its persistent render grids for the small kernels of the frames in flight (default 8 for N<=2, else 16)")
    ap.add_argument("--quick", action="store_true", help="skip the stand-alone kernel microbenches and the CPU baseline (development runs)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--cpu-stride", type=int, default=16)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
