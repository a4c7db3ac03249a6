"""The frame loop as a device-side pipeline: several frames in flight per GPU, every per-rank frame one CUDA graph, and —
on N GPUs of one NVLink node — the frame's two exchanges done by stores into peer memory instead of collectives.

What a frame is (reference order, nerf/trainer.py:284-329,531-602): read the IP state, advance the simulator, render the
state that was read BEFORE the step, hand the frame to the host.  Data flow here (rank 0 owns the simulator):

    rank 0, sim stream    [cam H2D] -> ip_info -> (wait: peers finished the slot's previous frame) -> pn_peer_put of
                          [cam | pos | F | dF] into every rank's state slot -> raise their `state` flags -> stepforward
    every rank, slot s    wait `state` flag (own memory) -> rays of its tiles -> pn_render_deformed_ex: the compositor
                          stores finished pixels straight into rank 0's frame slot over NVLink -> raise `done` flag on rank 0
    rank 0, copy stream   wait all `done` flags -> frame slot -> pinned host buffer (async, double buffered by slot)

Frames k, k+1, ... use slots k % S on their own streams, so the latency-bound parts of a frame (the small preparation
kernels, near-empty late passes, the flag waits, launch gaps) overlap the bulk of its neighbours; what bounds the frame
rate is the sum of kernel work per GPU, not the length of the dependency chain.  Every slot has its own workspace and
its own copy of the hash table (S x 46.7 MiB > L2), so no frame finds its inputs L2-warm from the previous one.
NCCL / torch.distributed is only used to exchange the 64-byte IPC handles at set-up.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import FrameIoT, check, dptr, lib, stream_ptr, vp


class _CudaView:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerBlock:
    """One cudaMalloc allocation that other processes on the node can map (CUDA IPC)."""

    def __init__(self, nbytes, device):
        self.nbytes = int((nbytes + 255) // 256 * 256)
        self.device = device
        p = vp()
        check(lib.pn_peer_alloc(self.nbytes, C.byref(p)))
        self.ptr = int(p.value)
        self.bytes = torch.as_tensor(_CudaView(self.ptr, self.nbytes), device=device)
        self._opened = []

    def handle(self):
        buf = C.create_string_buffer(64)
        check(lib.pn_peer_export(vp(self.ptr), buf))
        return buf.raw

    def view(self, offset, dtype, shape):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.bytes[offset:offset + n].view(dtype).view(*shape)

    @staticmethod
    def open(handle):
        p = vp()
        check(lib.pn_peer_open(handle, C.byref(p)))
        return int(p.value)

    def free(self):
        if self.ptr:
            self.bytes = None
            check(lib.pn_peer_free(vp(self.ptr)))
            self.ptr = 0


def green_streams(device_index, sim_sms, n_render_streams):
    """SM partition of one GPU with CUDA green contexts (driver API, CUDA >= 12.4): `sim_sms` SMs for the simulator's stream, the
    rest for the render streams.  Kernels launched into a green context's stream only run on its SMs, so the simulator's many small
    launches never wait for (or share issue slots with) persistent render CTAs.  Returns (sim_stream, [render streams], n_render_sms)
    as torch ExternalStreams, or None when the driver / bindings do not offer it."""
    try:
        from cuda.bindings import driver as drv
    except Exception:
        return None

    def ok(ret):
        if int(ret[0]) != 0:
            raise RuntimeError(f"driver call failed: {ret[0]}")
        return ret[1] if len(ret) == 2 else ret[1:]
    try:
        torch.zeros(1, device=f"cuda:{device_index}")                          # primary context exists and is current
        dev = ok(drv.cuDeviceGet(device_index))
        res = ok(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
        groups, nb, remaining = ok(drv.cuDevSmResourceSplitByCount(1, res, 0, int(sim_sms)))
        d_sim = ok(drv.cuDevResourceGenerateDesc([groups[0]], 1))
        d_ren = ok(drv.cuDevResourceGenerateDesc([remaining], 1))
        flags = drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM
        g_sim = ok(drv.cuGreenCtxCreate(d_sim, dev, flags)); g_ren = ok(drv.cuGreenCtxCreate(d_ren, dev, flags))
        nb_flag = int(drv.CUstream_flags.CU_STREAM_NON_BLOCKING)
        s_sim = ok(drv.cuGreenCtxStreamCreate(g_sim, nb_flag, -1))
        s_ren = [ok(drv.cuGreenCtxStreamCreate(g_ren, nb_flag, 0)) for _ in range(n_render_streams)]
        n_sim = int(groups[0].sm.smCount); n_ren = int(remaining.sm.smCount)
        keep = (g_sim, g_ren, s_sim, s_ren)                                     # contexts / streams live as long as the pipeline
        wrap = lambda h: torch.cuda.ExternalStream(int(h), device=torch.device("cuda", device_index))
        return wrap(s_sim), [wrap(h) for h in s_ren], n_sim, n_ren, keep
    except Exception as e:                                                      # noqa: BLE001 - any failure means "not available here"
        import warnings
        warnings.warn(f"green contexts unavailable ({e}); falling back to stream priorities")
        return None


def dist_backend():
    import torch.distributed as dist
    return dist.get_backend() if dist.is_initialized() else None


def _ptr_array(ptrs, device):
    return torch.tensor([int(p) for p in ptrs], dtype=torch.int64, device=device)


def _al(n, a=256):
    return (int(n) + a - 1) // a * a


class FramePipeline:
    """Pipelined frame driver (see the module docstring).  world == 1 needs no process group.

    frame(pose, intrinsics) enqueues one GUI frame and returns its slot; wait_host(slot) returns the pinned host frame once
    it has landed; drain() joins every stream into the current one.  The simulator steps once per frame unless paused."""

    def __init__(self, model, sim, opt, slots=3, tile=16, weights=None, timeout_ms=20000, table_copies=True, mode=3, sim_sm_reserve=None, lean_passes=True, green=None, merge_passes=None):
        import torch.distributed as dist
        from .dist import tile_partition
        self.model, self.sim, self.opt = model, sim, opt
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dev = next(model.parameters()).device
        self.S, self.W, self.H, self.mode = int(slots), int(opt.W), int(opt.H), mode
        self.timeout_ms = int(timeout_ms)
        # Every rank sizes its persistent march / field grids for (SMs - reserve) SMs: the small latency-bound kernels of the frames in
        # flight (IP-grid preparation, compositor, the simulator's launches on rank 0) then always find a free SM instead of queueing
        # behind CTAs that hold every SM until their kernel ends.  Measured on B200 (profiles/r2_measurements.md): 8 SMs are best for a
        # whole 800x800 frame on one GPU (2.36 -> 2.23 ms with the simulator), 16 when a rank renders 1/8 of it (0.414 -> 0.381 ms).
        if sim_sm_reserve is None:
            sim_sm_reserve = 8 if self.world <= 2 else 16
        self.sim_sm_reserve = int(sim_sm_reserve)
        # hard partition instead of a soft reserve (green contexts): measured slower — the simulator's kernels want the whole GPU for
        # a short time, not 16 SMs for a long time (0.73 ms per step on 16 SMs, 0.57 on 24, 0.21 on all) — kept as an option
        self.green = None
        if green and self.rank == 0 and self.sim_sm_reserve > 0:
            self.green = green_streams(self.dev.index or 0, self.sim_sm_reserve, int(slots))
            if self.green is not None:
                n_all = C.c_int(0)
                check(lib.pn_device_sm_count(C.byref(n_all)))
                self.sim_sm_reserve = int(n_all.value) - self.green[3]          # what the partition really took (granularity of 8)
        check(lib.pn_set_render_sm_reserve(self.sim_sm_reserve))
        self.n_ip = n_ip = int(sim.n_ip) if sim is not None else int(model.p_ori.shape[0])
        npx = self.W * self.H
        dev = self.dev
        # ---- static IP data: rest positions are the same on every rank (main_gui.py:50-56)
        if self.rank == 0:
            p_ori = sim.get_IP_info()[0]
        else:
            p_ori = torch.empty(n_ip, 3, dtype=torch.float32, device=dev)
        if self.world > 1:
            if dist.get_backend() == "nccl":
                dist.broadcast(p_ori, src=0)
            else:                                                               # gloo (two test ranks sharing one GPU): through the host
                h = p_ori.cpu(); dist.broadcast(h, src=0); p_ori = h.to(dev)
        self.p_ori = p_ori
        model.p_ori = p_ori
        model.IP_dx = float(opt.sim_dx) * 1.05
        # ---- peer block: [state slots | state flags | done flags | (rank 0) frame slots]
        self.state_floats = 32 + 39 * n_ip                                    # cam (20 used) | pos | F | dF
        self.state_bytes = _al(self.state_floats * 4, 256)
        self.off_state = 0
        self.off_sflag = self.S * self.state_bytes
        self.off_dflag = self.off_sflag + _al(self.S * 4)
        self.off_frame = self.off_dflag + _al(self.S * self.world * 4)
        self.frame_bytes = _al(npx * 5 * 4)
        total = self.off_frame + (self.S * self.frame_bytes if self.rank == 0 else 0)
        self.block = PeerBlock(total, dev)
        self.peer_ptr = [self.block.ptr] * self.world                          # base address of every rank's block as seen from here
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.block.handle())
            for r in range(self.world):
                if r != self.rank:
                    self.peer_ptr[r] = PeerBlock.open(handles[r])
        # ---- partition of the frame
        self.parts = tile_partition(self.H, self.W, self.world, tile, weights)
        self.n_my = len(self.parts[self.rank])
        self.pix = torch.from_numpy(self.parts[self.rank].astype(np.int32)).to(dev) if self.world > 1 else None
        # ---- per-slot resources
        need = model.workspace_bytes(self.n_my, n_ip, **opt)
        self.slots = []
        for s in range(self.S):
            sl = {}
            st = self.block.view(self.off_state + s * self.state_bytes, torch.float32, (self.state_floats,))
            sl["state"] = st
            sl["cam"] = st[:20]
            sl["pos"] = st[32:32 + 3 * n_ip].view(n_ip, 3); sl["F"] = st[32 + 3 * n_ip:32 + 12 * n_ip].view(n_ip, 9)
            sl["dF"] = st[32 + 12 * n_ip:32 + 39 * n_ip].view(n_ip, 27)
            sl["workspace"] = torch.empty(need, dtype=torch.uint8, device=dev)
            sl["stats"] = torch.zeros(8, dtype=torch.int64, device=dev)
            sl["rays_o"] = torch.empty(self.n_my, 3, dtype=torch.float32, device=dev)
            sl["rays_d"] = torch.empty(self.n_my, 3, dtype=torch.float32, device=dev)
            sl["wsum"] = torch.empty(self.n_my, dtype=torch.float32, device=dev)
            sl["table"] = model.encoder.embeddings.data if (s == 0 or not table_copies) else model.encoder.embeddings.data.clone()
            sl["epoch"] = torch.zeros(1, dtype=torch.int32, device=dev)        # bumped by the slot's frame graph
            sl["state_epoch"] = torch.zeros(1, dtype=torch.int32, device=dev)  # rank 0: bumped by the slot's state graph
            sl["stream"] = self.green[1][s] if self.green is not None else torch.cuda.Stream(device=dev)
            sl["render_done"] = None
            sl["state_ready"] = None
            sl["copy_done"] = None
            # the frame this rank's compositor writes into: rank 0's slot (peer memory for the other ranks)
            fbase = self.peer_ptr[0] + self.off_frame + s * self.frame_bytes
            if self.rank == 0:
                fb = self.block.view(self.off_frame + s * self.frame_bytes, torch.float32, (npx * 5,))
                sl["frame"] = {"image": fb[:3 * npx].view(npx, 3), "depth": fb[3 * npx:4 * npx], "depth_0": fb[4 * npx:5 * npx]}
                sl["cam_host"] = torch.zeros(20, dtype=torch.float32).pin_memory()
                # the frame slot is one contiguous [image | depth | depth_0] run: ONE device-to-host copy per frame into one pinned buffer
                sl["frame_flat"] = fb
                hb = torch.empty(npx * 5, dtype=torch.float32).pin_memory()
                sl["host_flat"] = hb
                sl["host"] = {"image": hb[:3 * npx].view(npx, 3), "depth": hb[3 * npx:4 * npx], "depth_0": hb[4 * npx:5 * npx]}
            sl["frame_ptr"] = (fbase, fbase + 12 * npx, fbase + 16 * npx)
            # flag addresses
            sl["my_state_flag"] = self.block.ptr + self.off_sflag + 4 * s
            sl["state_flags_all"] = _ptr_array([self.peer_ptr[r] + self.off_sflag + 4 * s for r in range(1, self.world)], dev) if self.world > 1 else None
            sl["state_dsts"] = _ptr_array([self.peer_ptr[r] + self.off_state + s * self.state_bytes for r in range(1, self.world)], dev) if self.world > 1 else None
            sl["done_local"] = _ptr_array([self.block.ptr + self.off_dflag + 4 * (s * self.world + r) for r in range(1, self.world)], dev) if self.world > 1 else None
            sl["my_done_flag"] = _ptr_array([self.peer_ptr[0] + self.off_dflag + 4 * (s * self.world + self.rank)], dev)
            sl["wait_arr"] = _ptr_array([sl["my_state_flag"]], dev)
            sl["graph"] = None
            sl["state_graph"] = None
            self.slots.append(sl)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sim_stream = (self.green[0] if self.green is not None else torch.cuda.Stream(device=dev, priority=-1)) if self.rank == 0 else None
        self.copy_stream = torch.cuda.Stream(device=dev) if self.rank == 0 else None
        self.frame_id = 0
        self.max_passes = None
        self.lean_passes = lean_passes
        self.merge_passes = merge_passes
        self.launches_per_frame = 0
        self._warm = False

    # ------------------------------------------------------------------------------------------ graph bodies
    def _io(self, sl):
        io = FrameIoT()
        io.pix = dptr(self.pix) if self.pix is not None else vp(0)
        io.epoch = dptr(sl["epoch"])
        if self.world > 1 and self.rank != 0:
            io.wait_flag, io.n_wait = dptr(sl["wait_arr"]), 1
            io.signal_flag, io.n_signal = dptr(sl["my_done_flag"]), 1
        else:
            io.wait_flag, io.n_wait, io.signal_flag, io.n_signal = vp(0), 0, vp(0), 0
        io.status, io.timeout_ms = dptr(self.status), self.timeout_ms
        return io

    def _frame_body(self, sl, sync=True, lean=False):
        """rays of this rank's tiles + the deformed render into rank 0's frame slot (+ on rank 0: wait for the peers' pixels)."""
        m = self.model
        check(lib.pn_get_rays_pix(dptr(sl["cam"]), self.H, self.W, dptr(self.pix) if self.pix is not None else vp(0), self.n_my,
                                  dptr(sl["rays_o"]), dptr(sl["rays_d"]), stream_ptr()))
        fp = sl["frame_ptr"]
        if self.rank == 0:
            out = {"image": sl["frame"]["image"], "depth": sl["frame"]["depth"], "depth_0": sl["frame"]["depth_0"], "weights_sum": sl["wsum"]}
        else:                                                                   # raw addresses of rank 0's frame slot (peer memory)
            out = {"image": fp[0], "depth": fp[1], "depth_0": fp[2], "weights_sum": sl["wsum"]}
        io = self._io(sl) if sync else None
        if io is None and self.pix is not None:
            io = FrameIoT(); io.pix = dptr(self.pix)
        # lean (the captured graphs): the weight image of the slot's workspace was built by the warm-up call, and only as many passes
        # as the warm-up frames needed (+1, the last one unbounded) are enqueued
        m.render_deformed(sl["rays_o"], sl["rays_d"], out=out, workspace=sl["workspace"], stats=sl["stats"], io=io, embeddings=sl["table"],
                          ip_state=(sl["pos"], self.p_ori, sl["F"], sl["dF"]), mode=self.mode, max_passes=self.max_passes if lean else None,
                          weights_ready=lean, **self.opt)
        n = 1 + m._render_launches
        if sync and self.rank == 0 and self.world > 1:                          # the frame is complete when every peer's pixels are in
            check(lib.pn_epoch_wait(dptr(sl["epoch"]), 0, dptr(sl["done_local"]), self.world - 1, 0, dptr(self.status), self.timeout_ms, stream_ptr()))
            n += 1
        return n

    def _state_body(self, sl, sync=True):
        """rank 0: camera + IP state of this frame into the slot, then out to every rank."""
        sl["cam"].copy_(sl["cam_host"], non_blocking=True)
        self.sim.get_IP_info(out=(sl["pos"], sl["F"], sl["dF"]))                # state BEFORE the step (trainer.py:303-306)
        n = 2
        if self.world > 1 and sync:
            # the slot's previous frame must be finished everywhere before its state (and rank 0's frame slot) are overwritten
            check(lib.pn_epoch_wait(dptr(sl["state_epoch"]), 1, dptr(sl["done_local"]), self.world - 1, 1, dptr(self.status), self.timeout_ms, stream_ptr()))
            check(lib.pn_peer_put(dptr(sl["state"]), _al(self.state_floats * 4, 16), dptr(sl["state_dsts"]), self.world - 1, stream_ptr()))
            check(lib.pn_epoch_signal(dptr(sl["state_epoch"]), dptr(sl["state_flags_all"]), self.world - 1, stream_ptr()))
            n += 3
        return n

    def _capture(self, fn, stream):
        g = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream()
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            with torch.cuda.graph(g, stream=stream):
                n = fn()
        cur.wait_stream(stream)
        return g, n

    def build(self, pose, intrinsics):
        """Warm every kernel up eagerly (no flags, so epochs stay in step across ranks), then capture the graphs."""
        cam = self._cam(pose, intrinsics)
        for sl in self.slots:
            if self.rank == 0:
                sl["cam_host"].copy_(cam)
                self._state_body(sl, sync=False)
            else:
                sl["cam"].copy_(cam.to(self.dev))
                sl["pos"].copy_(self.p_ori); sl["F"].zero_(); sl["F"][:, 0] = 1; sl["F"][:, 4] = 1; sl["F"][:, 8] = 1; sl["dF"].zero_()
            self._frame_body(sl, sync=False)
        if self.rank == 0 and self.sim is not None:
            self.sim.capture_stream = self.sim_stream                           # the step's graph nodes belong to the simulator's stream (priority / SM partition)
            dof, vel = self.sim.dof.clone(), self.sim.dof_vel.clone()
            self.sim.stepforward(); self.sim.stepforward()                      # plain call + graph capture inside the simulator
            self.sim.dof.copy_(dof); self.sim.dof_vel.copy_(vel)
        torch.cuda.synchronize()
        # passes the warm-up frames actually used (stats[7]), agreed over the ranks; one spare, and the last pass is unbounded anyway
        used = max(int(sl["stats"][7].item()) for sl in self.slots)
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([used], dtype=torch.int64, device=self.dev if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            used = int(t.item())
        self.max_passes = None if self.lean_passes is False else min(used + 1, int(lib.pn_render_pass_count(int(self.opt.max_steps))))
        # Small ray shares (a rank of a 4- or 8-GPU frame): every pass costs a ramp-up and a drain of three kernels, so when the
        # warm-up frames show no early termination to speak of (field evaluations ~ composited samples: nothing is wasted by marching
        # a ray to its end) the second pass is the last one and takes everything the first 32 samples per ray left.
        if self.max_passes is not None and self.merge_passes is not False:
            st = [sl["stats"].tolist() for sl in self.slots]
            waste = max(float(s[2]) / max(float(s[0]), 1.0) for s in st)
            small = self.n_my <= (self.merge_passes if isinstance(self.merge_passes, int) and self.merge_passes > 1 else 200_000)
            flag = torch.tensor([1 if (waste < 1.02 and small and used >= 2) else 0], dtype=torch.int64, device=self.dev if (self.world == 1 or dist_backend() == "nccl") else "cpu")
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()):
                self.max_passes = 2
        self._barrier()
        n_frame = n_state = 0
        for sl in self.slots:
            sl["graph"], n_frame = self._capture(lambda sl=sl: self._frame_body(sl, lean=True), sl["stream"])
            if self.rank == 0:
                sl["state_graph"], n_state = self._capture(lambda sl=sl: self._state_body(sl), self.sim_stream)
        self.launches_per_frame = n_frame + n_state + (self.sim.step_launches if self.rank == 0 else 0)
        torch.cuda.synchronize()
        self._barrier()
        self._warm = True

    @staticmethod
    def calibrate_rank0_share(model, sim, opt, pose, intrinsics, slots=3, frames=12, **kw):
        """Rank 0 also runs the simulator.  Its chain (state + step) is sequential — one frame per chain latency — and that latency
        grows with the render work sharing rank 0's GPU (the field kernel saturates the L2 the step's small kernels read through):
        c(x) ~ c0 + (c1 - c0) x  for a tile share x relative to an equal split, c0 = the chain alone, c1 = beside a full share.
        The other ranks need r (N - x) / (N - 1) per frame, r = render time of an equal share.  Measured here: r (paused frames
        through an equal-split pipeline, slowest rank), c1 (the same pipeline with the simulator running), c0 (the step alone +
        the state kernels).  For large shares (2 GPUs) what matters instead is the GPU time the step takes from rank 0's render:
        x r + S = r (N - x) / (N - 1).  Returned: the smaller of the two balancing shares, clamped to [0.1, 1].  Collective: every rank calls it and
        gets the same x."""
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_initialized() else 1
        if world == 1:
            return 1.0
        pipe = FramePipeline(model, sim, opt, slots=slots, **kw)
        pipe.build(pose, intrinsics)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def period(paused):
            for _ in range(4):
                pipe.frame(pose, intrinsics, to_host=False, paused=paused)
            pipe.drain(); torch.cuda.synchronize(); dist.barrier()
            e0.record()
            for _ in range(frames):
                pipe.frame(pose, intrinsics, to_host=False, paused=paused)
            pipe.drain(); e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / frames
        dof = vel = None
        if pipe.rank == 0:
            dof, vel = sim.dof.clone(), sim.dof_vel.clone()
        r = period(True)
        c1 = period(False)
        S = 0.0
        if pipe.rank == 0:
            e0.record()
            for _ in range(20):
                sim.stepforward()
            e1.record(); torch.cuda.synchronize()
            S = e0.elapsed_time(e1) / 20
            sim.dof.copy_(dof); sim.dof_vel.copy_(vel)
        t = torch.tensor([r, c1, S], dtype=torch.float64, device=pipe.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r, c1, S = float(t[0]), float(t[1]), float(t[2])
        dist.barrier()
        pipe.close()
        c0 = S + 0.04                                                           # + camera upload, ip_info, state push
        c1 = max(c1, c0 + 1e-3)
        n = world
        x_chain = (r * n / (n - 1) - c0) / ((c1 - c0) + r / (n - 1))          # the chain's latency must not exceed the others' frame time
        x_share = 1.0 - S * (n - 1) / (n * r)                                   # nor rank 0's render + the step's own GPU time (large shares)
        return float(min(1.0, max(0.1, min(x_chain, x_share))))

    def _barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()

    @staticmethod
    def _cam(pose, intrinsics):
        p = torch.as_tensor(np.asarray(pose, dtype=np.float32)).reshape(-1)[:16]
        return torch.cat([p, torch.as_tensor(np.asarray(intrinsics, dtype=np.float32)).reshape(4)])

    # ------------------------------------------------------------------------------------------ per frame
    @torch.no_grad()
    def frame(self, pose, intrinsics, to_host=True, paused=False):
        """Enqueue one GUI frame (host pose -> frame in rank 0's slot [-> pinned host buffer]).  Returns the slot index."""
        if not self._warm:
            self.build(pose, intrinsics)
        s = self.frame_id % self.S
        sl = self.slots[s]
        cur = torch.cuda.current_stream()
        if self.rank == 0:
            if sl["state_ready"] is not None:
                sl["state_ready"].synchronize()                                  # bounds the host's run-ahead to S frames (cam_host reuse)
            sl["cam_host"].copy_(self._cam(pose, intrinsics))
            sim_st = self.sim_stream
            sim_st.wait_stream(cur)
            if sl["render_done"] is not None:
                sim_st.wait_event(sl["render_done"])                             # rank 0's previous frame in this slot has read the old state
            if sl["copy_done"] is not None:
                sim_st.wait_event(sl["copy_done"])                               # ... and its pixels have left for the host
            with torch.cuda.stream(sim_st):
                sl["state_graph"].replay()
                ev = torch.cuda.Event(); ev.record(sim_st)
                sl["state_ready"] = ev
                if not paused:
                    self.sim.stepforward()                                       # trainer.py:308 — concurrent with the render of the state read above
            sl["stream"].wait_stream(cur)
            sl["stream"].wait_event(sl["state_ready"])
        else:
            if sl["render_done"] is not None:
                sl["render_done"].synchronize()                                  # host run-ahead bound
            sl["stream"].wait_stream(cur)
        with torch.cuda.stream(sl["stream"]):
            sl["graph"].replay()
            ev = torch.cuda.Event(); ev.record(sl["stream"])
            sl["render_done"] = ev
        if self.rank == 0 and to_host:
            cs = self.copy_stream
            cs.wait_event(sl["render_done"])
            with torch.cuda.stream(cs):
                sl["host_flat"].copy_(sl["frame_flat"], non_blocking=True)       # trainer.py:589-593 .cpu().numpy() of image, depth, depth_0
                ev = torch.cuda.Event(); ev.record(cs)
                sl["copy_done"] = ev
        self.frame_id += 1
        return s

    def drain(self):
        """Make the current stream wait for everything enqueued so far (frames, simulator, host copies)."""
        cur = torch.cuda.current_stream()
        for sl in self.slots:
            cur.wait_stream(sl["stream"])
        if self.rank == 0:
            cur.wait_stream(self.sim_stream); cur.wait_stream(self.copy_stream)

    def wait_host(self, slot):
        if self.rank != 0:
            return None
        sl = self.slots[slot]
        if sl["copy_done"] is not None:
            sl["copy_done"].synchronize()
        return sl["host"]

    def close(self):
        """Release the graphs, workspaces and the peer block (after a synchronize; peers must have closed their mappings' users)."""
        torch.cuda.synchronize()
        for sl in self.slots:
            sl.clear()
        self.slots = []
        for r, p in enumerate(self.peer_ptr):
            if r != self.rank and p:
                lib.pn_peer_close(vp(p))
        self.peer_ptr = []
        self._barrier()                                                          # nobody frees a block another rank still has mapped
        self.block.free()
        lib.pn_set_render_sm_reserve(0)

    def check(self):
        """After a synchronize: raise if a flag wait timed out or a frame reported an error."""
        code = int(self.status.item())
        if code:
            raise RuntimeError(f"rank {self.rank}: peer flag wait timed out (code {code}) — a rank died or fell more than {self.timeout_ms} ms behind")
        return [self.model.check_stats(sl["stats"]) for sl in self.slots]
