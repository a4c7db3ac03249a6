"""Host-side mirror of simulator/solver.py `Simulator` backed by the `_qgmls` CUDA operators.

Same constructor arguments, method names and return layouts as the reference class
(solver.py:12-617): `initialize()`, `InitializeFromPly()`, `get_IP_info()`, `stepforward()`,
`update_force()`, `clear_force()`, `update_pos()`, `OutputToPly()`; attributes `dx`, `IP_pos`, `dof`, ...
What differs is layout, not meaning: the system matrices are kept compact ([n,n] with n = 10*n_k acting on
[n,3] DOFs instead of (Mat (x) I3) on [3n]), the rhs is a deterministic gather, and one `stepforward()` is a
single C-ABI call that enqueues the whole local-global loop with no host synchronisation.
"""
import numpy as np
import torch

from . import _qgmls
from ._lib import QgmlsStepT, dptr

torchfloat = torch.float64


class Simulator:
    def __init__(self, dt=1e-2, iters=20, bbox=None, kres=7, dx=1, gravity=None, stiff=1e5, base=None,
                 solver="inverse", pcg_iters=200, device="cuda", use_graph=True):
        # solver.py:17-25: the reference scales the caller's tensors IN PLACE in the caller's dtype
        # (main_gui.py passes float32) — reproduce the rounding, not the aliasing bug (fresh tensors here).
        bbox = torch.tensor([1.0, 1.0, 1.0], dtype=torchfloat) if bbox is None else torch.as_tensor(bbox).clone()
        base = torch.tensor([-0.5, -0.5, -0.5], dtype=torchfloat) if base is None else torch.as_tensor(base).clone()
        gravity = torch.tensor([0.0, -9.8, 0.0], dtype=torchfloat) if gravity is None else torch.as_tensor(gravity).clone()
        bbox *= 1.02
        base *= 1.01
        self.device = torch.device(device)
        bbox = bbox.to(dtype=torchfloat, device=self.device)
        self.gravity = gravity.to(dtype=torchfloat, device=self.device)
        self.base = base.to(dtype=torchfloat, device=self.device)
        self.dt = dt
        self.iters = iters
        self.res = (bbox // dx).to(dtype=torch.int32)
        self.dx = dx
        self.kres = kres
        self.stiff = stiff
        self.solver = {"inverse": 0, "pcg": 1}[solver]
        self.use_graph = use_graph
        self.pcg_iters = pcg_iters
        self.pos = self.mass = self.mu = self.lam = self.is_pin = None
        self._step_desc = None
        self._graph = self._graph_key = self._graph_warm = None

    # ------------------------------------------------------------------ I/O (solver.py:109-137)
    def InitializeFromPly(self, path):
        from .ply import read_ply_vertices
        v = read_ply_vertices(path)
        dev = self.device
        self.pos = torch.from_numpy(np.stack((v["x"], v["y"], v["z"]), axis=1).astype(np.float64)).to(dev)
        assert self.pos.shape[0] > 0
        self.mass = torch.from_numpy(np.asarray(v["mass"]).astype(np.float64)).to(dev)
        self.mu = torch.from_numpy(np.asarray(v["mu"]).astype(np.float64)).to(dev)
        self.lam = torch.from_numpy(np.asarray(v["lam"]).astype(np.float64)).to(dev)
        self.is_pin = torch.from_numpy(np.asarray(v["pin"]).astype(bool)).to(dev)
        self.initialize()

    def OutputToPly(self, path):
        from .ply import write_ply_xyz
        self.update_pos()
        write_ply_xyz(path, self.pos.cpu().numpy().astype(np.float64))

    def set_points(self, pos, mass, mu, lam, is_pin):
        dev = self.device
        self.pos = torch.as_tensor(pos, dtype=torchfloat).to(dev).contiguous()
        self.mass = torch.as_tensor(mass, dtype=torchfloat).to(dev).contiguous()
        self.mu = torch.as_tensor(mu, dtype=torchfloat).to(dev).contiguous()
        self.lam = torch.as_tensor(lam, dtype=torchfloat).to(dev).contiguous()
        self.is_pin = torch.as_tensor(np.asarray(is_pin).astype(bool)).to(dev)
        return self

    # ------------------------------------------------------------------ init (solver.py:139-331)
    @torch.no_grad()
    def initialize(self):
        dev = self.device
        res = [int(v) for v in self.res.tolist()]
        K = self.kres
        self.grid_idx = ((self.pos - self.base) // self.dx).to(dtype=torch.int32)
        gi = self.grid_idx.long()
        if (gi < 0).any() or (gi >= torch.tensor(res, device=dev)).any():
            raise RuntimeError("sample points fall outside the simulator bbox")
        self.IP_mask = torch.zeros(res, dtype=torch.bool, device=dev)
        self.IP_mask[gi[:, 0], gi[:, 1], gi[:, 2]] = True
        n_ip = int(self.IP_mask.sum())
        self.IP_idx = -torch.ones(res, dtype=torch.int32, device=dev)
        self.IP_idx[self.IP_mask] = torch.arange(0, n_ip, 1, dtype=torch.int32, device=dev)
        self.pts_IP = self.IP_idx[gi[:, 0], gi[:, 1], gi[:, 2]].contiguous()
        # kornia meshgrid + [1,2] swap == integer (i,j,k) in C-order of the mask (solver.py:162-177)
        self.IP_grid = torch.nonzero(self.IP_mask).to(torch.int32)
        self.IP_pos = ((self.IP_grid + 0.5) * self.dx + self.base).contiguous()

        # solver.py:184 — int32 0-dim tensor times python floats: evaluated by torch in float32
        self.kdx = (self.res.max() * self.dx) / (K - 1)
        IP2K = ((self.IP_pos - self.base) // self.kdx).to(dtype=torch.int32).long()
        pts2K = ((self.pos - self.base) // self.kdx).to(dtype=torch.int32).long()
        if (IP2K < 0).any() or (IP2K + 1 >= K).any() or (pts2K < 0).any() or (pts2K + 1 >= K).any():
            raise RuntimeError("points fall outside the kernel lattice")
        self.kernel_mask = torch.zeros((K, K, K), dtype=torch.bool, device=dev)
        offs = [(S >> 2 & 1, S >> 1 & 1, S & 1) for S in range(8)]
        for (x, y, z) in offs:
            self.kernel_mask[IP2K[:, 0] + x, IP2K[:, 1] + y, IP2K[:, 2] + z] = True
        n_k = int(self.kernel_mask.sum())
        # 0 (not -1) where unmasked: a sample point in a different kernel cell than its IP may alias kernel 0
        self.kernel_idx = torch.zeros((K, K, K), dtype=torch.int32, device=dev)
        self.kernel_idx[self.kernel_mask] = torch.arange(0, n_k, 1, dtype=torch.int32, device=dev)
        self.IP_kernel = torch.stack([self.kernel_idx[IP2K[:, 0] + x, IP2K[:, 1] + y, IP2K[:, 2] + z] for (x, y, z) in offs], 1).contiguous()
        self.pts_kernel = torch.stack([self.kernel_idx[pts2K[:, 0] + x, pts2K[:, 1] + y, pts2K[:, 2] + z] for (x, y, z) in offs], 1).contiguous()
        self.kernel_grid = torch.nonzero(self.kernel_mask).to(torch.int32)
        # solver.py:248 — int32 * float32 0-dim is a float32 product, widened by `+ base`
        self.kernel_pos = (self.kernel_grid * self.kdx + self.base).contiguous()
        self.n_ip, self.n_k, self.n = n_ip, n_k, 10 * n_k
        kdx = float(self.kdx)

        # shape functions (solver.py:250-252, 334-399)
        self.pts_Nx, _, _ = _qgmls.shape_functions(kdx, self.pos, self.pts_kernel, self.kernel_pos, want_derivatives=False)
        self.pts_dNx = self.pts_ddNx = None                      # never read by the reference either
        self.IP_Nx, self.IP_dNx, self.IP_ddNx = _qgmls.shape_functions(kdx, self.IP_pos, self.IP_kernel, self.kernel_pos)
        self.IP_mu, self.IP_lam, self.IP_rho = _qgmls.collect_param(self.pts_IP, self.mu, self.lam, self.mass, n_ip, self.dx)
        self.build_global()

        # DOF layout (solver.py:258-277), kept as [n,3]: dof[k*10+s] = vec3
        self.dof = torch.zeros((self.n, 3), dtype=torchfloat, device=dev)
        slot0 = torch.arange(n_k, device=dev) * 10
        self.dof[slot0] = self.kernel_pos
        for x in range(3):
            self.dof[slot0 + 1 + x, x] = 1
        self.dof_rest = self.dof.clone()
        self.dof_tilde = self.dof.clone()
        self.dof_vel = torch.zeros_like(self.dof)
        self.dof_f = torch.zeros_like(self.dof)

        # kernel -> (ip, corner) CSR (solver.py:279-313; deterministic: ascending ip*8+corner per kernel)
        flat = self.IP_kernel.reshape(-1).long()
        order = torch.sort(flat, stable=True).indices
        self.kernel_cnt = torch.bincount(flat, minlength=n_k).to(torch.int32)
        self.kernel_bg = torch.zeros(n_k + 1, dtype=torch.int32, device=dev)
        self.kernel_bg[1:] = torch.cumsum(self.kernel_cnt, 0).to(torch.int32)
        self.buffer = order.to(torch.int32).contiguous()        # code = ip*8 + corner
        self.tot = int(self.buffer.numel())

        self.adj_slices = max(1, (int(self.kernel_cnt.max()) + 127) // 128)
        self._scratch = torch.empty(_qgmls.step_scratch_doubles(n_ip, n_k, self.adj_slices), dtype=torchfloat, device=dev)
        self._ip_stress = torch.empty(n_ip, 9, dtype=torchfloat, device=dev)
        self._partial = torch.empty(n_k, self.adj_slices, 30, dtype=torchfloat, device=dev)
        # rhs_rest = build_rhs() + M/dt^2 @ dof (solver.py:314)
        tmp = torch.empty_like(self.dof)
        _qgmls.matvec3(self.mass_matrix_invt2, self.dof, tmp)
        self.rhs_rest = self.build_rhs() + tmp
        self.rhs_gravity = torch.zeros_like(self.dof)
        _qgmls.collect_gravity(self.dx, self.IP_kernel, self.IP_Nx, self.gravity.tolist(), self.IP_rho, self.rhs_gravity)
        self._step_desc = None
        self._graph = self._graph_key = self._graph_warm = None
        return self

    @torch.no_grad()
    def build_global(self):
        """solver.py:453-538 in compact form: A [n,n], global_matrix = inv(A_active + 1e-3 I) scattered, M [n,n]."""
        dev, n = self.device, self.n
        mat = torch.zeros((n, n), dtype=torchfloat, device=dev)
        _qgmls.build_ip_global(self.dx, self.dt, self.IP_kernel, self.IP_mu, self.IP_lam, self.IP_rho, self.IP_Nx, self.IP_dNx, self.IP_ddNx, mat)
        vid = torch.nonzero(self.is_pin).reshape(-1).to(torch.int32).contiguous()
        assert int(self.pts_kernel.min()) >= 0 and int(self.pts_kernel.max()) < self.n_k
        if vid.numel():
            _qgmls.build_pin_global(self.stiff, vid, self.pts_kernel, self.pts_Nx, mat)
        self.system_matrix = mat
        diag0 = mat.diagonal()[0::10]
        self.kernel_active = (diag0 > 0.0)
        lst = (torch.nonzero(self.kernel_active).reshape(-1)[:, None] * 10 + torch.arange(10, device=dev)[None, :]).reshape(-1)
        sub = mat[lst][:, lst].clone()
        sub.diagonal().add_(1e-3)
        inv = torch.linalg.inv(sub)                                # init-time library call, as in the reference (solver.py:508)
        self.global_matrix = torch.zeros_like(mat)
        self.global_matrix[lst[:, None], lst[None, :]] = inv
        self.mass_matrix_invt2 = torch.zeros((n, n), dtype=torchfloat, device=dev)
        _qgmls.build_ip_global(self.dx, self.dt, self.IP_kernel, None, None, self.IP_rho, self.IP_Nx, self.IP_dNx, self.IP_ddNx,
                               self.mass_matrix_invt2)
        self._active_u8 = self.kernel_active.to(torch.uint8).contiguous()
        self._graph = self._graph_key = self._graph_warm = None          # new matrices: any captured step is stale

    # ------------------------------------------------------------------ per-frame API
    @torch.no_grad()
    def build_rhs(self):
        """solver.py:541-571 -> [n,3]."""
        rhs = torch.empty_like(self.dof)
        _qgmls.build_rhs(self.dx, self.IP_kernel, self.IP_mu, self.IP_lam, self.IP_dNx, self.dof, self.n_k, self.kernel_bg,
                         self.buffer, self.adj_slices, self._ip_stress, self._partial, rhs)
        return rhs

    def _desc(self):
        """pn_qgmls_step_t over the CURRENT buffers (rebuilt on every call: it is a handful of pointer reads, and a cached copy
        would keep pointing at freed memory after `sim.dof = ...`, a second build_global() or a changed dt)."""
        d = QgmlsStepT()
        d.n_ip, d.n_k, d.iters, d.dt, d.dx = self.n_ip, self.n_k, int(self.iters), float(self.dt), float(self.dx)
        d.topo, d.mu, d.lam, d.dNx = dptr(self.IP_kernel), dptr(self.IP_mu), dptr(self.IP_lam), dptr(self.IP_dNx)
        d.adj_bgn, d.adj, d.adj_slices = dptr(self.kernel_bg), dptr(self.buffer), int(self.adj_slices)
        d.Ainv, d.M, d.A, d.active = dptr(self.global_matrix), dptr(self.mass_matrix_invt2), dptr(self.system_matrix), dptr(self._active_u8)
        d.pcg_iters = int(self.pcg_iters)
        d.dof_rest, d.rhs_rest, d.rhs_gravity = dptr(self.dof_rest), dptr(self.rhs_rest), dptr(self.rhs_gravity)
        d.dof, d.dof_vel, d.scratch = dptr(self.dof, "dof", torchfloat), dptr(self.dof_vel, "dof_vel", torchfloat), dptr(self._scratch)
        d.dof_f = dptr(self.dof_f, "dof_f", torchfloat)
        self._step_desc = d
        return d

    @property
    def step_launches(self):
        """kernels one stepforward() enqueues (1 when the whole step runs as the thread-block-cluster kernel)"""
        from ._lib import lib
        return int(lib.pn_qgmls_step_launches(self.n_ip, self.n_k, int(self.iters), self.solver, int(self.pcg_iters)))

    def _graph_signature(self):
        ts = (self.dof, self.dof_vel, self.dof_f, self.dof_rest, self.rhs_rest, self.rhs_gravity, self.global_matrix, self.mass_matrix_invt2,
              self.system_matrix, self._scratch)
        return (int(self.iters), int(self.pcg_iters), self.solver, float(self.dt), _qgmls.step_mode_value()) + tuple(t.data_ptr() for t in ts)

    @torch.no_grad()
    def stepforward(self, graph=None):
        """solver.py:595-602: momentum, `iters` local-global iterations, damped velocity — one enqueue-only call: one
        thread-block-cluster kernel for systems up to n = 1280, otherwise a fixed chain of 3 + 4*iters kernels replayed as ONE
        CUDA graph launch after the first call (`graph=False` forces the plain enqueue; `self.use_graph` is the default).
        The graph is keyed on every buffer address, dt and the iteration counts, so rebinding a tensor re-captures."""
        use = self.use_graph if graph is None else graph
        if not use or not self.dof.is_cuda:
            _qgmls.step(self._desc(), self.solver)
            return
        key = self._graph_signature()
        if self._graph is None or self._graph_key != key:
            if self._graph_warm != key:                                   # first call with this configuration: plain, warms everything up
                _qgmls.step(self._desc(), self.solver)
                self._graph_warm = key
                return
            cur = torch.cuda.current_stream()
            # captured on a HIGH-priority stream: graph kernel nodes keep the priority of the stream they were captured on, and on the
            # GPU that also renders the step's small kernels must get the first CTA slot that frees up (pipeline.py)
            side = getattr(self, "capture_stream", None) or torch.cuda.Stream(device=self.dof.device, priority=-1)
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    _qgmls.step(self._desc(), self.solver)            # captured, not executed
            cur.wait_stream(side)
            self._graph, self._graph_key = g, key
        self._graph.replay()

    @torch.no_grad()
    def get_IP_info(self, out=None):
        """solver.py:402-424: (pos [n,3], F [n,9], dF [n,27]) float32 in the renderer's layouts.  `out` = three
        preallocated contiguous tensors to fill (the multi-GPU driver passes views of its broadcast buffer)."""
        dev = self.device
        if out is not None:
            pos, F, dF = out
        else:
            pos = torch.empty(self.n_ip, 3, dtype=torch.float32, device=dev)
            F = torch.empty(self.n_ip, 9, dtype=torch.float32, device=dev)
            dF = torch.empty(self.n_ip, 27, dtype=torch.float32, device=dev)
        _qgmls.ip_info(self.IP_kernel, self.dof, self.IP_Nx, self.IP_dNx, self.IP_ddNx, pos, F, dF)
        return pos, F, dF

    @torch.no_grad()
    def update_force(self, vid, f):
        """solver.py:578-588."""
        f = torch.as_tensor(f, dtype=torchfloat).reshape(3).tolist()
        _qgmls.update_force(int(vid), f, self.IP_kernel, self.IP_Nx, self.IP_rho, self.dx, self.dof_f)

    @torch.no_grad()
    def clear_force(self):
        """solver.py:590-593."""
        _qgmls.update_force(-1, None, self.IP_kernel, self.IP_Nx, self.IP_rho, self.dx, self.dof_f)

    @torch.no_grad()
    def update_pos(self):
        """solver.py:604-617."""
        self.pos = torch.empty_like(self.pos)
        _qgmls.update_pos(self.pts_kernel, self.dof, self.pts_Nx, self.pos)
        return self.pos
