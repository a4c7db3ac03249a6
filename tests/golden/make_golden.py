"""Generates tests/golden/ref_kernels.npz by running the REFERENCE's own CUDA kernels (oracle/_ref/*.so, compiled
unmodified from /root/reference for sm_100a) on small seeded inputs.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_kernels.npz'

and copy the file to tests/golden/.  The CPU test tests/test_golden.py then pins oracle/render_oracle.py to these
vectors without needing a GPU or the reference tree."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.build_ref import load_ref  # noqa: E402
from pienerf_b200.synthetic import grid_offsets  # noqa: E402
from tests.util import deformed_ip_state, small_scene  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_kernels.npz")
rg, rs, rm = load_ref("_ref_gridencoder"), load_ref("_ref_shencoder"), load_ref("_ref_raymarching")
assert rg and rs and rm, "oracle/_ref not built"
dev = "cuda"
g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
rng = np.random.default_rng(2024)
G = {}

# ---- hash grid: 8 levels, 2^12 table (levels 0-1 dense, 2-7 hashed), C=2, D=3
off, s = grid_offsets(num_levels=8, log2_hashmap_size=12, desired_resolution=512)
emb = rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)
x = rng.uniform(0, 1, size=(256, 3)).astype(np.float32)
x[0] = 0; x[1] = 1; x[2] = [1.0001, .5, .5]; x[3] = [.5, -1e-6, .5]; x[4] = .5
outp = torch.empty(8, 256, 2, device=dev)
rg.grid_encode_forward(g(x), g(emb), g(off), outp, 256, 3, 2, 8, float(np.log2(s)), 16, None, 0, False, 0)
G.update(grid_off=off, grid_scale=np.float64(s), grid_emb=emb, grid_x=x, grid_out=outp.cpu().numpy())
# tiled + align_corners + smoothstep variant, C=4, D=2
off2, s2 = grid_offsets(input_dim=2, num_levels=4, log2_hashmap_size=10, desired_resolution=128, align_corners=True)
emb2 = rng.uniform(-1, 1, size=(int(off2[-1]), 4)).astype(np.float32)
x2 = rng.uniform(0, 1, size=(128, 2)).astype(np.float32)
out2 = torch.empty(4, 128, 4, device=dev)
rg.grid_encode_forward(g(x2), g(emb2), g(off2), out2, 128, 2, 4, 4, float(np.log2(s2)), 16, None, 1, True, 1)
G.update(grid2_off=off2, grid2_scale=np.float64(s2), grid2_emb=emb2, grid2_x=x2, grid2_out=out2.cpu().numpy())

# ---- SH degree 4
d = rng.normal(size=(256, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
sh = torch.empty(256, 16, device=dev)
rs.sh_encode_forward(g(d), sh, 256, 3, 4, None)
G.update(sh_d=d, sh_out=sh.cpu().numpy())

# ---- ray marching on a small deformed scene
body, field, bits, pose, intr = small_scene(W=16, H=16)
p_ori, p_def, F, dF = deformed_ip_state(body, amp=0.03)
rays_o, rays_d = ro.get_rays(pose, intr, 16, 16)
N = rays_o.shape[0]
bbmin = p_def.min(0) - np.float32(1e-3); bbmax = p_def.max(0) + np.float32(1e-3)
res = np.ceil((bbmax - bbmin) * (np.float32(1) / np.float32(0.06))).astype(np.int32)
nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)
rm.near_far_from_aabb(g(rays_o), g(rays_d), g(np.concatenate([bbmin, bbmax])), N, 0.2, nears, fars)
cnt, bgn, idx = ro.get_pnts_in_grids(p_def, bbmin, 0.06, res)
occ_idx = np.nonzero(bits)[0].astype(np.int32); occ_val = bits[occ_idx]
G.update(rm_rays_o=rays_o, rm_rays_d=rays_d, rm_p_ori=p_ori, rm_p_def=p_def, rm_F=F, rm_dF=dF, rm_bbmin=bbmin, rm_bbmax=bbmax, rm_res=res,
         rm_nears=nears.cpu().numpy(), rm_fars=fars.cpu().numpy(), rm_bits_idx=occ_idx, rm_bits_val=occ_val, rm_bits_len=np.int64(bits.size))
for K, mi in ((1, 1), (3, 1), (3, 100)):
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    n_step = 6
    M = N * n_step + 128 - (N * n_step) % 128
    xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
    rm.march_rays_quadratic_bending(g(cnt), g(bgn), g(idx), p_ori.shape[0], int(np.prod(res)), g(p_def), g(p_ori), g(F), g(dF), mi, g(bbmin), g(bbmax),
                                    0.06, g(res), K, 0.0525, False, torch.zeros(6, device=dev), N, n_step, alive, nears.clone(), g(rays_o), g(rays_d),
                                    1.0, 0.0, 256, 1, 128, g(bits), nears, fars, xyzs, dirs, deltas, torch.zeros(N, device=dev))
    torch.cuda.synchronize()
    G[f"rm_xyzs_K{K}_it{mi}"] = xyzs.cpu().numpy(); G[f"rm_deltas_K{K}_it{mi}"] = deltas.cpu().numpy()
    if K == 3 and mi == 1:
        sig = rng.uniform(0, 60, size=M).astype(np.float32); rgb = rng.uniform(0, 1, size=(M, 3)).astype(np.float32)
        t = nears.clone(); ws = torch.zeros(N, device=dev); dp = torch.zeros(N, device=dev); im = torch.zeros(N, 3, device=dev)
        rm.composite_rays(N, n_step, 1e-2, alive, t, g(sig), g(rgb), deltas, ws, dp, im)
        torch.cuda.synchronize()
        G.update(cp_sig=sig, cp_rgb=rgb, cp_alive=alive.cpu().numpy(), cp_t=t.cpu().numpy(), cp_ws=ws.cpu().numpy(), cp_depth=dp.cpu().numpy(), cp_image=im.cpu().numpy())
# plain march with dt_gamma
aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
n2 = torch.empty(N, device=dev); f2 = torch.empty(N, device=dev)
rm.near_far_from_aabb(g(rays_o), g(rays_d), g(aabb), N, 0.2, n2, f2)
alive = torch.arange(N, dtype=torch.int32, device=dev)
xyzs = torch.zeros(N * 4, 3, device=dev); dirs = torch.zeros(N * 4, 3, device=dev); deltas = torch.zeros(N * 4, 2, device=dev)
rm.march_rays(N, 4, alive, n2.clone(), g(rays_o), g(rays_d), 1.0, 1 / 128, 256, 1, 128, g(bits), n2, f2, xyzs, dirs, deltas, torch.zeros(N, device=dev))
torch.cuda.synchronize()
G.update(mr_nears=n2.cpu().numpy(), mr_fars=f2.cpu().numpy(), mr_xyzs=xyzs.cpu().numpy(), mr_deltas=deltas.cpu().numpy())
# morton / packbits
c = rng.integers(0, 128, size=(64, 3)).astype(np.int32)
mi_ = torch.empty(64, dtype=torch.int32, device=dev); rm.morton3D(g(c), 64, mi_)
cb = torch.empty(64, 3, dtype=torch.int32, device=dev); rm.morton3D_invert(mi_, 64, cb)
dg = rng.uniform(0, 20, size=(1, 512)).astype(np.float32); pb = torch.empty(64, dtype=torch.uint8, device=dev); rm.packbits(g(dg), 64, 10.0, pb)
torch.cuda.synchronize()
G.update(mo_c=c, mo_idx=mi_.cpu().numpy(), mo_back=cb.cpu().numpy(), pk_grid=dg, pk_bits=pb.cpu().numpy())
os.makedirs(os.path.dirname(out_path), exist_ok=True)
np.savez_compressed(out_path, **G)
print("wrote", out_path, os.path.getsize(out_path), "bytes;", len(G), "arrays")
