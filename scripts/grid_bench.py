"""Hash-grid microbench (BASELINE.json configs[4]): 2^22 samples x 16 levels; uniform-random and ray-coherent inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pienerf_b200 import _gridencoder
from pienerf_b200.synthetic import grid_offsets
B = 1 << 22
off, s = grid_offsets(desired_resolution=2048)
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
emb = (torch.rand(int(off[-1]), 2, device=dev, generator=g) * 2 - 1)
offs = torch.from_numpy(off).to(dev)
pts = torch.rand(B, 3, device=dev, generator=g)
# ray-coherent: 2^22/128 rays x 128 consecutive samples, dt = 0.0017 in [0,1] units
nr = B // 128
o = torch.rand(nr, 1, 3, device=dev, generator=g) * 0.5 + 0.1
d = torch.nn.functional.normalize(torch.rand(nr, 1, 3, device=dev, generator=g) + 0.1, dim=-1)
coh = (o + d * (torch.arange(128, device=dev).view(1, 128, 1) * 0.0017)).reshape(B, 3).contiguous().clamp(0, 1)
out = torch.empty(16, B, 2, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
S = float(np.log2(s))
def run(x):
    ts = []
    for i in range(8):
        flush.fill_(1.0)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); _gridencoder.grid_encode_forward(x, emb, offs, out, B, 3, 2, 16, S, 16, None, 0, False, 0); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts[2:]))
for name, x in (("random", pts), ("ray-coherent", coh)):
    ms = run(x)
    print(f"variant {os.environ.get('PN_GRID_VARIANT','0')} {name}: {ms:.3f} ms  {B*1164/ms/1e6:.0f} GB/s algorithmic  frac {B*1164/ms/1e6/6538.9:.3f}")
