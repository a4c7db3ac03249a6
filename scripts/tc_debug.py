import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pienerf_b200.network import NeRFNetwork
from pienerf_b200.synthetic import make_field
M = int(sys.argv[1]) if len(sys.argv) > 1 else 384
field = make_field(bound=1.0, seed=5)
model = NeRFNetwork(bound=1).cuda().load_field(field)
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.uniform(-1, 1, size=(M, 3)).astype(np.float32)).cuda()
d = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda"), dim=-1)
s0, c0 = model.forward_fused(x, d, mode=0)
torch.cuda.synchronize()
s1, c1 = model.forward_fused(x, d, mode=1)
torch.cuda.synchronize()
print("sigma rel", float((s1 / s0 - 1).abs().max()), "rgb abs", float((c1 - c0).abs().max()))
print(s0[:4].tolist(), s1[:4].tolist()); print(c0[:2].tolist(), c1[:2].tolist())
