"""One rank of a FramePipeline run (launched by tests/test_gpu_pipeline.py and scripts/; not a test module).

    RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT / LOCAL_RANK in the env; argv: out.npz n_frames W H [kind] [slots]
Rank 0 writes the host frames of every frame to out.npz.  With fewer GPUs than ranks the ranks share GPU 0 (gloo process
group; CUDA IPC and the flag protocol work the same, the GPU time-slices between the processes)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out, n_frames, W, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    kind = sys.argv[5] if len(sys.argv) > 5 else "block64"
    slots = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    ngpu = torch.cuda.device_count()
    local = rank if ngpu >= world else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        if ngpu >= world:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")
    from pienerf_b200.frame import Options
    from pienerf_b200.network import NeRFNetwork
    from pienerf_b200.pipeline import FramePipeline
    from pienerf_b200.simulator import Simulator
    from tests.util import small_scene
    body, field, bits, pose, intr = small_scene(kind=kind, W=W, H=H)
    model = NeRFNetwork(bound=1, density_scale=20.0).to(dev).load_field(field)
    model.density_bitfield.copy_(torch.from_numpy(bits).to(dev))
    sim = Simulator(dt=1e-2, iters=10, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]), device=dev)
    sim.set_points(body["pos"], body["mass"], body["mu"], body["lam"], body["pin"]).initialize()
    opt = Options.defaults(bound=1.0, W=W, H=H, max_steps=256, T_thresh=1e-2, dt_gamma=0.0, min_near=0.2, max_iter_num=1, num_seek_IP=3, sim_dx=0.05)
    pipe = FramePipeline(model, sim, opt, slots=slots, tile=8, timeout_ms=60000)
    frames = []
    for k in range(n_frames):
        if rank == 0 and k == 2:
            sim.update_force(5, torch.tensor([4e4, 1e4, -2e4]))
        s = pipe.frame(pose, intr, to_host=True)
        if rank == 0:
            h = pipe.wait_host(s)
            frames.append({k_: v.clone().numpy() for k_, v in h.items()})
    pipe.drain()
    torch.cuda.synchronize()
    stats = pipe.check()
    if world > 1:
        dist.barrier()
    if rank == 0:
        np.savez(out, image=np.stack([f["image"] for f in frames]), depth=np.stack([f["depth"] for f in frames]),
                 depth_0=np.stack([f["depth_0"] for f in frames]), stats=np.asarray(stats), launches=pipe.launches_per_frame)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
