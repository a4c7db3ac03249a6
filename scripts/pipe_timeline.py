"""Timeline of the frame pipeline from CUDA events (no nsys in the image): for K frames, when each rank-frame graph, state
graph and simulator step started and ended on this GPU, relative to the first event.   usage (1 GPU or under torchrun):
    python scripts/pipe_timeline.py [config] [frames] [slots]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import torch.distributed as dist

config = sys.argv[1] if len(sys.argv) > 1 else "chair"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
slots = int(sys.argv[3]) if len(sys.argv) > 3 else 3
paused = len(sys.argv) > 4 and sys.argv[4] == "paused"
reserve = int(sys.argv[5]) if len(sys.argv) > 5 else None
green = len(sys.argv) > 6 and sys.argv[6] == "green"
merge = False if (len(sys.argv) > 7 and sys.argv[7] == "nomerge") else None
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from pienerf_b200.frame import build_scene
from pienerf_b200.pipeline import FramePipeline

model, sim, opt, pose, intr, body, field = build_scene(config, device=dev)
pipe = FramePipeline(model, sim, opt, slots=slots, sim_sm_reserve=reserve, green=green, merge_passes=merge)
pipe.build(pose, intr)
for _ in range(6):
    pipe.frame(pose, intr, to_host=False, paused=paused)
pipe.drain(); torch.cuda.synchronize()
if world > 1:
    dist.barrier()

ev = lambda: torch.cuda.Event(enable_timing=True)
rec = []
orig_frame_replay = {}
torch.cuda.profiler.start()
t0 = ev(); t0.record()
# instrument: wrap graph replays with events on their streams
for k in range(K):
    s = pipe.frame_id % pipe.S
    sl = pipe.slots[s]
    e = {"slot": s}
    g, sg = sl["graph"], sl.get("state_graph")

    class Wrap:
        def __init__(self, graph, tag, stream_of):
            self.graph, self.tag, self.stream_of = graph, tag, stream_of

        def replay(self):
            a, b = ev(), ev()
            a.record(torch.cuda.current_stream()); self.graph.replay(); b.record(torch.cuda.current_stream())
            e[self.tag] = (a, b)
    sl["graph"] = Wrap(g, "frame", None)
    if sg is not None:
        sl["state_graph"] = Wrap(sg, "state", None)
    if rank == 0:
        old_step = sim.stepforward

        def stepf():
            a, b = ev(), ev()
            a.record(torch.cuda.current_stream()); old_step(); b.record(torch.cuda.current_stream())
            e["step"] = (a, b)
        sim.stepforward = stepf
    pipe.frame(pose, intr, to_host=False, paused=paused)
    sl["graph"] = g
    if sg is not None:
        sl["state_graph"] = sg
    if rank == 0:
        sim.stepforward = old_step
    rec.append(e)
pipe.drain()
t1 = ev(); t1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if world > 1:
    dist.barrier()
total = t0.elapsed_time(t1)
lines = [f"rank {rank}/{world} {config}: {K} frames in {total:.3f} ms = {total / K:.3f} ms/frame ({1e3 * K / total:.1f} fps), {slots} slots, paused={paused}, reserve={pipe.sim_sm_reserve}, green={pipe.green is not None}, passes={pipe.max_passes}"]
for k, e in enumerate(rec):
    parts = [f"frame {k:2d} slot {e['slot']}"]
    for tag in ("state", "step", "frame"):
        if tag in e:
            a, b = e[tag]
            parts.append(f"{tag} [{t0.elapsed_time(a):7.3f} -> {t0.elapsed_time(b):7.3f}] ({a.elapsed_time(b):.3f})")
    lines.append("  ".join(parts))
for r in range(world):
    if r == rank:
        print("\n".join(lines), flush=True)
    if world > 1:
        dist.barrier()
pipe.check()
if world > 1:
    dist.destroy_process_group()
