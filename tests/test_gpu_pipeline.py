"""The pipelined frame driver (pienerf_b200/pipeline.py): frames in flight on several slots, each a CUDA graph, the
multi-GPU exchanges as peer-memory stores + epoch flags.  Bar: the frames are BIT-IDENTICAL to the synchronous one-GPU
driver (same kernels on the same rays: the tile split and the pipelining must be invisible), frame after frame, with the
simulator stepping and a force switched on in between (the reference's order, trainer.py:299-308)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
W = H = 64
N_FRAMES = 7


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run(world, out, slots=3):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_DEVICE_MAX_CONNECTIONS="32")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "pipeline_worker.py"), out, str(N_FRAMES), str(W), str(H), "block64", str(slots)],
                                      env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o.decode(errors="replace"))
    for r, (p, o) in enumerate(zip(procs, logs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    return np.load(out)


@pytest.fixture(scope="module")
def sync_frames():
    """The same 7 frames from the synchronous single-GPU driver (FrameDriver.test_gui, one frame at a time)."""
    from pienerf_b200.frame import FrameDriver, Options
    from pienerf_b200.network import NeRFNetwork
    from pienerf_b200.simulator import Simulator
    from tests.util import small_scene
    body, field, bits, pose, intr = small_scene(kind="block64", W=W, H=H)
    model = NeRFNetwork(bound=1, density_scale=20.0).cuda().load_field(field)
    model.density_bitfield.copy_(torch.from_numpy(bits).cuda())
    sim = Simulator(dt=1e-2, iters=10, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]))
    sim.set_points(body["pos"], body["mass"], body["mu"], body["lam"], body["pin"]).initialize()
    opt = Options.defaults(bound=1.0, W=W, H=H, max_steps=256, T_thresh=1e-2, dt_gamma=0.0, min_near=0.2, max_iter_num=1, num_seek_IP=3, sim_dx=0.05)
    drv = FrameDriver(model, sim, opt)
    frames = []
    for k in range(N_FRAMES):
        if k == 2:
            sim.update_force(5, torch.tensor([4e4, 1e4, -2e4]))
        f = drv.test_gui(pose, intr, W, H)
        frames.append({k_: np.array(v).reshape(H * W, -1).squeeze() for k_, v in f.items()})
    return frames


def _same(got, want):
    for k in range(N_FRAMES):
        for name in ("image", "depth", "depth_0"):
            a, b = got[name][k], want[k][name]
            assert np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0)), (k, name, float(np.nanmax(np.abs(a - b))))
    # frames differ from one another (the body moves): the comparison above is not vacuous
    assert np.abs(got["image"][N_FRAMES - 1] - got["image"][0]).max() > 1e-3


def test_pipeline_one_rank_is_bit_identical_to_the_sync_driver(sync_frames, tmp_path):
    got = _run(1, str(tmp_path / "w1.npz"))
    _same(got, sync_frames)
    assert (got["stats"][:, 4] == 0).all() and int(got["launches"]) > 30


@pytest.mark.parametrize("world,slots", [(2, 3), (3, 2)])
def test_pipeline_n_ranks_frame_equals_one_rank_frame(sync_frames, tmp_path, world, slots):
    """N ranks (N GPUs over NCCL-bootstrapped IPC when the box has them, otherwise N processes sharing GPU 0): the frame
    assembled in rank 0's memory by the peers' compositor stores equals the one-rank frame bit for bit."""
    got = _run(world, str(tmp_path / f"w{world}.npz"), slots=slots)
    _same(got, sync_frames)
