"""Point sampling (pienerf_b200/sampling.py) against tests/golden/ref_sampling.npz: the output of the reference's
own, unmodified main_sample.py (AdaptiveUniformSampling.sample) run through the numpy warp stand-in
(tests/golden/make_golden_sampling.py) on the same analytic density field, options and torch seed.  Runs on the CPU device:
the module is device-agnostic torch (a once-per-asset tool); both sides do float32 torch / numpy arithmetic, so points and
volumes are compared bit for bit."""
import argparse
import os

import numpy as np
import pytest
import torch

from pienerf_b200.ply import read_ply_vertices
from pienerf_b200.sampling import AdaptiveUniformSampling
from tests.sampling_cases import CASES, BlobField

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_sampling.npz")


@pytest.mark.parametrize("tag", sorted(CASES))
def test_sample_matches_reference(tag, tmp_path):
    G = np.load(GOLD)
    o = dict(CASES[tag]); o["cut_bounds"] = list(o["cut_bounds"])
    opt = argparse.Namespace(**o)
    torch.manual_seed(int(o["seed"]))
    s = AdaptiveUniformSampling(opt, BlobField(), device="cpu", out_dir=str(tmp_path))
    pts, vols = s.sample()
    want_p, want_v = G[f"{tag}_points"], G[f"{tag}_volumes"]
    assert pts.shape == want_p.shape and pts.shape[0] > 3000
    assert np.array_equal(pts.numpy(), want_p)                       # same points, same order
    assert np.array_equal(vols.numpy(), want_v)
    # the ply carries the simulator's input schema (x, y, z, vp as f8), main_sample.py:14-23
    ply = read_ply_vertices(os.path.join(str(tmp_path), "blob", o["exp_name"] + ".ply"))
    assert sorted(ply) == ["vp", "x", "y", "z"] and ply["x"].dtype == np.float64
    assert np.array_equal(np.stack([ply["x"], ply["y"], ply["z"]], 1), want_p.astype(np.float64)) and np.array_equal(ply["vp"], want_v.astype(np.float64))


def test_sampling_properties():
    """Size-independent properties: every kept point is above the density threshold; a hash cell's points share its volume,
    hgs^3 / count, so the volumes of each occupied cell sum to hgs^3; explicit offsets make the result reproducible."""
    o = dict(CASES["a"]); opt = argparse.Namespace(**o)
    s = AdaptiveUniformSampling(opt, BlobField(), device="cpu")
    offsets = torch.rand((64, 3), dtype=torch.float32, generator=torch.Generator().manual_seed(5))
    pts, vols = s.sample(write=False, points_tmp=offsets)
    assert float(s.get_density(pts).min()) > opt.density_threshold
    bbmin = pts.min(0).values - 1e-3
    cell = torch.floor((pts - bbmin) / torch.tensor(opt.hash_grid_size, dtype=torch.float32)).to(torch.int64)
    key = (cell[:, 2] * 1000 + cell[:, 1]) * 1000 + cell[:, 0]
    uniq, inv = torch.unique(key, return_inverse=True)
    sums = torch.zeros(uniq.shape[0], dtype=torch.float64).index_add_(0, inv, vols.double())
    assert float((sums - opt.hash_grid_size ** 3).abs().max()) < 1e-9
    p2, v2 = s.sample(write=False, points_tmp=offsets.clone())
    assert torch.equal(p2, pts) and torch.equal(v2, vols)
    with pytest.raises(AssertionError, match="No points sampled"):
        AdaptiveUniformSampling(argparse.Namespace(**dict(o, density_threshold=0.999)), BlobField(), device="cpu").sample(write=False)
