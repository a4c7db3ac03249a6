"""Inputs shared by tests/golden/make_golden_sampling.py (the reference's AdaptiveUniformSampling through the warp shim) and
tests/test_sampling.py (ours): an analytic density field and the option sets."""
import torch


class BlobField(torch.nn.Module):
    """sigma(x): two soft blobs and a slab, float32 torch ops only (identical on both sides of the comparison)."""

    def to(self, *a, **k):
        return self

    def density(self, x):
        x = x.to(torch.float32)
        c0 = torch.tensor([0.15, -0.1, 0.05], device=x.device); c1 = torch.tensor([-0.35, 0.3, -0.2], device=x.device)
        b0 = torch.exp(-((x - c0) ** 2).sum(-1) / 0.08) * 400.0
        b1 = torch.exp(-((x - c1) ** 2).sum(-1) / 0.03) * 250.0
        slab = 120.0 * torch.sigmoid((0.12 - (x[..., 1] + 0.55).abs()) * 40.0) * torch.sigmoid((0.6 - x[..., 0].abs()) * 30.0)
        return {"sigma": b0 + b1 + slab}


CASES = {
    # the defaults of get_opts.py:77-79,96 (sim_dx 0.05 -> hash_grid_size 0.06) at bound 1, with a larger sub_coeff so cells add points
    "a": dict(bound=1.0, density_threshold=0.05, sub_res=20, sub_coeff=0.6, hash_grid_size=0.06, cut=False,
              cut_bounds=[0.0, 2.0, -2.0, 1.0, -1.42, 0.92], workspace="ws/blob", exp_name="exp", seed=11),
    # cut box + a different resolution / threshold
    "b": dict(bound=1.0, density_threshold=0.2, sub_res=16, sub_coeff=1.1, hash_grid_size=0.09, cut=True,
              cut_bounds=[-0.8, 0.6, -2.0, 0.7, -0.5, 0.92], workspace="ws/blob", exp_name="cut", seed=12),
}
