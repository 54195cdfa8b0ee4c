"""oracle/crf.py -- TEST INFRASTRUCTURE, not product code.

torch restatement of the reference's EmorCRF (crf/model_crf.py):
  forward            :68-86    clip(hdr*exposure,0,1), per-channel interpolation of crf = f0 + weight @ basis on linspace(0,1,n_bins)
  mono_increase_constraint :22-30, get_inv_crf :45-55, inverse :88-105
  cal_weight_fitting_crf :61-66 (least-squares fit of the weights to a given response)
The reference delegates the 1-D interpolation to `torch_interpolations.RegularGridInterpolator` (third-party, git HEAD,
environment.yml:43; ABSENT from /root/reference and from this image).  `regular_grid_interp_1d` restates that package's published
algorithm (sbarratt/torch_interpolations, multilinear.py: bucketize, clamped neighbours, distance-weighted mean with the
0/0 -> 1/1 guard); it also serves as the stub under which the reference's own module is imported to make tests/golden/crf.npz
(oracle/refharness.load_reference_crf).  Parity with torch_interpolations proper is unpinned; everything else is pinned by
that golden, produced by the reference's EmorCRF on the real EMoR tables (crf/emor.txt)."""
import numpy as np
import torch


def regular_grid_interp_1d(points, values, x):
    """RegularGridInterpolator([points], values)([x]) for one dimension.  points ascending (n,), values (n,), x any shape."""
    n = points.shape[0]
    right = torch.bucketize(x.detach().contiguous(), points.detach().contiguous())
    right = torch.where(right >= n, torch.full_like(right, n - 1), right)
    left = (right - 1).clamp(0, n - 1)
    dl = x - points[left]
    dr = points[right] - x
    dl = torch.where(dl < 0, torch.zeros_like(dl), dl)
    dr = torch.where(dr < 0, torch.zeros_like(dr), dr)
    both = (dl == 0) & (dr == 0)
    dl = torch.where(both, torch.ones_like(dl), dl)
    dr = torch.where(both, torch.ones_like(dr), dr)
    return (values[left] * dr + values[right] * dl) / (dl + dr)


class RegularGridInterpolator:
    """Drop-in for the 1-D use the reference makes of torch_interpolations.RegularGridInterpolator."""

    def __init__(self, points, values):
        assert len(points) == 1, "only the 1-D case is restated"
        self.points, self.values = points[0], values

    def __call__(self, xs):
        return regular_grid_interp_1d(self.points, self.values, xs[0])


def emor_forward(hdr, exposure, f0, basis, weight):
    crf = f0 + weight @ basis                                   # (3, n_bins)
    n = crf.shape[1]
    x = torch.clip(hdr * exposure, 0, 1)
    s = x * (n - 1)
    b = s.detach().floor().long().clamp(0, n - 2)
    w = s - b
    cols = torch.arange(3).expand_as(b)
    return crf[cols, b] * (1 - w) + crf[cols, b + 1] * w


def mono_increase_constraint(crf):
    """crf/model_crf.py:22-30: shift the finite differences so that none is negative, renormalise, integrate."""
    diff = crf[1:] - crf[:-1]
    dmin = diff.min()
    gap = -dmin if dmin < 0 else 0
    diff = diff + gap
    diff = diff / diff.sum()
    return torch.cat([torch.zeros(1), torch.cumsum(diff, 0)])


def inv_crf(f0, basis, weight):
    """crf/model_crf.py:45-55 -> (3, n_bins): the inverse response sampled on linspace(0,1,n_bins)."""
    crf = f0 + weight @ basis
    x = torch.linspace(0, 1, crf.shape[1])
    return torch.stack([regular_grid_interp_1d(mono_increase_constraint(crf[i]), x, x) for i in range(3)], 0)


def emor_inverse(ldr, exposure, f0, basis, weight):
    """crf/model_crf.py:88-105."""
    table = inv_crf(f0, basis, weight)
    x = torch.linspace(0, 1, table.shape[1])
    ldr = torch.clip(ldr, 0, 1)
    return torch.stack([regular_grid_interp_1d(x, table[i], ldr[:, i]) for i in range(3)], -1) / exposure


def fit_weight(crf, f0, basis):
    """crf/model_crf.py:61-66: weight (3,dim) = argmin |f0 + w @ basis - crf|, by the normal equations."""
    B = np.asarray(basis, np.float64).T
    return (np.linalg.inv(B.T @ B) @ B.T @ (np.asarray(crf, np.float64) - np.asarray(f0, np.float64)).T).T
