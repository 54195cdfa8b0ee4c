"""oracle/crf.py -- TEST INFRASTRUCTURE.  torch restatement of EmorCRF.forward (reference crf/model_crf.py:68-86): clip(hdr*exposure,0,1)
then per-channel linear interpolation of crf = f0 + weight @ basis on linspace(0,1,n_bins).  The reference delegates the interpolation to
torch_interpolations (absent third-party, git HEAD, environment.yml:43): parity with it is unpinned; linear interpolation on a regular
grid is restated directly and differentiated by autograd."""
import torch


def emor_forward(hdr, exposure, f0, basis, weight):
    crf = f0 + weight @ basis                                   # (3, n_bins)
    n = crf.shape[1]
    x = torch.clip(hdr * exposure, 0, 1)
    s = x * (n - 1)
    b = s.detach().floor().long().clamp(0, n - 2)
    w = s - b
    cols = torch.arange(3).expand_as(b)
    return crf[cols, b] * (1 - w) + crf[cols, b + 1] * w
