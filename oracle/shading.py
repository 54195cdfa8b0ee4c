"""oracle/shading.py -- TEST INFRASTRUCTURE.  torch restatement of the training-step shading of the reference:

    lerp_specular                     utils/ops.py:99-119
    kd, ks, Ld, Ls, L                 train_brdf_crf.py:197-206

Pinned: `lerp_specular` is checked against the reference's own function (utils/ops.py imports with torch alone) by
tests/golden/make_golden.py, which also writes tests/golden/shading.npz (L and the autograd gradients w.r.t. albedo / roughness /
metallic for seeded inputs).  train_brdf_crf.py itself needs pytorch_lightning / mitsuba and cannot be imported; its six
arithmetic lines are restated verbatim below."""
import torch


def lerp_specular(specular, roughness):
    r_min, r_max = 0.02, 1.0
    r_num = specular.shape[-2]
    r = (roughness - r_min) / (r_max - r_min) * (r_num - 1)
    r1 = r.ceil().long()
    r0 = r.floor().long()
    r_ = r - r0
    s0 = torch.gather(specular, 1, r0[..., None].expand(r0.shape[0], 1, 3))[:, 0]
    s1 = torch.gather(specular, 1, r1[..., None].expand(r1.shape[0], 1, 3))[:, 0]
    return s0 * (1 - r_) + s1 * r_


def brdf_shading(albedo, roughness, metallic, diffuse, specular0, specular1, lerp=lerp_specular):
    kd = albedo * (1 - metallic)
    ks = 0.04 * (1 - metallic) + albedo * metallic
    Ld = kd * diffuse
    Ls = ks * lerp(specular0, roughness) + lerp(specular1, roughness)
    return Ld + Ls
