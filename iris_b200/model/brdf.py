"""BSDF classes with the reference's interface (model/brdf.py).

`BaseBRDF` carries the GGX / Lambert samplers and evaluators with the reference's signatures for callers that drive the
sampling loop themselves (bake_shading.py:113-114,173-177): on CUDA tensors `sample_diffuse` / `sample_specular` and the
gradient-free `sample_brdf` run on the library's kernel (iris_bsdf_sample -- the same device code, and the same fixed
sin / cos / asin / acos definitions, as the fused estimators); the evaluators, and `sample_brdf` when the material carries
gradients, are batched torch expressions; `NGPBRDF(voxel_min, voxel_max)` is the hash-grid + MLP material field whose forward and
adjoint run on the CUDA path (iris_field_forward / iris_field_backward) and whose state_dict key is the reference's
`mlp.params` (one flat fp32 vector [MLP 9216 | grid 27 954 112], tiny-cuda-nn layout)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as NF

N_MLP = 9216
N_GRID = 27954112


def _frame(n):
    # utils/ops.py:12-30
    ax = torch.zeros_like(n)
    use_x = n[..., 0].abs() <= 0.1
    ax[..., 0] = use_x.float()
    ax[..., 1] = (~use_x).float()
    t = NF.normalize(torch.cross(ax, n, dim=-1), dim=-1)
    return t, torch.cross(n, t, dim=-1)


def _to_world(theta, phi, n):
    st = torch.sin(theta)
    l = NF.normalize(torch.stack([st * torch.cos(phi), st * torch.sin(phi), torch.cos(theta)], -1), dim=-1)
    t, b = _frame(n)
    return l[..., 0:1] * t + l[..., 1:2] * b + l[..., 2:3] * n


def diffuse_sampler(sample2, normal):
    return _to_world(torch.asin(sample2[..., 0].sqrt()), math.pi * 2 * sample2[..., 1], normal)


def specular_sampler(sample2, roughness, wo, normal):
    alpha = (roughness * roughness).squeeze(-1).detach() if torch.is_tensor(roughness) else float(roughness) ** 2
    c2 = (1 - sample2[..., 0]) / (sample2[..., 0] * (alpha * alpha - 1) + 1)
    wh = _to_world(torch.acos(c2.sqrt()), 2 * math.pi * sample2[..., 1], normal)
    return NF.normalize(2 * (wo * wh).sum(-1, keepdim=True) * wh - wo, dim=-1)


def _cosines(wi, wo, n):
    h = NF.normalize(wi + wo, dim=-1)
    d = lambda a, b: (a * b).sum(-1, keepdim=True).relu()
    return d(wi, n), d(wo, n), d(wo, h), d(n, h)


def _D(NoH, r):
    a2 = (r * r) * (r * r)
    den = NoH * NoH * (a2 - 1.0) + 1.0
    return a2 / (math.pi * den * den)


def _G(NoV, NoL, r):
    k = (r + 1) * (r + 1) / 8
    return 1 / ((NoL * (1 - k) + k) * (NoV * (1 - k) + k))


class BaseBRDF(nn.Module):
    def forward(self):
        pass

    def eval_diffuse(self, wi, normal):
        pdf = (normal * wi).sum(-1, keepdim=True).relu() / math.pi
        return pdf.expand(len(wi), 3), pdf

    def sample_diffuse(self, sample2, normal):
        if normal.is_cuda:
            from .. import core
            wi, pdf, w0, _ = core.bsdf_sample(0, sample2.reshape(-1, 2), None, normal.reshape(-1, 3))
            return wi, pdf[:, None], w0
        wi = diffuse_sampler(sample2, normal)
        return wi, (normal * wi).sum(-1, keepdim=True).relu() / math.pi, torch.ones_like(normal)

    def eval_specular(self, wi, wo, normal, roughness):
        NoL, NoV, VoH, NoH = _cosines(wi, wo, normal)
        D = _D(NoH, roughness)
        x = (1 - VoH).pow(5)
        base = D * _G(NoV, NoL, roughness) / 4.0 * NoL
        return base * (1 - x), base * x, D.detach() / (4 * VoH.clamp_min(1e-4)) * NoH

    def sample_specular(self, sample2, wo, normal, roughness):
        if normal.is_cuda and not (torch.is_tensor(roughness) and roughness.requires_grad):
            from .. import core
            per_lane = torch.is_tensor(roughness) and roughness.numel() > 1      # (B,1) tensor, or the scalar level of bake_shading.py:147
            mat = None
            if per_lane:
                mat = torch.zeros(normal.shape[0], 5, device=normal.device)
                mat[:, 3] = roughness.reshape(-1)
            wi, pdf, w0, w1 = core.bsdf_sample(1, sample2.reshape(-1, 2), wo.reshape(-1, 3), normal.reshape(-1, 3), mat=mat,
                                               roughness=0.0 if per_lane else float(roughness))
            return wi, pdf[:, None], w0[:, :1], w1[:, :1]
        wi = specular_sampler(sample2, roughness, wo, normal)
        NoL, NoV, VoH, NoH = _cosines(wi, wo, normal)
        pdf = _D(NoH, roughness).detach() / (4 * VoH.clamp_min(1e-4)) * NoH
        x = (1 - VoH).pow(5)
        fac = _G(NoV, NoL, roughness) * VoH * NoL / NoH.clamp_min(1e-4)
        return wi, pdf, (1 - x) * fac, x * fac

    def eval_brdf(self, wi, wo, normal, mat):
        a, r, m = mat["albedo"], mat["roughness"], mat["metallic"]
        NoL, NoV, VoH, NoH = _cosines(wi, wo, normal)
        D = _D(NoH, r)
        pdf = 0.5 * (D.detach() / (4 * VoH.clamp_min(1e-4)) * NoH) + 0.5 * NoL / math.pi
        ks = 0.04 * (1 - m) + a * m
        F = ks + (1 - ks) * (1 - VoH).pow(5)
        return a * (1 - m) / math.pi * NoL + D * _G(NoV, NoL, r) * F / 4.0 * NoL, pdf

    def sample_brdf(self, sample1, sample2, wo, normal, mat):
        if normal.is_cuda and not any(v.requires_grad for v in mat.values()):
            from .. import core
            u = torch.cat([sample1.reshape(-1, 1), sample2.reshape(-1, 2)], 1)
            m = torch.cat([mat["albedo"], mat["roughness"], mat["metallic"]], -1)
            wi, pdf, w, _ = core.bsdf_sample(2, u, wo.reshape(-1, 3), normal.reshape(-1, 3), mat=m)
            return wi, pdf[:, None], w
        pick_diffuse = (sample1 > 0.5)[:, None]
        wi = torch.where(pick_diffuse, diffuse_sampler(sample2, normal), specular_sampler(sample2, mat["roughness"], wo, normal))
        brdf, pdf = self.eval_brdf(wi, wo, normal, mat)
        ok = pdf > 0
        w = torch.where(ok, brdf / torch.where(ok, pdf, torch.ones_like(pdf)), torch.zeros_like(brdf))   # zero (not NaN) gradient where pdf == 0
        return wi, pdf, torch.nan_to_num(w, nan=0.0)


class _FieldParams(nn.Module):
    """Parameter container in tiny-cuda-nn's layout: `params` = [W1 64x64 | W2 64x64 | W3 16x64 | 32 grid levels x 2 features]."""

    def __init__(self, seed=1337):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        p = torch.empty(N_MLP + N_GRID)
        p[:8192] = (torch.rand(8192, generator=g) * 2 - 1) * math.sqrt(6.0 / 128)       # Xavier uniform
        p[8192:N_MLP] = (torch.rand(1024, generator=g) * 2 - 1) * math.sqrt(6.0 / 80)
        p[N_MLP:] = (torch.rand(N_GRID, generator=g) * 2 - 1) * 1e-4                     # tcnn grid init U(-1e-4,1e-4)
        self.params = nn.Parameter(p)


class NGPBRDF(BaseBRDF):
    """Hash-grid BRDF field (model/brdf.py:213-260): forward(position) -> dict(albedo, roughness in [0.02,1], metallic)."""

    def __init__(self, voxel_min, voxel_max):
        super().__init__()
        self.voxel_min = voxel_min
        self.voxel_max = voxel_max
        self.mlp = _FieldParams()

    def forward(self, position):
        from .. import ops
        return ops.field(self, position)
