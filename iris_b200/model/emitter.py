"""Emitter classes with the reference's interface (model/emitter.py): `SLFEmitter(emitter_path, slf_path)`,
`SLFEmitterLearn` (radiance is an nn.Parameter, `update_slf`), `AreaEmitter(emitter_path)`.

The fused estimators read these buffers through iris_b200.ops.tables_for; the per-call methods below
(`eval_emitter`, `sample_emitter`, `forward`) keep the reference's signatures for callers outside the fused path
(train_brdf_crf.py:395-409 logs `emitter_net.eval_emitter`, slf_refine.py queries `forward`)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .slf import VoxelSLF


def _load(path):
    return torch.load(path, map_location="cpu", weights_only=False)


class _TriangleEmitters(nn.Module):
    def _register_emitters(self, weight):
        is_emitter = weight["is_emitter"].bool()
        self.register_buffer("is_emitter", is_emitter)
        self.register_buffer("emitter_vertices", weight["emitter_vertices"].float())
        self.register_buffer("emitter_area", weight["emitter_area"].float())
        self.register_buffer("radiance", weight["emitter_radiance"].float())
        F, K = len(is_emitter), int(is_emitter.sum())
        emitter_idx = torch.full((F,), -1, dtype=torch.long)
        emitter_idx[is_emitter] = torch.arange(K)
        self.register_buffer("emitter_idx", emitter_idx)                       # face -> emitter row
        self.register_buffer("triangle_idx", torch.arange(F)[is_emitter])      # emitter row -> face
        pdf = torch.ones(K) / max(K, 1)                                         # uniform pick (model/emitter.py:170-171)
        self.register_buffer("emitter_pdf", pdf)
        self.register_buffer("emitter_cdf", pdf.cumsum(-1).contiguous())

    def _area_lights(self, triangle_idx):
        vis = triangle_idx != -1
        is_area = self.is_emitter[triangle_idx] & vis
        e = self.emitter_idx[triangle_idx].clamp_min(0)
        n = triangle_idx.shape[0]
        Le = torch.where(is_area[:, None], self.radiance[e], self.radiance.new_zeros(n, 3))
        pdf = torch.where(is_area, self.emitter_pdf[e] / self.emitter_area[e].clamp_min(1e-12), self.radiance.new_zeros(n))
        return Le, pdf, is_area, vis

    def sample_emitter(self, sample1, sample2, position):
        """model/emitter.py:224-255 -> wi (B,3), pdf (B,1) in area measure, sampled triangle id (B)."""
        K = len(self.emitter_pdf)
        e = torch.searchsorted(self.emitter_cdf, sample1.clamp_min(1e-12).contiguous()).clamp_max(K - 1)
        s = sample2[..., 0].sqrt()
        b0, b1 = (1 - s)[:, None], (s * sample2[..., 1])[:, None]
        tri = self.emitter_vertices[e]
        p = tri[:, 0] * b0 + tri[:, 1] * b1 + tri[:, 2] * (1 - b0 - b1)
        wi = torch.nn.functional.normalize(p - position, dim=-1)
        return wi, (self.emitter_pdf[e] / self.emitter_area[e].clamp_min(1e-12))[:, None], self.triangle_idx[e]


class AreaEmitter(_TriangleEmitters):
    """Triangle emitters without the radiance cache (model/emitter.py:15-131)."""

    def __init__(self, emitter_path):
        super().__init__()
        self._register_emitters(_load(emitter_path))

    def eval_emitter(self, position, light_dir, triangle_idx, *args):
        Le, pdf, is_area, vis = self._area_lights(triangle_idx)
        return Le, pdf[:, None], (~is_area) & vis


class SLFEmitter(_TriangleEmitters):
    """Triangle emitters + voxel surface light field (model/emitter.py:134-255)."""

    def __init__(self, emitter_path, slf_path):
        super().__init__()
        sd = _load(slf_path)
        self.slf = VoxelSLF(sd["mask"], sd["voxel_min"], sd["voxel_max"])
        self.slf.load_state_dict(sd["weight"])
        self._register_emitters(_load(emitter_path))

    def forward(self, position):
        return self.slf(position)["rgb"]

    def eval_emitter(self, position, light_dir, triangle_idx, roughness=None, trace_roughness=0.6):
        """model/emitter.py:180-221 -> Le (B,3), emit_pdf (B,1), valid_next (B)."""
        Le, pdf, is_area, vis = self._area_lights(triangle_idx)
        valid_next = (~is_area) & vis
        if roughness is not None:
            diffuse = valid_next & (roughness.squeeze(-1) > trace_roughness)
            S = torch.where(diffuse[:, None], self.slf(position)["rgb"], torch.zeros_like(Le))
            Le = Le + S
            valid_next = valid_next & ~(diffuse & (S.sum(-1) > 0))
        return Le, pdf[:, None], valid_next


class SLFEmitterLearn(SLFEmitter):
    """SLFEmitter whose `radiance` is trainable (model/emitter.py:257-275); gradients reach rows [0,K)."""

    def __init__(self, emitter_path, slf_path):
        super().__init__(emitter_path, slf_path)
        r = self.radiance
        del self._buffers["radiance"]
        self.radiance = nn.Parameter(r.clone())

    def update_slf(self, slf_path):
        sd = torch.load(slf_path, map_location=self.slf.radiance.device, weights_only=False)
        self.slf.load_state_dict(sd["weight"])
