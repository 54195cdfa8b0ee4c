"""Emitter extraction on the device (reference extract_emitter_ldr.py:76-116, mode 'export'):

    ex = EmitterExtractor(n_faces, device)
    for view in views:  ex.accumulate(triangle_idxs, valid, rgbs)      # :79-90   (triangle_idxs, valid from ray_intersect)
    emitter = ex.finalize(vertices, faces, threshold)                   # :92-110  -> the dict torch.save() writes as emitter.pth

The result has the reference's keys, dtypes and shapes: is_emitter (F,) bool, emitter_vertices (K,3,3), emitter_area (K,),
emitter_normal (K,3), emitter_radiance (F,3) zeros -- what SLFEmitter / iris_b200.core.ShadingTables.set_emitter load.
"""
from __future__ import annotations

import torch

from . import _capi as C


class EmitterExtractor:
    def __init__(self, n_faces, device="cuda:0"):
        self.F = int(n_faces)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("EmitterExtractor runs on a CUDA device (no CPU path)")
        self.tri_sum = torch.zeros(max(self.F, 1), 3, device=self.device)[: self.F]
        self.tri_count = torch.zeros(max(self.F, 1), dtype=torch.int32, device=self.device)[: self.F]

    def accumulate(self, triangle_idxs, valid, radiance):
        prim = triangle_idxs.reshape(-1).to(device=self.device, dtype=torch.int32).contiguous()
        radiance = radiance.reshape(-1, 3).to(self.device).float().contiguous()
        if valid is not None:
            valid = valid.reshape(-1).to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            C.check(C.lib().iris_tri_accumulate(C.ptr(prim), C.ptr(valid), C.ptr(radiance), prim.shape[0], self.F, C.ptr(self.tri_sum), C.ptr(self.tri_count),
                                                C.stream_ptr()))

    def finalize(self, vertices, faces, threshold):
        lib = C.lib()
        vertices = vertices.to(self.device).float().contiguous()
        faces32 = faces.to(device=self.device, dtype=torch.int32).contiguous()
        mask = torch.zeros(max(self.F, 1), dtype=torch.uint8, device=self.device)[: self.F]
        with torch.cuda.device(self.device):
            C.check(lib.iris_emitter_classify(C.ptr(self.tri_sum), C.ptr(self.tri_count), self.F, float(threshold), C.ptr(mask), C.stream_ptr()))
            is_emitter = mask.bool()
            idx = torch.nonzero(is_emitter).reshape(-1).contiguous()                 # increasing face index: boolean-mask order
            K = int(idx.shape[0])
            ev = torch.empty(K, 3, 3, device=self.device)
            area = torch.empty(K, device=self.device)
            normal = torch.empty(K, 3, device=self.device)
            C.check(lib.iris_emitter_geometry(C.ptr(vertices), C.ptr(faces32), C.ptr(idx), K, C.ptr(ev), C.ptr(area), C.ptr(normal), C.stream_ptr()))
        return {"is_emitter": is_emitter, "emitter_vertices": ev, "emitter_area": area, "emitter_normal": normal,
                "emitter_radiance": torch.zeros(self.F, 3, device=self.device)}
