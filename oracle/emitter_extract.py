"""oracle/emitter_extract.py -- TEST INFRASTRUCTURE.  torch (CPU) restatement of the reference's emitter extraction,
extract_emitter_ldr.py:76-110.  The reference reduces with torch_scatter.scatter(..., reduce='sum') (third-party, absent here); a
scatter-sum over an index is restated with index_add_, everything else is the reference's own torch expression order.
Pinned by tests/golden/emitter_extract.npz: the output of the reference's own script executed unmodified through
oracle/refharness.run_extract_emitter_script (torch_scatter stubbed by its documented scatter-sum semantics)."""
import torch
import torch.nn.functional as NF


def extract(views, vertices, faces, threshold):
    """views: list of (triangle_idxs (n,) long, valid (n,) bool, rgbs (n,3))."""
    n_face = len(faces)
    triangle_radiance = torch.zeros(n_face, 3)
    triangle_count = torch.zeros(n_face)
    for triangle_idxs, valid, rgbs in views:
        idx = triangle_idxs[valid]
        triangle_radiance.index_add_(0, idx, rgbs[valid])
        triangle_count.index_add_(0, idx, torch.ones(len(idx)))
    mean = triangle_radiance / triangle_count.unsqueeze(-1).clamp_min(1)
    mean = torch.max(mean, dim=-1)[0]
    is_emitter = mean > threshold
    emitter_vertices = vertices[faces[is_emitter]]
    emitter_area = torch.cross(emitter_vertices[:, 1] - emitter_vertices[:, 0], emitter_vertices[:, 2] - emitter_vertices[:, 0], -1)
    emitter_normal = NF.normalize(emitter_area, dim=-1)
    emitter_area = emitter_area.norm(dim=-1) / 2.0
    return {"is_emitter": is_emitter, "emitter_vertices": emitter_vertices, "emitter_area": emitter_area, "emitter_normal": emitter_normal,
            "emitter_radiance": torch.zeros(n_face, 3), "triangle_count": triangle_count}
