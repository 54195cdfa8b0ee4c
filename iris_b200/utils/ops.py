"""Mirror of the part of the reference's utils/ops.py that sits on the training-step path (same name, arguments and result):

    lerp_specular(specular (B,R,3), roughness (B,1)) -> (B,3)          utils/ops.py:99-119

backed by the CUDA kernel behind `iris_brdf_shading_forward/backward` (gradient flows to `roughness`).  The fused form of the whole
block train_brdf_crf.py:193-206 is `iris_b200.ops.brdf_shading`.
"""
from ..ops import lerp_specular, brdf_shading  # noqa: F401
