"""Mirrors of the reference's scene-model classes (model/brdf.py, model/emitter.py, model/slf.py): same names, constructor
arguments, buffers / state_dict keys and method signatures, backed by the CUDA path where the reference's method is on the
hot path."""
from .brdf import BaseBRDF, NGPBRDF          # noqa: F401
from .emitter import AreaEmitter, SLFEmitter, SLFEmitterLearn   # noqa: F401
from .slf import VoxelSLF                     # noqa: F401
