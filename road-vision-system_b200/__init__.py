"""B200-native implementation of road-vision-system's preprocessing chain.

Hot path only: `CLAHEDehaze -> MedianDerain` behind the reference's plugin API
(/root/reference/src/preprocess), computed by hand-written sm_100a CUDA kernels through the
C ABI in include/rv_b200.h.  There is no CPU fallback: without the built library and a
Blackwell GPU every compute call raises.

The directory name contains a hyphen, so import it through the `rvb200` shim at the repo
root (`import rvb200`), or place this directory on a path under an importable name (for the
drop-in case: as `src` next to main_preview.py).
"""
from . import _native
from ._native import chunk_schedule, kernel_source_hash, kernel_sass_hash, kernel_sass_hashes, Context, DeviceArray, Params, RvError, default_context, library_path, build_library
from .preprocess import PreprocessPipeline
from .preprocess.base import PreprocessOp
from .preprocess.registry import REGISTRY, get_op_class
from .preprocess.ops import CLAHEDehaze, MedianDerain
from .io_video import VideoSource, Frame, FPSMeter, BatchFeeder, SyntheticReader

__all__ = [
    "chunk_schedule", "kernel_source_hash", "kernel_sass_hash", "kernel_sass_hashes", "Context", "DeviceArray", "Params", "RvError", "default_context", "library_path", "build_library",
    "PreprocessPipeline", "PreprocessOp", "REGISTRY", "get_op_class",
    "CLAHEDehaze", "MedianDerain", "VideoSource", "Frame", "FPSMeter", "BatchFeeder", "SyntheticReader",
]
