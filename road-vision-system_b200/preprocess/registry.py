"""name -> operator class, as /root/reference/src/preprocess/registry.py:14-28."""
from typing import Dict, Type

from .base import PreprocessOp
from .ops import CLAHEDehaze, MedianDerain
from .ops_cuda import CUDACLAHEDehaze, CUDAMedianDerain

REGISTRY: Dict[str, Type[PreprocessOp]] = {
    "CLAHEDehaze": CLAHEDehaze,
    "MedianDerain": MedianDerain,
    "CUDACLAHEDehaze": CUDACLAHEDehaze,
    "CUDAMedianDerain": CUDAMedianDerain,
}


def get_op_class(name: str) -> Type[PreprocessOp]:
    if name not in REGISTRY:
        raise KeyError(f"Preprocess op '{name}' not found. Available: {list(REGISTRY.keys())}")
    return REGISTRY[name]
