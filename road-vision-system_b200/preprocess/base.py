"""Operator base class -- same contract as /root/reference/src/preprocess/base.py:4-16."""
from abc import ABC, abstractmethod
from typing import Any

import numpy as np


class PreprocessOp(ABC):
    """`op(image) -> image`; image is BGR (H,W,3) uint8; unknown params are kept and ignored."""

    def __init__(self, **params: Any):
        self.params = params

    @abstractmethod
    def __call__(self, image):
        pass


def as_bgr_u8(image):
    """Validate the per-frame contract: (H,W,3) uint8, any strides.

    The reference hands the array to cv2, which accepts strided / read-only / Fortran-order
    views and fails with cv2.error on other shapes or dtypes; here those failures are
    ValueError / TypeError raised before anything touches the GPU.
    """
    if not isinstance(image, np.ndarray):
        raise TypeError(f"image must be a numpy.ndarray, got {type(image).__name__}")
    if image.dtype != np.uint8:
        raise ValueError(f"image must be uint8, got {image.dtype}")
    if image.ndim != 3 or image.shape[2] != 3:
        raise ValueError(f"image must have shape (H,W,3), got {image.shape}")
    if image.shape[0] < 1 or image.shape[1] < 1:
        raise ValueError("empty image")
    return np.ascontiguousarray(image)
