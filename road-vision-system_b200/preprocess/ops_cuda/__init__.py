"""`CUDACLAHEDehaze` / `CUDAMedianDerain` -- the names the reference registers for its cv2.cuda
variants (/root/reference/src/preprocess/ops_cuda/cuda_clahe_dehaze.py:9, cuda_median_derain.py:8).
Here every op already runs on the GPU, so they are the same classes; unlike the reference's
(`cuda_clahe_dehaze.py:38-39`) there is no soft fallback to a CPU path."""
from ..ops import CLAHEDehaze, MedianDerain


class CUDACLAHEDehaze(CLAHEDehaze):
    pass


class CUDAMedianDerain(MedianDerain):
    pass


__all__ = ["CUDACLAHEDehaze", "CUDAMedianDerain"]
