"""MedianDerain -- drop-in for /root/reference/src/preprocess/ops/median_derain.py:4-14."""
from ..._native import default_context
from ..base import PreprocessOp, as_bgr_u8


def coerce(params):
    """median_derain.py:11-13: int, even -> +1, clamp to [3, 9]."""
    k = int(params.get("ksize", 3))
    if k % 2 == 0:
        k += 1
    return max(3, min(k, 9))


class MedianDerain(PreprocessOp):
    """k x k median per channel, BORDER_REPLICATE (cv2.medianBlur semantics). params: ksize 3/5/7/9."""

    def __call__(self, image):
        k = coerce(self.params)
        img = as_bgr_u8(image)
        ctx = default_context(self.params.get("device"))
        return ctx.median(img[None], k)[0]
