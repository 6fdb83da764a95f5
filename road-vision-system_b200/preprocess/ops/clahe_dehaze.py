"""CLAHEDehaze -- drop-in for /root/reference/src/preprocess/ops/clahe_dehaze.py:4-32.

Same name, params and per-frame contract; the cv2.cvtColor / split / createCLAHE().apply /
merge / cvtColor sequence (:19-30) runs as the sm_100a kernels behind rv_clahe_dehaze.
"""
from ..._native import default_context
from ..base import PreprocessOp, as_bgr_u8


def coerce(params):
    """clahe_dehaze.py:14-17 -- evaluated on every call, so mutating op.params takes effect."""
    space = str(params.get("space", "YCrCb")).upper()
    clip_limit = float(params.get("clip_limit", 2.0))
    grid = max(2, int(params.get("tile_grid", 8)))
    return ("LAB" if space == "LAB" else "YCrCb"), clip_limit, grid


class CLAHEDehaze(PreprocessOp):
    """CLAHE on the luminance channel. params: space "YCrCb"|"LAB", clip_limit float, tile_grid int."""

    def __call__(self, image):
        space, clip_limit, grid = coerce(self.params)
        img = as_bgr_u8(image)
        ctx = default_context(self.params.get("device"))
        return ctx.clahe_dehaze(img[None], space, clip_limit, grid)[0]
