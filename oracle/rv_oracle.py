"""ctypes front-end of oracle/librv_oracle.so (rv_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Every function restates one stage of the reference chain
(/root/reference/src/preprocess/ops/clahe_dehaze.py:19-30, ops/median_derain.py:10-14,
pipeline.py:24-30); see rv_oracle.c for the per-function citations.  The oracle is pinned
against the live cv2 and against fixtures produced by the reference's own classes
(tests/test_oracle_*.py, tests/golden/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SPACE_YCRCB, SPACE_LAB = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "librv_oracle.so")
    src = os.path.join(_HERE, "rv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "librv_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        for name in ("rvo_bgr2ycrcb", "rvo_ycrcb2bgr", "rvo_bgr2lab", "rvo_lab2bgr", "rvo_bgr2gray"):
            getattr(L, name).argtypes = [u8p, u8p, C.c_long]
            getattr(L, name).restype = None
        L.rvo_clahe_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.rvo_clahe_hist.argtypes = [u8p, C.c_int, C.c_int, C.c_long, C.c_int, i32p]
        L.rvo_clahe_lut.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_double, u8p]
        L.rvo_clahe_apply.argtypes = [u8p, C.c_int, C.c_int, C.c_long, C.c_int, u8p, u8p, C.c_long]
        L.rvo_median.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_int, u8p, C.c_long]
        L.rvo_clahe_dehaze.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, u8p]
        L.rvo_clahe_dehaze.restype = C.c_int
        L.rvo_chain.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, u8p]
        L.rvo_chain.restype = C.c_int
        L.rvo_gray_span.argtypes = [u8p, C.c_long]
        L.rvo_gray_span.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def _u8c(a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        raise TypeError("uint8 expected")
    return a


def _cvt(name, img):
    img = _u8c(img)
    out = np.empty_like(img)
    getattr(lib(), name)(_p(img), _p(out), img.size // 3)
    return out


def bgr2ycrcb(img): return _cvt("rvo_bgr2ycrcb", img)
def ycrcb2bgr(img): return _cvt("rvo_ycrcb2bgr", img)
def bgr2lab(img): return _cvt("rvo_bgr2lab", img)
def lab2bgr(img): return _cvt("rvo_lab2bgr", img)


def bgr2gray(img):
    img = _u8c(img)
    out = np.empty(img.shape[:2], np.uint8)
    lib().rvo_bgr2gray(_p(img), _p(out), out.size)
    return out


def gray_span(img):
    img = _u8c(img)
    return int(lib().rvo_gray_span(_p(img), img.size // 3))


def luma(img, space):
    """Luminance plane the CLAHE runs on: L (LAB) or Y (YCrCb)."""
    return np.ascontiguousarray((bgr2lab(img) if space == SPACE_LAB else bgr2ycrcb(img))[..., 0])


def clahe_geometry(H, W, grid):
    tw, th = C.c_int(), C.c_int()
    lib().rvo_clahe_geometry(H, W, grid, C.byref(tw), C.byref(th))
    return tw.value, th.value


def clahe_hist(plane, grid):
    plane = _u8c(plane)
    H, W = plane.shape
    hist = np.empty((grid * grid, 256), np.int32)
    lib().rvo_clahe_hist(_p(plane), H, W, W, grid, _p(hist, C.c_int32))
    return hist


def clahe_lut(hist, H, W, grid, clip_limit):
    hist = np.ascontiguousarray(hist, np.int32)
    lut = np.empty((grid * grid, 256), np.uint8)
    lib().rvo_clahe_lut(_p(hist, C.c_int32), H, W, grid, float(clip_limit), _p(lut))
    return lut


def clahe_apply(plane, grid, lut):
    plane = _u8c(plane)
    lut = _u8c(lut)
    H, W = plane.shape
    out = np.empty_like(plane)
    lib().rvo_clahe_apply(_p(plane), H, W, W, grid, _p(lut), _p(out), W)
    return out


def clahe_plane(plane, clip_limit, grid):
    """cv2.createCLAHE(clip_limit, (grid, grid)).apply(plane)"""
    H, W = plane.shape
    return clahe_apply(plane, grid, clahe_lut(clahe_hist(plane, grid), H, W, grid, clip_limit))


def median(img, k):
    img = _u8c(img)
    H, W = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    lib().rvo_median(_p(img), H, W, ch, W * ch, k, _p(out), W * ch)
    return out


def clahe_dehaze(img, space, clip_limit, grid):
    img = _u8c(img)
    H, W = img.shape[:2]
    out = np.empty_like(img)
    if lib().rvo_clahe_dehaze(_p(img), H, W, space, float(clip_limit), grid, _p(out)) != 0:
        raise MemoryError
    return out


def chain(img, space, clip_limit, grid, ksize):
    img = _u8c(img)
    H, W = img.shape[:2]
    out = np.empty_like(img)
    if lib().rvo_chain(_p(img), H, W, space, float(clip_limit), grid, ksize, _p(out)) != 0:
        raise MemoryError
    return out


# ---------------------------------------------------------------- detector-input stage
def letterbox_geometry(h, w, size=640):
    """ultralytics-style square letterbox (src/detect/yolo_ultralytics.py:28-35 delegates to it): Python round()."""
    r = min(size / h, size / w)
    nw, nh = max(1, int(round(w * r))), max(1, int(round(h * r)))
    dw, dh = (size - nw) / 2, (size - nh) / 2
    return nw, nh, int(round(dh - 0.1)), int(round(dw - 0.1))


def _resize_tab(ssize, dsize, is_x):
    scale = 1.0 / (float(dsize) / ssize)
    o0 = np.zeros(dsize, np.int64); o1 = np.zeros(dsize, np.int64); a = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f)); f = np.float32(f - np.float32(s))
        if is_x:
            if s < 0: s, f = 0, np.float32(0)
            if s >= ssize - 1: s, f = ssize - 1, np.float32(0)
            o0[d], o1[d] = s, min(s + 1, ssize - 1)
        else:
            o0[d], o1[d] = min(max(s, 0), ssize - 1), min(max(s + 1, 0), ssize - 1)
        a[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return o0, o1, a


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for 8-bit images (OpenCV resize.cpp: HResizeLinear /
    VResizeLinear<uchar,int,short>, coefficients scaled by 2048; pinned against cv2 in tests/test_oracle.py)."""
    H, W = src.shape[:2]
    x0, x1, xa = _resize_tab(W, dw, True)
    y0, y1, ya = _resize_tab(H, dh, False)
    s = src.astype(np.int64)
    hrow = s[:, x0] * xa[:, 0][None, :, None] + s[:, x1] * xa[:, 1][None, :, None]
    out = (((ya[:, 0][:, None, None] * (hrow[y0] >> 4)) >> 16) + ((ya[:, 1][:, None, None] * (hrow[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_f16(img, size=640, pad_value=114):
    """BGR (H,W,3) uint8 -> (3,size,size) float16: letterbox, BGR->RGB, HWC->CHW, float32(v)/255 rounded to half."""
    h, w = img.shape[:2]
    nw, nh, top, left = letterbox_geometry(h, w, size)
    canvas = np.full((size, size, 3), pad_value, np.uint8)
    canvas[top:top + nh, left:left + nw] = resize_linear_u8(img, nw, nh) if (nw, nh) != (w, h) else img
    return (canvas[:, :, ::-1].transpose(2, 0, 1).astype(np.float32) / np.float32(255)).astype(np.float16)
