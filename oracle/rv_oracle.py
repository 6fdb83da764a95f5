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
