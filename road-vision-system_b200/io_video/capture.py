"""Frame source feeding the preprocessing chain.

`VideoSource.read()` keeps the contract of /root/reference/src/io_video/capture.py:10-24
(`Frame(ok, image, ts)` with `ts = time.time()` at capture).  `read_batch` is new: it fills a
pinned (B,H,W,3) buffer so PreprocessPipeline.process_batch can stream it to the GPU.

The reader is pluggable: by default `cv2.VideoCapture(source)`; any object with
`read() -> (ok, image)` and `release()` works (tests and the benchmark use synthetic readers).
"""
import time

import numpy as np


class Frame:
    __slots__ = ("ok", "image", "ts")

    def __init__(self, ok, image, ts):
        self.ok = ok
        self.image = image
        self.ts = ts


class SyntheticReader:
    """Cycles through a pool of pre-generated frames (no camera, no files)."""

    def __init__(self, frames, limit=None):
        self.frames = frames
        self.limit = limit
        self.i = 0

    def read(self):
        if self.limit is not None and self.i >= self.limit:
            return False, None
        img = self.frames[self.i % len(self.frames)]
        self.i += 1
        return True, img

    def release(self):
        pass


class VideoSource:
    def __init__(self, source=0, width=1280, height=720, fps_request=30, backend="auto", reader=None, pinned="auto"):
        """Same constructor as the reference (capture.py:11-16).  New, optional: `reader` (any object with read() / release())
        and `pinned`: a rvb200.Context (or "auto" = the process-wide default context when a GPU is usable, None / False = off).
        With it, read() hands out frames that already live in page-locked memory (fresh, caller-owned arrays whose blocks
        are recycled when they are garbage-collected), so `pipeline(raw)` uploads them with a plain DMA."""
        if reader is not None:
            self.cap = reader
        else:
            import cv2  # capture only; no cv2 arithmetic on the chain's path
            self.cap = cv2.VideoCapture(source)
            self.cap.set(cv2.CAP_PROP_FRAME_WIDTH, width)
            self.cap.set(cv2.CAP_PROP_FRAME_HEIGHT, height)
            self.cap.set(cv2.CAP_PROP_FPS, fps_request)
        self._pin_ctx = None
        self._shape = None
        self._into = None              # does the reader's read() accept a destination array (cv2.VideoCapture does)?
        if pinned == "auto":
            try:
                from .._native import default_context
                self._pin_ctx = default_context()
            except Exception:          # no GPU / library not built: frames stay in pageable memory (placement only)
                self._pin_ctx = None
        elif pinned:
            self._pin_ctx = pinned

    def _read_pinned(self):
        buf = self._pin_ctx._pooled_pinned(self._shape, keep=8)
        if self._into is not False:
            try:
                ok, img = self.cap.read(buf)
                self._into = True
            except TypeError:
                self._into = False
                ok, img = self.cap.read()
        else:
            ok, img = self.cap.read()
        if not ok or img is None:
            return ok, img
        if img is buf:
            return ok, img
        if isinstance(img, np.ndarray) and img.shape == self._shape and img.dtype == np.uint8:
            np.copyto(buf, img)
            return ok, buf
        self._shape = None             # geometry changed mid-stream: hand the frame out as it is
        return ok, img

    def read(self) -> Frame:
        if self._pin_ctx is not None and self._shape is not None:
            ok, img = self._read_pinned()
        else:
            ok, img = self.cap.read()
            if ok and self._pin_ctx is not None and isinstance(img, np.ndarray) and img.dtype == np.uint8 and img.ndim == 3:
                self._shape = img.shape        # from the next frame on, frames are captured into page-locked blocks
        return Frame(ok, img, time.time())

    def read_batch(self, n, out=None):
        """Read up to `n` frames into `out` (B,H,W,3) uint8 (e.g. from Context.pinned_empty).

        Returns (count, frames_view, timestamps): `count` frames were captured (fewer than n at end
        of stream), `frames_view = out[:count]`, `timestamps[i]` is time.time() at capture of frame i.
        """
        ts = np.zeros(n, np.float64)
        count = 0
        for i in range(n):
            ok, img = self.cap.read()
            if not ok or img is None:
                break
            if out is None:
                out = np.empty((n,) + img.shape, np.uint8)
            if img.shape != out.shape[1:]:
                raise ValueError(f"frame shape {img.shape} does not match the batch buffer {out.shape[1:]}")
            out[i] = img
            ts[i] = time.time()
            count += 1
        if out is None:
            return 0, None, ts[:0]
        return count, out[:count], ts[:count]

    def release(self):
        if self.cap:
            self.cap.release()
