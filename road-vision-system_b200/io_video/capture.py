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
    def __init__(self, source=0, width=1280, height=720, fps_request=30, backend="auto", reader=None):
        if reader is not None:
            self.cap = reader
        else:
            import cv2  # capture only; no cv2 arithmetic on the chain's path
            self.cap = cv2.VideoCapture(source)
            self.cap.set(cv2.CAP_PROP_FRAME_WIDTH, width)
            self.cap.set(cv2.CAP_PROP_FRAME_HEIGHT, height)
            self.cap.set(cv2.CAP_PROP_FPS, fps_request)

    def read(self) -> Frame:
        ok, img = self.cap.read()
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
