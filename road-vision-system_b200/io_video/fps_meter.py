"""EMA frame-rate meter with the interface of /root/reference/src/io_video/fps_meter.py:3-18."""
import time


class FPSMeter:
    def __init__(self, alpha=0.1):
        self.alpha = alpha
        self._prev = None
        self.fps = 0.0

    def tick(self, now=None):
        now = now or time.time()
        if self._prev is not None:
            inst = 1.0 / max(1e-6, now - self._prev)
            self.fps += self.alpha * (inst - self.fps)
        self._prev = now
        return self.fps
