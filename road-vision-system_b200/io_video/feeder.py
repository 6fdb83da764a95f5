"""Capture -> pinned batch feeder (no reference counterpart: main_preview.py:88-142 keeps one frame in flight).

A reader thread fills a ring of pinned (B,H,W,3) buffers from a VideoSource while the consumer runs
PreprocessPipeline.process_batch on the previous one, so capture, H2D/kernels/D2H and the consumer overlap.
Per-frame capture timestamps (time.time() at read, as src/io_video/capture.py:20) travel with each batch because the
tracker downstream consumes them (main_preview.py:103).
"""
import queue
import threading

import numpy as np


class Batch:
    __slots__ = ("frames", "ts", "count", "slot")

    def __init__(self, frames, ts, count, slot):
        self.frames, self.ts, self.count, self.slot = frames, ts, count, slot


class BatchFeeder:
    """Iterate over batches captured in the background.

        feeder = BatchFeeder(source, batch=16, shape=(1080, 1920, 3), alloc=ctx.pinned_empty)
        for b in feeder:                       # b.frames: (count,H,W,3) view of a pinned buffer, b.ts: capture times
            out = pipeline.process_batch(b.frames, out=my_out[:b.count])
            feeder.release(b)                  # hand the buffer back to the reader

    `alloc(shape)` returns a uint8 array (Context.pinned_empty for page-locked memory; numpy.empty works for tests).
    `depth` buffers are cycled; the reader blocks when the consumer holds them all (back-pressure, no frame drops).
    """

    def __init__(self, source, batch, shape, alloc=None, depth=3, fps=None, clock=None):
        import time
        self.source, self.batch, self.shape = source, int(batch), tuple(shape)
        alloc = alloc or (lambda s: np.empty(s, np.uint8))
        self.buffers = [alloc((self.batch,) + self.shape) for _ in range(depth)]
        self.free = queue.Queue()
        for i in range(depth):
            self.free.put(i)
        self.ready = queue.Queue()
        self.period = (1.0 / fps) if fps else 0.0          # optional pacing: one frame every 1/fps seconds (camera rate)
        self.clock = clock or time
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            t_next = self.clock.time()
            while True:
                slot = self.free.get()
                if slot is None:
                    break
                buf = self.buffers[slot]
                ts = np.zeros(self.batch, np.float64)
                count = 0
                for i in range(self.batch):
                    if self.period:
                        now = self.clock.time()
                        if now < t_next:
                            self.clock.sleep(t_next - now)
                        t_next = max(t_next + self.period, self.clock.time() - self.period)
                    fr = self.source.read()
                    if not fr.ok or fr.image is None:
                        break
                    if fr.image.shape != self.shape:
                        raise ValueError(f"frame shape {fr.image.shape} does not match {self.shape}")
                    buf[i] = fr.image
                    ts[i] = fr.ts
                    count += 1
                if count:
                    self.ready.put(Batch(buf[:count], ts[:count], count, slot))
                if count < self.batch:                      # end of stream
                    break
        except Exception as e:                              # surfaced to the consumer
            self.error = e
        finally:
            self.ready.put(None)

    def __iter__(self):
        while True:
            b = self.ready.get()
            if b is None:
                if self.error:
                    raise self.error
                return
            yield b

    def release(self, batch):
        self.free.put(batch.slot)

    def close(self):
        self.free.put(None)
