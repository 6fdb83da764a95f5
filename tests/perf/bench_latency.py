#!/usr/bin/env python3
"""Per-frame call latency of the drop-in contract `proc = pipeline(raw)` (main_preview.py:94) next to the reference's cv2 chain.

Variants: input frame in pageable memory (any numpy array) or in page-locked memory (what this package's VideoSource.read()
hands out), CUDA graph replay on / off (option "frame_graphs").  Also prints the PCIe floor of the two copies alone."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rvb200
from rvb200 import synth
from oracle import cv2_chain

ctx = rvb200.default_context()


def p50(fn, n=200, warm=10):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return round(1e3 * ts[len(ts) // 2], 4), round(1e3 * ts[0], 4)


for (h, w, space, k) in [(480, 640, "YCrCb", 3), (720, 1280, "YCrCb", 3), (720, 1280, "LAB", 3), (1080, 1920, "YCrCb", 5)]:
    img = synth.road_frame(h, w, 1)
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": space, "clip_limit": 2.0, "tile_grid": 8}},
                     {"name": "MedianDerain", "params": {"ksize": k}}]}
    pipe = rvb200.PreprocessPipeline(cfg)
    want = cv2_chain.chain(img, space, 2.0, 8, k)
    pin = ctx.pinned_empty(img.shape)
    pin[:] = img
    row = {"shape": [h, w], "space": space, "ksize": k}
    for graphs in (1, 0):
        ctx.set_option("frame_graphs", graphs)
        for name, src in (("pageable", img), ("pinned", pin)):
            for _ in range(3):
                out = pipe(src)
            assert np.array_equal(out, want), (name, graphs)
            assert out is not src and out.flags.writeable
            row[f"{name}_{'graph' if graphs else 'direct'}_ms_p50"], row[f"{name}_{'graph' if graphs else 'direct'}_ms_min"] = p50(lambda: pipe(src))
    ctx.set_option("frame_graphs", 1)
    ctx.set_option("stage_threads", 3)                       # experiment: helper threads stage pageable input (default: the driver does)
    for _ in range(3):
        out = pipe(img)
    assert np.array_equal(out, want)
    row["pageable_graph_helper_threads_ms_p50"], _ = p50(lambda: pipe(img))
    ctx.set_option("stage_threads", 0)
    # floor: the two PCIe copies of one frame, back to back, nothing else
    dev = rvb200.DeviceArray(ctx, img.shape)
    res = ctx.pinned_empty(img.shape)
    lib, hh = ctx._lib, ctx._h

    def copies():
        lib.rv_memcpy(hh, dev.ptr, pin.ctypes.data, img.nbytes, 0)
        lib.rv_memcpy(hh, res.ctypes.data, dev.ptr, img.nbytes, 1)
    row["two_copies_only_ms_p50"], _ = p50(copies)
    tc = []
    for _ in range(10):
        t0 = time.perf_counter(); cv2_chain.chain(img, space, 2.0, 8, k); tc.append(time.perf_counter() - t0)
    tc.sort()
    row["cv2_call_ms_p50"] = round(1e3 * tc[5], 3)
    print(json.dumps(row), flush=True)
