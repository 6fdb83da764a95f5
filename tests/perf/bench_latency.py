#!/usr/bin/env python3
"""Per-frame call latency of the drop-in contract `proc = pipeline(raw)` (main_preview.py:94) next to the reference's cv2 chain."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rvb200
from rvb200 import synth
from oracle import cv2_chain

for (h, w, space, k) in [(480, 640, "YCrCb", 3), (720, 1280, "YCrCb", 3), (720, 1280, "LAB", 3), (1080, 1920, "YCrCb", 5)]:
    img = synth.road_frame(h, w, 1)
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": space, "clip_limit": 2.0, "tile_grid": 8}},
                     {"name": "MedianDerain", "params": {"ksize": k}}]}
    pipe = rvb200.PreprocessPipeline(cfg)
    for _ in range(5):
        out = pipe(img)
    assert np.array_equal(out, cv2_chain.chain(img, space, 2.0, 8, k))
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); pipe(img); ts.append(time.perf_counter() - t0)
    ts.sort()
    tc = []
    for _ in range(10):
        t0 = time.perf_counter(); cv2_chain.chain(img, space, 2.0, 8, k); tc.append(time.perf_counter() - t0)
    tc.sort()
    print(json.dumps({"shape": [h, w], "space": space, "ksize": k, "gpu_call_ms_p50": round(1e3 * ts[25], 3), "gpu_call_ms_min": round(1e3 * ts[0], 3),
                      "cv2_call_ms_p50": round(1e3 * tc[5], 3)}), flush=True)
