#!/usr/bin/env python3
"""A/B of the chunk schedule of the host pipeline (option "chunk_taper") on the end-to-end legs of bench.py: one process, the
two settings alternate round by round so that box-to-box and minute-to-minute differences cancel.  usage: tools/exp_taper.py [rounds]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvb200  # noqa: E402
from rvb200 import synth  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    H, W, B, S = 1080, 1920, 64, 640
    ctx = rvb200.Context(0)
    pool = synth.frame_pool(H, W, 4, base_seed=3000)
    host = np.stack([pool[i % len(pool)] for i in range(B)])
    pl = rvb200.PreprocessPipeline({"enabled": True, "chain": [{"name": "CLAHEDehaze", "params": {"space": "YCrCb", "clip_limit": 2.0, "tile_grid": 8}},
                                                               {"name": "MedianDerain", "params": {"ksize": 5}}]}, context=ctx)
    pin_in, pin_out = ctx.pinned_empty(host.shape), ctx.pinned_empty(host.shape)
    pin_in[:] = host
    pin_t = ctx.pinned_empty((B, 3, S, S), np.float16)
    pin_t[:] = 0
    ctx.fill_tensor_padding(pin_t, H, W, S)
    legs = {
        "e2e": lambda: pl.process_batch(pin_in, out=pin_out),
        "e2e_tensor": lambda: pl.process_batch_to_tensor(pin_in, size=S, out=pin_t, padding_present=True),
        "e2e_keep": lambda: pl.process_batch_to_tensor(pin_in, size=S, out="device"),
    }
    ref = {}
    res = {(leg, t): [] for leg in legs for t in (0, 1)}
    for r in range(rounds):
        for taper in (0, 1):
            ctx.set_option("chunk_taper", taper)
            for leg, fn in legs.items():
                for _ in range(2):
                    fn()
                t0 = time.perf_counter()
                n = 10
                for _ in range(n):
                    fn()
                dt = time.perf_counter() - t0
                res[(leg, taper)].append(B * n / dt)
                if leg == "e2e":                                   # same bytes whatever the schedule
                    cur = pin_out.copy()
                    assert ref.setdefault("e2e", cur) is cur or np.array_equal(ref["e2e"], cur)
                if leg == "e2e_tensor":
                    cur = pin_t.copy()
                    assert ref.setdefault("t", cur) is cur or np.array_equal(ref["t"].view(np.uint16), cur.view(np.uint16))
    for leg in legs:
        a, b = np.array(res[(leg, 0)]), np.array(res[(leg, 1)])
        print(json.dumps({"leg": leg, "uniform_fps": [round(x) for x in a], "taper_fps": [round(x) for x in b],
                          "uniform_median": round(float(np.median(a))), "taper_median": round(float(np.median(b))),
                          "gain_pct": round(100 * (float(np.median(b)) / float(np.median(a)) - 1), 2)}))


if __name__ == "__main__":
    main()
