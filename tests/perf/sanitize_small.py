#!/usr/bin/env python3
"""Smallest end-to-end exercise of every kernel for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tests/perf/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rvb200
from oracle import rv_oracle as O

ctx = rvb200.default_context(0)
rng = np.random.RandomState(0)
for (h, w) in [(64, 160), (37, 53), (130, 250)]:
    img = rng.randint(0, 256, (2, h, w, 3)).astype(np.uint8)
    for space, grid, k in [("YCrCb", 8, 5), ("LAB", 4, 3), ("YCrCb", 2, 7), ("LAB", 8, 9), ("YCrCb", 8, 0)]:
        got = ctx.chain(img, rvb200.Params.make(space, 2.0, grid, k))
        want = O.chain(img[1], O.SPACE_LAB if space == "LAB" else O.SPACE_YCRCB, 2.0, grid, k)
        assert np.array_equal(got[1], want), (h, w, space, grid, k)
    assert np.array_equal(ctx.median(img, 5)[0], O.median(img[0], 5))
    t, full = ctx.chain_letterbox(img, rvb200.Params.make("YCrCb", 2.0, 8, 3), 64, want_full=True)
    assert np.array_equal(t[0].view(np.uint16), O.letterbox_f16(full[0], 64).view(np.uint16))
    assert int(ctx.gray_span(img)[0]) == O.gray_span(img[0])
print("sanitize_small ok")
