#!/usr/bin/env python3
"""Fog frames of the loaded library for a fixed set of seeds / levels / sizes -> npz (for an A/B between two builds), and its timing.
usage: RV_B200_LIB=... tools/exp_fog_ab.py out.npz"""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
import rvb200  # noqa: E402
from rvb200 import synth  # noqa: E402
from rvb200.augment import EnhancedFogSynthesizer  # noqa: E402
from test_fog import KW  # noqa: E402

out = {}
ctx = rvb200.default_context()
for (h, w) in ((1080, 1920), (360, 640), (97, 131)):
    clean = synth.clean_scene(h, w, 950)
    for level in ("light", "medium", "heavy"):
        for seed in (5, 6):
            fog = EnhancedFogSynthesizer(level=level, seed=seed, context=ctx, **KW)
            hazy, meta = fog.synthesize(clean)
            out[f"{h}x{w}_{level}_{seed}"] = hazy
            out[f"{h}x{w}_{level}_{seed}_t"] = meta["t"]
            out[f"{h}x{w}_{level}_{seed}_A"] = meta["A_map"]
np.savez_compressed(sys.argv[1], **out)
clean = [synth.clean_scene(1080, 1920, 950 + i) for i in range(4)]
fog = EnhancedFogSynthesizer(level="medium", seed=5, context=ctx, **KW)
for i in range(3):
    fog.synthesize(clean[i % 4], meta=False)
n, t0 = 0, time.perf_counter()
while time.perf_counter() - t0 < 3.0:
    fog.synthesize(clean[n % 4], meta=False)
    n += 1
print("frames/s without meta maps:", round(n / (time.perf_counter() - t0), 1), os.environ.get("RV_B200_LIB", "librv_b200.so"))
