#!/usr/bin/env python3
"""Six fog frames (1080p, medium, fog_batch.py parameters, no meta maps) for an ncu launch list: where does the synthesis spend its GPU time?"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
import rvb200  # noqa: E402
from rvb200 import synth  # noqa: E402
from rvb200.augment import EnhancedFogSynthesizer  # noqa: E402
from test_fog import KW  # noqa: E402

clean = [synth.clean_scene(1080, 1920, 950 + i) for i in range(2)]
fog = EnhancedFogSynthesizer(level="medium", seed=5, context=rvb200.default_context(), **KW)
for i in range(2):
    fog.synthesize(clean[i % 2], meta=False)
t0 = time.perf_counter()
for i in range(4):
    fog.synthesize(clean[i % 2], meta=False)
print("ms per frame (wall):", 1e3 * (time.perf_counter() - t0) / 4)
