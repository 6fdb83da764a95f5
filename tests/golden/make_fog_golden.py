#!/usr/bin/env python3
"""Fog-synthesis fixtures made by the REFERENCE's own EnhancedFogSynthesizer (this container only; the result is committed).

    python tests/golden/make_fog_golden.py

For a handful of (shape, level, seed) cases -- fog_batch.py's parameters (tools/fog_batch.py:19-27) plus a seed, chosen so that the
gamma and sensor-noise branches (fog.py:286-291) are each taken and skipped -- records the reference's hazy frame (lossless PNG
bytes), its transmission map (float16) and the means of its beta / airlight maps.  Inputs are regenerated from their seeds
(rvb200.synth.clean_scene).  tests/test_fog.py replays the cases on the GPU.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)
sys.dont_write_bytecode = True

import cv2  # noqa: E402
from src.augment.fog import EnhancedFogSynthesizer  # noqa: E402  (the reference)

import rvb200  # noqa: E402,F401
from rvb200 import synth  # noqa: E402

KW = dict(y_h_ratio=0.42, perlin_scale_ratio=0.18, perlin_octaves=2, horizon_softness=0.07, global_veil=0.5, depth_blur_max=4.0)


def branches(level, seed):
    """Which optional branches a seed takes (replays the draws of fog.py without computing anything)."""
    rng = np.random.RandomState(seed)
    rng.rand(); rng.randint(1e9); rng.uniform(-0.02, 0.02, size=3); rng.rand(); rng.rand(); rng.rand(); rng.uniform(-0.015, 0.02, size=3)
    g = rng.rand() < 0.35
    if g:
        rng.uniform(-0.04, 0.05)
    return g, rng.rand() < 0.3


def main():
    want = {(False, False): None, (True, False): None, (False, True): None, (True, True): None}
    for seed in range(1, 200):
        b = branches("medium", seed)
        if want[b] is None:
            want[b] = seed
    cases = [((270, 480), "light", want[(False, False)]), ((270, 480), "medium", want[(True, False)]),
             ((216, 384), "heavy", want[(False, True)]), ((243, 431), "medium", want[(True, True)]),
             ((540, 960), "heavy", want[(False, False)] + 1000)]
    out, meta = {}, []
    for i, ((h, w), level, seed) in enumerate(cases):
        clean = synth.clean_scene(h, w, 500 + i)
        hazy, m = EnhancedFogSynthesizer(level=level, seed=seed, **KW).synthesize(clean)
        ok, png = cv2.imencode(".png", hazy, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        assert ok
        out[f"hazy_{i}"] = png
        out[f"t_{i}"] = m["t"].astype(np.float16)
        g, nz = branches(level, seed)
        meta.append(f"{i}|{h}|{w}|{level}|{seed}|{500 + i}|{int(g)}|{int(nz)}|{m['beta_map'].mean():.8f}|{m['A_map'].mean():.8f}|{m['t'].mean():.8f}")
        print(meta[-1], len(png))
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "fog_small.npz"), **out)
    print("fog_small.npz:", os.path.getsize(os.path.join(HERE, "fog_small.npz")), "bytes")


if __name__ == "__main__":
    main()
