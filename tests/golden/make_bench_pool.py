#!/usr/bin/env python3
"""Benchmark frame pool made by the REFERENCE's own fog synthesiser (this container only; the PNGs are committed).

    python tests/golden/make_bench_pool.py

Eight 1920x1080 frames: seeded clean road scene (rvb200.synth.clean_scene) -> `EnhancedFogSynthesizer`
(/root/reference/src/augment/fog.py:84-299) with exactly the parameters of tools/fog_batch.py:19-27 plus an explicit seed,
levels light / medium / heavy in turn -> rain streaks (rvb200.synth.add_rain; the reference has no rain generator).
Written as lossless PNG under tests/golden/pool_1080p/, with the SHA-1 of each decoded frame and of the reference
`PreprocessPipeline` output for the headline chain (YCrCb, clip 2, grid 8, k5) in pool_1080p/index.txt.
bench.py uses these frames as its pool; tests pin the GPU output against the recorded reference hashes.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)
sys.dont_write_bytecode = True

import cv2  # noqa: E402
from src.preprocess import PreprocessPipeline  # noqa: E402  (the reference)
from src.augment.fog import EnhancedFogSynthesizer  # noqa: E402  (the reference)

import rvb200  # noqa: E402,F401
from rvb200 import synth  # noqa: E402

H, W, COUNT = 1080, 1920, 8
LEVELS = ("light", "medium", "heavy")


def main():
    dest = os.path.join(HERE, "pool_1080p")
    os.makedirs(dest, exist_ok=True)
    cfg = {"enabled": True, "chain": [{"name": "CLAHEDehaze", "params": {"space": "YCrCb", "clip_limit": 2.0, "tile_grid": 8}},
                                      {"name": "MedianDerain", "params": {"ksize": 5}}]}
    pipe = PreprocessPipeline(cfg)
    lines = ["# file|level|seed|sha1(frame)|sha1(reference PreprocessPipeline output, YCrCb clip 2.0 grid 8 k5)|gray mean|gray std",
             f"# cv2 {cv2.__version__}; fog parameters = tools/fog_batch.py:19-27 + seed"]
    total = 0
    for i in range(COUNT):
        level, seed = LEVELS[i % 3], 2000 + i
        clean = synth.clean_scene(H, W, seed)
        fog = EnhancedFogSynthesizer(level=level, y_h_ratio=0.42, perlin_scale_ratio=0.18, perlin_octaves=2,
                                     horizon_softness=0.07, global_veil=0.5, depth_blur_max=4.0, seed=seed)
        frame = synth.add_rain(fog.synthesize(clean)[0], seed=seed)
        ok, png = cv2.imencode(".png", frame, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        assert ok and np.array_equal(cv2.imdecode(png, cv2.IMREAD_COLOR), frame)
        name = f"frame_{i}.png"
        with open(os.path.join(dest, name), "wb") as fh:
            fh.write(png.tobytes())
        total += len(png)
        ref = pipe(frame)
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        lines.append(f"{name}|{level}|{seed}|{hashlib.sha1(frame.tobytes()).hexdigest()}|{hashlib.sha1(ref.tobytes()).hexdigest()}|"
                     f"{gray.mean():.1f}|{gray.std():.1f}")
        print(lines[-1], len(png))
    with open(os.path.join(dest, "index.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("total PNG bytes", total)


if __name__ == "__main__":
    main()
