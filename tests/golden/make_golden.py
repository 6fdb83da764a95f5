#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE's own code (this container only).

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports, unchanged, from the read-only mount:
  src.preprocess.PreprocessPipeline      (/root/reference/src/preprocess/pipeline.py:7-45)
  src.augment.fog.EnhancedFogSynthesizer (/root/reference/src/augment/fog.py:84-299; parameters of
                                          tools/fog_batch.py:19-27 plus an explicit seed)
and records input frames + the reference's outputs.  /root/reference does not exist on the GPU
box, so the vectors are committed; tests/test_golden.py replays them against the oracle (CPU)
and tests/test_gpu_parity.py against the CUDA path (GPU).
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


def ref_cfg(space, clip, grid, k, gate=False, thresh=20.0):
    chain = [{"name": "CLAHEDehaze", "params": {"space": space, "clip_limit": clip, "tile_grid": grid}}]
    if k:
        chain.append({"name": "MedianDerain", "params": {"ksize": k}})
    return {"enabled": True, "chain": chain,
            "auto_gate": {"enable_low_contrast_gate": gate, "contrast_thresh": thresh}}


def main():
    inputs = {}
    # frames fogged by the reference's synthesiser (fog_batch.py parameters), then rained by ours
    for i, (level, shape) in enumerate([("light", (180, 320)), ("medium", (216, 384)), ("heavy", (123, 457))]):
        clean = synth.clean_scene(shape[0], shape[1], 100 + i)
        fog = EnhancedFogSynthesizer(level=level, y_h_ratio=0.42, perlin_scale_ratio=0.18, perlin_octaves=2,
                                     horizon_softness=0.07, global_veil=0.5, depth_blur_max=4.0, seed=1000 + i)
        inputs[f"fog_{level}"] = synth.add_rain(fog.synthesize(clean)[0], seed=i)
    rng = np.random.RandomState(42)
    inputs["uniform"] = rng.randint(0, 256, (96, 160, 3)).astype(np.uint8)
    inputs["constant"] = np.full((64, 96, 3), 137, np.uint8)
    inputs["lowcontrast"] = (120 + rng.randint(0, 12, (72, 128, 3))).astype(np.uint8)
    inputs["tiny"] = rng.randint(0, 256, (9, 11, 3)).astype(np.uint8)

    cases = [
        # (input, space, clip, grid, ksize)
        ("fog_light", "YCrCb", 2.0, 8, 3),      # configs/default.yaml chain
        ("fog_light", "LAB", 2.0, 8, 3),        # BASELINE.json config 1 wording
        ("fog_medium", "YCrCb", 2.0, 8, 5),     # config 2 chain
        ("fog_medium", "LAB", 2.0, 16, 3),      # config 3 chain
        ("fog_heavy", "LAB", 3.7, 8, 5),        # both dimensions ragged
        ("fog_heavy", "YCrCb", 2.0, 7, 7),
        ("uniform", "YCrCb", 40.0, 4, 9),
        ("uniform", "LAB", 0.0, 2, 0),          # clip 0: plain AHE, no median
        ("constant", "YCrCb", 2.0, 8, 3),
        ("lowcontrast", "LAB", 0.001, 8, 5),
        ("tiny", "YCrCb", 2.0, 8, 3),           # tiles of 2x2 after REFLECT_101 padding
    ]
    out = {f"in_{k}": v for k, v in inputs.items()}
    meta = []
    for idx, (name, space, clip, grid, k) in enumerate(cases):
        ref = PreprocessPipeline(ref_cfg(space, clip, grid, k))(inputs[name])
        out[f"out_{idx}"] = ref
        meta.append(f"{idx}|{name}|{space}|{clip}|{grid}|{k}")
    # the gate (pipeline.py:37-40): low-contrast frame is processed, a contrasty one is returned untouched
    gate = []
    for name in ("lowcontrast", "uniform"):
        pl = PreprocessPipeline(ref_cfg("YCrCb", 2.0, 8, 3, gate=True, thresh=20.0))
        res = pl(inputs[name])
        gate.append(f"{name}|{int(res is not inputs[name])}")
    out["meta"] = np.array(meta)
    out["gate"] = np.array(gate)
    np.savez_compressed(os.path.join(HERE, "chain_small.npz"), **out)

    # SHA-1 pins at benchmark shapes on integer-random frames (regenerated from the seed at test time)
    pins = []
    for (h, w, space, grid, k, seed) in [(720, 1280, "YCrCb", 8, 3, 1), (720, 1280, "LAB", 8, 3, 2),
                                         (1080, 1920, "YCrCb", 8, 5, 3), (1080, 1920, "LAB", 16, 3, 4),
                                         (1080, 1923, "YCrCb", 8, 3, 5), (540, 964, "LAB", 16, 5, 6)]:
        img = np.random.RandomState(seed).randint(0, 256, (h, w, 3)).astype(np.uint8)
        ref = PreprocessPipeline(ref_cfg(space, 2.0, grid, k))(img)
        pins.append(f"{h}|{w}|{space}|{grid}|{k}|{seed}|{hashlib.sha1(ref.tobytes()).hexdigest()}")
    with open(os.path.join(HERE, "chain_sha1.txt"), "w") as fh:
        fh.write("# h|w|space|grid|ksize|seed|sha1 of reference PreprocessPipeline output on RandomState(seed).randint(0,256,(h,w,3))\n")
        fh.write(f"# cv2 {cv2.__version__}\n")
        fh.write("\n".join(pins) + "\n")
    print("wrote", len(cases), "cases,", len(pins), "pins;", os.path.getsize(os.path.join(HERE, "chain_small.npz")), "bytes")




def full_size_fog_fixture():
    """One 1280x720 frame fogged by the reference's synthesiser (fog_batch.py parameters, seed 7) + rain, stored as PNG
    bytes, with SHA-1 of the reference PreprocessPipeline output for the default.yaml chain and the LAB/k5 chain."""
    clean = synth.clean_scene(720, 1280, 777)
    fog = EnhancedFogSynthesizer(level="medium", y_h_ratio=0.42, perlin_scale_ratio=0.18, perlin_octaves=2,
                                 horizon_softness=0.07, global_veil=0.5, depth_blur_max=4.0, seed=7)
    frame = synth.add_rain(fog.synthesize(clean)[0], seed=7)
    ok, png = cv2.imencode(".png", frame, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    assert ok and np.array_equal(cv2.imdecode(png, cv2.IMREAD_COLOR), frame)
    shas = []
    for space, grid, k in (("YCrCb", 8, 3), ("LAB", 8, 5)):
        ref = PreprocessPipeline(ref_cfg(space, 2.0, grid, k))(frame)
        shas.append(f"{space}|{grid}|{k}|{hashlib.sha1(ref.tobytes()).hexdigest()}")
    np.savez(os.path.join(HERE, "fog_720p.npz"), png=png, shas=np.array(shas))
    print("fog_720p.npz:", os.path.getsize(os.path.join(HERE, "fog_720p.npz")), "bytes")


if __name__ == "__main__":
    main()
    full_size_fog_fixture()
