"""Property tests (hypothesis): random frame shapes (including non-multiples of the grid, widths below one vector, heights
below the median radius), grids, clip limits and kernel sizes.  CPU: oracle vs live cv2.  GPU: CUDA path vs oracle."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import rv_oracle as O

cv2 = pytest.importorskip("cv2")
from oracle import cv2_chain as R  # noqa: E402

import os

BIG = int(os.environ.get("RV_PROP_SCALE", "1"))          # RV_PROP_SCALE=4: one-off deeper stress (bigger frames, 4x the examples)
shapes = st.tuples(st.integers(1, 96 * BIG), st.integers(1, 140 * BIG))
grids = st.sampled_from([2, 3, 5, 7, 8, 13, 16, 32])
clips = st.sampled_from([0.0, -1.0, 0.001, 0.5, 2.0, 3.7, 40.0, 1000.0])
ksizes = st.sampled_from([0, 3, 5, 7, 9])
spaces = st.sampled_from(["YCrCb", "LAB"])
kinds = st.sampled_from(["uniform", "narrow", "constant", "gradient", "binary"])


def make_frame(h, w, kind, seed):
    rng = np.random.RandomState(seed)
    if kind == "uniform":
        return rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    if kind == "narrow":
        return np.clip(rng.normal(180, 6, (h, w, 3)), 0, 255).astype(np.uint8)
    if kind == "constant":
        return np.full((h, w, 3), rng.randint(0, 256), np.uint8)
    if kind == "gradient":
        yy, xx = np.mgrid[0:h, 0:w]
        return np.stack([(yy * 3 + xx) % 256, (xx * 2) % 256, (yy + 2 * xx) % 256], -1).astype(np.uint8)
    return (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(shapes, grids, clips, ksizes, spaces, kinds, st.integers(0, 10_000))
def test_oracle_equals_cv2(shape, grid, clip, k, space, kind, seed):
    img = make_frame(shape[0], shape[1], kind, seed)
    want = R.chain(img, space, clip, grid, k)
    got = O.chain(img, O.SPACE_LAB if space == "LAB" else O.SPACE_YCRCB, clip, grid, k)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@settings(max_examples=120 * BIG, deadline=None, suppress_health_check=list(HealthCheck))
@given(shapes, grids, clips, ksizes, spaces, kinds, st.integers(0, 10_000), st.integers(1, 3))
def test_gpu_equals_oracle(shape, grid, clip, k, space, kind, seed, n):
    import rvb200
    ctx = rvb200.default_context()
    frames = np.stack([make_frame(shape[0], shape[1], kind, seed + i) for i in range(n)])
    got = ctx.chain(frames, rvb200.Params.make(space, clip, grid, k))
    for i in range(n):
        want = O.chain(frames[i], O.SPACE_LAB if space == "LAB" else O.SPACE_YCRCB, clip, grid, k)
        assert np.array_equal(got[i], want), (shape, grid, clip, k, space, kind, seed, i)


@pytest.mark.gpu
@settings(max_examples=40 * BIG, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.tuples(st.integers(1, 200), st.integers(1, 300)), st.sampled_from([32, 64, 96, 160]), kinds, st.integers(0, 10_000))
def test_gpu_letterbox_equals_oracle(shape, size, kind, seed):
    import rvb200
    ctx = rvb200.default_context()
    img = make_frame(shape[0], shape[1], kind, seed)
    got = ctx.letterbox_f16(img[None], size)[0]
    assert np.array_equal(got.view(np.uint16), O.letterbox_f16(img, size).view(np.uint16)), (shape, size, kind, seed)
