"""Pin the CPU oracle (oracle/rv_oracle.c) against the live OpenCV and the reference-made fixtures.

The arithmetic of the path lives in OpenCV (the reference only calls it:
/root/reference/src/preprocess/ops/clahe_dehaze.py:19-30, ops/median_derain.py:14), so the
ground truth is cv2 itself plus tests/golden/ (outputs of the reference's own classes).
"""
import hashlib

import numpy as np
import pytest

from conftest import load_golden_cases, load_sha_pins
from oracle import rv_oracle as O

cv2 = pytest.importorskip("cv2")
from oracle import cv2_chain as R  # noqa: E402


def all_colours():
    """Every 24-bit colour once, as a 4096 x 4096 BGR image."""
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.empty((1 << 24, 3), np.uint8)
    img[:, 0] = v & 255
    img[:, 1] = (v >> 8) & 255
    img[:, 2] = v >> 16
    return img.reshape(4096, 4096, 3)


@pytest.mark.parametrize("name,code", [("bgr2ycrcb", "COLOR_BGR2YCrCb"), ("ycrcb2bgr", "COLOR_YCrCb2BGR"),
                                       ("bgr2lab", "COLOR_BGR2LAB"), ("lab2bgr", "COLOR_LAB2BGR")])
def test_colour_exhaustive(name, code):
    img = all_colours()
    want = cv2.cvtColor(img, getattr(cv2, code))
    got = getattr(O, name)(img)
    assert np.array_equal(got, want)


def test_gray_exhaustive():
    img = all_colours()
    assert np.array_equal(O.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def _planes():
    rng = np.random.RandomState(7)
    yield "uniform", rng.randint(0, 256, (270, 480)).astype(np.uint8)
    yield "foggy", np.clip(rng.normal(185, 12, (216, 384)), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[0:200, 0:300]
    yield "gradient", ((yy + xx) % 256).astype(np.uint8)
    yield "constant", np.full((96, 128), 200, np.uint8)
    yield "ragged", rng.randint(0, 256, (123, 457)).astype(np.uint8)
    yield "one_div", rng.randint(0, 256, (135, 243)).astype(np.uint8)      # H divisible by 9/5/3, W not by 8
    yield "small", rng.randint(0, 256, (16, 16)).astype(np.uint8)
    yield "thin", rng.randint(0, 256, (5, 300)).astype(np.uint8)


@pytest.mark.parametrize("grid", [2, 7, 8, 16])
@pytest.mark.parametrize("clip", [0.0, 0.001, 2.0, 3.7, 40.0])
def test_clahe_plane_matrix(grid, clip):
    for name, pl in _planes():
        want = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid)).apply(pl)
        got = O.clahe_plane(pl, clip, grid)
        assert np.array_equal(got, want), (name, grid, clip)


def test_clahe_intermediates_hist_and_lut():
    rng = np.random.RandomState(3)
    pl = np.clip(rng.normal(150, 30, (240, 320)), 0, 255).astype(np.uint8)
    grid, clip = 8, 2.0
    tw, th = O.clahe_geometry(240, 320, grid)
    assert (tw, th) == (40, 30)
    hist = O.clahe_hist(pl, grid)
    lut = O.clahe_lut(hist, 240, 320, grid, clip)
    for ty in range(grid):
        for tx in range(grid):
            crop = np.ascontiguousarray(pl[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw])
            assert np.array_equal(hist[ty * grid + tx], np.bincount(crop.ravel(), minlength=256))
            # a 1x1 CLAHE on the crop applies exactly this tile's LUT to every present grey level
            want = cv2.createCLAHE(clipLimit=clip, tileGridSize=(1, 1)).apply(crop)
            assert np.array_equal(lut[ty * grid + tx][crop], want)


def test_geometry_quirk_both_dims_padded():
    # A.3: if either dimension is ragged BOTH get padded, a divisible one by a full `grid`
    assert O.clahe_geometry(1080, 1923, 8) == (241, 136)
    assert O.clahe_geometry(1080, 1920, 8) == (240, 135)
    assert O.clahe_geometry(123, 457, 8) == (58, 16)


@pytest.mark.parametrize("k", [3, 5, 7, 9])
def test_median(k):
    rng = np.random.RandomState(k)
    for shape in [(64, 80, 3), (7, 5, 3), (1, 9, 3), (33, 1, 3), (40, 41)]:
        img = rng.randint(0, 256, shape).astype(np.uint8)
        assert np.array_equal(O.median(img, k), cv2.medianBlur(img, k)), (k, shape)
    img = (rng.randint(0, 4, (50, 60, 3)) * 60).astype(np.uint8)        # heavy ties
    assert np.array_equal(O.median(img, k), cv2.medianBlur(img, k))


def test_chain_vs_cv2_chain():
    rng = np.random.RandomState(11)
    img = rng.randint(0, 256, (180, 320, 3)).astype(np.uint8)
    for space, grid, k in [("YCrCb", 8, 3), ("LAB", 8, 5), ("LAB", 7, 7), ("YCrCb", 16, 9), ("LAB", 2, 0)]:
        want = R.chain(img, space, 2.0, grid, k)
        got = O.chain(img, O.SPACE_LAB if space == "LAB" else O.SPACE_YCRCB, 2.0, grid, k)
        assert np.array_equal(got, want), (space, grid, k)


def test_golden_cases_oracle_and_cv2_chain():
    cases, gate, _ = load_golden_cases()
    assert len(cases) >= 10
    for c in cases:
        got = O.chain(c["inp"], O.SPACE_LAB if c["space"] == "LAB" else O.SPACE_YCRCB, c["clip"], c["grid"], c["k"])
        assert np.array_equal(got, c["out"]), c["idx"]
        assert np.array_equal(R.chain(c["inp"], c["space"], c["clip"], c["grid"], c["k"]), c["out"]), c["idx"]


def test_golden_gate():
    _, gate, z = load_golden_cases()
    for name, processed in gate:
        assert (O.gray_span(z[f"in_{name}"]) < 20.0) == bool(int(processed))
        assert R.low_contrast(z[f"in_{name}"], 20.0) == bool(int(processed))


def test_sha_pins_full_size():
    """Reference PreprocessPipeline outputs at benchmark shapes (SHA-1 recorded by make_golden.py)."""
    for p in load_sha_pins()[:4]:
        img = np.random.RandomState(p["seed"]).randint(0, 256, (p["h"], p["w"], 3)).astype(np.uint8)
        got = O.chain(img, O.SPACE_LAB if p["space"] == "LAB" else O.SPACE_YCRCB, 2.0, p["grid"], p["k"])
        assert hashlib.sha1(got.tobytes()).hexdigest() == p["sha"], p


def test_ycrcb_forward_ranges():
    """The CUDA kernel drops two saturations the data can never trigger: over all 2^24 colours the un-saturated
    forward Cb stays inside [0,255] and the un-saturated Cr never goes negative (it does exceed 255)."""
    img = all_colours().reshape(-1, 3).astype(np.int64)
    B, G, R = img[:, 0], img[:, 1], img[:, 2]
    Y = (4899 * R + 9617 * G + 1868 * B + 8192) >> 14
    cr = ((R - Y) * 11682 + (128 << 14) + 8192) >> 14
    cb = ((B - Y) * 9241 + (128 << 14) + 8192) >> 14
    assert Y.min() >= 0 and Y.max() <= 255
    assert cb.min() >= 0 and cb.max() <= 255
    assert cr.min() >= 0 and cr.max() > 255


def test_letterbox_restatement_vs_cv2():
    """resize_linear_u8 / letterbox_f16 restate cv2.resize(INTER_LINEAR) + copyMakeBorder bit for bit (down, up, ragged)."""
    rng = np.random.RandomState(2)
    for (h, w, dh, dw) in [(1080, 1920, 360, 640), (720, 1280, 360, 640), (1080, 1923, 359, 640), (123, 457, 172, 640),
                           (600, 800, 480, 640), (97, 33, 640, 218), (36, 64, 72, 128), (36, 64, 50, 100), (2, 2, 5, 7)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        assert np.array_equal(O.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)), (h, w, dh, dw)
    for (h, w, size) in [(1080, 1920, 640), (720, 1280, 640), (480, 640, 640), (1080, 1923, 640), (123, 457, 320), (300, 200, 640), (640, 640, 640)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        a, b = O.letterbox_f16(img, size), R.letterbox_f16(img, size)
        assert a.dtype == np.float16 and a.shape == (3, size, size)
        assert np.array_equal(a.view(np.uint16), b.view(np.uint16)), (h, w, size)
    # the two ways of normalising a byte to half agree for every value: float32(v)/255 -> half  ==  half(v)/half(255)
    v = np.arange(256)
    assert np.array_equal((v.astype(np.float32) / np.float32(255)).astype(np.float16), (v.astype(np.float16) / np.float16(255)))


def test_lab_forward_ranges():
    """The CUDA kernel does not saturate the forward a/b channels: over all 2^24 colours they stay well inside [0,255]."""
    from oracle.gen_lab_tables import build_tables
    t = build_tables()
    G8, CB = t["G8"].astype(np.int64), t["CB"].astype(np.int64)
    v = np.arange(1 << 24, dtype=np.int64)
    B, G, R = G8[v & 255], G8[(v >> 8) & 255], G8[v >> 16]
    fX = CB[(1777 * R + 1541 * G + 778 * B + 2048) >> 12]
    fY = CB[(871 * R + 2929 * G + 296 * B + 2048) >> 12]
    fZ = CB[(73 * R + 448 * G + 3575 * B + 2048) >> 12]
    a = (500 * (fX - fY) + (128 << 15) + 16384) >> 15
    b = (200 * (fY - fZ) + (128 << 15) + 16384) >> 15
    assert (a.min(), a.max(), b.min(), b.max()) == (42, 226, 20, 223)


def _fog_720p():
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "fog_720p.npz"))
    frame = cv2.imdecode(z["png"], cv2.IMREAD_COLOR)
    return frame, [str(s).split("|") for s in z["shas"]]


def test_full_size_reference_fog_fixture():
    """A 1280x720 frame fogged by the reference's own EnhancedFogSynthesizer; SHA-1 of the reference pipeline's outputs."""
    frame, shas = _fog_720p()
    assert frame.shape == (720, 1280, 3)
    for space, grid, k, sha in shas:
        got = O.chain(frame, O.SPACE_LAB if space == "LAB" else O.SPACE_YCRCB, 2.0, int(grid), int(k))
        assert hashlib.sha1(got.tobytes()).hexdigest() == sha, (space, grid, k)
