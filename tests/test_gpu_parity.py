"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, the live cv2 and the
reference-made fixtures.  Bit-exact everywhere -- integer stages by construction, the float32
bilinear blend because every product/sum is individually rounded (no FMA contraction), so the
tolerance for final pixels is 0 LSB."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden_cases, load_sha_pins
from oracle import rv_oracle as O

pytestmark = pytest.mark.gpu


def ospace(space):
    return O.SPACE_LAB if space == "LAB" else O.SPACE_YCRCB


def frames_small():
    import rvb200
    from rvb200 import synth
    rng = np.random.RandomState(1)
    yield "fog", synth.road_frame(216, 384, 5)
    yield "uniform", rng.randint(0, 256, (180, 320, 3)).astype(np.uint8)
    yield "constant", np.full((64, 96, 3), 90, np.uint8)
    yield "ragged", rng.randint(0, 256, (123, 457, 3)).astype(np.uint8)
    yield "one_div", rng.randint(0, 256, (135, 243, 3)).astype(np.uint8)
    yield "narrow", rng.randint(0, 256, (300, 5, 3)).astype(np.uint8)
    yield "tiny", rng.randint(0, 256, (3, 2, 3)).astype(np.uint8)
    yield "one", rng.randint(0, 256, (1, 1, 3)).astype(np.uint8)
    yield "wide", rng.randint(0, 256, (33, 1000, 3)).astype(np.uint8)


# ------------------------------------------------------------------ stage level
@pytest.mark.parametrize("space", ["YCrCb", "LAB"])
@pytest.mark.parametrize("grid", [2, 7, 8, 16])
def test_luma_and_histograms(ctx, space, grid):
    for name, img in frames_small():
        hist, luma, mm = ctx.luma_hist(img[None], space, grid, want_luma=True, want_gray=True)
        pl = O.luma(img, ospace(space))
        assert np.array_equal(luma[0], pl), (name, "luma")
        assert np.array_equal(hist[0], O.clahe_hist(pl, grid)), (name, "hist")
        gray = O.bgr2gray(img)
        assert (int(mm[0, 0]), int(mm[0, 1])) == (int(gray.min()), int(gray.max())), (name, "gray")


@pytest.mark.parametrize("clip", [0.0, 0.001, 2.0, 3.7, 40.0])
def test_lut(ctx, clip):
    for name, img in frames_small():
        for grid in (2, 8, 16):
            pl = O.luma(img, O.SPACE_YCRCB)
            hist = O.clahe_hist(pl, grid)
            h, w = pl.shape
            got = ctx.build_lut(hist[None], h, w, grid, clip)[0]
            assert np.array_equal(got, O.clahe_lut(hist, h, w, grid, clip)), (name, grid, clip)


@pytest.mark.parametrize("k", [3, 5, 7, 9])
def test_median_alone(ctx, k):
    for name, img in frames_small():
        assert np.array_equal(ctx.median(img[None], k)[0], O.median(img, k)), (name, k)
    ties = (np.random.RandomState(k).randint(0, 3, (70, 130, 3)) * 100).astype(np.uint8)
    assert np.array_equal(ctx.median(ties[None], k)[0], O.median(ties, k))


@pytest.mark.parametrize("space", ["YCrCb", "LAB"])
def test_clahe_dehaze_alone(ctx, space):
    for name, img in frames_small():
        for grid, clip in [(8, 2.0), (2, 0.0), (16, 40.0), (7, 3.7)]:
            got = ctx.clahe_dehaze(img[None], space, clip, grid)[0]
            assert np.array_equal(got, O.clahe_dehaze(img, ospace(space), clip, grid)), (name, grid, clip)


def test_colour_round_trip_all_colours(ctx):
    """Exhaustive 2^24 colours through forward+inverse colour code: CLAHE with a 1-value-per-tile identity is not
    available, so compare the luminance plane exhaustively and the full chain on the same image."""
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([v & 255, (v >> 8) & 255, v >> 16], -1).astype(np.uint8).reshape(4096, 4096, 3)
    for space in ("YCrCb", "LAB"):
        _, luma, _ = ctx.luma_hist(img[None], space, 8, want_luma=True)
        assert np.array_equal(luma[0], O.luma(img, ospace(space))), space
        got = ctx.clahe_dehaze(img[None], space, 2.0, 8)[0]
        assert np.array_equal(got, O.clahe_dehaze(img, ospace(space), 2.0, 8)), space


# ------------------------------------------------------------------ whole chain
@pytest.mark.parametrize("space,grid,k", [("YCrCb", 8, 3), ("LAB", 8, 3), ("YCrCb", 8, 5), ("LAB", 16, 5),
                                          ("YCrCb", 7, 7), ("LAB", 2, 9)])
def test_chain_vs_oracle(ctx, space, grid, k):
    import rvb200
    for name, img in frames_small():
        p = rvb200.Params.make(space, 2.0, grid, k)
        got = ctx.chain(img[None], p)[0]
        assert np.array_equal(got, O.chain(img, ospace(space), 2.0, grid, k)), (name, space, grid, k)


def test_chain_vs_live_cv2(ctx):
    cv2 = pytest.importorskip("cv2")
    import rvb200
    from oracle import cv2_chain as R
    from rvb200 import synth
    img = synth.road_frame(720, 1280, 11)
    for space, grid, k in [("YCrCb", 8, 3), ("LAB", 8, 3), ("YCrCb", 8, 5)]:
        got = ctx.chain(img[None], rvb200.Params.make(space, 2.0, grid, k))[0]
        assert np.array_equal(got, R.chain(img, space, 2.0, grid, k)), (space, grid, k)


def test_golden_fixtures(ctx):
    import rvb200
    cases, gate, z = load_golden_cases()
    for c in cases:
        p = rvb200.Params.make(c["space"], c["clip"], c["grid"], c["k"])
        assert np.array_equal(ctx.chain(c["inp"][None], p)[0], c["out"]), c["idx"]
    for name, processed in gate:
        img = z[f"in_{name}"]
        pl = rvb200.PreprocessPipeline({"chain": [{"name": "CLAHEDehaze"}, {"name": "MedianDerain"}],
                                        "auto_gate": {"enable_low_contrast_gate": True, "contrast_thresh": 20.0}})
        res = pl(img)
        assert (res is not img) == bool(int(processed)), name


def test_sha_pins_benchmark_shapes(ctx):
    """Reference outputs at 720p / 1080p / ragged 1080x1923 / 540x964 (SHA-1 from the reference's own pipeline)."""
    import rvb200
    for p in load_sha_pins():
        img = np.random.RandomState(p["seed"]).randint(0, 256, (p["h"], p["w"], 3)).astype(np.uint8)
        got = ctx.chain(img[None], rvb200.Params.make(p["space"], 2.0, p["grid"], p["k"]))[0]
        assert hashlib.sha1(got.tobytes()).hexdigest() == p["sha"], p


def test_config2_full_size_vs_oracle(ctx):
    """BASELINE config 2 shape (1080p, YCrCb, k5) and config 3 shape (4K, grid 16) on fogged frames."""
    import rvb200
    from rvb200 import synth
    img = synth.road_frame(1080, 1920, 21)
    got = ctx.chain(img[None], rvb200.Params.make("YCrCb", 2.0, 8, 5))[0]
    assert np.array_equal(got, O.chain(img, O.SPACE_YCRCB, 2.0, 8, 5))
    big = np.tile(img, (2, 2, 1))
    got = ctx.chain(big[None], rvb200.Params.make("LAB", 2.0, 16, 3))[0]
    assert np.array_equal(got, O.chain(big, O.SPACE_LAB, 2.0, 16, 3))


# ------------------------------------------------------------------ plugin contract and batch entry
def test_plugin_ops_and_pipeline(ctx):
    import rvb200
    from rvb200 import synth
    img = synth.road_frame(240, 320, 3)
    ro = img.copy(); ro.setflags(write=False)
    a = rvb200.CLAHEDehaze(space="LAB", clip_limit=2.0, tile_grid=8)(ro)
    assert a.flags.writeable and a.flags.c_contiguous and a.shape == img.shape and a is not img
    assert np.array_equal(a, O.clahe_dehaze(img, O.SPACE_LAB, 2.0, 8))
    b = rvb200.MedianDerain(ksize=4)(a)
    assert np.array_equal(b, O.median(a, 5))
    cfg = {"enabled": True, "chain": [{"name": "CLAHEDehaze", "params": {"space": "LAB", "clip_limit": 2.0, "tile_grid": 8}},
                                      {"name": "MedianDerain", "params": {"ksize": 4}}]}
    fused = rvb200.PreprocessPipeline(cfg)(img, ts=0.0)
    assert np.array_equal(fused, b)                     # fused pass == op-by-op
    assert np.array_equal(img, synth.road_frame(240, 320, 3))      # input untouched
    view = np.asfortranarray(img)
    assert np.array_equal(rvb200.PreprocessPipeline(cfg)(view), b)


def test_batch_equals_per_frame_and_pinned(ctx):
    import rvb200
    from rvb200 import synth
    pool = synth.frame_pool(180, 320, 5, base_seed=30)
    frames = np.concatenate([pool, pool[::-1], pool[:3]])           # 13 frames
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": "YCrCb"}}, {"name": "MedianDerain", "params": {"ksize": 5}}]}
    pl = rvb200.PreprocessPipeline(cfg)
    want = np.stack([pl(f) for f in frames])
    ctx.set_option("chunk_frames", 4)                               # force several pipeline chunks
    ctx.set_option("group_frames", 3)
    try:
        got = pl.process_batch(frames)
        assert np.array_equal(got, want)
        pin_in = ctx.pinned_empty(frames.shape); pin_out = ctx.pinned_empty(frames.shape)
        pin_in[:] = frames
        res = pl.process_batch(pin_in, out=pin_out)
        assert res is pin_out and np.array_equal(pin_out, want)
        p = rvb200.Params.make("YCrCb", 2.0, 8, 5)
        pin_out[:] = 0
        ctx.submit(pin_in, pin_out, p); ctx.wait()
        assert np.array_equal(pin_out, want)
    finally:
        ctx.set_option("chunk_frames", 0)
        ctx.set_option("group_frames", 0)


def test_batch_gate_passthrough(ctx):
    import rvb200
    rng = np.random.RandomState(9)
    low = (100 + rng.randint(0, 10, (3, 90, 160, 3))).astype(np.uint8)
    high = rng.randint(0, 256, (2, 90, 160, 3)).astype(np.uint8)
    frames = np.concatenate([low[:1], high[:1], low[1:], high[1:]])
    cfg = {"chain": [{"name": "CLAHEDehaze"}, {"name": "MedianDerain"}],
           "auto_gate": {"enable_low_contrast_gate": True, "contrast_thresh": 20.0}}
    pl = rvb200.PreprocessPipeline(cfg)
    got = pl.process_batch(frames)
    for i, f in enumerate(frames):
        want = O.chain(f, O.SPACE_YCRCB, 2.0, 8, 3) if O.gray_span(f) < 20.0 else f
        assert np.array_equal(got[i], want), i


def test_device_buffers_with_pitch(ctx):
    """RV_MEM_DEVICE with a row pitch larger than 3*w and an odd base offset (unaligned slow paths)."""
    torch = pytest.importorskip("torch")
    import rvb200
    rng = np.random.RandomState(4)
    n, h, w = 3, 77, 203
    frames = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    for pitch, off in [(3 * w, 0), (640, 0), (3 * w + 7, 1)]:
        buf = torch.zeros(n * h * pitch + 16, dtype=torch.uint8, device="cuda")
        out = torch.zeros_like(buf)
        host = np.zeros((n, h, pitch), np.uint8)
        host[:, :, :3 * w] = frames.reshape(n, h, 3 * w)
        buf[off:off + host.size] = torch.from_numpy(host.reshape(-1)).cuda()
        p = rvb200.Params.make("LAB", 2.0, 8, 5)
        ctx.chain_device(buf.data_ptr() + off, out.data_ptr() + off, n, h, w, p, in_pitch=pitch, out_pitch=pitch)
        res = out[off:off + host.size].cpu().numpy().reshape(n, h, pitch)[:, :, :3 * w].reshape(n, h, w, 3)
        for i in range(n):
            assert np.array_equal(res[i], O.chain(frames[i], O.SPACE_LAB, 2.0, 8, 5)), (pitch, off, i)


def test_size_independent_properties_full_batch(ctx):
    """At BASELINE config 2's full size (64 x 1080p): every copy of a frame in the batch gives the identical result
    (checksum of checksums), a constant frame stays constant under the median, and the median is idempotent on it."""
    import rvb200
    from rvb200 import synth
    pool = synth.frame_pool(1080, 1920, 2, base_seed=50)
    frames = np.empty((64, 1080, 1920, 3), np.uint8)
    for i in range(64):
        frames[i] = pool[i % 2]
    p = rvb200.Params.make("YCrCb", 2.0, 8, 5)
    got = ctx.chain(frames, p)
    sums = [hashlib.sha1(got[i].tobytes()).hexdigest() for i in range(64)]
    assert len(set(sums[0::2])) == 1 and len(set(sums[1::2])) == 1
    assert np.array_equal(got[0], O.chain(pool[0], O.SPACE_YCRCB, 2.0, 8, 5))
    const = np.full((1, 1080, 1920, 3), 77, np.uint8)
    assert np.array_equal(ctx.median(const, 5), const)


def test_tma_and_plain_staging_agree(ctx):
    """k_chain stages its box with one TMA load per CTA on aligned buffers; the plain-load path must give the same bytes."""
    import rvb200
    from rvb200 import synth
    frames = synth.frame_pool(360, 640, 3, base_seed=70)
    p = rvb200.Params.make("YCrCb", 2.0, 8, 5)
    a = ctx.chain(frames, p)
    ctx.set_option("use_tma", 0)
    try:
        b = ctx.chain(frames, p)
    finally:
        ctx.set_option("use_tma", 1)
    assert np.array_equal(a, b)
    for i in range(3):
        assert np.array_equal(a[i], O.chain(frames[i], O.SPACE_YCRCB, 2.0, 8, 5))


def test_l2_prefetch_option_changes_nothing(ctx):
    """Option "prefetch_ctas" (k_chain also prefetches a later CTA's box into L2; off by default) must not change a byte."""
    import rvb200
    from rvb200 import synth
    frames = synth.frame_pool(360, 640, 5, base_seed=71)
    p = rvb200.Params.make("LAB", 2.0, 8, 3)
    a = ctx.chain(frames, p)
    for dist in (1, 3, 1000):              # 1000: every prefetch target lies beyond the grid
        ctx.set_option("prefetch_ctas", dist)
        try:
            b = ctx.chain(frames, p)
        finally:
            ctx.set_option("prefetch_ctas", 0)
        assert np.array_equal(a, b), dist


def test_letterbox_stage_and_fused_chain(ctx):
    """Detector-input stage: stand-alone letterbox and chain+letterbox, fused (integer scale: 1080p->640 /3,
    720p /2, 4K-ish /6) and unfused (ragged), all bit-exact in fp16 against the oracle restatement of cv2.resize."""
    import rvb200
    from rvb200 import synth
    rng = np.random.RandomState(6)
    for (h, w, size) in [(1080, 1920, 640), (720, 1280, 640), (480, 640, 640), (1080, 1923, 640), (123, 457, 320), (300, 200, 640)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        got = ctx.letterbox_f16(img[None], size)[0]
        assert np.array_equal(got.view(np.uint16), O.letterbox_f16(img, size).view(np.uint16)), (h, w, size)
        nw, nh, top, left, _ = ctx.letterbox_geometry(h, w, size)
        assert (nw, nh, top, left) == O.letterbox_geometry(h, w, size)
    cases = [(1080, 1920, 640, "YCrCb", 8, 5, 3), (720, 1280, 640, "LAB", 8, 3, 2), (1440, 3840, 640, "YCrCb", 16, 3, 6),
             (1080, 1923, 640, "YCrCb", 8, 3, 0), (480, 640, 640, "YCrCb", 8, 3, 1), (1080, 1920, 640, "YCrCb", 8, 0, 3)]
    for (h, w, size, space, grid, k, want_scale) in cases:
        frames = np.stack([synth.road_frame(h, w, 90 + i) for i in range(2)])
        assert ctx.letterbox_geometry(h, w, size)[4] == want_scale
        p = rvb200.Params.make(space, 2.0, grid, k)
        for want_full in (False, True):
            t, full = ctx.chain_letterbox(frames, p, size, want_full=want_full)
            for i in range(2):
                proc = O.chain(frames[i], ospace(space), 2.0, grid, k)
                assert np.array_equal(t[i].view(np.uint16), O.letterbox_f16(proc, size).view(np.uint16)), (h, w, k, want_full, i)
                if want_full:
                    assert np.array_equal(full[i], proc)


def test_process_batch_torch_cuda_tensor(ctx):
    """Device-resident entry: torch CUDA tensor in, torch CUDA tensor out, on torch's current stream."""
    torch = pytest.importorskip("torch")
    import rvb200
    from rvb200 import synth
    frames = synth.frame_pool(270, 480, 4, base_seed=120)
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": "LAB"}}, {"name": "MedianDerain", "params": {"ksize": 5}}]}
    pl = rvb200.PreprocessPipeline(cfg)
    want = pl.process_batch(frames)
    d = torch.from_numpy(frames).cuda()
    got = pl.process_batch(d)                                   # default stream -> synchronous path
    assert got.is_cuda and got.data_ptr() != d.data_ptr()
    assert np.array_equal(got.cpu().numpy(), want)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = torch.empty_like(d)
        res = pl.process_batch(d, out=out)                      # asynchronous on stream s
    s.synchronize()
    assert res is out and np.array_equal(out.cpu().numpy(), want)


def test_argument_errors(ctx):
    """Bad arguments fail loudly with ValueError (RV_ERR_ARG), nothing is computed."""
    torch = pytest.importorskip("torch")
    import rvb200
    img = np.zeros((1, 16, 16, 3), np.uint8)
    for bad in (dict(ksize=4), dict(ksize=11), dict(grid=1), dict(grid=1000)):
        kw = dict(space="YCrCb", clip_limit=2.0, grid=8, ksize=3)
        kw.update(bad)
        with pytest.raises(ValueError):
            ctx.chain(img, rvb200.Params.make(**kw))
    with pytest.raises(ValueError):
        ctx.chain(img, rvb200.Params.make(ksize=0, clahe=False))            # nothing to do
    d = torch.zeros(16 * 16 * 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ctx.chain_device(d.data_ptr(), d.data_ptr(), 1, 16, 16, rvb200.Params.make())   # in-place
    with pytest.raises(ValueError):
        ctx.chain(np.zeros((1, 16, 16, 3), np.float32), rvb200.Params.make())


def test_full_size_reference_fog_fixture_gpu(ctx):
    """The reference-fogged 720p frame (tests/golden/fog_720p.npz): CUDA output hashes equal the reference pipeline's."""
    cv2 = pytest.importorskip("cv2")
    import os
    import rvb200
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "fog_720p.npz"))
    frame = cv2.imdecode(z["png"], cv2.IMREAD_COLOR)
    for line in z["shas"]:
        space, grid, k, sha = str(line).split("|")
        got = ctx.chain(frame[None], rvb200.Params.make(space, 2.0, int(grid), int(k)))[0]
        assert hashlib.sha1(got.tobytes()).hexdigest() == sha, (space, grid, k)


def test_c_abi_from_plain_c(tmp_path):
    """A C program (gcc, no Python / torch types) drives the library through include/rv_b200.h and checks it against the C oracle."""
    import os
    import shutil
    import subprocess
    import rvb200
    from oracle import rv_oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    so, oso = rvb200.library_path(), rv_oracle.build()
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(root, "tests", "c_abi_smoke.c"), "-I", os.path.join(root, "include"),
                           so, oso, "-lm", "-Wl,-rpath," + os.path.dirname(so), "-Wl,-rpath," + os.path.dirname(oso)])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout
