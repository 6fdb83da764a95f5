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


@pytest.mark.parametrize("k", [3, 5, 7, 9])
def test_median_zero_one_frames(ctx, k):
    """0-1 principle on the device code itself: a min/max selection network is the median iff it is right on 0-1 inputs, and the
    inputs that can tell a wrong network apart have about half ones in the window.  Random binary frames (two grey levels) with
    densities around 1/2 put hundreds of thousands of windows per channel at exactly (k*k-1)/2 and (k*k+1)/2 ones; the frame is
    wider than one 120-pixel tile and taller than one 48-row tile, so every group, lane (rows s / s+24) and tile edge is hit."""
    h, w = 200, 380
    for dens, lo, hi, seed in [(0.5, 0, 255, 1), (0.47, 17, 18, 2), (0.53, 254, 255, 3), (0.5, 0, 1, 4)]:
        rng = np.random.RandomState(100 * k + seed)
        img = np.where(rng.rand(h, w, 3) < dens, hi, lo).astype(np.uint8)
        assert np.array_equal(ctx.median(img[None], k)[0], O.median(img, k)), (k, dens, lo, hi)


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


def test_host_pipeline_chunk_schedules(ctx):
    """Every chunk schedule of the host pipeline (uniform / tapered ends, chunk sizes 2..6, batch sizes around the multiples of the
    chunk size) returns the frames of the per-frame path, each exactly once and in order; the detector tensor likewise."""
    import rvb200
    from rvb200 import synth
    pool = synth.frame_pool(96, 168, 6, base_seed=77)
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": "LAB", "tile_grid": 4}}, {"name": "MedianDerain", "params": {"ksize": 3}}]}
    pl = rvb200.PreprocessPipeline(cfg)
    per_frame = [pl(f) for f in pool]
    try:
        for taper in (1, 0):
            ctx.set_option("chunk_taper", taper)
            for chunk in (2, 3, 4, 6):
                ctx.set_option("chunk_frames", chunk)
                for n in (1, 2, 3 * chunk, 3 * chunk + 1, 4 * chunk - 1, 5 * chunk + 2, 23):
                    idx = [(5 * i + n) % len(pool) for i in range(n)]
                    frames = np.stack([pool[i] for i in idx])
                    got = pl.process_batch(frames)
                    assert all(np.array_equal(got[j], per_frame[i]) for j, i in enumerate(idx)), (taper, chunk, n)
                    if chunk == 3:
                        t, _ = pl.process_batch_to_tensor(frames, size=64)
                        want, _ = pl.process_batch_to_tensor(frames[:1], size=64)
                        assert np.array_equal(t[0].view(np.uint16), want[0].view(np.uint16)) and t.shape == (n, 3, 64, 64), (taper, n)
    finally:
        ctx.set_option("chunk_frames", 0)
        ctx.set_option("chunk_taper", 1)


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


# ------------------------------------------------------------------ round 2: streams, limits, mixed host/device ends
def test_two_caller_streams_share_one_context(ctx):
    """two batches enqueued on different caller streams through one context use the same workspace set; the
    context orders them with events, so neither overwrites the other's histograms / quad tables mid-use."""
    torch = pytest.importorskip("torch")
    import rvb200
    from rvb200 import synth
    a = synth.frame_pool(360, 640, 6, base_seed=200)
    b = np.random.RandomState(5).randint(0, 256, (6, 360, 640, 3)).astype(np.uint8)
    pa, pb = rvb200.Params.make("YCrCb", 2.0, 8, 5), rvb200.Params.make("LAB", 3.0, 4, 3)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    oa, ob = torch.empty_like(da), torch.empty_like(db)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(4):                                           # several rounds back to back, no host sync in between
        ctx.submit_device(da.data_ptr(), oa.data_ptr(), 6, 360, 640, pa, stream=s1.cuda_stream)
        ctx.submit_device(db.data_ptr(), ob.data_ptr(), 6, 360, 640, pb, stream=s2.cuda_stream)
    span = ctx.gray_span(a[:2])                                  # a stage call on the context's own stream in between
    torch.cuda.synchronize()
    ra, rb = oa.cpu().numpy(), ob.cpu().numpy()
    for i in range(6):
        assert np.array_equal(ra[i], O.chain(a[i], O.SPACE_YCRCB, 2.0, 8, 5)), i
        assert np.array_equal(rb[i], O.chain(b[i], O.SPACE_LAB, 3.0, 4, 3)), i
    assert [int(v) for v in span] == [O.gray_span(a[0]), O.gray_span(a[1])]


def test_large_tile_grids(ctx):
    """tile_grid 32 (per-quad LUT kernel, grids > 16) and the documented maximum 256 (65,536 tiles on gridDim.x)."""
    import rvb200
    rng = np.random.RandomState(12)
    img = rng.randint(0, 256, (200, 333, 3)).astype(np.uint8)
    for grid in (17, 32):
        pl = O.luma(img, O.SPACE_YCRCB)
        hist = O.clahe_hist(pl, grid)
        assert np.array_equal(ctx.luma_hist(img[None], "YCrCb", grid, want_luma=False)[0][0], hist), grid
        for clip in (0.0, 2.0, 40.0):
            assert np.array_equal(ctx.build_lut(hist[None], 200, 333, grid, clip)[0], O.clahe_lut(hist, 200, 333, grid, clip)), (grid, clip)
        for space, k in (("YCrCb", 5), ("LAB", 3)):
            got = ctx.chain(img[None], rvb200.Params.make(space, 2.0, grid, k))[0]
            assert np.array_equal(got, O.chain(img, ospace(space), 2.0, grid, k)), (grid, space)
    small = rng.randint(0, 256, (70, 90, 3)).astype(np.uint8)
    got = ctx.chain(small[None], rvb200.Params.make("YCrCb", 2.0, 256, 3))[0]
    assert np.array_equal(got, O.chain(small, O.SPACE_YCRCB, 2.0, 256, 3))


def test_many_tiny_frames_one_call(ctx):
    """A batch larger than one launch's gridDim.z budget goes out in several launches (host and device paths)."""
    torch = pytest.importorskip("torch")
    import rvb200
    n = 33000
    frames = np.random.RandomState(8).randint(0, 256, (n, 4, 5, 3)).astype(np.uint8)
    p = rvb200.Params.make("YCrCb", 2.0, 2, 3)
    got = ctx.chain(frames, p)
    d = torch.from_numpy(frames).cuda()
    o = torch.empty_like(d)
    ctx.chain_device(d.data_ptr(), o.data_ptr(), n, 4, 5, p)
    assert np.array_equal(o.cpu().numpy(), got)
    ctx.set_option("group_frames", 40000)                        # one group: every kernel of the chain splits its launch
    try:
        o.zero_()
        ctx.chain_device(d.data_ptr(), o.data_ptr(), n, 4, 5, p)
    finally:
        ctx.set_option("group_frames", 0)
    assert np.array_equal(o.cpu().numpy(), got)
    for i in (0, 1, 32767, 32768, n - 1):
        assert np.array_equal(got[i], O.chain(frames[i], O.SPACE_YCRCB, 2.0, 2, 3)), i
    assert np.array_equal(ctx.gray_span(frames)[[0, 32768, n - 1]],
                          [O.gray_span(frames[0]), O.gray_span(frames[32768]), O.gray_span(frames[n - 1])])


def test_gate_threshold_boundary_is_decided_in_integers(ctx):
    """thresholds that binary32 cannot represent must give the same decision in the fused batch path, the per-frame
    path and the reference (int span against a Python float)."""
    import rvb200
    f = np.full((48, 64, 3), 100, np.uint8)
    f[0, 0] = (120, 120, 120)                                   # gray span exactly 20
    assert O.gray_span(f) == 20
    for thresh, processed in ((20.0000001, True), (20.0, False), (19.9999999, False), (21, True), (0.0, False), (-3.0, False)):
        cfg = {"chain": [{"name": "CLAHEDehaze"}, {"name": "MedianDerain"}],
               "auto_gate": {"enable_low_contrast_gate": True, "contrast_thresh": thresh}}
        pl = rvb200.PreprocessPipeline(cfg)
        assert (20 < thresh) == processed
        one = pl(f)
        assert (one is not f) == processed, thresh
        got = pl.process_batch(np.stack([f, f]))
        want = O.chain(f, O.SPACE_YCRCB, 2.0, 8, 3) if processed else f
        assert np.array_equal(got[0], want) and np.array_equal(got[1], want), thresh
        if processed:
            assert np.array_equal(one, want)


def test_overlapping_buffers_are_rejected(ctx):
    torch = pytest.importorskip("torch")
    import rvb200
    d = torch.zeros(3 * 32 * 32 * 3 + 64, dtype=torch.uint8, device="cuda")
    p = rvb200.Params.make()
    with pytest.raises(ValueError):
        ctx.chain_device(d.data_ptr(), d.data_ptr() + 96, 2, 32, 32, p)          # partial overlap, one row down
    pl = rvb200.PreprocessPipeline({"chain": [{"name": "MedianDerain"}]})
    x = torch.zeros((2, 32, 32, 3), dtype=torch.uint8, device="cuda")
    for bad in (torch.zeros((2, 32, 32, 3), dtype=torch.float16, device="cuda"), torch.zeros((2, 32, 31, 3), dtype=torch.uint8, device="cuda"),
                torch.zeros((2, 32, 64, 3), dtype=torch.uint8, device="cuda")[:, :, ::2]):
        with pytest.raises(ValueError):
            pl.process_batch(x, out=bad)


def test_results_that_stay_on_the_gpu(ctx):
    """process_batch(out="device") and process_batch_to_tensor(out="device" | pinned): same bytes as the host path; the
    device results are usable from torch through __cuda_array_interface__."""
    torch = pytest.importorskip("torch")
    import rvb200
    from rvb200 import synth
    frames = synth.frame_pool(1080, 1920, 7, base_seed=300)
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": "YCrCb"}}, {"name": "MedianDerain", "params": {"ksize": 5}}]}
    pl = rvb200.PreprocessPipeline(cfg)
    want = pl.process_batch(frames)
    for i in (0, 6):
        assert np.array_equal(want[i], O.chain(frames[i], O.SPACE_YCRCB, 2.0, 8, 5))
    pin = ctx.pinned_empty(frames.shape)
    pin[:] = frames
    ctx.set_option("chunk_frames", 2)                            # several pipeline chunks, last one partial
    try:
        dev = pl.process_batch(pin, out="device")
        assert isinstance(dev, rvb200.DeviceArray) and dev.shape == frames.shape
        t = torch.as_tensor(dev, device="cuda")
        assert t.data_ptr() == dev.ptr and np.array_equal(t.cpu().numpy(), want)
        # detector tensor: pageable, pinned and device outputs, fused (1080p -> 640) path
        t0, none = pl.process_batch_to_tensor(frames)
        assert none is None
        for i in (0, 6):
            assert np.array_equal(t0[i].view(np.uint16), O.letterbox_f16(want[i], 640).view(np.uint16)), i
        pt = ctx.pinned_empty((7, 3, 640, 640), np.float16)
        t1, full = pl.process_batch_to_tensor(pin, out=pt, want_frames=True)
        assert t1 is pt and np.array_equal(pt.view(np.uint16), t0.view(np.uint16)) and np.array_equal(full, want)
        t2, _ = pl.process_batch_to_tensor(pin, out="device")
        assert np.array_equal(t2.numpy().view(np.uint16), t0.view(np.uint16))
        # unfused geometry (ragged size) through the same pipeline
        rag = np.ascontiguousarray(frames[:3, :1079, :1917])
        r0, _ = pl.process_batch_to_tensor(rag)
        r2, rf = pl.process_batch_to_tensor(rag, out="device", want_frames=True)
        assert np.array_equal(r2.numpy().view(np.uint16), r0.view(np.uint16))
        for i in range(3):
            proc = O.chain(rag[i], O.SPACE_YCRCB, 2.0, 8, 5)
            assert np.array_equal(rf[i], proc)
            assert np.array_equal(r0[i].view(np.uint16), O.letterbox_f16(proc, 640).view(np.uint16)), i
    finally:
        ctx.set_option("chunk_frames", 0)


def test_batch_feeder_pinned_ring_to_process_batch(ctx):
    """capture -> BatchFeeder(pinned ring) -> process_batch equals the per-frame oracle; per-frame capture
    timestamps (capture.py:20) are monotonic; the last batch is partial."""
    import rvb200
    from rvb200 import synth
    from rvb200.io_video.capture import SyntheticReader
    h, w = 270, 480
    pool = synth.frame_pool(h, w, 5, base_seed=400)
    nframes, batch = 23, 8
    pctx = rvb200.Context(0)                                     # the stream's own context, bound to its pipeline
    pl = rvb200.PreprocessPipeline({"chain": [{"name": "CLAHEDehaze", "params": {"space": "LAB"}}, {"name": "MedianDerain"}]},
                                   context=pctx)
    assert pl._ctx() is pctx
    vs = rvb200.VideoSource(reader=SyntheticReader(list(pool), limit=nframes))
    feeder = rvb200.BatchFeeder(vs, batch=batch, shape=(h, w, 3), alloc=pctx.pinned_empty, depth=3)
    out = pctx.pinned_empty((batch, h, w, 3))
    want = [O.chain(f, O.SPACE_LAB, 2.0, 8, 3) for f in pool]
    seen, counts, ts_all = 0, [], []
    for b in feeder:
        assert pctx.mem_kind(b.frames) == rvb200._native.MEM_PINNED
        res = pl.process_batch(b.frames, out=out[:b.count])
        for i in range(b.count):
            assert np.array_equal(res[i], want[(seen + i) % len(pool)]), (seen, i)
        ts_all.extend(b.ts)
        counts.append(b.count)
        seen += b.count
        feeder.release(b)
    assert seen == nframes and counts == [8, 8, 7]
    assert all(t > 0 for t in ts_all) and np.all(np.diff(ts_all) >= 0)
    pctx.close()


def test_single_frame_path_staging_and_graphs(ctx):
    """The per-frame contract's fast path: pageable frames (staged by the driver, or by the optional helper threads), page-locked frames uploaded
    directly, CUDA-graph replay vs direct launches, shapes and parameters alternating between calls (graph cache, staging-frame
    growth, workspace reallocation) -- every variant gives the oracle's bytes, call after call."""
    import rvb200
    from rvb200 import synth
    shapes = [(720, 1280), (1080, 1920), (480, 640), (1080, 1920), (2160, 3840), (720, 1280)]
    frames = {s: synth.road_frame(s[0], s[1], 600 + i) for i, s in enumerate(dict.fromkeys(shapes))}
    params = [("YCrCb", 8, 5), ("LAB", 8, 3), ("YCrCb", 16, 3)]
    want = {}
    for s, f in frames.items():
        for (space, grid, k) in params:
            if s == (2160, 3840) and space == "LAB":
                continue
            want[(s, space, grid, k)] = O.chain(f, ospace(space), 2.0, grid, k)
    pinned = {}
    for s, f in frames.items():
        pinned[s] = ctx.pinned_empty(f.shape)
        pinned[s][:] = f
    try:
        for rnd in range(3):
            ctx.set_option("frame_graphs", 0 if rnd == 1 else 1)
            ctx.set_option("stage_threads", 3 if rnd == 2 else 0)
            for s in shapes:
                for (space, grid, k) in params:
                    if (s, space, grid, k) not in want:
                        continue
                    p = rvb200.Params.make(space, 2.0, grid, k)
                    for src in (frames[s], pinned[s], frames[s]):
                        got = ctx.chain(src[None], p)[0]
                        assert np.array_equal(got, want[(s, space, grid, k)]), (rnd, s, space, grid, k)
        # back-to-back pageable frames of one shape with different contents: the staging frame is reused every call
        ctx.set_option("frame_graphs", 1)
        ctx.set_option("stage_threads", 3)                      # (the optional helper-thread staging; default is off)
        p = rvb200.Params.make("YCrCb", 2.0, 8, 5)
        a, b = frames[(1080, 1920)], synth.road_frame(1080, 1920, 777)
        wa, wb = want[((1080, 1920), "YCrCb", 8, 5)], O.chain(b, O.SPACE_YCRCB, 2.0, 8, 5)
        for i in range(40):
            src, w_ = (a, wa) if i % 2 == 0 else (b, wb)
            assert np.array_equal(ctx.chain(src[None], p)[0], w_), i
    finally:
        ctx.set_option("frame_graphs", 1)
        ctx.set_option("stage_threads", 0)


def test_contexts_are_independent_across_threads():
    """A context is for one thread at a time; DIFFERENT contexts may run concurrently from different threads (one per camera stream).
    Three threads, each with its own Context and pipeline, mix per-frame calls (graph capture and replay), batch calls and the gate."""
    import threading
    import rvb200
    from rvb200 import synth
    jobs = [("YCrCb", 8, 5, (360, 640)), ("LAB", 8, 3, (270, 480)), ("YCrCb", 4, 3, (480, 640))]
    data = []
    for i, (space, grid, k, (h, w)) in enumerate(jobs):
        fr = synth.frame_pool(h, w, 4, base_seed=800 + 10 * i)
        data.append((fr, [O.chain(f, ospace(space), 2.0, grid, k) for f in fr]))
    errors = []

    def work(i):
        try:
            space, grid, k, _ = jobs[i]
            fr, want = data[i]
            c = rvb200.Context(0)
            pl = rvb200.PreprocessPipeline({"chain": [{"name": "CLAHEDehaze", "params": {"space": space, "tile_grid": grid}},
                                                      {"name": "MedianDerain", "params": {"ksize": k}}]}, context=c)
            for rep in range(25):
                j = rep % len(fr)
                if not np.array_equal(pl(fr[j]), want[j]):
                    raise AssertionError(f"thread {i} frame call {rep}")
                if rep % 5 == 0 and not np.array_equal(pl.process_batch(fr), np.stack(want)):
                    raise AssertionError(f"thread {i} batch call {rep}")
            c.close()
        except Exception as e:          # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_context_lifecycle_returns_its_memory():
    """Contexts come and go with camera streams: creating one, running frames of several geometries through every kind of entry
    (per-frame with graphs, host batch, device-resident results, detector tensor) and closing it gives all device memory back."""
    import torch
    import rvb200
    from rvb200 import synth
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": "YCrCb"}}, {"name": "MedianDerain", "params": {"ksize": 5}}]}
    shapes = [(270, 480), (360, 640), (135, 300)]
    pools = [synth.frame_pool(h, w, 5, base_seed=900 + i) for i, (h, w) in enumerate(shapes)]

    def cycle():
        c = rvb200.Context(0)
        pl = rvb200.PreprocessPipeline(cfg, context=c)
        for fr in pools:
            pl(fr[0]); pl(fr[1])
            pl.process_batch(fr)
            pin = c.pinned_empty(fr.shape); pin[:] = fr
            dev, _ = pl.process_batch_to_tensor(pin, size=96, out="device")
            del dev, pin
        c.close()

    cycle()                                        # first cycle: one-time allocations of the CUDA runtime and the module's tables
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(8):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < (8 << 20), f"{(free0 - free1) >> 20} MiB of device memory not returned after 8 context life cycles"


def test_tensor_rows_only_download(ctx):
    """padding_present: a host tensor buffer that already holds the letterbox padding rows gets only its image rows back over PCIe;
    the complete tensor equals the ordinary one (fused 1080p, unfused ragged, and a portrait frame whose padding is left / right)."""
    import rvb200
    from rvb200 import synth
    cfg = {"chain": [{"name": "CLAHEDehaze", "params": {"space": "YCrCb"}}, {"name": "MedianDerain", "params": {"ksize": 3}}]}
    pl = rvb200.PreprocessPipeline(cfg, context=ctx)
    for (h, w) in [(1080, 1920), (539, 961), (640, 360)]:
        frames = np.stack([synth.road_frame(h, w, 850 + i) for i in range(5)])
        want, _ = pl.process_batch_to_tensor(frames)
        buf = ctx.pinned_empty(want.shape, np.float16)
        buf[:] = 7.0                                                 # poison: rows the copy skips keep whatever the buffer held
        ctx.fill_tensor_padding(buf, h, w, 640)
        ctx.set_option("chunk_frames", 2)
        try:
            got, _ = pl.process_batch_to_tensor(frames, out=buf, padding_present=True)
        finally:
            ctx.set_option("chunk_frames", 0)
        assert got is buf and np.array_equal(buf.view(np.uint16), want.view(np.uint16)), (h, w)
        nw, nh, top, left, _ = ctx.letterbox_geometry(h, w, 640)
        if nh < 640:                                                 # without the pre-fill the padding rows are simply not written
            buf[:] = 7.0
            pl.process_batch_to_tensor(frames, out=buf, padding_present=True)
            assert np.all(buf[:, :, :top, :] == 7.0) and np.all(buf[:, :, top + nh:, :] == 7.0)
            assert np.array_equal(buf[:, :, top:top + nh, :].view(np.uint16), want[:, :, top:top + nh, :].view(np.uint16))
