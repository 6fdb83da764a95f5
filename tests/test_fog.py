"""Fog synthesis on the GPU against the reference's EnhancedFogSynthesizer (/root/reference/src/augment/fog.py:84-299).

Parity is statistical by nature (float pipeline; library exp / pow; OpenCV's SIMD summation order; the reference's own output
moves by an LSB between OpenCV builds), so the bar is written here as tolerances:

  * per-pixel: mean absolute difference <= 0.002 LSB, at most 0.01 % of the values off by more than 2 LSB, none by more than 6
    (measured on B200: 5e-6 .. 4e-5 LSB mean, at most 3 LSB on a handful of pixels);
  * per-channel mean and standard deviation within 0.02 LSB of the reference frame's (measured: <= 4e-5);
  * transmission map RMS error <= 4e-4 against the fixtures (which store it as float16; measured 1.4e-4, all of it storage) and
    <= 1e-6 against the live reference (measured 6e-8);
  * with the sensor-noise field generated on the device instead of drawn from the reference's stream: statistics only
    (mean / std within 0.02 LSB; measured 0.005).

Fixtures (tests/golden/fog_small.npz) were produced by the reference itself (tests/golden/make_fog_golden.py); on a box where the
reference tree is available (here: /root/reference, GPU box: oracle/_ref) a second test runs the reference live with fresh seeds.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "fog_small.npz")
KW = dict(y_h_ratio=0.42, perlin_scale_ratio=0.18, perlin_octaves=2, horizon_softness=0.07, global_veil=0.5, depth_blur_max=4.0)


def reference_fog_class():
    """The reference's class from /root/reference or oracle/_ref, or None."""
    for cand in ("/root/reference", os.path.join(ROOT, "oracle", "_ref", "road-vision-system")):
        if os.path.isfile(os.path.join(cand, "src", "augment", "fog.py")):
            saved = {k: sys.modules.pop(k) for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]}
            sys.path.insert(0, cand)
            try:
                sys.dont_write_bytecode = True
                from src.augment.fog import EnhancedFogSynthesizer as Ref
                return Ref
            finally:
                sys.path.remove(cand)
                for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
                    del sys.modules[k]
                sys.modules.update(saved)
    return None


def cases():
    z = np.load(GOLDEN)
    import cv2
    for line in z["meta"]:
        i, h, w, level, seed, scene, g, nz, mb, ma, mt = str(line).split("|")
        yield dict(i=int(i), h=int(h), w=int(w), level=level, seed=int(seed), scene=int(scene), gamma=bool(int(g)), noise=bool(int(nz)),
                   mean_beta=float(mb), mean_A=float(ma), mean_t=float(mt),
                   hazy=cv2.imdecode(z[f"hazy_{i}"], cv2.IMREAD_COLOR), t=z[f"t_{i}"].astype(np.float32))


class Capture:
    """Stands in for the CUDA library on a box without a GPU: records what the host side hands to rv_fog_*."""

    def __init__(self):
        self.geo, self.frame, self.lattice = None, None, None

    def rv_fog_set_geometry(self, h_, h, w, depth, sky, vgrad, xgrad):
        self.geo = (h, w, np.ctypeslib.as_array(depth, (h, w)).copy(), np.ctypeslib.as_array(sky, (h, w)).copy(),
                    np.ctypeslib.as_array(vgrad, (h,)).copy(), np.ctypeslib.as_array(xgrad, (w,)).copy())
        return 0

    def rv_fog_u8(self, h_, inp, out, h, w, f, lattice, noise, t, beta, amap):
        from rvb200.augment.fog import FogFrame
        fr = FogFrame.from_buffer_copy(f._obj)
        n = sum((fr.lat_gh[o] + 1) * (fr.lat_gw[o] + 1) for o in range(fr.octaves))
        self.frame, self.lattice = fr, np.ctypeslib.as_array(lattice, (n,)).copy()
        self.noise = None if not noise else np.ctypeslib.as_array(noise, (h, w, 3)).copy()
        return 0

    def rv_last_error(self, h_):
        return b""


def value_noise(fr, lattice, h, w):
    """numpy statement of k_fog_noise / k_fog_trans's beta map (csrc/rv_fog.cu), for checking the host side's lattices."""
    base, off, amp, norm = np.zeros((h, w), np.float32), 0, 1.0, 0.0
    for o in range(fr.octaves):
        gh, gw = fr.lat_gh[o], fr.lat_gw[o]
        g = lattice[off:off + (gh + 1) * (gw + 1)].reshape(gh + 1, gw + 1).astype(np.float64)
        off += (gh + 1) * (gw + 1)
        ys, xs = np.arange(h) * (gh / h), np.arange(w) * (gw / w)
        y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
        y1, x1 = np.minimum(y0 + 1, gh), np.minimum(x0 + 1, gw)
        wy, wx = (ys - y0)[:, None], (xs - x0)[None, :]
        top = g[y0][:, x0] * (1 - wx) + g[y0][:, x1] * wx
        bot = g[y1][:, x0] * (1 - wx) + g[y1][:, x1] * wx
        base = (base.astype(np.float64) + amp * (top * (1 - wy) + bot * wy)).astype(np.float32)
        norm += amp
        amp *= fr.persistence
    base = base / np.float32(max(1e-6, norm))
    nz = (base - base.min()) / max(np.float32(1e-6), base.max() - base.min())
    return np.float32(fr.base_beta) * (np.float32(0.85) + np.float32(0.35) * nz)


def test_host_side_matches_the_reference_draw_for_draw():
    """Depth prior, horizon, noise lattices / beta map and the branch decisions of every fixture case equal the reference's."""
    Ref = reference_fog_class()
    if Ref is None:
        pytest.skip("reference tree not available")
    import rvb200
    from rvb200 import synth
    from rvb200.augment import EnhancedFogSynthesizer
    from rvb200 import _native
    for c in cases():
        clean = synth.clean_scene(c["h"], c["w"], c["scene"])
        _, meta = Ref(level=c["level"], seed=c["seed"], **KW).synthesize(clean)
        cap = Capture()
        fake = _native.Context.__new__(_native.Context)
        fake._lib, fake._h = cap, C.c_void_p(1)
        mine = EnhancedFogSynthesizer(level=c["level"], seed=c["seed"], context=fake, exact_noise=True, **KW)
        mine.synthesize(clean)
        h, w, depth, sky, vgrad, xgrad = cap.geo
        assert (h, w) == (c["h"], c["w"]) and np.array_equal(depth, meta["depth"]), c["i"]
        assert mine._geo["horizon"] == meta["y_h"]
        fr = cap.frame
        beta = value_noise(fr, cap.lattice, h, w)
        assert np.abs(beta - meta["beta_map"]).max() <= 2e-7, (c["i"], np.abs(beta - meta["beta_map"]).max())
        assert (fr.gamma > 0) == c["gamma"] and (fr.noise_sigma > 0) == c["noise"], c["i"]
        assert (cap.noise is not None) == c["noise"]
        assert fr.glow_k % 2 == 1 and fr.glow_k2 % 2 == 1 and fr.fade_d % 2 == 1 and fr.octaves == 2
        assert all(0.7 - 1e-6 <= fr.A_bgr[k] <= 1.0 for k in range(3)) and 0.8 <= fr.a_target <= 1.0
        assert abs(meta["A_map"].mean() - fr.a_target) < 0.02           # the airlight map is scaled to this mean, then clipped


def test_host_shortcuts_equal_the_plain_formulas():
    """The per-frame host work is done with less arithmetic than the reference spends (one partition instead of a full quantile,
    the band radii from per-geometry means) -- the results must be the plain formulas' bit for bit: fog.py:120-131 and :203-213."""
    from rvb200 import synth
    from rvb200.augment import EnhancedFogSynthesizer
    from rvb200.augment import fog as F

    def plain_airlight(rng, bgr):
        top = bgr[:max(10, int(0.12 * bgr.shape[0]))].astype(np.float32) / 255.0
        lum = 0.299 * top[:, :, 2] + 0.587 * top[:, :, 1] + 0.114 * top[:, :, 0]
        mask = lum >= np.quantile(lum, 0.9)
        a = (top.mean(axis=(0, 1)) if mask.sum() < 100 else top[mask].mean(axis=0)).astype(np.float32)
        return np.clip(a + rng.uniform(-0.02, 0.02, size=3).astype(np.float32), 0.7, 1.0)

    def plain_radii(dmax, geo, beta):
        out = []
        for count, vals in geo["bands"]:
            rad = 0
            if count >= 100:
                r = np.clip(vals * dmax * (0.5 + beta), 0.0, dmax * 1.5)
                rad = int(max(1, np.mean(r) * 1.5)) | 1
            out.append(rad if rad > 1 else 0)
        return out

    rng = np.random.RandomState(4)
    for (h, w) in [(1080, 1920), (270, 480), (97, 131), (84, 11)]:
        frames = [synth.clean_scene(h, w, 3), rng.randint(0, 256, (h, w, 3)).astype(np.uint8), np.full((h, w, 3), 200, np.uint8),
                  rng.randint(100, 104, (h, w, 3)).astype(np.uint8), np.clip(rng.normal(180, 3, (h, w, 3)), 0, 255).astype(np.uint8)]
        for img in frames:
            mine = EnhancedFogSynthesizer(seed=3, context=object(), **KW)
            got, want = mine._airlight_colour(img), plain_airlight(np.random.RandomState(3), img)
            assert got.dtype == want.dtype and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (h, w)
        for dmax in (4.0, 3.5, 9.0):
            kw = dict(KW, depth_blur_max=dmax)
            mine = EnhancedFogSynthesizer(seed=3, context=object(), **kw)
            geo = mine._geometry(h, w)
            for t in range(40):
                beta = rng.uniform(0.02, 0.25) if t % 4 else 3.912 / rng.uniform(2, 300)      # presets and visibility (mor) values
                assert mine._band_radii(geo, beta) == plain_radii(dmax, geo, beta), (h, w, dmax, beta)
    assert F._FAST_QUANTILE is True          # this numpy: the one-partition quantile reproduces np.quantile (else it is not used)
    for t in range(400):
        x = [rng.rand(999), rng.randint(0, 5, 1877) / 7.0, 0.5 + np.arange(2500) * 6e-8, rng.normal(0.7, 1e-6, 4001)][t % 4].astype(np.float32)
        rng.shuffle(x)
        assert F._quantile09_fast(x) == np.quantile(x, 0.9)


def test_fog_has_no_cpu_fallback():
    import rvb200
    from rvb200 import _native
    from rvb200.augment import EnhancedFogSynthesizer
    if _native.load_library().rv_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rvb200.RvError):
        EnhancedFogSynthesizer(seed=1).synthesize(np.zeros((32, 48, 3), np.uint8))
    with pytest.raises(ValueError):
        EnhancedFogSynthesizer(seed=1).synthesize(np.zeros((32, 48), np.uint8))


def compare(got, want, label, pixelwise=True):
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    rep = dict(label=label, mad=float(d.mean()), gt2=float((d > 2).mean()), max=int(d.max()),
               dmean=[float(abs(got[..., k].mean() - want[..., k].mean())) for k in range(3)],
               dstd=[float(abs(got[..., k].std() - want[..., k].std())) for k in range(3)])
    print(rep)
    assert max(rep["dmean"]) <= 0.02 and max(rep["dstd"]) <= 0.02, rep
    if pixelwise:
        assert rep["mad"] <= 0.002 and rep["gt2"] <= 1e-4 and rep["max"] <= 6, rep
    return rep


@pytest.mark.gpu
def test_fog_fixtures_made_by_the_reference(ctx):
    from rvb200 import synth
    from rvb200.augment import EnhancedFogSynthesizer
    for c in cases():
        clean = synth.clean_scene(c["h"], c["w"], c["scene"])
        fog = EnhancedFogSynthesizer(level=c["level"], seed=c["seed"], context=ctx, exact_noise=True, **KW)
        hazy, meta = fog.synthesize(clean)
        assert hazy.shape == clean.shape and hazy.dtype == np.uint8
        rms = float(np.sqrt(np.mean((meta["t"] - c["t"]) ** 2)))
        print(c["i"], "t rms", rms, "beta mean", float(meta["beta_map"].mean()), c["mean_beta"], "A mean", float(meta["A_map"].mean()), c["mean_A"])
        assert rms <= 4e-4, (c["i"], rms)
        assert abs(float(meta["beta_map"].mean()) - c["mean_beta"]) <= 1e-6 and abs(float(meta["A_map"].mean()) - c["mean_A"]) <= 2e-4
        compare(hazy, c["hazy"], f"fixture {c['i']} {c['level']} gamma={c['gamma']} noise={c['noise']}")
        if c["noise"]:                                   # the same case with the noise field generated on the device: statistics only
            hz2, _ = EnhancedFogSynthesizer(level=c["level"], seed=c["seed"], context=ctx, **KW).synthesize(clean)
            compare(hz2, c["hazy"], f"fixture {c['i']} device noise", pixelwise=False)
            assert not np.array_equal(hz2, hazy)


@pytest.mark.gpu
def test_fog_live_against_the_reference_at_full_size(ctx):
    """1080p, fresh seeds, all three levels: the reference runs on the host (1-2 s per frame), this package on the GPU."""
    Ref = reference_fog_class()
    if Ref is None:
        pytest.skip("reference tree not available (oracle/_ref is installed by __graft_entry__.build())")
    import time
    from rvb200 import synth
    from rvb200.augment import EnhancedFogSynthesizer
    clean = synth.clean_scene(1080, 1920, 900)
    for level, seed in (("light", 41), ("medium", 42), ("heavy", 43)):
        t0 = time.perf_counter()
        want, wmeta = Ref(level=level, seed=seed, **KW).synthesize(clean)
        t1 = time.perf_counter()
        fog = EnhancedFogSynthesizer(level=level, seed=seed, context=ctx, exact_noise=True, **KW)
        fog.synthesize(clean)                                            # first call uploads the geometry
        fog = EnhancedFogSynthesizer(level=level, seed=seed, context=ctx, exact_noise=True, **KW)
        fog._geo = None
        t2 = time.perf_counter()
        got, meta = fog.synthesize(clean)
        t3 = time.perf_counter()
        rms = float(np.sqrt(np.mean((meta["t"] - wmeta["t"]) ** 2)))
        print(level, "reference %.2f s, GPU %.3f s (incl. host maps), t rms %.2e" % (t1 - t0, t3 - t2, rms))
        assert rms <= 1.0e-6
        compare(got, want, f"1080p {level}")
