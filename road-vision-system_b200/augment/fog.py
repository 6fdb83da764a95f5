"""EnhancedFogSynthesizer on the GPU -- same class name, constructor and `synthesize(bgr) -> (hazy, meta)` contract as
/root/reference/src/augment/fog.py:84-299 (driven by tools/fog_batch.py:7-34), which costs 1-2 s per 1080p frame on the CPU.

Division of labour (details in csrc/rv_fog.cu):

* host, per synthesizer: the random stream.  Every draw is made here with numpy's RandomState, in the order the reference makes
  them (density, lattice seed, airlight tint, airlight target, glow, contrast drop, colour tint, gamma decision, noise decision),
  so a given seed selects the same fog parameters as in the reference;
* host, per geometry (cached): the depth prior and the sky weight (closed forms over the pixel grid) and the two airlight ramps;
* host, per frame: the airlight colour from the brightest tenth of the top band (a quantile over 12 % of the frame);
* device, per frame: everything else -- value noise, transmission, both edge-preserving filters, composition, veil, glow,
  depth blur, contrast fade, tint / gamma / sensor noise (rv_fog_u8).

There is no CPU fallback: without the CUDA library and a B200 `synthesize` raises.
"""
import ctypes as C
import os

import numpy as np

from .._native import default_context

FOG_PRESETS = {
    # level: fog density, airlight mean, glow strength, contrast drop -- the ranges of fog.py:72-76
    "light": {"beta": (0.03, 0.06), "airlight": (0.82, 0.93), "glow": (0.12, 0.22), "contrast_drop": (0.06, 0.12)},
    "medium": {"beta": (0.06, 0.12), "airlight": (0.86, 0.96), "glow": (0.18, 0.34), "contrast_drop": (0.10, 0.18)},
    "heavy": {"beta": (0.12, 0.22), "airlight": (0.90, 0.99), "glow": (0.28, 0.48), "contrast_drop": (0.15, 0.26)},
}
_MOR_RANGES = {"airlight": (0.86, 0.98), "glow": (0.12, 0.45), "contrast_drop": (0.08, 0.22)}      # fog.py:244-247


class FogFrame(C.Structure):
    """rv_fog_frame (include/rv_b200.h)"""
    _fields_ = [
        ("persistence", C.c_double), ("base_beta", C.c_float), ("A_bgr", C.c_float * 3), ("a_target", C.c_float),
        ("global_veil", C.c_float), ("glow", C.c_float), ("cdrop", C.c_float), ("tint", C.c_float * 3), ("gamma", C.c_float),
        ("noise_sigma", C.c_float), ("noise_seed", C.c_uint32), ("octaves", C.c_int32), ("lat_gh", C.c_int32 * 4),
        ("lat_gw", C.c_int32 * 4), ("band_rad", C.c_int32 * 3), ("glow_k", C.c_int32), ("glow_k2", C.c_int32), ("fade_d", C.c_int32),
        ("fade_sigma", C.c_float), ("edge_guided", C.c_int32),
    ]


_FAST_QUANTILE = None


def _quantile09_fast(flat):
    """np.quantile(flat, 0.9) for a 1-D float32 array from one np.partition: numpy's linear method interpolates between the order
    statistics k and k + 1 at the fractional part of the virtual index (n - 1) q, evaluated in the array's own precision; the
    interpolation itself is left to numpy (np.quantile of the two values at that fractional position)."""
    vi = (flat.size - 1) * np.asarray(0.9, dtype=flat.dtype)
    lo = np.floor(vi)
    k = int(lo)
    if k + 1 >= flat.size:
        return np.quantile(flat, 0.9)
    part = np.partition(flat, k)                                      # part[k] = order statistic k, everything right of it is >= it
    pair = np.array([part[k], part[k + 1:].min()], dtype=flat.dtype)
    return np.quantile(pair, float(vi - lo))


def _quantile09(flat):
    """The fast form where it reproduces this numpy's np.quantile bit for bit (checked once per process on arrays with distinct
    values, ties and neighbouring floats), np.quantile itself otherwise -- the threshold must be the reference's."""
    global _FAST_QUANTILE
    if _FAST_QUANTILE is None:
        rng, ok = np.random.RandomState(12345), flat.dtype == np.float32
        for n in (12, 101, 1000, 1877, 4096, 24769):
            for x in (rng.rand(n), rng.randint(0, 5, n) / 7.0, 0.5 + np.arange(n) * 6e-8, rng.normal(0.7, 1e-6, n)):
                x = x.astype(np.float32)
                rng.shuffle(x)
                a, b = _quantile09_fast(x), np.quantile(x, 0.9)
                ok = ok and a.dtype == b.dtype and a == b
        _FAST_QUANTILE = bool(ok)
    if _FAST_QUANTILE and flat.dtype == np.float32:
        return _quantile09_fast(flat)
    return np.quantile(flat, 0.9)


def _logistic(z):
    return 1.0 / (1.0 + np.exp(-z))


class EnhancedFogSynthesizer:
    def __init__(self, level="medium", mor=None, y_h_ratio=0.42, vanishing_x_ratio=0.5, perlin_scale_ratio=0.18, perlin_octaves=2,
                 sky_boost=1.25, road_damp=0.9, edge_guided=True, horizon_softness=0.06, depth_blur_max=3.5, global_veil=0.06,
                 seed=None, context=None, exact_noise=False):
        """Reference arguments (fog.py:92-105) plus two of this package's own: `context` (rvb200.Context; default: the process-wide
        one) and `exact_noise` (draw the sensor-noise field on the host from the same random stream as the reference, 6 M normals per
        1080p frame, instead of generating it on the device)."""
        self.level, self.mor = level, mor
        self.y_h_ratio, self.vx_ratio = y_h_ratio, vanishing_x_ratio
        self.perlin_scale_ratio, self.perlin_octaves = perlin_scale_ratio, perlin_octaves
        self.sky_boost, self.road_damp = sky_boost, road_damp
        self.edge_guided = edge_guided
        self.horizon_softness, self.depth_blur_max, self.global_veil = horizon_softness, depth_blur_max, global_veil
        self.rng = np.random.RandomState(seed) if seed is not None else np.random
        self.exact_noise = exact_noise
        self._context = context
        self._geo = None            # (h, w, key) of the maps currently uploaded to the context

    # ---- per-geometry maps (fog.py:144-170 and the ramps of :133-134) ------------------------------------------------------
    def _geometry(self, h, w):
        horizon = int(self.y_h_ratio * h)
        rows = np.arange(h, dtype=np.float32)[:, None] + np.zeros((1, w), np.float32)
        cols = np.arange(w, dtype=np.float32)[None, :] + np.zeros((h, 1), np.float32)
        below = np.maximum(rows - horizon, 1.0)                       # perspective term: 1 / distance below the horizon line
        persp = 1.0 / below
        vx, vy = float(self.vx_ratio * w), float(horizon)
        radial = 1.0 / (np.sqrt((cols - vx) ** 2 + (rows - vy) ** 2) + 1.0)     # farther towards the vanishing point
        depth = 0.7 * (persp / persp.max()) + 0.3 * (radial / radial.max())
        depth = (depth - depth.min()) / max(1e-6, (depth.max() - depth.min()))
        soft = max(1e-3, self.horizon_softness) * h
        sky = _logistic((horizon - rows) / soft).astype(np.float32)  # ~1 above the horizon, ~0 below, S-shaped across it
        depth = depth * ((1.0 + (self.sky_boost - 1.0) * sky) * (self.road_damp ** (1.0 - sky)))
        depth = np.clip(depth, 0, 1).astype(np.float32)
        vgrad = np.linspace(1.0, 0.85, h, dtype=np.float32)
        xgrad = np.linspace(0.95, 1.05, w, dtype=np.float32)
        # mean depth of the three blur bands (fog.py:206-212): the band radius is depth_blur_max (0.5 + beta) mean(depth in band)
        bands, lo = [], 0.0
        for hi in (0.33, 0.66, 1.0):
            sel = (depth >= np.float32(lo)) & (depth < np.float32(hi))
            bands.append((int(sel.sum()), depth[sel] if sel.any() else None))
            lo = hi
        return {"h": h, "w": w, "horizon": horizon, "depth": depth, "sky": sky, "vgrad": vgrad, "xgrad": xgrad, "bands": bands}

    def _ctx(self):
        return self._context if self._context is not None else default_context()

    def _ensure_geometry(self, h, w):
        key = (h, w, self.y_h_ratio, self.vx_ratio, self.sky_boost, self.road_damp, self.horizon_softness)
        ctx = self._ctx()
        if self._geo is None or self._geo["key"] != key:
            self._geo = self._geometry(h, w)
            self._geo["key"] = key
            self._geo["uploaded_to"] = None
        # one geometry lives on a context at a time; several synthesizers may share the context
        if self._geo["uploaded_to"] is not ctx or getattr(ctx, "_fog_geo_key", None) != (id(self), key):
            g = self._geo
            fp = C.POINTER(C.c_float)
            ctx._ck(ctx._lib.rv_fog_set_geometry(ctx._h, h, w, g["depth"].ctypes.data_as(fp), g["sky"].ctypes.data_as(fp),
                                                 g["vgrad"].ctypes.data_as(fp), g["xgrad"].ctypes.data_as(fp)))
            g["uploaded_to"] = ctx
            ctx._fog_geo_key = (id(self), key)
        return self._geo

    # ---- per-frame host work --------------------------------------------------------------------------------------------------
    def _uniform(self, lo, hi):
        return float(lo + (hi - lo) * self.rng.rand())

    def _lattices(self, h, w):
        """The uniform lattices of the value noise, drawn exactly like rand_perlin draws them (fog.py:13-23), and their sizes."""
        scale = max(16, int(self.perlin_scale_ratio * w))
        lattice_rng = np.random.RandomState(self.rng.randint(1e9))
        freq, shapes, chunks = 1.0 / max(1, scale), [], []
        for _ in range(max(1, self.perlin_octaves)):
            gh, gw = max(1, int(h * freq)), max(1, int(w * freq))
            chunks.append(lattice_rng.rand(gh + 1, gw + 1).astype(np.float32).ravel())
            shapes.append((gh, gw))
            freq *= 2.0
        return shapes, np.concatenate(chunks)

    def _airlight_colour(self, bgr):
        """Mean colour of the brightest tenth of the top 12 % of the frame, tinted and clipped (fog.py:120-131).  The same float32
        values as the reference's `lum >= np.quantile(lum, 0.9)` / `top[mask].mean(axis=0)`, obtained with less work: the two order
        statistics the 0.9-quantile interpolates between come from one `np.partition`, the interpolation itself is numpy's own
        (`np.quantile` of those two values at the same fractional position), and the selected pixels are gathered by index."""
        h = bgr.shape[0]
        top = bgr[:max(10, int(0.12 * h))].astype(np.float32) / 255.0
        lum = 0.299 * top[:, :, 2] + 0.587 * top[:, :, 1] + 0.114 * top[:, :, 0]
        flat = lum.ravel()
        thr = _quantile09(flat)
        sel = np.flatnonzero(flat >= thr)
        colour = (top.reshape(-1, 3)[sel].mean(axis=0) if sel.size >= 100 else top.mean(axis=(0, 1))).astype(np.float32)
        tint = self.rng.uniform(-0.02, 0.02, size=3).astype(np.float32)
        return np.clip(colour + tint, 0.7, 1.0)

    def _band_radii(self, geo, base_beta):
        """Gaussian sizes of the three depth-blur bands: int(max(1, 1.5 * mean(r in band))) | 1 with r = clip(depth * depth_blur_max *
        (0.5 + beta), 0, 1.5 depth_blur_max) (fog.py:203-213).  The mean over a band is a float32 reduction over up to 2 M pixels per
        frame in the reference; here depth * depth_blur_max and its float64 mean are kept per geometry, and the per-frame value
        s * mean decides the integer whenever it is not within 1e-4 of an integer boundary (and nothing is clipped) -- otherwise the
        reference's own reduction is evaluated, so the result is always the reference's integer."""
        dmax = float(self.depth_blur_max)
        cache = geo.setdefault("band_cache", {})
        if dmax not in cache:
            ent = []
            for count, vals in geo["bands"]:
                if count >= 100:
                    x = vals * dmax
                    ent.append((count, x, float(x.astype(np.float64).mean()), float(x.max()), float(x.min())))
                else:
                    ent.append((count, None, 0.0, 0.0, 0.0))
            cache[dmax] = ent
        s = 0.5 + base_beta
        s32 = float(np.float32(s))
        out = []
        for count, x, m64, xmax, xmin in cache[dmax]:
            rad = 0
            if count >= 100:
                est = s32 * m64 * 1.5
                safe = xmin >= 0.0 and s32 >= 0.0 and xmax * s32 * (1 + 1e-6) < dmax * 1.5 and abs(est - round(est)) > 1e-4 * max(1.0, est)
                if safe:
                    rad = int(max(1, est)) | 1
                else:
                    r = np.clip(x * s, 0.0, dmax * 1.5)
                    rad = int(max(1, np.mean(r) * 1.5)) | 1
            out.append(rad if rad > 1 else 0)
        return out

    def synthesize(self, bgr_uint8, level=None, meta=True):
        """BGR uint8 frame -> (hazy BGR uint8 frame, meta with beta_map / A_map / depth / y_h / t), as fog.py:227-299.
        `meta=False` (new) skips the download of the three float maps (41 MB at 1080p) and returns only depth and y_h."""
        if not isinstance(bgr_uint8, np.ndarray) or bgr_uint8.dtype != np.uint8 or bgr_uint8.ndim != 3 or bgr_uint8.shape[2] != 3:
            raise ValueError("synthesize expects a (H,W,3) uint8 BGR frame")
        frame = np.ascontiguousarray(bgr_uint8)
        h, w = frame.shape[:2]
        if level is not None:
            self.level = level
        ctx = self._ctx()
        octaves = max(1, int(self.perlin_octaves))
        if octaves > 4:
            raise ValueError("at most 4 noise octaves are supported")

        # the reference's draws, in its order
        if self.mor is not None and self.mor > 0:
            base_beta, ranges = 3.912 / float(self.mor), _MOR_RANGES             # Koschmieder, fog.py:243
        else:
            ranges = FOG_PRESETS[self.level]
            base_beta = self._uniform(*ranges["beta"])
        geo = self._ensure_geometry(h, w)
        shapes, lattice = self._lattices(h, w)
        colour = self._airlight_colour(frame)
        a_target = self._uniform(*ranges["airlight"])
        glow = self._uniform(*ranges["glow"])
        cdrop = self._uniform(*ranges["contrast_drop"])
        tint = (1.0 + self.rng.uniform(-0.015, 0.02, size=3)).astype(np.float32)
        gamma = (1.0 + self.rng.uniform(-0.04, 0.05)) if self.rng.rand() < 0.35 else 0.0
        noisy = self.rng.rand() < 0.3
        noise = None
        if noisy and self.exact_noise:
            noise = np.ascontiguousarray(self.rng.normal(0, 0.0035, size=frame.shape).astype(np.float32))

        f = FogFrame()
        f.persistence, f.base_beta = 0.5, base_beta
        f.A_bgr[:] = [float(v) for v in colour]
        f.a_target, f.global_veil, f.glow, f.cdrop = a_target, float(self.global_veil), glow, cdrop
        f.tint[:] = [float(v) for v in tint]
        f.gamma = float(gamma)
        f.noise_sigma = 0.0035 if noisy else 0.0
        f.noise_seed = int(self.rng.randint(1 << 31)) if (noisy and noise is None) else 0
        f.octaves = octaves
        for i, (gh, gw) in enumerate(shapes):
            f.lat_gh[i], f.lat_gw[i] = gh, gw
        # depth-blur band sizes: int(max(1, 1.5 * mean(r in band))) | 1 with r = depth * depth_blur_max * (0.5 + beta), fog.py:203-213
        for i, rad in enumerate(self._band_radii(geo, base_beta)):
            f.band_rad[i] = rad
        f.glow_k = int(9 + 20 * glow) | 1
        f.glow_k2 = int(max(7, (h + w) * (0.003 + 0.01 * glow))) | 1
        f.fade_d = int(5 + cdrop * 20) | 1
        f.fade_sigma = 25 + cdrop * 50
        f.edge_guided = 1 if self.edge_guided else 0

        out = np.empty_like(frame)
        fp = C.POINTER(C.c_float)
        info = {"depth": geo["depth"], "y_h": geo["horizon"]}
        if meta:
            info.update({"t": np.empty((h, w), np.float32), "beta_map": np.empty((h, w), np.float32), "A_map": np.empty((h, w, 3), np.float32)})
        ctx._ck(ctx._lib.rv_fog_u8(ctx._h, frame.ctypes.data, out.ctypes.data, h, w, C.byref(f), lattice.ctypes.data_as(fp),
                                   noise.ctypes.data_as(fp) if noise is not None else None,
                                   info["t"].ctypes.data_as(fp) if meta else None, info["beta_map"].ctypes.data_as(fp) if meta else None,
                                   info["A_map"].ctypes.data_as(fp) if meta else None))
        return out, info


def process_folder(inp, outp, levels=("light", "medium", "heavy"), limit=None, seed=None, context=None):
    """tools/fog_batch.py:7-34 on the GPU: every .jpg / .png / .jpeg below `inp` (recursively) at every level with fog_batch's
    parameters (:19-27), written to `outp/<level>/<relative path>`.  `seed` (new) makes the run reproducible.  Image decoding and
    encoding use cv2 (I/O only).  Returns the number of files written."""
    import cv2
    files = []
    for dirpath, _, names in os.walk(inp):
        files += [os.path.join(dirpath, n) for n in names if os.path.splitext(n)[1].lower() in (".jpg", ".png", ".jpeg")]
    files.sort()
    if limit:
        files = files[:limit]
    written = 0
    for i, path in enumerate(files, 1):
        img = cv2.imread(path)
        if img is None:
            print("Skip unreadable:", path)
            continue
        rel = os.path.relpath(path, inp)
        for j, level in enumerate(levels):
            synth = EnhancedFogSynthesizer(level=level, y_h_ratio=0.42, perlin_scale_ratio=0.18, perlin_octaves=2, horizon_softness=0.07,
                                           global_veil=0.5, depth_blur_max=4.0, context=context,
                                           seed=None if seed is None else seed + 3 * (i - 1) + j)
            hazy, _ = synth.synthesize(img, meta=False)
            dest = os.path.join(outp, level, rel)
            os.makedirs(os.path.dirname(dest), exist_ok=True)
            cv2.imwrite(dest, hazy)
            written += 1
        if i % 20 == 0:
            print(f"[{i}/{len(files)}] {path}")
    return written
