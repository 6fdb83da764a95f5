"""Seeded synthetic road frames: clean scene -> fog -> rain (input generator for tests and bench).

The reference synthesises fog offline with EnhancedFogSynthesizer (/root/reference/src/augment/fog.py:84-299,
driven by tools/fog_batch.py:19-27) and has no rain generator at all.  That code costs 1-2 s per frame and
does not travel to the GPU box, so this module is a fast numpy restatement of the same physical model
(I = J*t + A*(1-t), depth prior rising towards a horizon at 0.42*H, low-frequency density noise, global
veil) plus a new rain-streak overlay.  It only has to produce *inputs* with fog-like statistics (narrow
luminance band, gray mean ~180, std ~40-55); every parity check runs the oracle and the GPU path on the
SAME generated frames, and tests/golden/ additionally holds frames made by the reference's own synthesiser.
"""
import numpy as np


def _smooth_noise(rng, h, w, cells):
    """Bilinear up-sampling of a coarse random lattice: cheap Perlin-like field in [0, 1]."""
    gy, gx = cells
    lat = rng.rand(gy + 1, gx + 1).astype(np.float32)
    ys = np.linspace(0, gy, h, endpoint=False, dtype=np.float32)
    xs = np.linspace(0, gx, w, endpoint=False, dtype=np.float32)
    y0 = ys.astype(np.int32); x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
    a = lat[y0][:, x0]; b = lat[y0][:, x0 + 1]; c = lat[y0 + 1][:, x0]; d = lat[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def clean_scene(h, w, seed):
    rng = np.random.RandomState(seed)
    yh = int(0.42 * h)
    img = np.empty((h, w, 3), np.float32)
    yy = np.arange(h, dtype=np.float32)[:, None]
    sky_t = np.clip(yy / max(yh, 1), 0, 1)
    sky = np.stack([215 - 70 * sky_t, 180 - 50 * sky_t, 150 - 40 * sky_t], -1)          # BGR
    road_t = np.clip((yy - yh) / max(h - yh, 1), 0, 1)
    road = np.stack([70 + 40 * road_t, 72 + 42 * road_t, 75 + 45 * road_t], -1)
    img[:] = np.where(yy[..., None] < yh, sky, road)
    for _ in range(40):                                                                 # vehicles, signs, buildings
        rh, rw = rng.randint(max(h // 40, 2), max(h // 5, 3)), rng.randint(max(w // 60, 2), max(w // 6, 3))
        y = rng.randint(0, max(h - rh, 1)); x = rng.randint(0, max(w - rw, 1))
        img[y:y + rh, x:x + rw] = rng.randint(10, 245, 3)
    for lane in (0.35, 0.5, 0.65):                                                      # lane markings
        xs = (w * (0.5 + (lane - 0.5) * (np.arange(yh, h) - yh + 8) / max(h - yh, 1) * 2)).astype(np.int32)
        for i, y in enumerate(range(yh, h)):
            if (i // max(h // 36, 1)) % 2 == 0:
                x = int(np.clip(xs[i], 0, w - 3)); img[y, x:x + 3] = 230
    img += rng.normal(0, 2.0, (h, w, 1)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


FOG_LEVELS = {"light": (0.5, 0.80), "medium": (0.9, 0.82), "heavy": (1.5, 0.85)}       # (beta, airlight)


def add_fog(bgr, level="medium", seed=0):
    rng = np.random.RandomState(seed + 7919)
    h, w = bgr.shape[:2]
    beta, A = FOG_LEVELS[level]
    yh = 0.42 * h
    yy = np.arange(h, dtype=np.float32)[:, None]
    depth = np.where(yy < yh, 1.0, np.clip(1.0 - (yy - yh) / max(h - yh, 1.0), 0.05, 1.0) ** 1.5).astype(np.float32)
    depth = np.broadcast_to(depth, (h, w))
    density = 0.75 + 0.5 * _smooth_noise(rng, h, w, (max(int(1 / 0.18), 2), max(int(w / h / 0.18), 2)))
    t = np.exp(-beta * depth * density)[..., None]
    J = bgr.astype(np.float32)
    I = J * t + 255.0 * A * (1 - t)
    I = 0.85 * I + 0.15 * 255.0 * A                                 # global veil: pull contrast towards the airlight
    return np.clip(I, 0, 255).astype(np.uint8)


def add_rain(bgr, seed=0, coverage=0.01):
    """Bright 1-2 px wide, 10-30 px long streaks at -15..+15 degrees, alpha 0.5-0.9, plus 0.1 % salt noise."""
    rng = np.random.RandomState(seed + 104729)
    h, w = bgr.shape[:2]
    out = bgr.astype(np.float32)
    n = max(int(coverage * h * w / 20), 1)
    ys = rng.randint(0, h, n); xs = rng.randint(0, w, n)
    ln = rng.randint(10, 31, n); ang = np.deg2rad(rng.uniform(-15, 15, n))
    alpha = rng.uniform(0.5, 0.9, n); wid = rng.randint(1, 3, n)
    for i in range(n):
        tt = np.arange(ln[i])
        yy = np.clip(ys[i] + (tt * np.cos(ang[i])).astype(np.int32), 0, h - 1)
        xx = np.clip(xs[i] + (tt * np.sin(ang[i])).astype(np.int32), 0, w - 1)
        for d in range(wid[i]):
            xd = np.clip(xx + d, 0, w - 1)
            out[yy, xd] = out[yy, xd] * (1 - alpha[i]) + 250.0 * alpha[i]
    salt = rng.rand(h, w) < 0.001
    out[salt] = 255
    return np.clip(out, 0, 255).astype(np.uint8)


def road_frame(h, w, seed, level="medium", rain=True):
    """One fogged (and rained) frame; deterministic in (h, w, seed, level, rain)."""
    img = add_fog(clean_scene(h, w, seed), level, seed)
    return add_rain(img, seed) if rain else img


def frame_pool(h, w, count, base_seed=0):
    levels = ("light", "medium", "heavy")
    return np.stack([road_frame(h, w, base_seed + i, levels[i % 3], rain=True) for i in range(count)])
