"""Row +g of the coverage contract: the reference's OWN `main_preview.main()` (main_preview.py:36-142) and its own
`configs/default.yaml`, both unmodified, drive this package when it is mounted as `src/preprocess` + `src/io_video`
(INTEGRATION.md section 2).  Everything else under `src/` (config loader, detect, track, geometry, vis) is the reference's.

The reference's third-party dependencies that are absent from this image are stubbed, as SURVEY.md 8c prescribes:
`ultralytics.YOLO` (returns one fixed box per frame), `filterpy.kalman.KalmanFilter` (a plain linear Kalman filter),
`cv2.imshow / waitKey / destroyAllWindows` (headless wheel) and `cv2.VideoCapture` (a synthetic camera, since there is none).

The reference tree comes from /root/reference here and from oracle/_ref/ (installed by build(), git-ignored) on the GPU box.
CPU half: main() runs up to the first `pipeline(raw, ts=fr.ts)` and fails loudly there (no GPU, no CPU fallback).
GPU half: ten frames; every `proc` equals the oracle and is the array the reference then draws into.
"""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "road-vision-system_b200")


def reference_root():
    for cand in ("/root/reference", os.path.join(ROOT, "oracle", "_ref", "road-vision-system")):
        if os.path.isfile(os.path.join(cand, "main_preview.py")) and os.path.isfile(os.path.join(cand, "configs", "default.yaml")):
            return cand
    return None


def make_tree(tmp_path, ref):
    """<tmp>/main_preview.py, configs/, src/{config.py,detect,track,geometry,vis} -> reference (symlinks, unmodified);
    <tmp>/src/{preprocess,io_video,_native.py,csrc} -> this package; <tmp>/stubs/{ultralytics,filterpy} -> stand-ins."""
    src = tmp_path / "src"
    src.mkdir()
    (src / "__init__.py").write_text("")
    for name in ("preprocess", "io_video", "_native.py", "csrc", "synth.py"):
        os.symlink(os.path.join(PKG, name), src / name)
    for name in ("config.py", "detect", "track", "geometry", "vis"):
        os.symlink(os.path.join(ref, "src", name), src / name)
    os.symlink(os.path.join(ref, "main_preview.py"), tmp_path / "main_preview.py")
    os.symlink(os.path.join(ref, "configs"), tmp_path / "configs")
    stubs = tmp_path / "stubs"
    (stubs / "ultralytics").mkdir(parents=True)
    (stubs / "filterpy").mkdir()
    (stubs / "ultralytics" / "__init__.py").write_text(textwrap.dedent("""
        import numpy as np
        class _T:
            def __init__(self, a): self.a = np.asarray(a)
            def cpu(self): return self
            def numpy(self): return self.a
            @property
            def shape(self): return self.a.shape
        class _Boxes:
            def __init__(self, h, w):
                self.xyxy = _T([[0.25 * w, 0.25 * h, 0.6 * w, 0.7 * h]]); self.conf = _T([0.9]); self.cls = _T([2.0])
            @property
            def shape(self): return self.xyxy.shape
        class _Res:
            def __init__(self, h, w): self.boxes = _Boxes(h, w)
        class YOLO:
            seen = []
            def __init__(self, model): self.names = {2: "car"}
            def fuse(self): pass
            def predict(self, source=None, **kw):
                YOLO.seen.append(source)
                return [_Res(*source.shape[:2])]
    """))
    (stubs / "filterpy" / "__init__.py").write_text("")
    (stubs / "filterpy" / "kalman.py").write_text(textwrap.dedent("""
        import numpy as np
        class KalmanFilter:
            def __init__(self, dim_x, dim_z):
                self.x = np.zeros((dim_x, 1)); self.P = np.eye(dim_x); self.Q = np.eye(dim_x)
                self.F = np.eye(dim_x); self.H = np.zeros((dim_z, dim_x)); self.R = np.eye(dim_z)
            def predict(self):
                self.x = self.F @ self.x; self.P = self.F @ self.P @ self.F.T + self.Q
            def update(self, z):
                y = np.asarray(z, float).reshape(-1, 1) - self.H @ self.x
                S = self.H @ self.P @ self.H.T + self.R
                K = self.P @ self.H.T @ np.linalg.inv(S)
                self.x = self.x + K @ y; self.P = (np.eye(len(self.x)) - K @ self.H) @ self.P
    """))
    return tmp_path


DRIVER = """
import sys
sys.dont_write_bytecode = True
sys.path[:0] = [{tmp!r}, {stubs!r}, {root!r}]
import numpy as np
import cv2

NFRAMES = {nframes}
from src import synth                                   # this package's generator (mounted under src/)
POOL = [synth.road_frame(480, 640, 700 + i) for i in range(3)]   # configs/default.yaml:3-4 camera size

class FakeCapture:                                      # stands in for the camera behind cv2.VideoCapture(0)
    def __init__(self, source): self.i = 0; self.props = {{}}
    def set(self, k, v): self.props[k] = v; return True
    def read(self):
        if self.i >= NFRAMES: return False, None
        self.i += 1
        return True, POOL[(self.i - 1) % len(POOL)].copy()
    def release(self): pass

shown = []
cv2.VideoCapture = FakeCapture
cv2.imshow = lambda name, img: shown.append((name, img.shape))
cv2.waitKey = lambda ms: 0
cv2.destroyAllWindows = lambda: None

import main_preview                                     # the reference's file, unmodified
assert main_preview.PreprocessPipeline.__module__.startswith("src.preprocess"), main_preview.PreprocessPipeline.__module__
assert main_preview.VideoSource.__module__.startswith("src.io_video")

records = []
class Recording(main_preview.PreprocessPipeline):       # same class; remembers what it returned
    def __call__(self, image, ts=None):
        out = super().__call__(image, ts=ts)
        records.append((image.copy(), out, out.copy(), ts))
        return out
main_preview.PreprocessPipeline = Recording
canvases = []
_mk = main_preview.make_canvas
def make_canvas(raw, proc, **kw):
    canvases.append((raw, proc))
    return _mk(raw, proc, **kw)
main_preview.make_canvas = make_canvas
"""


def run(tmp, body, nframes=10):
    head = DRIVER.format(tmp=str(tmp), stubs=str(tmp / "stubs"), root=ROOT, nframes=nframes)
    return subprocess.run([sys.executable, "-c", head + textwrap.dedent(body)], cwd=str(tmp), capture_output=True, text=True, timeout=900)


def test_main_preview_reaches_the_pipeline_and_refuses_a_cpu_fallback(tmp_path):
    ref = reference_root()
    if ref is None:
        pytest.skip("reference tree not available (run __graft_entry__.build() where /root/reference is mounted)")
    import rvb200
    from rvb200 import _native
    if _native.load_library().rv_device_count() > 0:
        pytest.skip("a GPU is present: covered by the GPU half")
    tmp = make_tree(tmp_path, ref)
    r = run(tmp, """
        from src._native import RvError
        try:
            main_preview.main()
        except RvError as e:
            # load_config + default.yaml, VideoSource, pipeline construction from the YAML chain, detector, tracker all ran;
            # the first frame reached pipeline(raw, ts=fr.ts) and the GPU library refused to run without a B200
            assert "no CPU fallback" in str(e), e
            print("ok: refused at the first frame")
        else:
            raise SystemExit("main() finished without a GPU: something fell back to the CPU")
    """, nframes=2)
    assert r.returncode == 0 and "ok: refused" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


@pytest.mark.gpu
def test_main_preview_main_runs_unchanged_on_the_gpu(tmp_path):
    ref = reference_root()
    if ref is None:
        pytest.skip("oracle/_ref is missing: __graft_entry__.build() installs it where /root/reference is mounted (it travels with the snapshot)")
    tmp = make_tree(tmp_path, ref)
    r = run(tmp, """
        from oracle import rv_oracle as O
        main_preview.main()                                  # ten frames, then "read failed / end of video" -> break
        import ultralytics
        assert len(records) == NFRAMES == len(canvases) == len(shown), (len(records), len(canvases), len(shown))
        drawn = 0
        for i, (raw, out, out_then, ts) in enumerate(records):
            assert np.array_equal(raw, POOL[i % len(POOL)])                       # input untouched (RAW pane)
            want = O.chain(raw, O.SPACE_YCRCB, 2.0, 8, 3)                         # default.yaml:21-34 chain
            assert np.array_equal(out_then, want), i                              # proc == oracle
            assert out.flags.writeable and out.flags.c_contiguous and out is not raw
            assert isinstance(ts, float)
            assert ultralytics.YOLO.seen[i] is out                                # detector.infer(proc), main_preview.py:99
            assert canvases[i][1] is out                                          # the same array goes on to the canvas ...
            drawn += int(not np.array_equal(out, out_then))                       # ... after draw_detections wrote INTO it (:114)
        assert drawn >= NFRAMES - 3, drawn                                        # SORT shows a track after min_hits=3 frames
        assert all(s == ("Compare Preview", (480, 640 + 4 + 640, 3)) for s in shown), shown[:2]
        print("ok", drawn)
    """)
    assert r.returncode == 0 and "ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
