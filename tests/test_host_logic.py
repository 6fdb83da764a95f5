"""CPU-only checks of the host side: C-ABI surface, plugin contract, batch feeder, sharding."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    import rvb200
    from rvb200 import _native
    so = rvb200.build_library()
    assert os.path.exists(so)
    header = open(os.path.join(ROOT, "include", "rv_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rv_[a-z0-9_]+)\s*\(", body))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rv_b200.h but not exported"
    assert declared == set(_native.EXPORTS), "ctypes table and header disagree"
    assert b"sm_100a" in _native.load_library().rv_version()


def test_no_cpu_fallback_without_gpu():
    import rvb200
    from rvb200 import _native
    if _native.load_library().rv_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rvb200.RvError):
        rvb200.Context(0)
    op = rvb200.CLAHEDehaze(space="LAB")
    with pytest.raises(rvb200.RvError):
        op(np.zeros((8, 8, 3), np.uint8))


def test_only_tests_smoke_and_bench_use_the_oracle():
    """oracle/ is test infrastructure: nothing under tools/ or the product package may import it."""
    for sub in ("tools", "road-vision-system_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".sh", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "from oracle" not in text and "import oracle" not in text and "rv_oracle" not in text, (sub, f)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "road-vision-system_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "rv_lab_tables.h":
                text = open(os.path.join(dirpath, f)).read()
                assert "rv_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
                # cv2 may only appear for I/O: the camera in capture.py, image files in augment/fog.py's folder tool
                if f == "fog.py":
                    import re as _re
                    assert set(_re.findall(r"cv2\.(\w+)", text)) <= {"imread", "imwrite"}, f
                else:
                    assert "import cv2" not in text or f == "capture.py", f


def test_pipeline_identity_and_construction():
    import rvb200
    img = np.zeros((4, 4, 3), np.uint8)
    assert rvb200.PreprocessPipeline({"enabled": False, "chain": [{"name": "MedianDerain"}]})(img) is img
    assert rvb200.PreprocessPipeline({"chain": []})(img, ts=1.0) is img
    assert rvb200.PreprocessPipeline({"chain": None})(img) is img
    with pytest.raises(KeyError):
        rvb200.PreprocessPipeline({"chain": [{"name": "Missing"}]})
    p = rvb200.PreprocessPipeline({"chain": [{"name": "CUDACLAHEDehaze", "params": {"space": "Lab", "tile_grid": 1, "junk": 3}},
                                             {"name": "CUDAMedianDerain", "params": {"ksize": 6}}]})
    segs = p._segments()
    assert len(segs) == 1 and segs[0].space == 1 and segs[0].grid == 2 and segs[0].ksize == 7 and segs[0].clahe == 1
    p.ops[1].params["ksize"] = 11               # params are re-read on every call, like the reference
    assert p._segments()[0].ksize == 9
    q = rvb200.PreprocessPipeline({"chain": [{"name": "MedianDerain"}, {"name": "CLAHEDehaze"}]})
    s = q._segments()
    assert [x.clahe for x in s] == [0, 1] and [x.ksize for x in s] == [3, 0] and s[1].space == 0 and s[1].grid == 8


def test_input_validation():
    from rvb200.preprocess.base import as_bgr_u8
    with pytest.raises(TypeError):
        as_bgr_u8([[1, 2, 3]])
    for bad in [np.zeros((4, 4), np.uint8), np.zeros((4, 4, 4), np.uint8), np.zeros((0, 4, 3), np.uint8)]:
        with pytest.raises(ValueError):
            as_bgr_u8(bad)
    with pytest.raises(ValueError):
        as_bgr_u8(np.zeros((4, 4, 3), np.float32))
    base = np.arange(6 * 8 * 3, dtype=np.uint8).reshape(6, 8, 3)
    for view in [base[::2], base[:, ::2], np.asfortranarray(base)]:
        c = as_bgr_u8(view)
        assert c.flags.c_contiguous and np.array_equal(c, view)


def test_video_source_read_and_read_batch():
    import rvb200
    from rvb200.io_video.capture import SyntheticReader
    pool = [np.full((6, 8, 3), i, np.uint8) for i in range(3)]
    vs = rvb200.VideoSource(reader=SyntheticReader(pool, limit=5))
    fr = vs.read()
    assert fr.ok and fr.image is pool[0] and isinstance(fr.ts, float)
    buf = np.empty((8, 6, 8, 3), np.uint8)
    count, frames, ts = vs.read_batch(8, out=buf)
    assert count == 4 and frames.shape == (4, 6, 8, 3) and len(ts) == 4 and np.all(np.diff(ts) >= 0)
    assert [int(f[0, 0, 0]) for f in frames] == [1, 2, 0, 1]
    assert vs.read().ok is False
    m = rvb200.FPSMeter(alpha=0.5)
    m.tick(1.0); m.tick(1.5)
    assert abs(m.fps - 1.0) < 1e-9


def test_host_pipeline_chunk_schedule():
    """rv_chunk_schedule (the arithmetic chain_pipe cuts host jobs with): chunks partition the job in order, none exceeds the chunk
    size, the uniform schedule is n // C full chunks and a remainder, the tapered one starts and ends small once the job is longer
    than three chunks.  Pure host code: runs without a GPU."""
    import rvb200
    for C in range(1, 14):
        for n in list(range(0, 90)) + [127, 128, 1000]:
            uni, tap = rvb200.chunk_schedule(n, C, taper=False), rvb200.chunk_schedule(n, C, taper=True)
            for sched in (uni, tap):
                assert sum(sched) == n and all(1 <= g <= C for g in sched), (n, C, sched)
            assert uni == [C] * (n // C) + ([n % C] if n % C else [])
            if C >= 2 and n > 3 * C:
                small = max(1, C // 3)
                assert tap[0] == small and tap[1] == max(small, 2 * C // 3) and tap[-1] <= max(small, 1), (n, C, tap)
                assert len(tap) <= len(uni) + 3
            else:
                assert tap == uni
    assert rvb200.chunk_schedule(64, 3) == [1, 2] + [3] * 20 + [1]          # the benchmark's job: 64 x 1080p frames
    lib = rvb200._native.load_library()
    assert lib.rv_chunk_schedule(5, 0, 1, None, 0) < 0 and lib.rv_chunk_schedule(-1, 3, 1, None, 0) < 0


def test_median_networks_are_current(tmp_path):
    """The committed rv_median_net.h is what tools/gen_median_net.py generates (networks verified there: random vectors for every
    network, 0-1 vectors at the median threshold for k = 7, 9; the exhaustive 0-1 check of k <= 5 runs without --quick)."""
    path = os.path.join(ROOT, "road-vision-system_b200", "csrc", "rv_median_net.h")
    fresh = str(tmp_path / "rv_median_net.h")
    env = {k: v for k, v in os.environ.items() if not k.startswith("RV_MEDIAN")}
    env["RV_MEDIAN_NET_OUT"] = fresh
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_median_net.py"), "--quick"],
                          stdout=subprocess.DEVNULL, timeout=900, env=env)
    assert open(fresh).read() == open(path).read()


def test_shard_plan_and_gloo_world2(tmp_path):
    """bench.py's sharding helper: frames partition across ranks with no overlap; run it under gloo, world_size 2."""
    sys.path.insert(0, ROOT)
    import bench
    for n, world in [(64, 1), (64, 2), (64, 8), (10, 4), (3, 8)]:
        parts = [bench.shard_range(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    script = tmp_path / "w2.py"
    script.write_text(
        "import sys; sys.path.insert(0, %r)\n"
        "import torch, torch.distributed as dist, bench\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "a, b = bench.shard_range(64, r, w)\n"
        "t = bench.max_over_ranks(float(10 + r), use_cuda=False)\n"
        "tot = bench.sum_over_ranks(float(b - a), use_cuda=False)\n"
        "assert t == 11.0 and tot == 64.0, (t, tot)\n"
        "dist.destroy_process_group()\n" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)], env=env, timeout=300)


def test_batch_feeder_order_timestamps_backpressure():
    import rvb200
    from rvb200.io_video.capture import SyntheticReader
    pool = [np.full((4, 6, 3), i, np.uint8) for i in range(5)]
    vs = rvb200.VideoSource(reader=SyntheticReader(pool, limit=23))
    feeder = rvb200.BatchFeeder(vs, batch=4, shape=(4, 6, 3), depth=2)
    seen, held = [], []
    for b in feeder:
        assert b.frames.shape == (b.count, 4, 6, 3) and len(b.ts) == b.count and np.all(np.diff(b.ts) >= 0)
        seen.extend(int(f[0, 0, 0]) for f in b.frames)
        held.append(b)
        if len(held) == 2:                 # the reader is blocked now (depth 2): release both and go on
            for h in held:
                feeder.release(h)
            held = []
    assert seen == [i % 5 for i in range(23)]          # nothing dropped, nothing reordered, partial last batch delivered


def test_ycc_chroma_table_equals_oracle_round_trip_exhaustive():
    """k_chain replaces the Cr/Cb arithmetic of BGR2YCrCb / YCrCb2BGR by two look-ups in the table rv_create uploads
    (include/rv_b200.h: rv_ycc_table).  Host-only check over ALL 2^24 colours, with the new luminance Y' swept against Y:
    B' = Y' + f(B - Y), G' = Y' + ((tB + tR + 8192) >> 14), R' = Y' + f(R - Y) must equal the oracle's
    ycrcb2bgr(Y', Cr, Cb) with (Y, Cr, Cb) = bgr2ycrcb(B, G, R)."""
    import ctypes as C
    from rvb200 import _native
    from oracle import rv_oracle as O
    lib = _native.load_library()
    tab = np.zeros(1024, np.uint32)
    assert lib.rv_ycc_table(tab.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([v & 255, (v >> 8) & 255, v >> 16], -1).astype(np.uint8).reshape(4096, 4096, 3)
    ycc = O.bgr2ycrcb(img)
    Y = ycc[..., 0].astype(np.int64)
    eB = tab[:512][img[..., 0].astype(np.int64) - Y + 255].astype(np.int64)
    eR = tab[512:][img[..., 2].astype(np.int64) - Y + 255].astype(np.int64)
    # layout of round 2 (csrc/rv_colour.cuh, RV_YCC16): low half-word f + 256, high half-word the 16-bit G term of rv_ycc_g.h
    fB, fR = (eB & 0xFFFF) - 256, (eR & 0xFFFF) - 256
    assert ((eB & 0xFFFF) + (eR & 0xFFFF)).max() < 65536 and ((eB >> 16) + (eR >> 16)).max() < 65536      # no carries between the fields
    g = ((eB + eR) & 0xFFFFFFFF) >> 23
    for shift in (0, 1, 37, 128, 255):           # Y' = (Y + shift) mod 256 exercises every (Y', chroma) pairing that matters
        y2 = (Y + shift) & 255
        want = O.ycrcb2bgr(np.stack([y2.astype(np.uint8), ycc[..., 1], ycc[..., 2]], -1))
        got = np.stack([np.clip(y2 + fB, 0, 255), np.clip(y2 + g - 256, 0, 255), np.clip(y2 + fR, 0, 255)], -1).astype(np.uint8)
        assert np.array_equal(got, want), shift


def test_ycc_g_tables_are_current_and_exact():
    """csrc/rv_ycc_g.h is what tools/gen_ycc_g_tables.py generates (the script re-solves and re-verifies all 65,536 chroma pairs)."""
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_ycc_g_tables.py"), "--check"], stdout=subprocess.DEVNULL, timeout=600)


def test_profile_constants_belong_to_the_shipped_kernels():
    """bench.py reports DRAM traffic and instruction counts from profiles/final_k_chain.json; the capture must have been taken on
    exactly the machine code of that kernel in the built library (hash of its SASS listing), else the numbers are stale."""
    import json
    import rvb200
    tree = rvb200.kernel_source_hash()
    rvb200.build_library()
    for name in ("final_k_chain.json", "final_k_luma_hist.json", "final_k_chain_lab_k3.json", "final_k_chain_lab_k5.json"):
        j = json.load(open(os.path.join(ROOT, "profiles", name)))
        # valid while the measured kernel's machine code in the built library is what the capture ran on (or, trivially, while the
        # kernel sources are untouched)
        built = rvb200.kernel_sass_hash(j["kernel"])
        assert j["sass_hash"] == built or j["source_hash"] == tree, \
            f"profiles/{name} was captured on SASS {j['sass_hash']} / sources {j['source_hash']}; the build has {built} / {tree}: re-capture (tools/gpu_r2e.sh)"
        assert j["dram_bytes_read"] > 0 and j["warp_instructions"] > 0
    j = json.load(open(os.path.join(ROOT, "profiles", "final_k_chain.json")))
    assert "k_chain<0, 5>" in j["kernel"] and j["grid"] == "(16, 23, 64)"


def test_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the GPU arm's metric / unit / config, `impl: reference`, a cpu_baseline
    describing the run and an e2e object without copies; it runs the reference's own PreprocessPipeline when oracle/_ref is installed."""
    import json
    pytest.importorskip("cv2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    sys.path.insert(0, ROOT)
    import bench
    want = bench.base_line(1, 1, 1)
    for k in ("metric", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert line[k] == want[k], k
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] > 0 and cb["cores"] >= 1 and cb["unit"] == line["unit"]
    assert cb["kind"] == ("reference" if os.path.isfile(os.path.join(bench.REF_ROOT, "src", "preprocess", "pipeline.py")) else "port")
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
