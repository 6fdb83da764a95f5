#!/usr/bin/env python3
"""Device-resident throughput of the chain for every BASELINE.json config shape (one GPU), next to the
reference's cv2 chain on the host (as shipped: cv2's own thread pool).  Prints one JSON line per config.
Not the headline benchmark (that is bench.py); used for the tables in DESIGN.md / profiles/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CONFIGS = [
    # name, (H, W), batch, space, grid, ksize
    ("C1 default.yaml 720p YCrCb k3", (720, 1280), 64, "YCrCb", 8, 3),
    ("C1 BASELINE wording 720p LAB k3", (720, 1280), 64, "LAB", 8, 3),
    ("C2 1080p YCrCb k5 (headline)", (1080, 1920), 64, "YCrCb", 8, 5),
    ("1080p LAB k3", (1080, 1920), 64, "LAB", 8, 3),
    ("1080p LAB k5", (1080, 1920), 64, "LAB", 8, 5),
    ("1080p YCrCb k3", (1080, 1920), 64, "YCrCb", 8, 3),
    ("C3 4K LAB grid16 k3", (2160, 3840), 16, "LAB", 16, 3),
    ("C3 4K YCrCb grid16 k5", (2160, 3840), 16, "YCrCb", 16, 5),
    ("1080p YCrCb k7", (1080, 1920), 16, "YCrCb", 8, 7),
    ("1080p YCrCb k9", (1080, 1920), 16, "YCrCb", 8, 9),
    ("1080p CLAHE only YCrCb", (1080, 1920), 64, "YCrCb", 8, 0),
]


def main():
    import torch
    import rvb200
    from rvb200 import synth
    from oracle import cv2_chain
    import cv2
    cpu = "--no-cpu" not in sys.argv
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    ctx = rvb200.Context(0)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    pools = {}
    for name, (h, w), batch, space, grid, k in CONFIGS:
        if only and only not in name:
            continue
        if (h, w) not in pools:
            base = synth.frame_pool(1080, 1920, 4, base_seed=3000)
            if (h, w) == (1080, 1920):
                pools[(h, w)] = base
            elif (h, w) == (720, 1280):
                pools[(h, w)] = np.stack([cv2.resize(f, (w, h), interpolation=cv2.INTER_AREA) for f in base])
            else:
                pools[(h, w)] = np.stack([np.tile(f, (2, 2, 1)) for f in base])
        pool = pools[(h, w)]
        host = np.stack([pool[i % len(pool)] for i in range(batch)])
        d_in = torch.from_numpy(host).cuda(); d_out = torch.empty_like(d_in)
        p = rvb200.Params.make(space, 2.0, grid, k)
        run = lambda: ctx.submit_device(d_in.data_ptr(), d_out.data_ptr(), batch, h, w, p, stream=st.cuda_stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        want = cv2_chain.chain(host[1], space, 2.0, grid, k)
        ok = bool(np.array_equal(d_out[1].cpu().numpy(), want))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 10
        e0.record()
        for _ in range(steps):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        fps = batch / (ms * 1e-3)
        rec = {"config": name, "shape": [h, w], "batch": batch, "space": space, "grid": grid, "ksize": k,
               "gpu_fps": round(fps, 1), "ms_per_batch": round(ms, 4), "bit_exact_vs_cv2": ok,
               "hbm_roofline_frac": round(2 * 3 * h * w * fps / 1e9 / 6533.8, 4)}
        if cpu:
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < 1.5:
                cv2_chain.chain(host[n % batch], space, 2.0, grid, k); n += 1
            rec["cv2_fps_as_shipped"] = round(n / (time.perf_counter() - t0), 1)
            rec["cv2_threads"] = cv2.getNumThreads()
        print(json.dumps(rec), flush=True)
        del d_in, d_out
    # C5: chain fused with letterbox + fp16 NCHW normalise to 640x640, batch 128 (1080p, YCrCb, k3 = default.yaml chain)
    from oracle import rv_oracle as O
    for k, want_full in ((3, False), (3, True), (5, False)):
        if only:
            break
        h, w, batch, size = 1080, 1920, 128, 640
        pool = pools[(h, w)]
        host = np.stack([pool[i % len(pool)] for i in range(batch)])
        d_in = torch.from_numpy(host).cuda()
        d_t = torch.empty((batch, 3, size, size), dtype=torch.float16, device="cuda")
        d_full = torch.empty_like(d_in) if want_full else None
        p = rvb200.Params.make("YCrCb", 2.0, 8, k)
        run = lambda: ctx.chain_letterbox_device(d_in.data_ptr(), d_t.data_ptr(), batch, h, w, p, size, 114,
                                                 d_full.data_ptr() if want_full else None, stream=st.cuda_stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        want = O.letterbox_f16(O.chain(host[1], O.SPACE_YCRCB, 2.0, 8, k), size)
        ok = bool(np.array_equal(d_t[1].cpu().numpy().view(np.uint16), want.view(np.uint16)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fps = batch / (ms * 1e-3)
        algo = 3 * h * w + 3 * size * size * 2 + (3 * h * w if want_full else 0)
        print(json.dumps({"config": f"C5 1080p chain(k{k})+letterbox 640 fp16 NCHW batch 128" + (" +full-res out" if want_full else ""),
                          "gpu_fps": round(fps, 1), "ms_per_batch": round(ms, 4), "bit_exact_vs_oracle": ok,
                          "algorithmic_bytes_per_frame": algo, "hbm_roofline_frac": round(algo * fps / 1e9 / 6533.8, 4)}), flush=True)
        del d_in, d_t, d_full


if __name__ == "__main__":
    main()
