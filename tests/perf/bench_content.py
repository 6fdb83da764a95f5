#!/usr/bin/env python3
"""Performance bookends: the same kernels on fogged road scenes, uniform-random frames (every table entry and histogram
bin in use: the worst case for data-dependent shared-memory reads) and constant frames (the worst case for histogram atomics).
Device-resident, 1080p x 64, CUDA events, results checked against cv2 on one frame of each kind.  One JSON line per case."""
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import rvb200  # noqa: E402
from rvb200 import synth  # noqa: E402
from oracle import cv2_chain  # noqa: E402

ctx = rvb200.Context(0)
h, w, batch = 1080, 1920, 64
rng = np.random.RandomState(7)
pool = synth.frame_pool(h, w, 4, base_seed=3000)
kinds = {
    "fogged road scenes": np.stack([pool[i % 4] for i in range(batch)]),
    "uniform random": np.stack([rng.randint(0, 256, (h, w, 3)).astype(np.uint8) for _ in range(4)] * (batch // 4)),
    "constant": np.stack([np.full((h, w, 3), 40 + 3 * i, np.uint8) for i in range(batch)]),
    "dark (night) scenes": np.stack([(pool[i % 4] // 8) for i in range(batch)]),
}
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for space, k in (("YCrCb", 5), ("LAB", 3), ("YCrCb", 3)):
    p = rvb200.Params.make(space, 2.0, 8, k)
    for kind, host in kinds.items():
        d_in = torch.from_numpy(host).cuda(); d_out = torch.empty_like(d_in)
        run = lambda: ctx.submit_device(d_in.data_ptr(), d_out.data_ptr(), batch, h, w, p, stream=st.cuda_stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ok = bool(np.array_equal(d_out[1].cpu().numpy(), cv2_chain.chain(host[1], space, 2.0, 8, k)))
        ctx.set_option("kernel_timing", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 30
        for _ in range(n):
            run()
        e1.record(); torch.cuda.synchronize()
        kt = ctx.kernel_times(reset=True)
        ctx.set_option("kernel_timing", 0)
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"chain": f"{space} k{k}", "frames": kind, "gpu_fps": round(batch / ms * 1e3), "bit_exact_vs_cv2": ok,
                          "us_per_batch": {a: round(1e3 * v[0] / max(v[1], 1), 1) for a, v in kt.items()}}), flush=True)
        del d_in, d_out
