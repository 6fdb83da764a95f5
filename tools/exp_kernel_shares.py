#!/usr/bin/env python3
"""Per-kernel time of one batch call (CUDA events inside the library, option kernel_timing) for a few configs.
usage: [RV_B200_LIB=...] tools/exp_kernel_shares.py"""
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import rvb200  # noqa: E402
from rvb200 import synth  # noqa: E402

ctx = rvb200.Context(0)
pool = synth.frame_pool(1080, 1920, 4, base_seed=3000)
for name, (h, w), batch, space, grid, k in [("1080p LAB k3", (1080, 1920), 64, "LAB", 8, 3), ("4K LAB grid16 k3", (2160, 3840), 16, "LAB", 16, 3),
                                            ("1080p YCrCb k5", (1080, 1920), 64, "YCrCb", 8, 5)]:
    fr = pool if (h, w) == (1080, 1920) else np.stack([np.tile(f, (2, 2, 1)) for f in pool])
    host = np.stack([fr[i % len(fr)] for i in range(batch)])
    d_in = torch.from_numpy(host).cuda(); d_out = torch.empty_like(d_in)
    p = rvb200.Params.make(space, 2.0, grid, k)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        ctx.submit_device(d_in.data_ptr(), d_out.data_ptr(), batch, h, w, p, stream=st)
    torch.cuda.synchronize()
    ctx.set_option("kernel_timing", 1)
    n = 50
    for _ in range(n):
        ctx.submit_device(d_in.data_ptr(), d_out.data_ptr(), batch, h, w, p, stream=st)
    torch.cuda.synchronize()
    kt = ctx.kernel_times(reset=True)
    ctx.set_option("kernel_timing", 0)
    print(json.dumps({"config": name, "lib": os.environ.get("RV_B200_LIB", "librv_b200.so"), "us_per_batch": {k2: round(1e3 * v[0] / max(v[1], 1), 1) for k2, v in kt.items()}}))
