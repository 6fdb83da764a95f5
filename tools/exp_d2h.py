#!/usr/bin/env python3
"""Experiments on the full-frame end-to-end path at N ranks (VERDICT r1 task 1c): where do the result frames land?

Run one rank per GPU (torchrun) or single process.  Per rank, 64 x 1080p frames per step, headline chain, pinned input:
  base      D2H copy-engine copies into cudaHostAlloc'ed memory (the shipped path)
  thp       the same into transparent-huge-page backed memory page-locked with rv_host_register
  mapped    k_chain stores straight into the page-locked result buffer over PCIe (no device staging copy, no D2H copy)
  keep      results stay on the GPU (no D2H): the H2D-only ceiling
Prints one JSON line from rank 0 with the aggregate frames/s of every variant (max time over ranks)."""
import ctypes as C
import json
import mmap
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvb200  # noqa: E402
import bench  # noqa: E402

H, W, B = 1080, 1920, 64


def thp_buffer(nbytes):
    """Anonymous mapping aligned to 2 MiB with MADV_HUGEPAGE, touched page by page."""
    two = 2 << 20
    size = (nbytes + two - 1) // two * two
    m = mmap.mmap(-1, size + two, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = C.addressof(C.c_char.from_buffer(m))
    off = (-base) % two
    try:
        m.madvise(mmap.MADV_HUGEPAGE, off, size)
        adv = True
    except Exception:
        adv = False
    arr = np.frombuffer(m, np.uint8, size, off)
    arr[::4096] = 0
    return m, arr[:nbytes], adv


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        bench.bind_near_gpu(local)
    ctx = rvb200.Context(local)
    p = rvb200.Params.make("YCrCb", 2.0, 8, 5)
    pool = bench.load_pool()
    host = np.stack([pool[i % 8] for i in range(B)])
    pin_in = ctx.pinned_empty(host.shape)
    pin_in[:] = host
    pin_out = ctx.pinned_empty(host.shape)
    m, thp, adv = thp_buffer(host.nbytes)
    thp_out = thp.reshape(host.shape)
    ctx._ck(ctx._lib.rv_host_register(ctx._h, thp_out.ctypes.data, host.nbytes))
    ctx._pinned[thp_out.ctypes.data] = host.nbytes
    dev_out = rvb200.DeviceArray(ctx, host.shape)
    steps = int(os.environ.get("STEPS", "8"))

    def run(fn):
        fn(); fn()
        return bench.timed_loop(fn, steps)

    def base():
        ctx.submit_io(pin_in, p, out=pin_out); ctx.wait()

    def thp_fn():
        ctx.submit_io(pin_in, p, out=thp_out); ctx.wait()

    def mapped():
        ctx.submit_io(pin_in, p, out=int(pin_out.ctypes.data)); ctx.wait()      # host pointer handed over as a device pointer (UVA)

    def keep():
        ctx.submit_io(pin_in, p, out=dev_out); ctx.wait()

    res = {}
    want = None
    for name, fn, buf in (("base", base, pin_out), ("thp", thp_fn, thp_out), ("mapped", mapped, pin_out), ("keep", keep, None)):
        if buf is not None:
            buf[:] = 0
        t = run(fn)
        if buf is not None:
            if want is None:
                want = buf.copy()
            assert np.array_equal(buf, want), name
        res[name] = round(bench.sum_over_ranks(B * steps) / t, 1)
    thp_pages = None
    try:
        for line in open("/proc/self/smaps_rollup"):
            if line.startswith("AnonHugePages"):
                thp_pages = line.split()[1] + " kB"
    except OSError:
        pass
    if rank == 0:
        print(json.dumps({"exp": "full-frame e2e: where results land", "n_gpus": world, "frames_per_gpu_per_step": B, "steps": steps,
                          "fps": res, "madvise_hugepage": adv, "anon_huge_pages": thp_pages,
                          "thp_enabled": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()}), flush=True)
    ctx._lib.rv_host_unregister(ctx._h, thp_out.ctypes.data)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
