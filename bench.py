#!/usr/bin/env python3
"""Headline benchmark: 1080p preprocessing-chain frames/s on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of the hot path (CLAHEDehaze -> MedianDerain) over one batch of synthetic fogged +
rained frames.  Workload at every N: BASELINE.json configs[1] -- 1920x1080, batch 64 per GPU, YCrCb,
clip 2.0, tile grid 8, median k5 (frames/streams shard across GPUs with no inter-GPU traffic, so N GPUs
process N such batches: weak scaling).

`value`  : frames/s with inputs and outputs resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the public plugin API (PreprocessPipeline.process_batch) from pinned host
           buffers, H2D and D2H inside the timed region.
`roofline`: the dominant kernel (k_chain) against the measured HBM copy bandwidth; its time is measured
           live with CUDA events around every launch in a second pass of the same K steps
           (event pairs between back-to-back launches would perturb `value`).
`cpu_baseline`: the reference's own six cv2 calls (oracle/cv2_chain.py) on this box's host cores.
`--impl reference` prints the same line for that CPU implementation alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, BATCH = 1080, 1920, 64
SPACE, CLIP, GRID, KSIZE = "YCrCb", 2.0, 8, 5
POOL = 8                                   # distinct synthetic frames, tiled to the batch
FRAME_BYTES = 3 * H * W
ALGO_BYTES_PER_FRAME = 2 * FRAME_BYTES     # read BGR once + write BGR once
METRIC, UNIT = "1080p preproc-chain frames/s", "frames/s"


# ----------------------------------------------------------------------------- sharding / reductions
def shard_range(n, rank, world):
    """Contiguous slab [a, b) of n items for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def _reduce(x, op, use_cuda):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device="cuda" if use_cuda else "cpu")
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(x, use_cuda=True):
    import torch.distributed as dist
    return _reduce(x, dist.ReduceOp.MAX, use_cuda)


def sum_over_ranks(x, use_cuda=True):
    import torch.distributed as dist
    return _reduce(x, dist.ReduceOp.SUM, use_cuda)


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm (reference's cv2 calls)
_WORKER_FRAMES = None


def _cpu_init(frames):
    """Pool initializer: every spawned worker receives the frame pool once, outside any timed region."""
    global _WORKER_FRAMES
    import cv2
    cv2.setNumThreads(1)
    _WORKER_FRAMES = frames


def _cpu_worker(args):
    """Runs in a *spawned* worker (never forked: a forked child inherits cv2's thread-pool state and can deadlock)."""
    from oracle import cv2_chain
    idx, reps = args
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in idx:
            cv2_chain.chain(_WORKER_FRAMES[i % len(_WORKER_FRAMES)], SPACE, CLIP, GRID, KSIZE)
    return time.perf_counter() - t0


def cpu_chain_fps(frames, budget_s, mode):
    """frames/s of the reference's cv2 chain on the host. mode 'threads': cv2's own pool (as shipped);
    mode 'procs': one single-threaded cv2 per core, frame-parallel (highest CPU throughput)."""
    import cv2
    from oracle import cv2_chain
    if mode == "threads":
        cv2_chain.chain(frames[0], SPACE, CLIP, GRID, KSIZE)
        n, t0 = 0, time.perf_counter()
        while True:
            for f in frames:
                cv2_chain.chain(f, SPACE, CLIP, GRID, KSIZE)
            n += len(frames)
            if time.perf_counter() - t0 > budget_s:
                break
        return n / (time.perf_counter() - t0), cv2.getNumThreads(), n
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per = 2
    reps = max(1, int(budget_s / (0.04 * per)))          # ~40 ms per 1080p frame on one core
    reps = min(reps, 8)
    with mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(list(frames),)) as pool:
        pool.map(_cpu_worker, [([0], 1)] * cores)                 # warm the workers
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(list(range(c, c + per)), reps) for c in range(cores)])
        dt = time.perf_counter() - t0
    n = cores * per * reps
    return n / dt, cores, n


def bind_near_gpu(index):
    """Pin this rank to the host cores next to its GPU (PCI device's local_cpulist) so that pinned staging buffers are
    first-touched on the GPU's NUMA node; with 8 ranks the end-to-end leg is host-memory bound.  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"{len(use)} cores local to {bdf}"
        return "affinity unchanged (all allowed cores are local or none is)"
    except Exception as e:        # noqa: BLE001
        return f"affinity unchanged ({type(e).__name__})"


def make_pool():
    import rvb200  # noqa: F401  (package import only; no GPU needed for the generator)
    from rvb200 import synth
    return synth.frame_pool(H, W, POOL, base_seed=2000)


def base_line(n_gpus, steps, warmup):
    return {
        "metric": METRIC, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic (seeded road scenes + fog + rain streaks, 8 distinct 1080p frames tiled to the batch)",
        "config": {"workload": "BASELINE configs[1]: 1920x1080 batch-64 chain, YCrCb, clip 2.0, tile_grid 8, MedianDerain k5",
                   "frames_per_gpu_per_step": BATCH, "global_batch": BATCH * n_gpus, "parallelism": f"frames sharded x{n_gpus}, no collective",
                   "l2": "per-step inputs+outputs (796 MB per GPU) exceed the 126 MB L2; no explicit flush"},
    }


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (its cv2 calls) on all host cores; rank 0 only."""
    if rank != 0:
        return
    pool = make_pool()
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    sample = min(max(16, 2 * cores), 4 * BATCH)   # frames per step: two per worker so every host core stays busy
    frames = [pool[i % POOL] for i in range(sample)]
    with mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(list(pool),)) as pp:
        chunks = [list(range(c, sample, cores)) for c in range(cores)]
        chunks = [c for c in chunks if c]
        for _ in range(max(args.warmup, 1)):
            pp.map(_cpu_worker, [(c, 1) for c in chunks])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pp.map(_cpu_worker, [(c, 1) for c in chunks])
        dt = time.perf_counter() - t0
    fps_procs = sample * args.steps / dt
    fps_thr, nthr, _ = cpu_chain_fps(frames[:4], 3.0, "threads")
    best = max(fps_procs, fps_thr)
    line = base_line(args.gpus, args.steps, args.warmup)
    line.update({
        "impl": "reference", "value": best, "ms_per_step": 1e3 * sample / best,
        "config": dict(line["config"], sample=f"{sample} frames per step (bounded sample of the 64-frame batch)"),
        "cpu_baseline": {"value": best, "unit": UNIT, "cores": cores if fps_procs >= fps_thr else nthr, "kind": "port",
                         "sample": f"{sample} x 1080p frames/step; reference's six cv2 calls (oracle/cv2_chain.py, cv2 "
                                   f"{__import__('cv2').__version__}); frame-parallel {cores} procs x 1 thread = {fps_procs:.1f} fps, "
                                   f"as shipped ({nthr} cv2 threads) = {fps_thr:.1f} fps"},
        "e2e": {"value": best, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(line), flush=True)


def run_gpu(args, rank, world, local_rank):
    import numpy as np
    import torch
    import rvb200

    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(local_rank) if world > 1 else "single rank: not bound"
    ctx = rvb200.Context(local_rank)
    if args.group:
        ctx.set_option("group_frames", args.group)
    if args.chunk:
        ctx.set_option("chunk_frames", args.chunk)
    if args.overlap >= 0:
        ctx.set_option("overlap_groups", args.overlap)
    if args.prefetch >= 0:
        ctx.set_option("prefetch_ctas", args.prefetch)
    params = rvb200.Params.make(SPACE, CLIP, GRID, KSIZE)
    pool = make_pool()
    a, _ = shard_range(BATCH * world, rank, world)            # this rank's slab of the global frame index space
    host = np.stack([pool[(a + i) % POOL] for i in range(BATCH)])
    d_in = torch.from_numpy(host).cuda()
    d_out = torch.empty_like(d_in)
    tstream = torch.cuda.Stream()          # the stream every kernel of the step is launched on; events are recorded on it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        ctx.submit_device(d_in.data_ptr(), d_out.data_ptr(), BATCH, H, W, params, stream=stream)

    step(); torch.cuda.synchronize()
    gpu_first = d_out[0].cpu().numpy()          # checked against the CPU chain in the cpu_baseline leg below

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(); barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start(); time.sleep(0.25)
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(); barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - l0
    ms_max = max_over_ranks(ms)
    frames_total = sum_over_ranks(BATCH * args.steps)
    value = frames_total / (ms_max * 1e-3)

    # second pass: per-kernel device time with an event pair around every launch
    ctx.set_option("kernel_timing", 1)
    ctx.kernel_times(reset=True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    kt = ctx.kernel_times(reset=True)
    ctx.set_option("kernel_timing", 0)
    # keep the same load running until nvidia-smi (100 ms period) has seen it for >= 1.5 s, then read the clocks
    while time.time() - t0 < 1.5:
        step()
        torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed steps + per-kernel pass + same-load soak, %.2f s" % (t1 - t0)

    # end to end through the plugin API from pinned host memory
    pin_in, pin_out = ctx.pinned_empty(host.shape), ctx.pinned_empty(host.shape)
    pin_in[:] = host
    pl = rvb200.PreprocessPipeline({"enabled": True, "device": local_rank, "chain": [
        {"name": "CLAHEDehaze", "params": {"space": SPACE, "clip_limit": CLIP, "tile_grid": GRID}},
        {"name": "MedianDerain", "params": {"ksize": KSIZE}}]})
    e2e_steps = max(3, min(args.steps, 10))
    if args.no_e2e:
        e2e_steps = 0
    for _ in range(2 if e2e_steps else 0):
        pl.process_batch(pin_in, out=pin_out)
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        pl.process_batch(pin_in, out=pin_out)
        _ = int(pin_out[-1, -1, -1, 0])                       # host read of the step's result
    w1 = time.perf_counter()
    if e2e_steps and not np.array_equal(pin_out[0], gpu_first):
        raise SystemExit("bench: end-to-end output differs from the device-resident output")
    e2e_s = max_over_ranks(w1 - w0)
    e2e_value = sum_over_ranks(BATCH * e2e_steps) / e2e_s if e2e_steps else None

    if rank != 0:
        return
    peaks, peak_src = None, "fallback (B200_PROFILING.md: 6650 GB/s)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0
    chain_ms, chain_n = kt["k_chain"]
    frames_per_launch = BATCH * args.steps / max(chain_n, 1)
    avg_launch_s = chain_ms * 1e-3 / max(chain_n, 1)
    achieved = ALGO_BYTES_PER_FRAME * frames_per_launch / avg_launch_s / 1e9 if chain_n else None
    total_k = sum(v[0] for v in kt.values())
    line = base_line(world, args.steps, args.warmup)
    line.update({
        "value": value, "ms_per_step": ms_max / args.steps, "gpu_launches": launches, "clocks": clocks,
        "host_affinity": numa,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * FRAME_BYTES, "d2h_bytes_per_step": BATCH * FRAME_BYTES,
                "steps": e2e_steps, "api": "PreprocessPipeline.process_batch(pinned in, pinned out)"},
        "roofline": {"bound": "hbm", "kernel": "k_chain<YCrCb,5>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_frame": ALGO_BYTES_PER_FRAME, "frames_per_launch": frames_per_launch,
                     "avg_launch_ms": avg_launch_s * 1e3,
                     "kernel_share_of_step": {k: (v[0] / total_k if total_k else None) for k, v in kt.items()},
                     "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kt.items()},
                     "whole_chain_achieved_gbs": ALGO_BYTES_PER_FRAME * value / world / 1e9,
                     "traffic": TRAFFIC_NCU},
    })
    # why the HBM fraction is low: k_chain is bound by instruction issue, not by memory.  Warp instructions
    # per launch come from the committed ncu capture, the launch time is the live measurement above.
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_peak = sms * 4 * sm_clock * 1e6                      # one warp instruction per scheduler per clock
    line["roofline"]["issue"] = {
        "warp_instr_per_launch": WARP_INSTR_NCU * frames_per_launch / 64.0, "thread_instr_per_pixel": WARP_INSTR_NCU * 32 / (64.0 * H * W),
        "achieved_gwarp_instr_s": WARP_INSTR_NCU * frames_per_launch / 64.0 / avg_launch_s / 1e9 if chain_n else None,
        "peak_gwarp_instr_s": issue_peak / 1e9,
        "frac": (WARP_INSTR_NCU * frames_per_launch / 64.0 / avg_launch_s / issue_peak) if chain_n else None,
        "source": "smsp__inst_executed.sum of one k_chain launch (profiles/r1_v9_k_chain_summary.txt); peak = SMs x 4 schedulers x SM clock"}
    if world == 1 and not args.no_cpu:
        fps_p, cores_p, n_p = cpu_chain_fps(list(pool), 6.0, "procs")
        from oracle import cv2_chain            # the CPU leg doubles as the checker: never report a wrong kernel's speed
        if not np.array_equal(gpu_first, cv2_chain.chain(host[0], SPACE, CLIP, GRID, KSIZE)):
            raise SystemExit("bench: CUDA chain differs from the reference's cv2 chain; refusing to report")
        line["parity"] = "frame 0 of the timed batch bit-exact vs the reference's cv2 chain"
        fps_t, cores_t, n_t = cpu_chain_fps(list(pool[:4]), 4.0, "threads")
        import cv2
        best = max(fps_p, fps_t)
        line["cpu_baseline"] = {
            "value": best, "unit": UNIT, "cores": cores_p if fps_p >= fps_t else cores_t, "kind": "port",
            "sample": f"reference's six cv2 calls (oracle/cv2_chain.py, cv2 {cv2.__version__}) on the same 1080p frames: "
                      f"frame-parallel {cores_p} procs x 1 thread, {n_p} frames = {fps_p:.1f} fps; "
                      f"as shipped ({cores_t} cv2 threads), {n_t} frames = {fps_t:.1f} fps"}
    print(json.dumps(line), flush=True)


# DRAM bytes of one k_chain launch (64 x 1080p frames: dram__bytes_read.sum 403.5 MB + dram__bytes_write.sum 356.1 MB) and its
# executed warp instructions, from the committed `ncu --set full` capture profiles/r1_v9_k_chain_summary.txt; algorithmic
# bytes of that launch: 796.3 MB.  (The shipped 5x5 network has since lost 8 of its 748 operations: the instruction count is
# about 1 % lower than this capture, the traffic is unchanged.)
TRAFFIC_NCU = 759.6e6
WARP_INSTR_NCU = 782662144


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--group", type=int, default=0, help="frames per hist->lut->chain group (0 = library default)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per host pipeline chunk (0 = library default)")
    ap.add_argument("--overlap", type=int, default=-1, help="groups per batch for the histogram/chain overlap (-1 = library default)")
    ap.add_argument("--prefetch", type=int, default=-1, help="k_chain L2 prefetch distance in CTAs per SM (-1 = library default, 0 = off)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # plain `python bench.py --gpus N`: relaunch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29577")] + sys.argv
        raise SystemExit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_gpu(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
