#!/usr/bin/env python3
"""BASELINE.json config 4: synthetic 1080p@30 camera streams, one stream per GPU (torchrun: one rank per GPU), or
`--streams S` streams on this rank's GPU.  Each stream: paced SyntheticReader -> BatchFeeder (pinned ring) ->
PreprocessPipeline.process_batch -> latency = completion time - capture timestamp.  `--fps 0` replays unpaced (saturation).
Prints one JSON line per rank."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--fps", type=float, default=30.0)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--batch", type=int, default=1, help="frames per process_batch call (1 = lowest latency)")
    ap.add_argument("--skip", type=float, default=2.0, help="seconds at the start of every stream excluded from the statistics")
    args = ap.parse_args()
    import rvb200
    from rvb200 import synth
    from rvb200.io_video.capture import SyntheticReader
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    H, W = 1080, 1920
    pool = synth.frame_pool(H, W, 4, base_seed=4000 + rank)
    cfg = {"enabled": True, "device": local, "chain": [
        {"name": "CLAHEDehaze", "params": {"space": "YCrCb", "clip_limit": 2.0, "tile_grid": 8}},
        {"name": "MedianDerain", "params": {"ksize": 5}}]}
    nframes = int(args.seconds * (args.fps if args.fps else 2000))
    results = [None] * args.streams

    def stream(i):
        ctx = rvb200.Context(local)                       # one context (streams + workspaces) per camera stream
        pipe = rvb200.PreprocessPipeline(cfg, context=ctx)
        vs = rvb200.VideoSource(reader=SyntheticReader(list(pool), limit=nframes), pinned=None)   # the feeder's ring is the pinned memory
        feeder = rvb200.BatchFeeder(vs, batch=args.batch, shape=(H, W, 3), alloc=ctx.pinned_empty, depth=3, fps=args.fps or None)
        out = ctx.pinned_empty((args.batch, H, W, 3))
        warm = ctx.pinned_empty((args.batch, H, W, 3))
        warm[:] = pool[0]
        for _ in range(3):                                # allocations, tables, graph capture before the camera starts
            pipe.process_batch(warm, out=out)
        lat, stamps = [], []
        for b in feeder:
            pipe.process_batch(b.frames, out=out[:b.count])
            done = time.time()
            lat.extend(done - b.ts)
            stamps.extend(b.ts)
            feeder.release(b)
        lat, stamps = np.array(lat), np.array(stamps)
        keep = stamps >= stamps[0] + args.skip                      # steady state: the first seconds (thread start-up) are excluded
        lat, st = np.sort(lat[keep]), stamps[keep]
        span = float(st[-1] - st[0]) if len(st) > 1 else 0.0
        results[i] = {"frames": int(keep.sum()), "fps": (len(st) - 1) / span if span > 0 else 0.0, "lat_ms_p50": 1e3 * float(lat[len(lat) // 2]),
                      "lat_ms_p99": 1e3 * float(lat[int(len(lat) * 0.99)]), "lat_ms_max": 1e3 * float(lat[-1])}

    threads = [threading.Thread(target=stream, args=(i,)) for i in range(args.streams)]
    t0 = time.time()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    wall = time.time() - t0
    print(json.dumps({"config": "C4 1080p streams, chain YCrCb k5", "rank": rank, "gpu": local, "streams": args.streams,
                      "fps_requested": args.fps, "batch": args.batch, "wall_s": round(wall, 2),
                      "excluded_first_s": args.skip, "total_fps": round(sum(r["fps"] for r in results), 1), "per_stream": results}), flush=True)


if __name__ == "__main__":
    main()
