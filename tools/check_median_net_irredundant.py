#!/usr/bin/env python3
"""Are the shipped two-row selection networks locally minimal?  For every min/max node of the live DAG, try to replace it by
either of its operands and re-check all outputs on 40,000 vectors (wide range, heavy ties, 0/1 inputs).  A replacement that
survives would be a removable operation.  Result on the shipped networks: none (5x5: 740 operations, 3x3: 212).
    python tools/check_median_net_irredundant.py [5|3]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_median_net as G


def shipped(k, M=6):
    best = None
    for parity in (0, 1):
        for rpar in ((0, 1) if k == 5 else (0,)):
            d, inp, outs = G.build5_2rows(M, parity, rpar) if k == 5 else G.build3_2rows(M, parity)
            n = sum(1 for x in d.live(outs) if d.nodes[x][0] != "in")
            if best is None or n < best[0]:
                best = (n, d, inp, outs)
    return best


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    M, T = 6, 40000
    nops, d, inp, outs = shipped(k, M)
    ncol, nrow, half_rank = M + k - 1, k + 1, (k * k) // 2
    rng = np.random.RandomState(7)
    data = rng.randint(0, 256, (ncol, nrow, T)).astype(np.int16)
    data[:, :, :T // 3] //= 64
    data[:, :, T // 3:2 * (T // 3)] = rng.randint(0, 2, (ncol, nrow, T // 3)) * 255
    want = [np.sort(data[o:o + k, h:h + k].reshape(k * k, T), axis=0)[half_rank] for h in range(2) for o in range(M)]
    live = sorted(d.live(outs))

    def ok(alias):
        val = {}
        for n in live:
            if n in alias:
                val[n] = val[alias[n]]
                continue
            op, a, b = d.nodes[n]
            val[n] = data[a[0], a[1]] if op == "in" else (np.minimum if op == "min" else np.maximum)(val[a], val[b])
        return all(np.array_equal(val[o], w) for o, w in zip(outs, want))

    assert ok({}), "the network itself fails"
    t0, removable = time.time(), []
    for n in reversed(live):
        op, a, b = d.nodes[n]
        if op == "in":
            continue
        for src in (a, b):
            if ok({n: src}):
                removable.append((n, op, src))
                break
    print(f"k={k}: {nops} operations, removable: {len(removable)} {removable[:5]}  ({time.time() - t0:.0f} s)")
    return 1 if removable else 0


if __name__ == "__main__":
    sys.exit(main())
