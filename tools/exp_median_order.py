#!/usr/bin/env python3
"""Experiment: does the emission order of the 5x5 two-row network, or the choice of WHICH compare-exchanges take the
FMA-pipe form, change its issue efficiency on sm_100a?  Writes alternative rv_median_net.h files under tools/exp/ (the
networks are identical DAGs, verified by the generator; only order / form assignment differ) for tools/ubench_median_sol.cu
and for full-kernel builds (-DRV_MEDIAN_NET_FILE=...).
  order:  creation (shipped) | bfs (by DAG depth) | dfs (post-order from the outputs, short live ranges)
  forms:  alt (every other compare-exchange, shipped) | slack (FMA form -- twice the latency -- for the compare-exchanges
          with the most slack off the critical path) | level (alternate within each DAG depth level)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_median_net as G


def build():
    best = None
    for parity in (0, 1):
        for rpar in (0, 1):
            d, inp, outs = G.build5_2rows(6, parity, rpar)
            nops = sum(1 for n in d.live(outs) if d.nodes[n][0] != "in")
            if best is None or nops < best[0]:
                best = (nops, d, inp, outs)
    return best[1:]


def emit(d, inp, outs, fh, order_kind, form_kind, k=5, M=6):
    live = d.live(outs)
    nodes = sorted(live)
    depth = {}
    for n in nodes:
        op, a, b = d.nodes[n]
        depth[n] = 0 if op == "in" else 1 + max(depth[a], depth[b])
    users = {n: [] for n in nodes}
    for n in nodes:
        op, a, b = d.nodes[n]
        if op != "in":
            users[a].append(n); users[b].append(n)
    # height = longest path to an output
    height = {}
    for n in reversed(nodes):
        height[n] = 0 if not users[n] else 1 + max(height[u] for u in users[n])
    crit = max(depth[n] + height[n] for n in nodes)
    partner = {}
    for n in nodes:
        op, a, b = d.nodes[n]
        if op == "min" and ("max", a, b) in d.memo and d.memo[("max", a, b)] in live:
            m = d.memo[("max", a, b)]
            partner[n] = m; partner[m] = n
    if order_kind == "creation":
        order = nodes
    elif order_kind == "bfs":
        order = sorted(nodes, key=lambda n: (depth[n], n))
    else:
        order, seen = [], set()

        def visit(n):
            stack = [(n, False)]
            while stack:
                x, done = stack.pop()
                if done:
                    order.append(x); continue
                if x in seen:
                    continue
                seen.add(x)
                stack.append((x, True))
                op, a, b = d.nodes[x]
                if op != "in":
                    stack.append((b, False)); stack.append((a, False))
        for o in outs:
            visit(o)
        order = [n for n in order if n in live]
    ces = [n for n in order if n in partner and d.nodes[n][0] == "min"]
    if form_kind == "alt":
        fma = None
    elif form_kind == "slack":
        slack = {n: crit - max(depth[n] + height[n], depth[partner[n]] + height[partner[n]]) for n in ces}
        ranked = sorted(ces, key=lambda n: -slack[n])
        fma = set(ranked[:len(ces) // 2 + 1])
    else:
        fma, cnt = set(), {}
        for n in sorted(ces, key=lambda n: (depth[n], n)):
            c = cnt.get(depth[n], 0); cnt[depth[n]] = c + 1
            if c % 2 == 0:
                fma.add(n)
    ncol = M + k - 1
    fh.write("#ifndef RV_MEDIAN_NET_H\n#define RV_MEDIAN_NET_H\n#include <stdint.h>\n"
             "#define RV_MN(a, b) __vminu2((a), (b))\n#define RV_MX(a, b) __vmaxu2((a), (b))\n"
             "#define RV_CEX5X2 RV_CEX\n#define RV_CEX3X2 RV_CEX3\n"
             "#define RV_MEDIAN3_M 4\n#define RV_MEDIAN5_M 6\n#define RV_MEDIAN7_M 4\n#define RV_MEDIAN9_M 2\n")
    fh.write(f"#define RV_MEDIAN{k}X2_M {M}\n")
    fh.write(f"__device__ __forceinline__ void rv_median{k}x2_net(const uint32_t (&v)[{ncol}][{k + 1}], uint32_t (&out)[2][{M}])\n{{\n")
    name, done, even, odd = {}, set(), 0, 1
    for n in order:
        op, a, b = d.nodes[n]
        if op == "in":
            name[n] = f"v[{a[0]}][{a[1]}]"; continue
        if n in done:
            continue
        if n in partner:
            lo, hi = (n, partner[n]) if op == "min" else (partner[n], n)
            # both operands must already be named: in dfs order the partner is emitted with the first of the pair
            name[lo], name[hi] = f"t{lo}", f"t{hi}"
            if fma is None:
                idx = even + odd - 1 if False else None
            if fma is None:
                idx = len(done) // 2
            elif lo in fma:
                idx = even; even += 2
            else:
                idx = odd; odd += 2
            fh.write(f"    RV_CEX{k}X2({idx}, t{lo}, t{hi}, {name[a]}, {name[b]});\n")
            done.update((lo, hi))
        else:
            name[n] = f"t{n}"
            fh.write(f"    const uint32_t t{n} = {'RV_MN' if op == 'min' else 'RV_MX'}({name[a]}, {name[b]});\n")
    for half in range(2):
        for o in range(M):
            fh.write(f"    out[{half}][{o}] = {name[outs[half * M + o]]};\n")
    fh.write("}\n")
    # the 3x3 two-row network, unchanged emission (the micro-benchmark instantiates it too)
    best = None
    for parity in (0, 1):
        d3, inp3, outs3 = G.build3_2rows(6, parity)
        nops = sum(1 for n in d3.live(outs3) if d3.nodes[n][0] != "in")
        if best is None or nops < best[0]:
            best = (nops, d3, inp3, outs3)
    G.emit_2rows(best[1], best[2], best[3], 6, fh, k=3)
    fh.write("#endif\n")
    return crit


def main():
    d, inp, outs = build()
    assert G.verify_2rows(d, inp, outs, 6, zero_one=False)
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exp")
    os.makedirs(root, exist_ok=True)
    for order_kind in ("creation", "bfs", "dfs"):
        for form_kind in ("alt", "slack", "level"):
            path = os.path.join(root, f"net_{order_kind}_{form_kind}.h")
            with open(path, "w") as fh:
                crit = emit(d, inp, outs, fh, order_kind, form_kind)
            print(path, "critical path", crit, "levels")


if __name__ == "__main__":
    main()
