#!/usr/bin/env python3
"""Generate the min/max selection networks used by the CUDA median kernel.

The k x k median of /root/reference/src/preprocess/ops/median_derain.py:14 (cv2.medianBlur)
is computed per byte lane with packed 16-bit min/max (VIMNMX.U16x2 on sm_100a).  One call of
the generated function produces M horizontally adjacent outputs of one channel plane from
(M + k - 1) pixel columns of k rows each, so that the column sorts are shared by the k
windows that contain them.  The network is built as a hash-consed DAG:

  1. sort every column (optimal small sorting networks, verified below);
  2. per output: merge tree over its k sorted columns (Batcher odd-even merges; pairs of
     columns at even positions are merged once and shared by up to four outputs; the pairing parity is searched), reduced to
     the rank that matters with the two-list selection identity
        kth(A u B) = min_{i+j=k} max(a_i, b_j);
  3. dead-code elimination from the M median outputs.

Every generated network is verified before it is written: exhaustively with the 0-1
principle for k <= 5 (2^25 inputs per output, bit-parallel) and with random vectors for all k.

Writes road-vision-system_b200/csrc/rv_median_net.h.
"""
import itertools
import os
import sys

import numpy as np

SORTERS = {
    2: [(0, 1)],
    3: [(0, 2), (0, 1), (1, 2)],
    4: [(0, 2), (1, 3), (0, 1), (2, 3), (1, 2)],
    5: [(0, 3), (1, 4), (0, 2), (1, 3), (0, 1), (2, 4), (1, 2), (3, 4), (2, 3)],
    6: [(0, 5), (1, 3), (2, 4), (1, 2), (3, 4), (0, 3), (2, 5), (0, 1), (2, 3), (4, 5), (1, 2), (3, 4)],
    7: [(0, 6), (2, 3), (4, 5), (0, 2), (1, 4), (3, 6), (0, 1), (2, 5), (3, 4), (1, 2), (4, 6),
        (2, 3), (4, 5), (1, 2), (3, 4), (5, 6)],
    8: [(0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7), (0, 1), (2, 3), (4, 5), (6, 7), (2, 4), (3, 5),
        (1, 4), (3, 6), (1, 2), (3, 4), (5, 6)],
    9: [(0, 3), (1, 7), (2, 5), (4, 8), (0, 7), (2, 4), (3, 8), (5, 6), (0, 2), (1, 3), (4, 5), (7, 8),
        (1, 4), (3, 6), (5, 7), (0, 1), (2, 4), (3, 5), (6, 8), (2, 3), (4, 5), (6, 7), (1, 2), (3, 4), (5, 6)],
}


def check_sorter(n, net):
    for bits in itertools.product((0, 1), repeat=n):
        v = list(bits)
        for a, b in net:
            if v[a] > v[b]:
                v[a], v[b] = v[b], v[a]
        if v != sorted(v):
            return False
    return True


class Dag:
    """Hash-consed min/max expression DAG. Node ids are ints; inputs first."""

    def __init__(self):
        self.nodes = []      # (op, a, b) ; op in {"in","min","max"}
        self.memo = {}

    def inp(self, name):
        self.nodes.append(("in", name, None))
        return len(self.nodes) - 1

    def _op(self, op, a, b):
        if a == b:
            return a
        if a > b:
            a, b = b, a
        key = (op, a, b)
        if key not in self.memo:
            self.nodes.append(key)
            self.memo[key] = len(self.nodes) - 1
        return self.memo[key]

    def mn(self, a, b): return self._op("min", a, b)
    def mx(self, a, b): return self._op("max", a, b)

    def sort(self, xs):
        xs = list(xs)
        for a, b in SORTERS[len(xs)]:
            lo, hi = self.mn(xs[a], xs[b]), self.mx(xs[a], xs[b])
            xs[a], xs[b] = lo, hi
        return xs

    def merge(self, A, B):
        """Batcher odd-even merge of two ascending lists of arbitrary lengths (Knuth 5.3.4)."""
        m, n = len(A), len(B)
        if m == 0:
            return list(B)
        if n == 0:
            return list(A)
        if m == 1 and n == 1:
            return [self.mn(A[0], B[0]), self.mx(A[0], B[0])]
        V = self.merge(A[0::2], B[0::2])     # 1-based odd-indexed elements
        Wl = self.merge(A[1::2], B[1::2])    # 1-based even-indexed elements
        out = [V[0]]
        i = 0
        while i < len(Wl) and i + 1 < len(V):
            out.append(self.mn(Wl[i], V[i + 1]))
            out.append(self.mx(Wl[i], V[i + 1]))
            i += 1
        out.extend(Wl[i:])
        out.extend(V[i + 1:])
        return out

    def kth2(self, A, B, k):
        """k-th smallest (1-based) of the union of two ascending lists: min_{i+j=k} max(a_i,b_j)."""
        terms = []
        for i in range(0, k + 1):
            j = k - i
            if i > len(A) or j > len(B):
                continue
            if i == 0:
                terms.append(B[j - 1])
            elif j == 0:
                terms.append(A[i - 1])
            else:
                terms.append(self.mx(A[i - 1], B[j - 1]))
        r = terms[0]
        for t in terms[1:]:
            r = self.mn(r, t)
        return r

    def select_window(self, lists, lo, hi):
        """Ascending elements of ranks lo..hi (0-based, inclusive) of a merged pair, via full merge
        (dead code is eliminated later)."""
        return self.merge(lists[0], lists[1])[lo:hi + 1]

    def live(self, outs):
        seen = set()
        stack = list(outs)
        while stack:
            n = stack.pop()
            if n in seen:
                continue
            seen.add(n)
            op, a, b = self.nodes[n]
            if op != "in":
                stack.extend((a, b))
        return seen


def _reduce(d, groups, half, order):
    """Merge `groups` (ascending lists) following `order` (list of index pairs into the shrinking
    group list); returns the node holding the element of 0-based rank `half` of the union."""
    groups = [list(g) for g in groups]
    below = 0
    for (i, j) in order:
        a, b = groups[i], groups[j]
        others = [g for t, g in enumerate(groups) if t not in (i, j)]
        rest = sum(len(g) for g in others)
        t = half - below
        if not others:
            return d.kth2(a, b, t + 1)
        merged = d.merge(a, b)
        lo = max(0, t - rest)                 # rank r of merged ends at final rank in [r, r + rest]
        hi = min(len(merged) - 1, t)
        below += lo
        groups = [merged[lo:hi + 1]] + others
    return groups[0][half - below]


def _orders(n):
    if n == 1:
        yield []
        return
    for i in range(n):
        for j in range(i + 1, n):
            for rest in _orders(n - 1):
                yield [(i, j)] + rest


def reduce_best(d, groups, half):
    """Try every merge order; keep the one that adds the fewest new nodes to the shared DAG."""
    if len(groups) > 4:            # the order space explodes; merge the two shortest lists each time
        order, sizes = [], [len(g) for g in groups]
        while len(sizes) > 1:
            idx = sorted(range(len(sizes)), key=lambda t: sizes[t])[:2]
            i, j = min(idx), max(idx)
            order.append((i, j))
            rest = sum(sizes) - sizes[i] - sizes[j]
            sizes = [min(sizes[i] + sizes[j], rest + 1)] + [s for t, s in enumerate(sizes) if t not in (i, j)]
        return _reduce(d, groups, half, order)
    best = None
    for order in _orders(len(groups)):
        n0, keys0 = len(d.nodes), set(d.memo.keys())
        _reduce(d, groups, half, order)
        added = len(d.nodes) - n0
        # roll back
        for key in set(d.memo.keys()) - keys0:
            del d.memo[key]
        del d.nodes[n0:]
        if best is None or added < best[0]:
            best = (added, order)
    return _reduce(d, groups, half, best[1])


def build(k, M, strategy="pairs", parity=0):
    """Return (dag, inputs[col][row], outs[M])."""
    d = Dag()
    ncol = M + k - 1
    inputs = [[d.inp((c, r)) for r in range(k)] for c in range(ncol)]
    cols = [d.sort(col) for col in inputs]
    half = (k * k) // 2          # 0-based rank of the median
    pair_cache = {}

    def pair(c):                 # merged columns (c, c+1); shared when c is even
        if c not in pair_cache:
            pair_cache[c] = d.merge(cols[c], cols[c + 1])
        return pair_cache[c]

    outs = []
    for o in range(M):
        cs = list(range(o, o + k))          # columns of this window
        if strategy == "pairs":
            # greedy pairing on even column indices so neighbouring outputs reuse the pair merges
            groups, i = [], 0
            while i < len(cs):
                if cs[i] % 2 == parity and i + 1 < len(cs):
                    groups.append(pair(cs[i])); i += 2
                else:
                    groups.append(cols[cs[i]]); i += 1
        else:
            groups = [cols[c] for c in cs]
        outs.append(reduce_best(d, [list(g) for g in groups], half))
    return d, inputs, outs


def build5_quads(M, parity):
    """5x5: columns -> pairs P(a) = merge(col a, col a+1) for a of the given parity -> quads
    Q(a) = ranks 7..12 of merge(P(a), P(a+2)) (the only ranks of the 20 that can still be the median of
    25) -> median = 6th smallest of Q(a) u the fifth column.  Q(a) serves outputs a and a-1, P(a)
    serves Q(a) and Q(a-2), a sorted column serves five outputs."""
    k = 5
    d = Dag()
    ncol = M + k - 1
    inputs = [[d.inp((c, r)) for r in range(k)] for c in range(ncol)]
    cols = [d.sort(col) for col in inputs]
    P, Q = {}, {}

    def pair(a):
        if a not in P:
            P[a] = d.merge(cols[a], cols[a + 1])
        return P[a]

    def quad(a):
        if a not in Q:
            Q[a] = d.merge(pair(a), pair(a + 2))[7:13]
        return Q[a]

    outs = []
    for o in range(M):
        if o % 2 == parity:
            outs.append(d.kth2(quad(o), cols[o + 4], 6))
        else:
            outs.append(d.kth2(quad(o + 1), cols[o], 6))
    return d, inputs, outs


def build_hier(k, M, parity):
    """Any odd k: sorted columns -> pairs P(a) (a of the given parity) -> H(a) = the (k-1)/2 pairs a, a+2, ... merged left
    to right, after every merge keeping only the ranks that can still hold the median of k*k -> median = one rank of
    H(a) u the remaining column.  H(a) serves outputs a-1 and a; for k = 5 this is the quads scheme."""
    d = Dag()
    ncol = M + k - 1
    inputs = [[d.inp((c, r)) for r in range(k)] for c in range(ncol)]
    cols = [d.sort(col) for col in inputs]
    half = (k * k) // 2
    P, H = {}, {}

    def pair(a):
        if a not in P:
            P[a] = d.merge(cols[a], cols[a + 1])
        return P[a]

    def trim(lst, below, rest):
        t = half - below
        lo, hi = max(0, t - rest), min(len(lst) - 1, t)
        return lst[lo:hi + 1], below + lo

    def hexa(a):
        if a not in H:
            npairs = (k - 1) // 2
            acc, below, used = pair(a), 0, 2
            for i in range(1, npairs):
                used += 2
                acc, below = trim(d.merge(acc, pair(a + 2 * i)), below, (k - used) * k)
            if npairs == 1:
                acc, below = trim(acc, 0, k)
            H[a] = (acc, below)
        return H[a]

    outs = []
    for o in range(M):
        (acc, below), single = (hexa(o), cols[o + k - 1]) if o % 2 == parity else (hexa(o + 1), cols[o])
        outs.append(d.kth2(acc, single, half - below + 1))
    return d, inputs, outs


def build5_2rows(M, parity, rparity=0):
    """5x5, two vertically adjacent output rows per call (window rows 0..4 and 1..5 of six input rows): the four
    middle rows are shared.  Columns of the middle rows are sorted once (sort4), merged in pairs P4 and quads Q4
    (ranks 3..12 of 16 kept), joined with the fifth column (ranks 4..9 of 14 kept = the only ranks of the 20 shared
    elements that can be the median of 25); each output then takes the 6th smallest of those six and its own
    sorted outer row window (sliding sort4 -- two ordered pairs merged, the pairs shared between groups -- + insert,
    shared by horizontal neighbours)."""
    d = Dag()
    ncol = M + 4
    inp = [[d.inp((c, r)) for r in range(6)] for c in range(ncol)]
    col4 = [d.sort([inp[c][r] for r in (1, 2, 3, 4)]) for c in range(ncol)]
    P, Q, core, rowwin = {}, {}, {}, {}

    def pair(a):
        if a not in P:
            P[a] = d.merge(col4[a], col4[a + 1])
        return P[a]

    def quad(a):
        if a not in Q:
            Q[a] = d.merge(pair(a), pair(a + 2))[3:13]
        return Q[a]

    def core6(o):
        if o not in core:
            m = d.merge(quad(o), col4[o + 4]) if o % 2 == parity else d.merge(quad(o + 1), col4[o])
            core[o] = m[4:10]
        return core[o]

    def rw(r, o):
        if (r, o) not in rowwin:
            b = o if o % 2 == rparity else o - 1
            if b >= 0 and b + 5 < ncol:
                # sort4 as two ordered pairs merged: the pair (b+3, b+4) is also the first pair of the next group (b+2)
                e = [inp[c][r] for c in range(b + 1, b + 5)]
                c4 = d.merge(d.sort(e[:2]), d.sort(e[2:]))
                rowwin[(r, b)] = d.merge([inp[b][r]], c4)
                rowwin[(r, b + 1)] = d.merge([inp[b + 5][r]], c4)
            else:
                rowwin[(r, o)] = d.sort([inp[c][r] for c in range(o, o + 5)])
        return rowwin[(r, o)]

    outs = [d.kth2(core6(o), rw(0, o), 6) for o in range(M)] + [d.kth2(core6(o), rw(5, o), 6) for o in range(M)]
    return d, inp, outs


def build3_2rows(M, parity):
    """3x3, two vertically adjacent output rows per call (window rows 0..2 and 1..3 of four input rows): the two shared
    rows of every column are ordered once, each output row inserts its own outer element (merge(1,2)), then the
    pairs scheme: P(a) = ranks 1..4 of two merged columns, median = 4th smallest of P(a) u the third column."""
    k = 3
    d = Dag()
    ncol = M + 2
    inp = [[d.inp((c, r)) for r in range(4)] for c in range(ncol)]
    mid = [d.sort([inp[c][1], inp[c][2]]) for c in range(ncol)]
    rows = [[d.merge([inp[c][0]], mid[c]) for c in range(ncol)], [d.merge([inp[c][3]], mid[c]) for c in range(ncol)]]
    outs = []
    for cols in rows:
        P = {}

        def pair(a, cols=cols, P=P):
            if a not in P:
                m = d.merge(cols[a], cols[a + 1])
                P[a] = m[1:5]                      # ranks that can still be the median of 9 (target rank 4, 3 more elements)
            return P[a]
        for o in range(M):
            acc, single = (pair(o), cols[o + 2]) if o % 2 == parity else (pair(o + 1), cols[o])
            outs.append(d.kth2(acc, single, 4))    # 4 - 1 discarded below = 3 -> 4th smallest (1-based)
    return d, inp, outs


def build_hier_2rows(k, M, parity, rparity=0):
    """Any odd k >= 5, two vertically adjacent output rows per call (window rows 0..k-1 and 1..k of k+1 input rows): the k-1
    middle rows are shared.  Their columns are sorted once, merged in pairs P(a) (a of the given parity) and, pair after pair,
    into H(a) -- after every merge only the ranks that can still hold the median of k*k are kept -- and joined with the
    remaining column: core(o) = the k+1 candidates among the k(k-1) shared elements of window o.  Each output then takes one
    rank of core(o) u its own sorted outer row window; the row windows of two horizontal neighbours share the sorted k-1
    elements they have in common (aligned ordered pairs merged, the pairs shared between groups) and insert one element each."""
    d = Dag()
    ncol = M + k - 1
    half = (k * k) // 2
    inp = [[d.inp((c, r)) for r in range(k + 1)] for c in range(ncol)]
    colc = [d.sort([inp[c][r] for r in range(1, k)]) for c in range(ncol)]
    P, H, core, rowwin = {}, {}, {}, {}

    def pair(a):
        if a not in P:
            P[a] = d.merge(colc[a], colc[a + 1])
        return P[a]

    def trim(lst, below, rest):
        t = half - below
        lo, hi = max(0, t - rest), min(len(lst) - 1, t)
        return lst[lo:hi + 1], below + lo

    def hexa(a):
        if a not in H:
            npairs = (k - 1) // 2
            acc, below, used = pair(a), 0, 2
            for i in range(1, npairs):
                used += 2
                acc, below = trim(d.merge(acc, pair(a + 2 * i)), below, (k - used) * (k - 1) + k)
            H[a] = (acc, below)
        return H[a]

    def core_of(o):
        if o not in core:
            (acc, below), single = (hexa(o), colc[o + k - 1]) if o % 2 == parity else (hexa(o + 1), colc[o])
            core[o] = trim(d.merge(acc, single), below, k)
        return core[o]

    def rw(r, o):
        if (r, o) not in rowwin:
            b = o if o % 2 == rparity else o - 1
            if b >= 0 and b + k < ncol:
                e = [inp[c][r] for c in range(b + 1, b + k)]            # the k-1 elements windows b and b+1 share
                acc = d.sort(e[:2])
                for i in range(2, k - 1, 2):
                    acc = d.merge(acc, d.sort(e[i:i + 2]))
                rowwin[(r, b)] = d.merge([inp[b][r]], acc)
                rowwin[(r, b + 1)] = d.merge([inp[b + k][r]], acc)
            else:
                rowwin[(r, o)] = d.sort([inp[c][r] for c in range(o, o + k)])
        return rowwin[(r, o)]

    outs = []
    for r in (0, k):
        for o in range(M):
            acc, below = core_of(o)
            outs.append(d.kth2(acc, rw(r, o), half - below + 1))
    return d, inp, outs


def verify_2rows_generic(d, inp, outs, k, M, zero_one=True, trials=20000):
    rng = np.random.RandomState(2)
    ncol = M + k - 1
    data = rng.randint(0, 256, (ncol, k + 1, trials)).astype(np.int32)
    data[:, :, : trials // 2] //= 64
    vals = {inp[c][r]: data[c, r] for c in range(ncol) for r in range(k + 1)}
    got = evaluate(d, inp, outs, vals)
    for half in range(2):
        for o in range(M):
            want = np.sort(data[o:o + k, half:half + k].reshape(k * k, trials), axis=0)[(k * k) // 2]
            if not np.array_equal(got[half * M + o], want):
                return False
    if zero_one:
        for half in range(2):
            sub = [[inp[c][half + r] for r in range(k)] for c in range(ncol)]
            if not verify_zero_one(d, sub, outs[half * M:(half + 1) * M], k, M):
                return False
    return True


def verify_threshold_random(d, inp_windows, outs, k, trials=60000, seed=5):
    """0-1 vectors at the threshold, for the networks that are too wide for the exhaustive 0-1 check (k = 7, 9).  A min/max
    network computes a monotone Boolean function; it is the median iff that function is 'at least (k*k+1)/2 ones', so the inputs
    that can expose a wrong network are the ones with exactly (k*k-1)/2 ones inside a window (must give 0) and exactly
    (k*k+1)/2 (must give 1).  inp_windows[o] = the k*k input nodes of output o; the other inputs are random bits."""
    rng = np.random.RandomState(seed)
    all_inputs = sorted({n for n in d.live(outs) if d.nodes[n][0] == "in"})
    kk = k * k
    for o, (win, out) in enumerate(zip(inp_windows, outs)):
        for ones in (kk // 2, kk // 2 + 1):
            vals = {n: rng.randint(0, 2, trials).astype(np.uint8) for n in all_inputs}
            order = np.argsort(rng.rand(trials, kk), axis=1)                 # a random subset of `ones` positions per trial
            bits = (order < ones).astype(np.uint8)
            for j, n in enumerate(win):
                vals[n] = bits[:, j]
            got = evaluate(d, None, [out], vals)[0]
            if not np.all(got == (1 if ones > kk // 2 else 0)):
                return False
    return True


def verify_2rows(d, inp, outs, M, zero_one=True, trials=20000):
    rng = np.random.RandomState(1)
    ncol = M + 4
    data = rng.randint(0, 256, (ncol, 6, trials)).astype(np.int32)
    data[:, :, : trials // 2] //= 64
    vals = {inp[c][r]: data[c, r] for c in range(ncol) for r in range(6)}
    got = evaluate(d, inp, outs, vals)
    for half in range(2):
        for o in range(M):
            want = np.sort(data[o:o + 5, half:half + 5].reshape(25, trials), axis=0)[12]
            if not np.array_equal(got[half * M + o], want):
                return False
    if not zero_one:
        return True
    # 0-1 principle, exhaustive per output: reuse verify_zero_one on a view where each output sees its own 25 inputs
    for half in range(2):
        sub_inputs = [[inp[c][half + r] for r in range(5)] for c in range(ncol)]
        if not verify_zero_one(d, sub_inputs, outs[half * M:(half + 1) * M], 5, M):
            return False
    return True


def emit_2rows(d, inp, outs, M, fh, k=5):
    live = d.live(outs)
    order = sorted(live)
    nops = sum(1 for n in order if d.nodes[n][0] != "in")
    ncol = M + k - 1
    partner = {}
    for n in order:
        op, a, b = d.nodes[n]
        if op == "min" and ("max", a, b) in d.memo and d.memo[("max", a, b)] in live:
            partner[n] = d.memo[("max", a, b)]
            partner[d.memo[("max", a, b)]] = n
    fh.write(f"/* k={k}, two output rows per call: 2 x {M} outputs from {ncol} columns x {k + 1} rows; {nops} packed min/max ops "
             f"({nops / (2 * M):.1f} per output pair-lane), {len(partner) // 2} full compare-exchanges */\n")
    fh.write(f"#define RV_MEDIAN{k}X2_M {M}\n#define RV_MEDIAN{k}X2_OPS {nops}\n")
    fh.write(f"__device__ __forceinline__ void rv_median{k}x2_net(const uint32_t (&v)[{ncol}][{k + 1}], uint32_t (&out)[2][{M}])\n{{\n")
    name, done, ce = {}, set(), 0
    for n in order:
        op, a, b = d.nodes[n]
        if op == "in":
            name[n] = f"v[{a[0]}][{a[1]}]"
            continue
        if n in done:
            continue
        if n in partner:
            lo, hi = (n, partner[n]) if op == "min" else (partner[n], n)
            name[lo], name[hi] = f"t{lo}", f"t{hi}"
            fh.write(f"    RV_CEX{k}X2({ce}, t{lo}, t{hi}, {name[a]}, {name[b]});\n")
            done.update((lo, hi))
            ce += 1
        else:
            name[n] = f"t{n}"
            fh.write(f"    const uint32_t t{n} = {'RV_MN' if op == 'min' else 'RV_MX'}({name[a]}, {name[b]});\n")
    for half in range(2):
        for o in range(M):
            fh.write(f"    out[{half}][{o}] = {name[outs[half * M + o]]};\n")
    fh.write("}\n\n")
    return nops


def evaluate(d, inputs, outs, values):
    """values: dict input-node -> numpy array (any dtype supporting minimum/maximum or bit ops)."""
    live = d.live(outs)
    val = {}
    for n in sorted(live):
        op, a, b = d.nodes[n]
        if op == "in":
            val[n] = values[n]
        elif op == "min":
            val[n] = np.minimum(val[a], val[b])
        else:
            val[n] = np.maximum(val[a], val[b])
    return [val[o] for o in outs]


def verify_random(d, inputs, outs, k, M, trials=20000, seed=0):
    rng = np.random.RandomState(seed)
    ncol = M + k - 1
    # mix of wide-range and few-distinct-values (ties) data
    data = rng.randint(0, 256, (ncol, k, trials)).astype(np.int32)
    data[:, :, : trials // 2] //= 64
    vals = {inputs[c][r]: data[c, r] for c in range(ncol) for r in range(k)}
    got = evaluate(d, inputs, outs, vals)
    for o in range(M):
        win = data[o:o + k].reshape(k * k, trials)
        want = np.sort(win, axis=0)[(k * k) // 2]
        if not np.array_equal(got[o], want):
            return False
    return True


def verify_zero_one(d, inputs, outs, k, M):
    """0-1 principle, exhaustive over the k*k inputs of each output (bit-parallel on uint64 words)."""
    n = k * k
    if n > 25:
        return None
    nbits = 1 << n
    words = nbits // 64
    # input i takes bit i of the case index
    idx = np.arange(words, dtype=np.uint64)
    pats = []
    for i in range(n):
        if i < 6:
            base = np.uint64(sum(((j >> i) & 1) << j for j in range(64)))
            pats.append(np.full(words, base, np.uint64))
        else:
            pats.append(np.where((idx >> np.uint64(i - 6)) & np.uint64(1), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0)))
    # popcount of case index >= half+1  <=> median is 1
    pc = np.zeros(nbits, np.uint8)
    ar = np.arange(nbits, dtype=np.uint32)
    for i in range(n):
        pc += ((ar >> i) & 1).astype(np.uint8)
    want_bits = (pc >= (n // 2 + 1))
    want = np.packbits(want_bits, bitorder="little").view(np.uint64)
    for o in range(M):
        live = d.live([outs[o]])
        val = {}
        vi = 0
        mapping = {}
        for c in range(o, o + k):
            for r in range(k):
                mapping[inputs[c][r]] = pats[vi]; vi += 1
        for nid in sorted(live):
            op, a, b = d.nodes[nid]
            if op == "in":
                val[nid] = mapping[nid]
            elif op == "min":
                val[nid] = val[a] & val[b]
            else:
                val[nid] = val[a] | val[b]
        if not np.array_equal(val[outs[o]], want):
            return False
    return True


def emit(d, inputs, outs, k, M, fh):
    """Straight-line code. A min and a max of the same operand pair that are both live form a
    compare-exchange and are emitted through RV_CEX(n, lo, hi, a, b) so the kernel can choose, per
    compile-time index n, between the ALU form (2 x VIMNMX.U16x2) and the FMA-pipe form
    (HFMA2.RELU + 2 x HADD2 on 0x6400|v half-floats); lone mins / maxes stay on the ALU pipe."""
    live = d.live(outs)
    order = [n for n in sorted(live)]
    nops = sum(1 for n in order if d.nodes[n][0] != "in")
    ncol = M + k - 1
    partner = {}
    for n in order:
        op, a, b = d.nodes[n]
        if op == "min" and ("max", a, b) in d.memo and d.memo[("max", a, b)] in live:
            partner[n] = d.memo[("max", a, b)]
            partner[d.memo[("max", a, b)]] = n
    nce = len(partner) // 2
    fh.write(f"/* k={k}: {M} outputs from {ncol} columns; {nops} packed min/max ops "
             f"({nops / M:.1f} per output pair-lane), of which {nce} full compare-exchanges */\n")
    fh.write(f"#define RV_MEDIAN{k}_M {M}\n#define RV_MEDIAN{k}_OPS {nops}\n#define RV_MEDIAN{k}_CES {nce}\n")
    fh.write(f"__device__ __forceinline__ void rv_median{k}_net(const uint32_t (&v)[{ncol}][{k}], uint32_t (&out)[{M}])\n{{\n")
    name, done, ce = {}, set(), 0
    for n in order:
        op, a, b = d.nodes[n]
        if op == "in":
            name[n] = f"v[{a[0]}][{a[1]}]"
            continue
        if n in done:
            continue
        if n in partner:
            lo, hi = (n, partner[n]) if op == "min" else (partner[n], n)
            name[lo], name[hi] = f"t{lo}", f"t{hi}"
            fh.write(f"    RV_CEX{k}({ce}, t{lo}, t{hi}, {name[a]}, {name[b]});\n")
            done.update((lo, hi))
            ce += 1
        else:
            name[n] = f"t{n}"
            fn = "RV_MN" if op == "min" else "RV_MX"
            fh.write(f"    const uint32_t t{n} = {fn}({name[a]}, {name[b]});\n")
    for o in range(M):
        fh.write(f"    out[{o}] = {name[outs[o]]};\n")
    fh.write("}\n\n")
    return nops


CONFIG = {3: 4, 5: int(os.environ.get('RV_MEDIAN5_M', '6')), 7: int(os.environ.get('RV_MEDIAN7_M', '4')),
          9: int(os.environ.get('RV_MEDIAN9_M', '2'))}     # k -> outputs per call


def main():
    for n, net in SORTERS.items():
        assert check_sorter(n, net), f"sorter {n} is wrong"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.environ.get("RV_MEDIAN_NET_OUT") or os.path.join(root, "road-vision-system_b200", "csrc", "rv_median_net.h")
    quick = "--quick" in sys.argv
    with open(path, "w") as fh:
        fh.write("/* GENERATED by tools/gen_median_net.py -- do not edit.\n"
                 " * Selection networks for the k x k median (cv2.medianBlur semantics) on packed u16x2 lanes.\n"
                 " * v[c][r]: column c (pixel x0-k/2+c of one channel plane), row r of the window rows; any order per column.\n"
                 " * out[o]: median of columns o..o+k-1.  Verified by the generator (0-1 principle for k<=5, random for all). */\n"
                 "#ifndef RV_MEDIAN_NET_H\n#define RV_MEDIAN_NET_H\n#include <stdint.h>\n"
                 "#ifndef RV_MN\n#define RV_MN(a, b) __vminu2((a), (b))\n#define RV_MX(a, b) __vmaxu2((a), (b))\n#endif\n"
                 "#ifndef RV_CEX\n#define RV_CEX(n, lo, hi, a, b) const uint32_t lo = RV_MN(a, b), hi = RV_MX(a, b)\n#endif\n"
                 "/* per-network compare-exchange macros (the kernel may give each network its own ALU/FMA mix) */\n"
                 "#ifndef RV_CEX3\n#define RV_CEX3 RV_CEX\n#endif\n#ifndef RV_CEX5\n#define RV_CEX5 RV_CEX\n#endif\n"
                 "#ifndef RV_CEX7\n#define RV_CEX7 RV_CEX\n#endif\n#ifndef RV_CEX9\n#define RV_CEX9 RV_CEX\n#endif\n"
                 "#ifndef RV_CEX5X2\n#define RV_CEX5X2 RV_CEX\n#endif\n#ifndef RV_CEX3X2\n#define RV_CEX3X2 RV_CEX\n#endif\n"
                 "#ifndef RV_CEX7X2\n#define RV_CEX7X2 RV_CEX\n#endif\n#ifndef RV_CEX9X2\n#define RV_CEX9X2 RV_CEX\n#endif\n\n")
        for k, M in CONFIG.items():
            best = None
            for strat, parity in (("pairs", 0), ("pairs", 1), ("flat", 0), ("quads", 0), ("quads", 1), ("hier", 0), ("hier", 1)):
                if strat == "hier":
                    d, inputs, outs = build_hier(k, M, parity)
                elif strat == "quads":
                    if k != 5:
                        continue
                    d, inputs, outs = build5_quads(M, parity)
                else:
                    d, inputs, outs = build(k, M, strat, parity)
                live = d.live(outs)
                nops = sum(1 for n in live if d.nodes[n][0] != "in")
                if best is None or nops < best[0]:
                    best = (nops, f"{strat}/parity{parity}", d, inputs, outs)
            nops, strat, d, inputs, outs = best
            assert verify_random(d, inputs, outs, k, M), f"k={k}: random verification failed"
            if k > 5:
                wins = [[inputs[c][r] for c in range(o, o + k) for r in range(k)] for o in range(M)]
                assert verify_threshold_random(d, wins, outs, k, trials=20000 if quick else 60000), f"k={k}: threshold 0-1 vectors failed"
            if not quick:
                z = verify_zero_one(d, inputs, outs, k, M)
                assert z in (True, None), f"k={k}: 0-1 verification failed"
            else:
                z = "skipped"
            emit(d, inputs, outs, k, M, fh)
            print(f"k={k} M={M} strategy={strat}: {nops} ops ({nops / M:.1f}/output), zero-one={z}")
        M2 = int(os.environ.get("RV_MEDIAN5X2_M", "6"))
        best = None
        for parity in (0, 1):
            for rpar in (0, 1):
                d, inp, outs = build5_2rows(M2, parity, rpar)
                nops = sum(1 for n in d.live(outs) if d.nodes[n][0] != "in")
                if best is None or nops < best[0]:
                    best = (nops, parity, rpar, d, inp, outs)
        nops, parity, rpar, d, inp, outs = best
        assert verify_2rows(d, inp, outs, M2, zero_one=not quick), "k=5 two-row network failed verification"
        emit_2rows(d, inp, outs, M2, fh)
        print(f"k=5 two rows M={M2} parity={parity}/{rpar}: {nops} ops ({nops / (2 * M2):.1f}/output), zero-one={'skipped' if quick else True}")
        M3 = int(os.environ.get("RV_MEDIAN3X2_M", "6"))
        best = None
        for parity in (0, 1):
            d, inp, outs = build3_2rows(M3, parity)
            nops = sum(1 for n in d.live(outs) if d.nodes[n][0] != "in")
            if best is None or nops < best[0]:
                best = (nops, parity, d, inp, outs)
        nops, parity, d, inp, outs = best
        assert verify_2rows_generic(d, inp, outs, 3, M3, zero_one=not quick), "k=3 two-row network failed verification"
        emit_2rows(d, inp, outs, M3, fh, k=3)
        print(f"k=3 two rows M={M3} parity={parity}: {nops} ops ({nops / (2 * M3):.1f}/output), zero-one={'skipped' if quick else True}")
        for k in (7, 9):
            Mk = int(os.environ.get(f"RV_MEDIAN{k}X2_M", {7: "6", 9: "4"}[k]))       # measured: profiles/r2_y_median79_two_rows.txt
            best = None
            for parity in (0, 1):
                for rpar in (0, 1):
                    d, inp, outs = build_hier_2rows(k, Mk, parity, rpar)
                    nops = sum(1 for n in d.live(outs) if d.nodes[n][0] != "in")
                    if best is None or nops < best[0]:
                        best = (nops, parity, rpar, d, inp, outs)
            nops, parity, rpar, d, inp, outs = best
            assert verify_2rows_generic(d, inp, outs, k, Mk, zero_one=False), f"k={k} two-row network failed verification"
            wins = [[inp[c][hrow + r] for c in range(o, o + k) for r in range(k)] for hrow in (0, 1) for o in range(Mk)]
            assert verify_threshold_random(d, wins, outs, k, trials=20000 if quick else 60000), f"k={k} two rows: threshold 0-1 vectors failed"
            emit_2rows(d, inp, outs, Mk, fh, k=k)
            print(f"k={k} two rows M={Mk} parity={parity}/{rpar}: {nops} ops ({nops / (2 * Mk):.1f}/output), random vectors + 0-1 vectors at the threshold")
        fh.write("#endif\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
