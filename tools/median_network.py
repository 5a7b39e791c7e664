"""Search for a small min/max selection network: median of 25 from five pre-sorted columns of five.

Wires are numbered w = 5*col + rank (rank 0 = smallest of its column).  By the 0-1 principle restricted
to monotone-closed input sets, a min/max network selects rank 12 for every input whose columns are
sorted iff it does so for every 0-1 input whose columns are sorted: 6^5 = 7776 vectors, evaluated
bit-parallel on Python big ints.  Start: full row sorts + a sorting network on the 13 candidates that
survive the row/column dominance argument; then greedy pruning, dead-half elimination and random
restarts.  Output: an op list  (dst, kind, a, b)  with kind in {min, max}.
"""
from __future__ import annotations

import itertools
import random
import sys

N = 25
PATS = list(itertools.product(range(6), repeat=5))  # ones per column
NV = len(PATS)
MASK = (1 << NV) - 1


def input_wires():
    w = [0] * N
    for v, pat in enumerate(PATS):
        for c, ones in enumerate(pat):
            for r in range(5):
                if r >= 5 - ones:  # sorted ascending: the top `ones` ranks are 1
                    w[5 * c + r] |= 1 << v
    want = 0
    for v, pat in enumerate(PATS):
        if sum(pat) >= 13:
            want |= 1 << v
    return w, want


W0, WANT = input_wires()


def run(net, out_wire):
    w = list(W0)
    for i, j in net:  # compare-exchange: min -> i, max -> j
        a, b = w[i], w[j]
        w[i], w[j] = a & b, a | b
    return w[out_wire] == WANT


SORT5 = [(0, 1), (3, 4), (2, 4), (2, 3), (0, 3), (0, 2), (1, 4), (1, 3), (1, 2)]


def oddeven_merge_sort(n):
    """Batcher's odd-even merge sort for n = power of two; returns comparator list."""
    net = []

    def merge(lo, n_, r):
        step = r * 2
        if step < n_:
            merge(lo, n_, step)
            merge(lo + r, n_, step)
            for i in range(lo + r, lo + n_ - r, step):
                net.append((i, i + r))
        else:
            net.append((lo, lo + r))

    def sort(lo, n_):
        if n_ > 1:
            m = n_ // 2
            sort(lo, m)
            sort(lo + m, m)
            merge(lo, n_, 1)

    sort(0, n)
    return net


def initial_network():
    net = []
    for r in range(5):  # sort row r across the columns: wires 5*c + r
        net += [(5 * a + r, 5 * b + r) for a, b in SORT5]
    cand = []
    for r, bs in enumerate([(3, 4), (2, 3, 4), (1, 2, 3), (0, 1, 2), (0, 1)]):
        cand += [5 * b + r for b in bs]
    assert len(cand) == 13
    # sort the 13 candidates with a 16-wire Batcher network (3 virtual +inf wires dropped), take rank 6
    bat = [(a, b) for a, b in oddeven_merge_sort(16) if a < 13 and b < 13]
    net += [(cand[a], cand[b]) for a, b in bat]
    return net, cand[6]


def oddeven_merge_lists(a, b):
    """Batcher odd-even merge of two sorted wire lists (any lengths); returns (comparators, merged wire order)."""
    if not a:
        return [], list(b)
    if not b:
        return [], list(a)
    if len(a) == 1 and len(b) == 1:
        return [(a[0], b[0])], [a[0], b[0]]
    ce, ev = oddeven_merge_lists(a[0::2], b[0::2])
    co, od = oddeven_merge_lists(a[1::2], b[1::2])
    net = ce + co
    out = [ev[0]]
    i = 1
    j = 0
    # interleave: compare od[j] with ev[i]
    while i < len(ev) and j < len(od):
        net.append((od[j], ev[i]))
        out += [od[j], ev[i]]
        i += 1
        j += 1
    out += ev[i:] + od[j:]
    return net, out


def merge_tree_network():
    """merge the five sorted columns pairwise with odd-even merges, take rank 12 of the final order"""
    cols = [[5 * c + r for r in range(5)] for c in range(5)]
    n1, s01 = oddeven_merge_lists(cols[0], cols[1])
    n2, s23 = oddeven_merge_lists(cols[2], cols[3])
    n3, s0123 = oddeven_merge_lists(s01, s23)
    n4, sall = oddeven_merge_lists(s0123, cols[4])
    return n1 + n2 + n3 + n4, sall[12]


def prune(net, out, rng):
    """remove comparators (random order) while the network stays correct"""
    net = list(net)
    changed = True
    while changed:
        changed = False
        order = list(range(len(net)))
        rng.shuffle(order)
        for k in sorted(order, reverse=True):
            trial = net[:k] + net[k + 1:]
            if run(trial, out):
                net = trial
                changed = True
    return net


def to_ops(net, out):
    """dead-code elimination at min/max granularity: returns ops [(kind, i, j)] writing wire i (min) or j (max)"""
    live = {out}
    ops = []
    for i, j in reversed(net):
        need_min, need_max = i in live, j in live
        if not (need_min or need_max):
            continue
        if need_min:
            ops.append(("min", i, j))
        if need_max:
            ops.append(("max", i, j))
        live.discard(i) if need_min else None
        live.discard(j) if need_max else None
        # inputs of this CE are both needed
        live.add(i)
        live.add(j)
    ops.reverse()
    return ops


def count_ops(net, out):
    return len(to_ops(net, out))


def mutate(net, out, rng):
    """swap two adjacent independent comparators / re-target one comparator, keep if correct"""
    net = list(net)
    k = rng.randrange(len(net))
    i, j = net[k]
    choice = rng.random()
    if choice < 0.5:
        a = rng.randrange(N)
        b = rng.randrange(N)
        if a == b:
            return None
        net[k] = (a, b)
    else:
        a, b = rng.randrange(N), rng.randrange(N)
        if a == b:
            return None
        net.insert(rng.randrange(len(net) + 1), (a, b))
    return net if run(net, out) else None


def search(seconds, seed):
    import time

    rng = random.Random(seed)
    net, out = merge_tree_network() if seed % 2 else initial_network()
    assert run(net, out), "initial network wrong"
    best = prune(net, out, rng)
    best_cost = count_ops(best, out)
    t_end = time.time() + seconds
    cur, cur_cost = best, best_cost
    it = 0
    while time.time() < t_end:
        it += 1
        cand = mutate(cur, out, rng)
        if cand is None:
            continue
        cand = prune(cand, out, rng)
        c = count_ops(cand, out)
        if c <= cur_cost:
            if c < cur_cost:
                print(f"  it {it}: ops {c} (CE {len(cand)})", file=sys.stderr)
            cur, cur_cost = cand, c
            if c < best_cost:
                best, best_cost = cand, c
    return best, out, best_cost


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    net, out, cost = search(secs, seed)
    print("CE", len(net), "ops", cost, "out", out)
    print("NET =", net)
