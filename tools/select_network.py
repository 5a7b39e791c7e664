"""Search for small compare-exchange networks on inputs made of pre-sorted lists.

    python tools/select_network.py --lists 10,10,5 --ranks 12 [--seconds 120] [--seed 0] [--procs 8]
    python tools/select_network.py --lists 5,5 --ranks all

Wires are numbered list by list (list 0 first), ascending inside a list.  By the 0-1 principle restricted to
monotone-closed input sets, a min/max network puts rank r on its output wire for every input whose lists are sorted
iff it does so for every 0-1 input whose lists are sorted: prod(len + 1) vectors, evaluated bit-parallel on Python
ints.  `--ranks all` asks for a full merge (output rank r on the wire the start network leaves it on); otherwise the
listed ranks (0-based) must come out sorted on some wires.  Start network: Batcher odd-even merges of the lists;
then random comparator removal, dead-half elimination, and mutate-and-prune local search.  Cost = min/max ops
(half compare-exchanges) that are live.
"""
from __future__ import annotations

import argparse
import itertools
import random
import sys
import time


def oddeven_merge_lists(a, b):
    """Batcher odd-even merge of two sorted wire lists (any lengths): (comparators, merged wire order)."""
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
    i, j = 1, 0
    while i < len(ev) and j < len(od):
        net.append((od[j], ev[i]))
        out += [od[j], ev[i]]
        i += 1
        j += 1
    out += ev[i:] + od[j:]
    return net, out


class Problem:
    def __init__(self, sizes, ranks):
        self.sizes = sizes
        self.n = sum(sizes)
        self.base = [sum(sizes[:k]) for k in range(len(sizes))]
        pats = list(itertools.product(*[range(s + 1) for s in sizes]))  # ones per list
        self.nv = len(pats)
        w = [0] * self.n
        for v, pat in enumerate(pats):
            for k, ones in enumerate(pat):
                for r in range(sizes[k] - ones, sizes[k]):
                    w[self.base[k] + r] |= 1 << v
        self.w0 = w
        self.ranks = ranks
        # rank r (0-based) of a 0-1 vector with `tot` ones among n wires is 1 iff r >= n - tot
        self.want = {}
        for r in ranks:
            m = 0
            for v, pat in enumerate(pats):
                if r >= self.n - sum(pat):
                    m |= 1 << v
            self.want[r] = m

    def start(self):
        lists = [[self.base[k] + r for r in range(s)] for k, s in enumerate(self.sizes)]
        net = []
        # merge the two longest first, then the rest in order
        order = sorted(range(len(lists)), key=lambda k: -len(lists[k]))
        cur = lists[order[0]]
        for k in order[1:]:
            n_, cur = oddeven_merge_lists(cur, lists[k])
            net += n_
        outs = {r: cur[r] for r in self.ranks}
        return net, outs

    def run(self, net, outs):
        w = list(self.w0)
        for i, j in net:
            a, b = w[i], w[j]
            w[i], w[j] = a & b, a | b
        return all(w[outs[r]] == self.want[r] for r in self.ranks)

    def ops(self, net, outs):
        live = set(outs.values())
        n = 0
        kept = []
        for i, j in reversed(net):
            nmin, nmax = i in live, j in live
            if not (nmin or nmax):
                continue
            n += nmin + nmax
            kept.append((i, j))
            live.add(i)
            live.add(j)
        kept.reverse()
        return n, kept


def prune(pb, net, outs, rng):
    net = list(net)
    changed = True
    while changed:
        changed = False
        order = list(range(len(net)))
        rng.shuffle(order)
        for k in sorted(order, reverse=True):
            trial = net[:k] + net[k + 1:]
            if pb.run(trial, outs):
                net = trial
                changed = True
    return pb.ops(net, outs)[1]


def mutate(pb, net, outs, rng):
    net = list(net)
    c = rng.random()
    a, b = rng.randrange(pb.n), rng.randrange(pb.n)
    if a == b:
        return None
    if c < 0.4 and net:
        net[rng.randrange(len(net))] = (a, b)
    elif c < 0.8:
        net.insert(rng.randrange(len(net) + 1), (a, b))
    else:
        if len(net) < 2:
            return None
        k = rng.randrange(len(net) - 1)
        net[k], net[k + 1] = net[k + 1], net[k]
    return net if pb.run(net, outs) else None


def search(pb, seconds, seed, verbose=False):
    rng = random.Random(seed)
    net, outs = pb.start()
    assert pb.run(net, outs), "start network wrong"
    cur = prune(pb, net, outs, rng)
    cur_cost = pb.ops(cur, outs)[0]
    best, best_cost = cur, cur_cost
    t_end = time.time() + seconds
    while time.time() < t_end:
        cand = mutate(pb, cur, outs, rng)
        if cand is None:
            continue
        cand = prune(pb, cand, outs, rng)
        c = pb.ops(cand, outs)[0]
        if c <= cur_cost:
            if c < cur_cost and verbose:
                print(f"  seed {seed}: ops {c} (CE {len(cand)})", file=sys.stderr, flush=True)
            cur, cur_cost = cand, c
            if c < best_cost:
                best, best_cost = cand, c
    return best_cost, len(best), best, outs


def _worker(args):
    sizes, ranks, seconds, seed = args
    return search(Problem(sizes, ranks), seconds, seed)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--lists", required=True)
    ap.add_argument("--ranks", required=True)
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--procs", type=int, default=1)
    a = ap.parse_args()
    sizes = [int(x) for x in a.lists.split(",")]
    ranks = list(range(sum(sizes))) if a.ranks == "all" else [int(x) for x in a.ranks.split(",")]
    pb = Problem(sizes, ranks)
    net0, outs0 = pb.start()
    print(f"vectors {pb.nv}, start network: CE {len(net0)}, ops {pb.ops(net0, outs0)[0]}", file=sys.stderr)
    if a.procs > 1:
        import multiprocessing as mp

        with mp.Pool(a.procs) as pool:
            res = pool.map(_worker, [(sizes, ranks, a.seconds, a.seed + k) for k in range(a.procs)])
    else:
        res = [search(pb, a.seconds, a.seed, verbose=True)]
    cost, nce, net, outs = min(res, key=lambda t: (t[0], t[1]))
    print("all seeds:", sorted((r[0], r[1]) for r in res), file=sys.stderr)
    print(f"# lists {sizes} ranks {ranks}: {nce} compare-exchanges, {cost} min/max ops")
    print("OUTS =", outs)
    print("NET =", net)
