"""Deterministic Louvain specification (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Stands in for ``louvain.find_partition(g, RBConfigurationVertexPartition,
resolution_parameter=gamma, seed=random_state)`` which ``sc.tl.louvain`` calls from
``/root/reference/doubletdetection/doubletdetection.py:337-342``.  The
``louvain`` / ``igraph`` packages are not in the image, so their exact move order cannot be
reproduced: PARITY UNPINNED for this stage.  What is specified here, and what the product's C++
implementation (``doubletdetection_b200/csrc/louvain.cpp``) must reproduce label-for-label:

Quality function (RB configuration, resolution gamma, undirected, weights w):
    Q = sum_ij (A_ij - gamma * k_i * k_j / (2m)) * delta(c_i, c_j)

Two flavours of the first level exist (``level0=`` argument of :func:`louvain`):

* ``"sequential"`` -- the classic queue-driven sweep described below, used for every level above the first and
  for explicit (possibly weighted) graphs (``dd_louvain_csr``);
* ``"parallel"`` -- what the kNN pipeline uses (``dd_louvain_knn`` / ``dd_fit_iterations``): the first level of an
  UNWEIGHTED graph is optimised by synchronous coloured rounds (:func:`level0_parallel`) so that it can run on
  the GPU; the communities it finds are aggregated and the remaining levels are sequential.

Synchronous coloured rounds (first level, unweighted, order-independent by construction):
  * colour(i) = SplitMix64-finaliser(seed + (i + 1) * 0x9E3779B97F4A7C15) mod 8;
  * a round visits the colours 0..7; all nodes of the current colour decide simultaneously from the state at
    the start of the sub-round (comm, tot, community sizes): with w(i, c) = number of neighbours in c and
    tot'(c) = tot[c] - (k_i if c == comm[i] else 0),
        gain(c) = w(i, c) - ((gamma * k_i) * tot'(c)) / two_m            (IEEE double, this order)
    the node moves to the neighbouring community c != comm[i] with the largest gain (ties: smallest id) iff
    that gain is strictly larger than gain(comm[i]) -- except that a singleton never moves into a singleton
    with a larger id (this breaks the swap oscillation of synchronous updates);
  * all moves of the sub-round are applied at once; the level ends after a round that moved at most
    floor(n / 512) nodes (i.e. no node for n < 512), or after 32 rounds;
  * WEIGHTED graphs (``louvain(..., weights, level0="parallel")``; not yet used by the product): the
    level works on fixed-point weights ``wq = rint(w * 2**32)`` held in int64, so that w(i, c), k_i, tot[c] and two_m are
    exact integer sums -- the same in any summation order, which is what makes simultaneous updates (atomics on a
    GPU) bit-reproducible.  The gain is the expression above evaluated in double on those integers converted to double.
    The levels above the first use the original double weights.

Sequential algorithm (one "level"):
  * every node starts in its own community; ``tot[c]`` = sum of weighted degrees in c;
  * a FIFO queue is filled with all nodes in a seeded random order (SplitMix64 + Fisher-Yates,
    ``j = next() % (i + 1)`` for i = n-1 .. 1);
  * pop node i, remove it from its community, evaluate for its own community first and then for
    every neighbouring community in order of first appearance in i's adjacency list
        gain(c) = w(i, c) - ((gamma * k_i) * tot[c]) / two_m          (IEEE double, this order)
    and move to the FIRST community with the strictly largest gain (staying wins ties);
  * when i moves, every neighbour that is not in i's new community and not queued is appended;
  * the level ends when the queue is empty.
Levels: communities are renumbered by first appearance over node index, the graph is aggregated
(neighbour lists ascending by id, internal weight kept as a self-loop counted twice), and the
procedure repeats until a level moves no node.  Final labels are renumbered by decreasing
community size (ties: smaller first-appearance id first), which is what louvain-igraph's
``renumber_communities`` does.
"""

from collections import deque

import numpy as np

_MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = int(seed) & _MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)


def _permutation(n, rng):
    p = list(range(n))
    for i in range(n - 1, 0, -1):
        j = rng.next() % (i + 1)
        p[i], p[j] = p[j], p[i]
    return p


def _one_level(n, indptr, indices, weights, selfw, gamma, two_m, rng):
    k = [0.0] * n
    for i in range(n):
        s = selfw[i]
        for e in range(indptr[i], indptr[i + 1]):
            s += weights[e]
        k[i] = s
    comm = list(range(n))
    tot = list(k)
    queue = deque(_permutation(n, rng))
    inq = [True] * n
    neigh_w = [0.0] * n
    seen = [False] * n
    moved_any = False
    while queue:
        i = queue.popleft()
        inq[i] = False
        ci = comm[i]
        ki = k[i]
        cands = [ci]
        seen[ci] = True
        neigh_w[ci] = 0.0
        for e in range(indptr[i], indptr[i + 1]):
            c = comm[indices[e]]
            if not seen[c]:
                seen[c] = True
                neigh_w[c] = 0.0
                cands.append(c)
            neigh_w[c] += weights[e]
        tot[ci] -= ki
        best = ci
        best_gain = neigh_w[ci] - ((gamma * ki) * tot[ci]) / two_m
        for c in cands[1:]:
            g = neigh_w[c] - ((gamma * ki) * tot[c]) / two_m
            if g > best_gain:
                best = c
                best_gain = g
        for c in cands:
            seen[c] = False
        tot[best] += ki
        if best != ci:
            comm[i] = best
            moved_any = True
            for e in range(indptr[i], indptr[i + 1]):
                j = indices[e]
                if comm[j] != best and not inq[j]:
                    inq[j] = True
                    queue.append(j)
    return comm, moved_any


def _aggregate(n, indptr, indices, weights, selfw, comm):
    new_id = {}
    for i in range(n):
        c = comm[i]
        if c not in new_id:
            new_id[c] = len(new_id)
    nc = len(new_id)
    node2new = [new_id[comm[i]] for i in range(n)]
    members = [[] for _ in range(nc)]
    for i in range(n):
        members[node2new[i]].append(i)
    n_indptr = [0]
    n_indices = []
    n_weights = []
    n_selfw = [0.0] * nc
    for a in range(nc):
        acc = {}
        s = 0.0
        for i in members[a]:
            s += selfw[i]
            for e in range(indptr[i], indptr[i + 1]):
                b = node2new[indices[e]]
                if b == a:
                    s += weights[e]
                else:
                    acc[b] = acc.get(b, 0.0) + weights[e]
        n_selfw[a] = s
        for b in sorted(acc):
            n_indices.append(b)
            n_weights.append(acc[b])
        n_indptr.append(len(n_indices))
    return nc, n_indptr, n_indices, n_weights, n_selfw, node2new


N_COLOURS = 8
MAX_ROUNDS = 32


def node_colours(n, seed):
    """colour(i) of the synchronous rounds (uint64 arithmetic, vectorised)."""
    with np.errstate(over="ignore"):
        z = np.uint64(int(seed) & _MASK) + (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(N_COLOURS)).astype(np.int64)


FIXED_POINT = 2.0**32


def level0_parallel(indptr, indices, gamma, seed, weights=None):
    """First level by synchronous coloured rounds on a symmetric graph (see the module docstring); unweighted, or with
    ``weights`` quantised to multiples of 2**-32.  Returns the community id (a node id) of every node."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = indptr.size - 1
    deg = np.diff(indptr)
    if weights is None:
        wq = np.ones(indices.size, dtype=np.int64)
    else:
        wq = np.rint(np.asarray(weights, dtype=np.float64) * FIXED_POINT).astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    k = np.zeros(n, dtype=np.int64)
    np.add.at(k, rows, wq)
    two_m = float(int(wq.sum()))
    comm = np.arange(n, dtype=np.int64)
    tot = k.copy()
    size = np.ones(n, dtype=np.int64)
    if two_m == 0.0:
        return comm
    colour = node_colours(n, seed)
    by_colour = [np.nonzero((colour == c) & (deg > 0))[0] for c in range(N_COLOURS)]
    gamma = float(gamma)
    for _ in range(MAX_ROUNDS):
        moved = 0
        for act in by_colour:
            if act.size == 0:
                continue
            d = deg[act]
            local = np.repeat(np.arange(act.size, dtype=np.int64), d)
            starts = np.repeat(indptr[act], d)
            offs = np.arange(d.sum(), dtype=np.int64) - np.repeat(np.cumsum(d) - d, d)
            edge = starts + offs
            c_nb = comm[indices[edge]]
            key, inv = np.unique(local * n + c_nb, return_inverse=True)
            w = np.zeros(key.size, dtype=np.int64)
            np.add.at(w, inv.ravel(), wq[edge])  # exact integer sums
            a = key // n  # local index of the node
            c = key % n  # candidate community
            node = act[a]
            ci = comm[node]
            ki = k[node].astype(np.float64)
            own = c == ci
            gain = w.astype(np.float64) - ((gamma * ki) * (tot[c] - np.where(own, k[node], 0)).astype(np.float64)) / two_m
            # gain of staying (w(i, ci) may be 0: then the pair is absent from `key`)
            w_stay = np.zeros(act.size, dtype=np.float64)
            w_stay[a[own]] = w[own].astype(np.float64)
            ci_act = comm[act]
            k_act = k[act].astype(np.float64)
            gain_stay = w_stay - ((gamma * k_act) * (tot[ci_act] - k[act]).astype(np.float64)) / two_m
            # best other community: largest gain, ties -> smallest id
            oth = ~own
            if not oth.any():
                continue
            ao, co, go = a[oth], c[oth], gain[oth]
            order = np.lexsort((co, -go, ao))
            ao, co, go = ao[order], co[order], go[order]
            first = np.ones(ao.size, dtype=bool)
            first[1:] = ao[1:] != ao[:-1]
            ab, cb, gb = ao[first], co[first], go[first]
            src = ci_act[ab]
            move = gb > gain_stay[ab]
            move &= ~((size[src] == 1) & (size[cb] == 1) & (cb > src))
            if not move.any():
                continue
            mv_node, mv_to, mv_from = act[ab[move]], cb[move], src[move]
            comm[mv_node] = mv_to
            np.add.at(tot, mv_from, -k[mv_node])
            np.add.at(tot, mv_to, k[mv_node])
            np.add.at(size, mv_from, -1)
            np.add.at(size, mv_to, 1)
            moved += int(mv_node.size)
        if moved <= (n >> 9):  # at most n / 512 moves: the level is settled
            break
    return comm


def louvain(indptr, indices, weights=None, resolution=1.0, seed=0, max_levels=64, level0="sequential"):
    """Cluster a symmetric, self-loop-free CSR graph.  Returns int64 labels, 0 = largest.
    ``level0="parallel"`` is the kNN-pipeline flavour (weights, if any, in fixed point during that level)."""
    comm0 = None
    if level0 == "parallel":
        comm0 = [int(x) for x in level0_parallel(indptr, indices, resolution, seed, weights)]
    indptr = [int(x) for x in indptr]
    indices = [int(x) for x in indices]
    n = len(indptr) - 1
    if weights is None:
        weights = [1.0] * len(indices)
    else:
        weights = [float(x) for x in weights]
    gamma = float(resolution)
    selfw = [0.0] * n
    two_m = 0.0
    for w in weights:
        two_m += w
    rng = SplitMix64(seed)
    membership = list(range(n))
    if two_m > 0.0 and comm0 is not None:
        nc, indptr, indices, weights, selfw, node2new = _aggregate(n, indptr, indices, weights, selfw, comm0)
        membership = [node2new[c] for c in membership]
    if two_m > 0.0:
        for _ in range(max_levels):
            comm, moved = _one_level(len(indptr) - 1, indptr, indices, weights, selfw, gamma, two_m, rng)
            if not moved:
                break
            nc, indptr, indices, weights, selfw, node2new = _aggregate(
                len(indptr) - 1, indptr, indices, weights, selfw, comm
            )
            membership = [node2new[c] for c in membership]
    return relabel_by_size(np.asarray(membership, dtype=np.int64))


def relabel_by_size(membership):
    """Renumber by decreasing size; ties keep first-appearance order."""
    membership = np.asarray(membership, dtype=np.int64)
    if membership.size == 0:
        return membership
    _, first_idx, inv = np.unique(membership, return_index=True, return_inverse=True)
    # ids by first appearance
    order_fa = np.argsort(first_idx, kind="stable")
    rank_fa = np.empty_like(order_fa)
    rank_fa[order_fa] = np.arange(order_fa.size)
    fa = rank_fa[inv]
    sizes = np.bincount(fa)
    order = np.argsort(-sizes, kind="stable")
    new = np.empty_like(order)
    new[order] = np.arange(order.size)
    return new[fa].astype(np.int64)
