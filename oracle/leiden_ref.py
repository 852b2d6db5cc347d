"""Deterministic Leiden specification (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Stands in for ``leidenalg.find_partition(g, RBConfigurationVertexPartition, weights=...,
resolution_parameter=gamma, n_iterations=-1, seed=random_state)`` which ``sc.tl.leiden`` calls from
``/root/reference/doubletdetection/doubletdetection.py:340-342`` on the UMAP-weighted neighbour graph
(``use_weights=True``).  ``leidenalg`` / ``igraph`` are not in the image, so their
exact move order and random stream cannot be reproduced: PARITY UNPINNED for this stage (the reference
itself warns that "Leiden clustering is experimental and results have not been validated", :105-106).
What is specified here, and what the product's C++ implementation
(``doubletdetection_b200/csrc/leiden.cpp``) must reproduce label for label, is the algorithm of
Traag, Waltman & van Eck (2019), "From Louvain to Leiden", with every free choice fixed:

Quality (RB configuration, resolution gamma, undirected, weights w, 2m = sum of all CSR weights):
    Q = sum_ij (A_ij - gamma * k_i * k_j / (2m)) * delta(c_i, c_j)

One *iteration* (``n_iterations=-1``: iterations repeat, each starting from the previous partition, until one
moves no node at any level; at most ``max_iterations`` = 32):
  1. fast local moving on the current (aggregate) graph from the current partition: FIFO queue of all nodes in
     a seeded order (SplitMix64 + Fisher-Yates, ``j = next() % (i + 1)`` for i = n-1 .. 1); a popped node is
     removed from its community and evaluates, for its own community first and then every neighbouring
     community in order of first appearance in its adjacency list,
         gain(c) = w(i, c) - ((gamma * k_i) * tot[c]) / two_m              (IEEE double, this order)
     moving to the FIRST community with the strictly largest gain (staying wins ties); if the node is not alone
     and every gain is negative it opens an empty community (gain 0; the smallest free id); neighbours outside
     the new community that are not queued are appended;
  2. stop if every node is its own community; otherwise *refine*: inside every community C every node starts
     as a singleton; nodes are visited in a fresh seeded permutation; a node v that is still a singleton and
     well connected, ``w(v, C - v) >= (gamma * k_v * (tot[C] - k_v)) / two_m``, gathers the refined communities
     R of its neighbours in C (first appearance); R is eligible if it is well connected,
     ``E(R, C - R) >= (gamma * rtot[R] * (tot[C] - rtot[R])) / two_m``, and
     ``gain(R) = w(v, R) - ((gamma * k_v) * rtot[R]) / two_m >= 0``; staying (gain 0) is always eligible and
     listed first.  With more than one eligible choice, one is drawn with probability proportional to
     ``exp((gain - max_gain) / theta)`` (theta = 0.01; ``r = (next() >> 11) * 2**-53 * total``, first choice whose
     running sum exceeds r);
  3. aggregate the graph by the REFINED partition (ids by first appearance over node index, neighbour lists
     ascending, internal weight kept as a self-loop counted twice); the aggregate nodes start in the
     communities of the non-refined partition; repeat from 1 unless the aggregation merged nothing.
Final labels are renumbered by decreasing community size (ties: first appearance), like ``louvain_ref``.
"""

import math
from collections import deque

import numpy as np

from .louvain_ref import SplitMix64, _aggregate, _permutation, relabel_by_size

THETA = 0.01
MAX_ITERATIONS = 32


def _first_appearance(ids):
    new_id = {}
    out = []
    for c in ids:
        if c not in new_id:
            new_id[c] = len(new_id)
        out.append(new_id[c])
    return out, len(new_id)


def _degrees(n, indptr, weights, selfw):
    k = [0.0] * n
    for i in range(n):
        s = selfw[i]
        for e in range(indptr[i], indptr[i + 1]):
            s += weights[e]
        k[i] = s
    return k


def _move_nodes(n, indptr, indices, weights, k, part, gamma, two_m, rng):
    """Fast local moving from the partition ``part`` (ids in [0, n)), in place.  Returns moved_any."""
    tot = [0.0] * n
    size = [0] * n
    for i in range(n):
        tot[part[i]] += k[i]
        size[part[i]] += 1
    free = [c for c in range(n - 1, -1, -1) if size[c] == 0]  # pop() yields the smallest free id
    queue = deque(_permutation(n, rng))
    inq = [True] * n
    neigh_w = [0.0] * n
    seen = [False] * n
    moved_any = False
    while queue:
        i = queue.popleft()
        inq[i] = False
        ci = part[i]
        ki = k[i]
        cands = [ci]
        seen[ci] = True
        neigh_w[ci] = 0.0
        for e in range(indptr[i], indptr[i + 1]):
            c = part[indices[e]]
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
        if size[ci] > 1 and 0.0 > best_gain:
            best = free.pop()  # an empty community: gain 0
        tot[best] += ki
        if best != ci:
            part[i] = best
            size[ci] -= 1
            size[best] += 1
            if size[ci] == 0:
                free.append(ci)
            moved_any = True
            for e in range(indptr[i], indptr[i + 1]):
                j = indices[e]
                if part[j] != best and not inq[j]:
                    inq[j] = True
                    queue.append(j)
    return moved_any


def _refine(n, indptr, indices, weights, k, part, gamma, two_m, theta, rng):
    """Refined partition (ids = node ids) inside the communities of ``part``."""
    ptot = [0.0] * n
    for i in range(n):
        ptot[part[i]] += k[i]
    refined = list(range(n))
    rtot = list(k)
    rsize = [1] * n
    ext = [0.0] * n  # E(R, C - R) per refined community; singletons: weight to the rest of their community
    for i in range(n):
        s = 0.0
        for e in range(indptr[i], indptr[i + 1]):
            if part[indices[e]] == part[i]:
                s += weights[e]
        ext[i] = s
    lw = [0.0] * n
    seen = [False] * n
    for v in _permutation(n, rng):
        if rsize[refined[v]] != 1:
            continue
        C = part[v]
        kv = k[v]
        if ext[v] < (gamma * kv * (ptot[C] - kv)) / two_m:
            continue
        cands = [v]
        seen[v] = True
        lw[v] = 0.0
        for e in range(indptr[v], indptr[v + 1]):
            u = indices[e]
            if part[u] != C:
                continue
            R = refined[u]
            if not seen[R]:
                seen[R] = True
                lw[R] = 0.0
                cands.append(R)
            lw[R] += weights[e]
        elig = [v]
        gains = [0.0]
        for R in cands[1:]:
            if ext[R] < (gamma * rtot[R] * (ptot[C] - rtot[R])) / two_m:
                continue
            g = lw[R] - ((gamma * kv) * rtot[R]) / two_m
            if g < 0.0:
                continue
            elig.append(R)
            gains.append(g)
        for R in cands:
            seen[R] = False
        if len(elig) == 1:
            continue
        gmax = max(gains)
        probs = [math.exp((g - gmax) / theta) for g in gains]
        total = 0.0
        for p in probs:
            total += p
        r = (rng.next() >> 11) * (2.0**-53) * total
        chosen = elig[-1]
        acc = 0.0
        for R, p in zip(elig, probs):
            acc += p
            if r < acc:
                chosen = R
                break
        if chosen != v:
            refined[v] = chosen
            rtot[chosen] += kv
            rtot[v] -= kv
            rsize[chosen] += 1
            rsize[v] = 0
            ext[chosen] = (ext[chosen] + ext[v]) - 2.0 * lw[chosen]
    return refined


def _iteration(n0, indptr0, indices0, weights0, membership, gamma, two_m, theta, rng):
    n, indptr, indices, weights, selfw = n0, indptr0, indices0, weights0, [0.0] * n0
    part, _ = _first_appearance(membership)
    node_of = list(range(n0))
    improved = False
    while True:
        k = _degrees(n, indptr, weights, selfw)
        if _move_nodes(n, indptr, indices, weights, k, part, gamma, two_m, rng):
            improved = True
        membership = [part[node_of[v]] for v in range(n0)]
        if len(set(part)) == n:
            break
        refined = _refine(n, indptr, indices, weights, k, part, gamma, two_m, theta, rng)
        nc, n_indptr, n_indices, n_weights, n_selfw, node2new = _aggregate(n, indptr, indices, weights, selfw, refined)
        if nc == n:
            break
        part2 = [0] * nc
        for i in range(n):
            part2[node2new[i]] = part[i]
        part, _ = _first_appearance(part2)
        node_of = [node2new[x] for x in node_of]
        n, indptr, indices, weights, selfw = nc, n_indptr, n_indices, n_weights, n_selfw
    return membership, improved


def leiden(indptr, indices, weights=None, resolution=1.0, seed=0, theta=THETA, max_iterations=MAX_ITERATIONS):
    """Cluster a symmetric, self-loop-free CSR graph.  Returns int64 labels, 0 = largest community."""
    indptr = [int(x) for x in indptr]
    indices = [int(x) for x in indices]
    n = len(indptr) - 1
    weights = [1.0] * len(indices) if weights is None else [float(x) for x in weights]
    gamma = float(resolution)
    two_m = 0.0
    for w in weights:
        two_m += w
    rng = SplitMix64(seed)
    membership = list(range(n))
    if two_m > 0.0:
        for _ in range(max_iterations):
            membership, improved = _iteration(n, indptr, indices, weights, membership, gamma, two_m, float(theta), rng)
            if not improved:
                break
    return relabel_by_size(np.asarray(membership, dtype=np.int64))


def quality(indptr, indices, weights, labels, resolution=1.0):
    """RB-configuration quality of a partition divided by 2m (the usual modularity when gamma = 1)."""
    indptr = np.asarray(indptr)
    indices = np.asarray(indices)
    w = np.ones(indices.size) if weights is None else np.asarray(weights, dtype=np.float64)
    labels = np.asarray(labels)
    n = indptr.size - 1
    rows = np.repeat(np.arange(n), np.diff(indptr))
    two_m = w.sum()
    k = np.bincount(rows, weights=w, minlength=n)
    internal = w[labels[rows] == labels[indices]].sum()
    tot = np.bincount(labels, weights=k)
    return (internal - resolution * (tot**2).sum() / two_m) / two_m
