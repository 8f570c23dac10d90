"""Host-side morphology tables: limb tree -> traversal ranks + pairwise relation features.

Produces the ``graph`` dict that ``SEPolicy/SECritic.change_morphology`` consume, with the
same keys and values as the reference's ``utils.getGraphDict`` (src/utils.py:449-484):

* ``parents``    pre-order parent list (torso = -1)
* ``traversals`` 3 int64 vectors (N,): rank of each limb in pre-order, in-order of the
                 left-child/right-sibling (LCRS) binary tree, post-order of the LCRS tree
                 (src/utils.py:368-409, 357-366)
* ``relation``   (N,N,3) fp32 = [personalised PageRank (damping 0.9, closed form,
                 src/utils.py:431-447), symmetric normalised Laplacian, BFS hops / N]

This is tiny, once-per-morphology host work; it stays in
numpy/torch-CPU by design and is checked against the reference in tests/test_graph.py.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, List, Sequence

import numpy as np
import torch

TRAVERSALS = ("pre", "inlcrs", "postlcrs")


def children_of(parents: Sequence[int]) -> List[List[int]]:
    kids: List[List[int]] = [[] for _ in parents]
    for i, p in enumerate(parents):
        if p >= 0:
            kids[p].append(i)
    return kids


def _lcrs_links(kids: List[List[int]]) -> List[List[int]]:
    """LCRS binary tree as adjacency lists.  A node's list is [first child, next sibling]
    when it has both, and a *single* entry when it has only one of them — the reference's
    in-order walk treats that single entry as the left link whichever it is
    (src/utils.py:357-366,382-391); we keep that behaviour since the learned positional
    embeddings are indexed by the resulting ranks."""
    links: List[List[int]] = [[] for _ in kids]
    for node, ch in enumerate(kids):
        if not ch:
            continue
        links[node].insert(0, ch[0])
        for a, b in zip(ch[:-1], ch[1:]):
            links[a].append(b)
    return links


def traversal_ranks(parents: Sequence[int], kinds: Sequence[str] = TRAVERSALS) -> List[List[int]]:
    kids = children_of(parents)
    n = len(parents)
    out = []
    for kind in kinds:
        if kind == "pre":
            out.append(list(range(n)))
            continue
        links = _lcrs_links(kids)
        order: List[int] = []
        if kind == "inlcrs":
            def walk(v):
                if links[v]:
                    walk(links[v][0])
                order.append(v)
                if len(links[v]) == 2:
                    walk(links[v][1])
        elif kind == "postlcrs":
            def walk(v):
                for c in links[v]:
                    walk(c)
                order.append(v)
        else:
            raise ValueError(f"unknown traversal {kind!r}")
        walk(0)
        rank = [0] * n
        for pos, v in enumerate(order):
            rank[v] = pos
        out.append(rank)
    return out


def adjacency(parents: Sequence[int]) -> torch.Tensor:
    n = len(parents)
    a = torch.zeros(n, n)
    for i, p in enumerate(parents):
        if p >= 0:
            a[i, p] = 1.0
            a[p, i] = 1.0
    return a


def _hops(adj: torch.Tensor) -> np.ndarray:
    n = adj.shape[0]
    nb = [torch.nonzero(adj[i]).flatten().tolist() for i in range(n)]
    d = np.full((n, n), -1.0)
    for s in range(n):
        d[s, s] = 0
        q = deque([s])
        while q:
            v = q.popleft()
            for u in nb[v]:
                if d[s, u] < 0:
                    d[s, u] = d[s, v] + 1
                    q.append(u)
    return d / n


def relation_features(parents: Sequence[int], damping: float = 0.9) -> torch.Tensor:
    """(N,N,3) fp32, computed with the same fp32 torch ops as the reference so the tables
    agree to rounding."""
    n = len(parents)
    adj = adjacency(parents)
    eye = torch.eye(n)
    a1 = adj + eye
    trans = (a1 * (1 / a1.sum(1).reshape(-1, 1))).T
    inv = torch.inverse(eye - damping * trans)
    ppr = torch.cat([(1 - damping) * inv @ eye[i].reshape(n, 1) for i in range(n)], dim=1).T
    deg = adj.sum(1)
    lap = torch.diag(deg) - adj
    sym = torch.diag(deg ** -0.5) @ lap @ torch.diag(deg ** -0.5)
    dist = torch.from_numpy(_hops(adj)).float()
    return torch.stack([ppr, sym, dist], dim=2)


def build_graph(parents: Sequence[int], device=None, kinds: Sequence[str] = TRAVERSALS) -> Dict:
    parents = list(parents)
    if len(parents) == 1:
        return {"parents": parents}
    dev = torch.device("cpu") if device is None else device
    return {
        "parents": parents,
        "traversals": [torch.tensor(r, dtype=torch.long, device=dev) for r in traversal_ranks(parents, kinds)],
        "relation": relation_features(parents).to(dev),
    }
