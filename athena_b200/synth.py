"""Synthetic graph batches of the shapes named in BASELINE.json `configs`
(SURVEY.md section 8d).  Host-side, vectorised numpy; produces the packed
layout the C ABI takes (graph.PackedGraphs).  Every undirected edge appears in
both endpoint rows with the same 1-based edge id; self-loops carry edge id 0
unless self_loop_features=True.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .graph import PackedGraphs


def packed_from_edges(nv: np.ndarray, src: np.ndarray, dst: np.ndarray, *,
                      add_self_loops: bool = True, self_loop_features: bool = False,
                      directed: bool = False, unique_pairs: bool = False
                      ) -> Tuple[PackedGraphs, np.ndarray]:
    """Build the packed per-graph CSR from GLOBAL vertex ids.

    nv       [B] vertices per graph (graphs occupy consecutive global id ranges)
    src,dst  [M] undirected edges (global ids, both ends in the same graph); each
             is listed once and expanded to both directions unless directed=True
    Returns (packed, edge_graph) where edge_graph[k] is the graph of edge k and
    packed.ne counts edge-feature columns per graph (undirected edges [+ self loops]).
    Rows are sorted by neighbour id (then edge id).
    """
    nv = np.asarray(nv, np.int64)
    B = nv.size
    V = int(nv.sum())
    voff = np.concatenate([[0], np.cumsum(nv)])
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    M = src.size
    vgraph = np.repeat(np.arange(B), nv)
    edge_graph = vgraph[src] if M else np.zeros(0, np.int64)
    # local 1-based edge ids: edges are numbered in input order within their graph
    order = np.argsort(edge_graph, kind="stable")
    ne_edges = np.bincount(edge_graph, minlength=B).astype(np.int64)
    eoff = np.concatenate([[0], np.cumsum(ne_edges)])
    local_id = np.empty(M, np.int64)
    local_id[order] = np.arange(M) - eoff[edge_graph[order]] + 1
    if directed:
        rows, cols, eids = src, dst, local_id
    else:
        rows = np.concatenate([src, dst])
        cols = np.concatenate([dst, src])
        eids = np.concatenate([local_id, local_id])
    ne = ne_edges.copy()
    if add_self_loops:
        allv = np.arange(V, dtype=np.int64)
        rows = np.concatenate([rows, allv])
        cols = np.concatenate([cols, allv])
        if self_loop_features:
            self_ids = ne_edges[vgraph] + (allv - voff[vgraph]) + 1
            ne = ne + nv
        else:
            self_ids = np.zeros(V, np.int64)
        eids = np.concatenate([eids, self_ids])
    if unique_pairs:   # no multi-edges: one 64-bit key sort instead of a three-key lexsort
        key = np.argsort(rows * np.int64(V) + cols, kind="stable")
    else:
        key = np.lexsort((eids, cols, rows))
    rows, cols, eids = rows[key], cols[key], eids[key]
    deg = np.bincount(rows, minlength=V).astype(np.int64)
    row_ptr = np.concatenate([[0], np.cumsum(deg)])
    nz = (row_ptr[voff[1:]] - row_ptr[voff[:-1]]).astype(np.int64)
    zoff = row_ptr[voff[:-1]]
    # per-graph 1-based adj_ia (nv+1 entries each)
    ia = np.empty(V + B, np.int32)
    pos = np.arange(V) + vgraph            # slot of row v inside the concatenated ia
    ia[pos] = row_ptr[:-1] - zoff[vgraph] + 1
    ia[voff[1:] + np.arange(B)] = nz + 1
    ja = np.empty((rows.size, 2), np.int32)
    ja[:, 0] = cols - voff[vgraph[rows]] + 1
    ja[:, 1] = eids
    packed = PackedGraphs(nv=nv.astype(np.int32), ne=ne.astype(np.int32), nz=nz.astype(np.int32),
                          ia=ia, ja=ja)
    return packed, edge_graph


def regular_batch(B: int, n: int, half_degree: int, F: int, rng: np.random.Generator,
                  Fe: int = 0, self_loop_features: bool = False) -> PackedGraphs:
    """cfg2: B graphs of n vertices, every vertex has exactly 2*half_degree distinct
    neighbours inside its graph (randomly relabelled circulant) + a self-loop."""
    assert 2 * half_degree < n
    offs = np.stack([rng.permutation((n - 1) // 2)[:half_degree] + 1 for _ in range(B)])  # [B,h]
    perm = np.stack([rng.permutation(n) for _ in range(B)])                                # [B,n]
    base = np.arange(n)[None, :, None]
    a = np.broadcast_to(base, (B, n, half_degree))
    b = (base + offs[:, None, :]) % n
    gi = np.arange(B)[:, None, None]
    src = perm[gi, a] + gi * n
    dst = perm[gi, b] + gi * n
    packed, _ = packed_from_edges(np.full(B, n), src.ravel(), dst.ravel(),
                                  self_loop_features=self_loop_features)
    packed.x = rng.standard_normal((B * n, F), dtype=np.float32)
    if Fe:
        packed.e = rng.random((packed.E, Fe), dtype=np.float32)
    return packed


def ragged_batch(B: int, n_min: int, n_max: int, half_degree: int, F: int,
                 rng: np.random.Generator) -> PackedGraphs:
    """cfg2 with ragged graph sizes: graph s has n_s ~ U[n_min, n_max] vertices, every vertex
    2*half_degree distinct neighbours inside its graph (randomly relabelled circulant) + a
    self-loop.  Graph-aligned 128-row tiles are then partially filled and unequal."""
    assert 2 * half_degree < n_min <= n_max
    nv = rng.integers(n_min, n_max + 1, B)
    voff = np.concatenate([[0], np.cumsum(nv)])
    srcs, dsts = [], []
    for n in np.unique(nv):
        gs = np.nonzero(nv == n)[0]
        k = gs.size
        offs = np.stack([rng.permutation((n - 1) // 2)[:half_degree] + 1 for _ in range(k)])
        perm = np.stack([rng.permutation(n) for _ in range(k)])
        base = np.arange(n)[None, :, None]
        a = np.broadcast_to(base, (k, n, half_degree))
        b = (base + offs[:, None, :]) % n
        gi = np.arange(k)[:, None, None]
        srcs.append((perm[gi, a] + voff[gs][:, None, None]).ravel())
        dsts.append((perm[gi, b] + voff[gs][:, None, None]).ravel())
    packed, _ = packed_from_edges(nv, np.concatenate(srcs), np.concatenate(dsts))
    packed.x = rng.standard_normal((int(voff[-1]), F), dtype=np.float32)
    return packed


def random_graph(V: int, out_degree: int, F: int, rng: np.random.Generator) -> PackedGraphs:
    """cfg3: one large graph; `out_degree` random out-neighbours per vertex, symmetrised,
    duplicates removed, + self-loops."""
    src = np.repeat(np.arange(V, dtype=np.int64), out_degree)
    dst = rng.integers(0, V, src.size, dtype=np.int64)
    keep = src != dst
    lo = np.minimum(src[keep], dst[keep])
    hi = np.maximum(src[keep], dst[keep])
    key = np.unique(lo * V + hi)
    packed, _ = packed_from_edges(np.array([V]), key // V, key % V, unique_pairs=True)
    packed.x = rng.standard_normal((V, F), dtype=np.float32)
    return packed


def molecular_batch(B: int, F: int, Fe: int, rng: np.random.Generator, nv_range=(10, 50),
                    self_loop_features: bool = True) -> PackedGraphs:
    """cfg4: bonded-molecule-like graphs, V ~ U[nv_range], degree <= 4 (+ self):
    a ring plus two partial perfect matchings of chords."""
    nv = rng.integers(nv_range[0], nv_range[1] + 1, B)
    voff = np.concatenate([[0], np.cumsum(nv)])
    V = int(voff[-1])
    vgraph = np.repeat(np.arange(B), nv)
    local = np.arange(V) - voff[vgraph]
    n_of = nv[vgraph]
    # ring
    src = [np.arange(V)]
    dst = [voff[vgraph] + (local + 1) % n_of]
    # chords i <-> i + n//2 for i < n//2 (random 50 %), and i <-> i + n//3 pattern (random 25 %)
    half = n_of // 2
    m1 = (local < half) & (half >= 2) & (rng.random(V) < 0.5) & ((local + half) % n_of != (local + 1) % n_of) \
        & ((local + half + 1) % n_of != local)
    src.append(np.nonzero(m1)[0])
    dst.append((voff[vgraph] + (local + half) % n_of)[m1])
    src = np.concatenate(src)
    dst = np.concatenate(dst)
    lo, hi = np.minimum(src, dst), np.maximum(src, dst)
    key = np.unique(lo * (V + 1) + hi)
    lo, hi = key // (V + 1), key % (V + 1)
    keep = lo != hi
    packed, _ = packed_from_edges(nv, lo[keep], hi[keep], self_loop_features=self_loop_features)
    packed.x = rng.random((V, F), dtype=np.float32)
    if Fe:
        packed.e = rng.random((packed.E, Fe), dtype=np.float32)
    return packed


def powerlaw_batch(B: int, n: int, F: int, rng: np.random.Generator, alpha: float = 2.1,
                   max_degree: int = 10000, Fe: int = 0) -> PackedGraphs:
    """cfg5: degree-skewed graphs: degrees ~ Zipf(alpha) truncated at max_degree,
    configuration-model wiring, self-pairs and duplicate pairs dropped, + self-loops."""
    srcs, dsts = [], []
    for g in range(B):
        d = np.minimum(rng.zipf(alpha, n), min(max_degree, n - 1)).astype(np.int64)
        if d.sum() % 2:
            d[0] += 1
        stubs = np.repeat(np.arange(n, dtype=np.int64), d)
        rng.shuffle(stubs)
        a, b = stubs[0::2], stubs[1::2]
        keep = a != b
        lo, hi = np.minimum(a[keep], b[keep]), np.maximum(a[keep], b[keep])
        key = np.unique(lo * n + hi)
        srcs.append(key // n + g * n)
        dsts.append(key % n + g * n)
    packed, _ = packed_from_edges(np.full(B, n), np.concatenate(srcs), np.concatenate(dsts),
                                  self_loop_features=bool(Fe))
    packed.x = rng.standard_normal((B * n, F), dtype=np.float32)
    if Fe:
        packed.e = rng.random((packed.E, Fe), dtype=np.float32)
    return packed


def chemical_batch(B: int, rng: np.random.Generator, n: int = 8, Fv: int = 6, Fe: int = 1
                   ) -> PackedGraphs:
    """cfg1 stand-in for example/msgpass_chemical (198 periodic 8-atom carbon cells,
    main.f90:75-77,129-157): n atoms, multi-edges between atom pairs (periodic images)
    so that degrees fall in ~6..17 (mean ~11.5), one edge feature in (0.17, 1) = r / 3 A,
    self-loops added by the caller's add_self_loops (edge id 0)."""
    iu, ju = np.triu_indices(n, 1)
    srcs, dsts = [], []
    for g in range(B):
        mult = rng.choice([0, 1, 2, 3], size=iu.size, p=[0.1, 0.35, 0.4, 0.15])
        srcs.append(np.repeat(iu, mult) + g * n)
        dsts.append(np.repeat(ju, mult) + g * n)
    packed, _ = packed_from_edges(np.full(B, n), np.concatenate(srcs), np.concatenate(dsts))
    packed.x = rng.random((B * n, Fv), dtype=np.float32)
    packed.e = (0.17 + 0.83 * rng.random((packed.E, Fe))).astype(np.float32)
    return packed


def edge_lists(p: PackedGraphs) -> Tuple[np.ndarray, np.ndarray]:
    """The undirected edge lists a packed batch was generated from, as graphstruc holds them
    before generate_adjacency: (num_edges [B], index_list [sum num_edges, 2]) with 1-based
    vertex pairs per graph, edge k of a graph = its k-th pair (edge id k of adj_ja(2, :)).
    Self loops with edge id 0 (add_self_loops) are not edges."""
    nv = p.nv.astype(np.int64)
    voff = np.concatenate([[0], np.cumsum(nv)])
    zoff = np.concatenate([[0], np.cumsum(p.nz.astype(np.int64))])
    deg = np.concatenate([np.diff(p.ia[voff[s] + s:voff[s + 1] + s + 1]) for s in range(p.B)]) \
        if p.B < 64 else None
    if deg is None:   # vectorised: row lengths from the concatenated per-graph row pointers
        ia = p.ia.astype(np.int64)
        idx = np.arange(voff[-1]) + np.repeat(np.arange(p.B), nv)
        deg = ia[idx + 1] - ia[idx]
    row_local = np.repeat(np.arange(voff[-1]) - np.repeat(voff[:-1], nv), deg) + 1
    graph = np.repeat(np.repeat(np.arange(p.B), nv), deg)
    nb, eid = p.ja[:, 0].astype(np.int64), p.ja[:, 1].astype(np.int64)
    keep = (eid > 0) & (nb >= row_local)      # each undirected edge once (self edges: nb == row)
    g, e, a, b = graph[keep], eid[keep], row_local[keep], nb[keep]
    order = np.lexsort((e, g))
    ne = np.bincount(g, minlength=p.B).astype(np.int32)
    il = np.stack([a[order], b[order]], axis=1).astype(np.int32)
    return ne, np.ascontiguousarray(il)
