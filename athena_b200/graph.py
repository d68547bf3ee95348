"""Host-side mirror of graphstruc's graph_type as athena consumes it.

Fields used by the message-passing path (docs/source/tutorials/network_outputs.rst:155-176,
athena_msgpass_layer_sub.f90:144-174, athena_input_layer.f90:511-556):
num_vertices, num_edges, num_vertex_features, num_edge_features,
vertex_features(Fv,V), edge_features(Fe,E), adj_ia(V+1), adj_ja(2,Z), is_sparse.

Arrays are stored in the reference's memory order: a Fortran (F, V) array is a
C-contiguous numpy array of shape [V, F]; adj_ja(2, Z) is [Z, 2].  Indices are
1-based int32, exactly what the Fortran side holds, so the same buffers cross
the C ABI unchanged.

graphstruc itself (generate_adjacency / add_self_loops) is an out-of-tree
dependency of the reference (fpm.toml:21) whose neighbour ordering no athena
test pins; the helpers below are conveniences for building inputs and are NOT
part of the parity contract -- the layers take adj_ia / adj_ja as given.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


class graph_type:
    def __init__(self):
        self.num_vertices = 0
        self.num_edges = 0
        self.num_vertex_features = 0
        self.num_edge_features = 0
        self.is_sparse = True
        self.vertex_features: Optional[np.ndarray] = None  # [V, Fv]
        self.edge_features: Optional[np.ndarray] = None    # [E, Fe]
        self.edge_weights: Optional[np.ndarray] = None
        self.adj_ia: Optional[np.ndarray] = None           # [V+1], 1-based
        self.adj_ja: Optional[np.ndarray] = None           # [Z, 2], 1-based {neighbour, edge id}

    # -- graphstruc-style setters -----------------------------------------
    def set_num_vertices(self, num_vertices: int, num_vertex_features: int = 0):
        self.num_vertices = int(num_vertices)
        self.num_vertex_features = int(num_vertex_features)
        self.vertex_features = np.zeros((self.num_vertices, self.num_vertex_features), np.float32)

    def set_num_edges(self, num_edges: int, num_edge_features: int = 0):
        self.num_edges = int(num_edges)
        self.num_edge_features = int(num_edge_features)
        self.edge_features = np.zeros((self.num_edges, self.num_edge_features), np.float32)
        self.edge_weights = np.ones(self.num_edges, np.float32)

    def generate_adjacency(self, index_list: Sequence[Sequence[int]]):
        """index_list: [num_edges][2] 1-based undirected vertex pairs.  Each
        undirected edge k appears in both endpoint rows with edge id k."""
        rows: List[List[tuple]] = [[] for _ in range(self.num_vertices)]
        for k, (i, j) in enumerate(index_list, start=1):
            rows[i - 1].append((j, k))
            if i != j:
                rows[j - 1].append((i, k))
        self._from_rows(rows)

    def add_self_loops(self):
        """Append a self-loop to every row that has none; edge id 0 (the marker
        seen in test/test_diffstruc_extd_kipf.f90:30-31)."""
        rows = self._rows()
        for v in range(self.num_vertices):
            if all(nb != v + 1 for nb, _ in rows[v]):
                rows[v].append((v + 1, 0))
        self._from_rows(rows)

    def _rows(self):
        rows = []
        for v in range(self.num_vertices):
            seg = self.adj_ja[self.adj_ia[v] - 1:self.adj_ia[v + 1] - 1]
            rows.append([(int(a), int(b)) for a, b in seg])
        return rows

    def _from_rows(self, rows):
        ia = [1]
        ja = []
        for r in rows:
            ja.extend(r)
            ia.append(ia[-1] + len(r))
        self.adj_ia = np.asarray(ia, np.int32)
        self.adj_ja = np.asarray(ja, np.int32).reshape(-1, 2)
        self.is_sparse = True

    @property
    def num_entries(self) -> int:
        return 0 if self.adj_ja is None else int(self.adj_ja.shape[0])


@dataclass
class PackedGraphs:
    """graph(:) packed for athena_cuda_batch_create (include/athena_cuda.h)."""
    nv: np.ndarray   # [B] int32
    ne: np.ndarray   # [B] int32
    nz: np.ndarray   # [B] int32
    ia: np.ndarray   # [sum(nv+1)] int32, 1-based per graph
    ja: np.ndarray   # [Z, 2] int32, 1-based per graph
    x: Optional[np.ndarray] = None  # [V, Fv] float32
    e: Optional[np.ndarray] = None  # [E, Fe] float32

    @property
    def B(self) -> int:
        return int(self.nv.size)

    @property
    def V(self) -> int:
        return int(self.nv.sum())

    @property
    def Z(self) -> int:
        return int(self.nz.sum())

    @property
    def E(self) -> int:
        return int(self.ne.sum())

    def slice(self, g0: int, g1: int) -> "PackedGraphs":
        """Graphs [g0, g1) as an independent batch (views where possible)."""
        voff = np.concatenate([[0], np.cumsum(self.nv, dtype=np.int64)])
        zoff = np.concatenate([[0], np.cumsum(self.nz, dtype=np.int64)])
        eoff = np.concatenate([[0], np.cumsum(self.ne, dtype=np.int64)])
        return PackedGraphs(
            nv=self.nv[g0:g1], ne=self.ne[g0:g1], nz=self.nz[g0:g1],
            ia=self.ia[voff[g0] + g0:voff[g1] + g1], ja=self.ja[zoff[g0]:zoff[g1]],
            x=None if self.x is None else self.x[voff[g0]:voff[g1]],
            e=None if self.e is None else self.e[eoff[g0]:eoff[g1]])


def pack_graphs(graphs: Sequence[graph_type], with_features: bool = True) -> PackedGraphs:
    nv = np.asarray([g.num_vertices for g in graphs], np.int32)
    ne = np.asarray([g.num_edges for g in graphs], np.int32)
    nz = np.asarray([g.num_entries for g in graphs], np.int32)
    ia = np.ascontiguousarray(np.concatenate([g.adj_ia for g in graphs]), np.int32)
    ja = np.ascontiguousarray(np.concatenate([g.adj_ja.reshape(-1, 2) for g in graphs]), np.int32)
    x = e = None
    if with_features:
        x = np.ascontiguousarray(np.concatenate([g.vertex_features for g in graphs]), np.float32)
        if graphs[0].num_edge_features > 0:
            e = np.ascontiguousarray(np.concatenate([g.edge_features for g in graphs]), np.float32)
    return PackedGraphs(nv, ne, nz, ia, ja, x, e)
