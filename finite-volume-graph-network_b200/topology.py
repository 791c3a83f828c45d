"""MeshTopology: the int32 index tables + destination-sorted CSR incidence tables the kernels walk, derived ONCE per mesh
batch from the reference-layout graph objects (int64 `edge_index` / `face` tensors) and cached.

Reference layouts:  graph.edge_index[2,F] = neighbour_cell, graph_node.edge_index[2,F] = face (nodes),
graph_node.face[3,C] = cells_node.T, graph_edge.face[3,C] = cells_face.T."""
from __future__ import annotations

import torch

from . import ops


class _Agg:
    __slots__ = ("ptr", "items", "rows", "idx", "face")

    def __init__(self, ptr, items, rows, idx, face):
        self.ptr, self.items, self.rows, self.idx, self.face = ptr, items, rows, idx, face


class MeshTopology:
    def __init__(self, neighbour_cell, face_node, cells_node_T, cells_face_T, n_nodes):
        self.nc = ops.as_i32(neighbour_cell)      # [2,F]
        self.fn = ops.as_i32(face_node)           # [2,F]
        self.cnT = ops.as_i32(cells_node_T)       # [3,C]; [4,C] with -1 padding on tri / quad hybrid meshes
        self.cfT = ops.as_i32(cells_face_T)       # [3,C] / [4,C]
        self.poly = self.cnT.shape[0] == 4
        self.F = int(self.nc.shape[1])
        self.C = int(self.cnT.shape[1])
        self.N = int(n_nodes)
        # node <- (face, half): stable CSR over cat(face[0], face[1])  ==  the reference's scatter_add order (GN_blocks.py:111-122)
        self.node_ptr, self.node_items = ops.build_csr(self.fn.reshape(-1), self.N)
        self._bwd = None

    def agg(self, mp_times=2):
        """What the CellBlock's edge aggregation walks (GN_blocks.py:110-138).  mp_times >= 2: faces -> nodes -> cells (rows = N, the cell
        kernel gathers through cells_node_T, backward scatters through face_node); mp_times = 1: faces -> their two cells directly (rows = C,
        stable CSR over cat(edge_index[0], edge_index[1]) == the reference's scatter_add order; no second gather: idx = None)."""
        if mp_times >= 2:
            return _Agg(self.node_ptr, self.node_items, self.N, self.cnT, self.fn)
        if self.poly:
            ptr, items = self.__dict__.get("_cff") or self.__dict__.setdefault("_cff", ops.build_csr(self.nc.reshape(-1), self.C))
            return _Agg(ptr, items, self.C, None, self.nc)
        ptr, items = self.backward_tables()["cell_from_face"]
        return _Agg(ptr, items, self.C, None, self.nc)

    def backward_tables(self):
        """CSR transposes needed by the backward kernels: cell<-(face,side), node<-(cell,k), face<-(cell,k).
        Tri / quad tables ([4,C], -1 padded): the -1 entries are keyed to an extra destination row that the kernels do not walk (they are given
        n_dst rows of an (n_dst + 2)-entry pointer array)."""
        if self._bwd is None:
            def csr(keys, n):
                if not self.poly:
                    return ops.build_csr(keys, n)
                return ops.build_csr(torch.where(keys < 0, torch.full_like(keys, n), keys), n + 1)

            self._bwd = dict(
                cell_from_face=ops.build_csr(self.nc.reshape(-1), self.C),
                node_from_cell=csr(self.cnT.reshape(-1), self.N),
                face_from_cell=csr(self.cfT.reshape(-1), self.F),
            )
        return self._bwd

    def inv_k(self):
        """[C,1] fp32: 1 / (number of nodes of the cell) — the transpose of the CellBlock's k-wide mean on tri / quad tables"""
        v = self.__dict__.get("_inv_k")
        if v is None:
            if self.poly:
                v = torch.where(self.cnT[3] >= 0, 0.25, 1.0 / 3.0).to(torch.float32).view(-1, 1).contiguous()
            else:
                v = torch.full((self.C, 1), 1.0 / 3.0, dtype=torch.float32, device=self.cnT.device)
            self.__dict__["_inv_k"] = v
        return v


def _train_tables(topo, mp_times):
    """fvgn_train_tables_t of this topology (cached; the struct borrows the topology's tensors)"""
    from ._lib import TrainTablesT

    key = "_tt%d" % (2 if mp_times >= 2 else 1)
    tt = topo.__dict__.get(key)
    if topo.poly:
        raise NotImplementedError("the C-side training executor takes [3,C] tables; tri / quad meshes train through the per-call path")
    if tt is None:
        bt = topo.backward_tables()
        p = lambda t: t.data_ptr()
        tt = TrainTablesT(topo.C, topo.F, topo.N, p(topo.nc), p(topo.fn), p(topo.cnT), p(topo.node_ptr), p(topo.node_items),
                          p(bt["cell_from_face"][0]), p(bt["cell_from_face"][1]), p(bt["node_from_cell"][0]), p(bt["node_from_cell"][1]))
        topo.__dict__[key] = tt
    return tt


MeshTopology.train_tables = _train_tables

_CACHE: dict = {}
_CACHE_MAX = 16


def _key(t):
    return (t.data_ptr(), tuple(t.shape), t._version, str(t.device))


def topology_of(graph, graph_node, graph_edge=None) -> MeshTopology:
    """Cached MeshTopology for a (graph_cell, graph_node[, graph_edge]) triple of reference-layout bags."""
    cfT = graph_edge.face if graph_edge is not None and hasattr(graph_edge, "face") else None
    k = (_key(graph.edge_index), _key(graph_node.edge_index), _key(graph_node.face), _key(cfT) if cfT is not None else None,
         int(graph_node.num_nodes))
    topo = _CACHE.get(k)
    if topo is None:
        if cfT is None:
            cfT = torch.zeros_like(graph_node.face)
        topo = MeshTopology(graph.edge_index, graph_node.edge_index, graph_node.face, cfT, graph_node.num_nodes)
        # keep the source tensors alive so data_ptr keys cannot be recycled while cached
        topo._src = (graph.edge_index, graph_node.edge_index, graph_node.face, cfT)
        if len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))
        _CACHE[k] = topo
    return topo


def clear_cache():
    _CACHE.clear()
