"""Locality-restoring renumbering of a mesh: cells (and optionally nodes) sorted along a Morton (Z-order) curve.

Why: the CellBlock gathers three node rows per cell and the EdgeBlock two cell rows per face (GN_blocks.py:83-145).  When neighbouring
cells carry neighbouring numbers those rows are still in L2 when the next cell asks for them; with an arbitrary numbering (mesh generators
and format converters do not promise one) every gather goes to DRAM — on the 1M-cell sweep mesh, whose cells are numbered at random on
purpose, the CellBlock kernel then moves 1.6x its algorithmic bytes (profiles/ncu_traffic.json).

* CELL renumbering (the default) is exact: the reference derives the faces, their orientation and the sender / receiver order from NODE
  numbers (parse_tfrecord_refactor.py:419-971: edges canonicalised as (min node, max node)), so a per-cell result of the renumbered mesh is
  the original's under `inverse_permutation(cell_perm)` up to the summation order inside the scatter-adds (tests/test_renumber_cpu.py:
  <= 1e-6 per step in the oracle, usually the same bits).
* NODE renumbering (nodes=True) also orders the faces (they are numbered in (min node, max node) order) and with them the EdgeBlock's
  gathers — but it changes those orientation conventions: the result is a different, equally valid dataset (the reference fixes no
  numbering; a network is not invariant to the flip of an edge's direction), not comparable value by value with the original one.

A one-time, host-side step of the dataset conversion (convert.py --renumber cells|all); numpy, O(n log n)."""
from __future__ import annotations

import numpy as np


def _spread16(v):
    v = v.astype(np.uint64)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
    return v


def morton_codes(xy, lo=None, hi=None):
    """32-bit Z-order code of 2-D points (16 bits per axis over the bounding box [lo, hi])"""
    xy = np.asarray(xy, dtype=np.float64)
    lo = xy.min(0) if lo is None else np.asarray(lo, dtype=np.float64)
    hi = xy.max(0) if hi is None else np.asarray(hi, dtype=np.float64)
    span = np.maximum(hi - lo, 1e-300)
    q = np.clip(np.floor((xy - lo) / span * 65536.0), 0, 65535).astype(np.uint64)
    return _spread16(q[:, 0]) | (_spread16(q[:, 1]) << np.uint64(1))


def morton_permutation(xy, lo=None, hi=None):
    """new -> old index order of the points along the Z-order curve (stable: ties keep their original order)"""
    return np.argsort(morton_codes(xy, lo, hi), kind="stable").astype(np.int64)


def inverse_permutation(perm):
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.shape[0], dtype=perm.dtype)
    return inv


def renumber_mesh(mesh, nodes=False, cells=True):
    """mesh: dict with `mesh_pos` [N,2], `cells` [C,3] (or [C,4] with -1 padding), `node_type` [N] and any number of node fields whose
    LAST-BUT-ONE axis is the node axis (`velocity` [T,N,2], `pressure` [T,N,1]) — the layout of the reference's h5 groups
    (Load_mesh.py:243-320) and of fvgn_b200.synthetic.  Returns (mesh', node_perm, cell_perm): new -> old orders;
    a per-cell result y' of the renumbered mesh is y'[inverse_permutation(cell_perm)] in the original numbering."""
    pos = np.asarray(mesh["mesh_pos"])
    cl = np.asarray(mesh["cells"])
    N, C = pos.shape[0], cl.shape[0]
    lo, hi = pos.min(0).astype(np.float64), pos.max(0).astype(np.float64)
    node_perm = morton_permutation(pos, lo, hi) if nodes else np.arange(N, dtype=np.int64)
    inv = inverse_permutation(node_perm)
    valid = cl >= 0
    cl2 = np.where(valid, inv[np.maximum(cl, 0)], -1).astype(cl.dtype)
    if cells:
        k = np.maximum(valid.sum(1, keepdims=True), 1)
        cen = (np.where(valid[..., None], pos[np.maximum(cl, 0)].astype(np.float64), 0.0)).sum(1) / k
        cell_perm = morton_permutation(cen, lo, hi)
    else:
        cell_perm = np.arange(C, dtype=np.int64)
    out = dict(mesh)
    out["cells"] = np.ascontiguousarray(cl2[cell_perm])
    for key, v in mesh.items():
        if key in ("cells", "meta") or not hasattr(v, "shape"):
            continue
        a = np.asarray(v)
        if a.ndim >= 1 and a.shape[0] == N and key in ("mesh_pos", "node_type"):
            out[key] = np.ascontiguousarray(a[node_perm])
        elif a.ndim >= 2 and a.shape[-2] == N:
            out[key] = np.ascontiguousarray(np.take(a, node_perm, axis=-2))
        elif a.ndim == 1 and a.shape[0] == N:
            out[key] = np.ascontiguousarray(a[node_perm])
    if "meta" in mesh and isinstance(mesh["meta"], dict):
        out["meta"] = dict(mesh["meta"], renumbered="morton")
    return out, node_perm, cell_perm
