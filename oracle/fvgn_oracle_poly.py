"""Test infrastructure: the oracle generalised to meshes that mix triangles and quadrilaterals (BASELINE configs[3] names a "hybrid tri/quad"
mesh; the reference itself is triangles only — SURVEY.md §0 item 5, §8(c): "for quad cells the oracle is the builder's own CSR-generalised
restatement ... it must reduce to the reference bit-for-bit when every cell has k = 3").  PARITY UNPINNED for k = 4: there is no reference
behaviour to compare with; what IS pinned is the reduction to the triangle oracle (tests/test_oracle_poly.py: tables `array_equal`, rollouts
`torch.equal`), which in turn is pinned against the reference.  The CUDA kernels do not take these tables yet.

Conventions for a cell with k nodes (k = 3: exactly extract_mesh_state's, parse_tfrecord_refactor.py:419-971):
  * cells come as int32 [C,4], 4th column -1 for triangles.  Canonical node order: triangles sorted ascending (:431); quads keep their cyclic
    order, rotated so that the smallest node id comes first and reflected so that node[1] < node[3].
  * edge j of a cell = (node[j], node[(j+1) % k]); the edge list is stacked column-major (all cells' edge 0, then edge 1, ...) as
    triangles_to_faces does (:1412-1437); faces = unique (max,min) pairs; cells_face[c,j] = index of edge j; padding -1.
  * centroid = vertex mean accumulated left to right, divided by k in fp32 (:437-445).  Area: Heron for triangles (:876-884), shoelace for quads.
  * unit normals, neighbour_cell (+ reorder_face), face / cell typing, cell_factor: the reference's rules with "3" read as "k".
  * network: CellBlock mean over the k node aggregates (sum left to right, / k); Intergrator sum over the k faces; interpolation node -> cell with
    the k cell_factor weights (oracle/fvgn_oracle.py takes [4,C] tables padded with -1 / weight 0)."""
from __future__ import annotations

import numpy as np

from . import fvgn_oracle as O

NORMAL, AIRFOIL, INFLOW, OUTFLOW, WALL_BOUNDARY = O.NORMAL, O.AIRFOIL, O.INFLOW, O.OUTFLOW, O.WALL_BOUNDARY


def canonical_cells(cells4: np.ndarray) -> np.ndarray:
    c = np.asarray(cells4).astype(np.int64)
    if c.shape[1] == 3:
        c = np.concatenate([c, np.full((c.shape[0], 1), -1, dtype=np.int64)], 1)
    out = c.copy()
    tri = c[:, 3] < 0
    out[tri, :3] = np.sort(c[tri, :3], axis=1)
    q = np.where(~tri)[0]
    if q.size:
        rows = c[q]
        r = np.argmin(rows, axis=1)
        rows = np.stack([rows[np.arange(q.size), (r + j) % 4] for j in range(4)], 1)  # smallest id first, cyclic order kept
        flip = rows[:, 1] > rows[:, 3]
        rows[flip] = rows[flip][:, [0, 3, 2, 1]]
        out[q] = rows
    return out.astype(np.int32)


def build_connectivity_poly(mesh_pos: np.ndarray, cells4: np.ndarray, node_type: np.ndarray, face_rule: str = "B") -> dict:
    pos = np.ascontiguousarray(mesh_pos, dtype=np.float32)
    nt = np.asarray(node_type).reshape(-1).astype(np.int32)
    cn = canonical_cells(cells4)
    C = cn.shape[0]
    k = (cn >= 0).sum(1).astype(np.int32)
    valid = np.arange(4)[None, :] < k[:, None]                                  # [C,4]
    cni = np.where(valid, cn, 0)
    quad = k == 4
    # centroid: ((p0+p1)+p2)(+p3) / k in fp32
    s3 = (pos[cni[:, 0]] + pos[cni[:, 1]]) + pos[cni[:, 2]]
    centroid = np.where(quad[:, None], (s3 + pos[cni[:, 3]]) / np.float32(4.0), s3 / np.float32(3.0)).astype(np.float32)
    # edges, stacked column-major over the cells that have an edge j
    nxt = np.stack([cni[:, 1], cni[:, 2], np.where(quad, cni[:, 3], cni[:, 0]), cni[:, 0]], 1)  # node after j in the cycle
    cols_c = [np.where(valid[:, j])[0] for j in range(4)]
    e0 = np.concatenate([cni[cc, j] for j, cc in enumerate(cols_c)])
    e1 = np.concatenate([nxt[cc, j] for j, cc in enumerate(cols_c)])
    key = (np.maximum(e0, e1).astype(np.int64) << 32) | np.minimum(e0, e1).astype(np.int64)
    ukey, inv = np.unique(key, return_inverse=True)
    F = ukey.shape[0]
    face = np.stack([(ukey >> 32), (ukey & 0xFFFFFFFF)], 0).astype(np.int32)
    cells_face = np.full((C, 4), -1, dtype=np.int32)
    o = 0
    for j, cc in enumerate(cols_c):
        cells_face[cc, j] = inv[o:o + cc.size]
        o += cc.size
    cfi = np.where(valid, cells_face, 0)
    d = pos[face[0]] - pos[face[1]]
    face_length = O._torch_norm2_f32(d[:, 0], d[:, 1])[:, None]
    # face types: the triangle oracle's code, unchanged
    a, b = nt[face[0]], nt[face[1]]
    fc = (pos[face[0]] + pos[face[1]]) / np.float32(2.0)
    outlet, inlet = np.max(fc[:, 0]), np.min(fc[:, 0])
    ft = np.zeros(F, dtype=np.int32)
    if face_rule == "A":
        for T in (AIRFOIL, NORMAL, INFLOW):
            ft[(a == b) & (a == T)] = T
    elif face_rule == "B":
        for T in (WALL_BOUNDARY, INFLOW, OUTFLOW, NORMAL):
            ft[(a == b) & (a == T)] = T
        wi = ((a == WALL_BOUNDARY) & (b == INFLOW)) | ((b == WALL_BOUNDARY) & (a == INFLOW))
        ft[wi & (fc[:, 0] < inlet + 1e-4) & (fc[:, 0] > inlet - 1e-4)] = INFLOW
        wo = ((a == WALL_BOUNDARY) & (b == OUTFLOW)) | ((b == WALL_BOUNDARY) & (a == OUTFLOW))
        ft[wo & (fc[:, 0] < outlet + 1e-4) & (fc[:, 0] > outlet - 1e-4)] = OUTFLOW
    else:
        raise ValueError(face_rule)
    tk = np.where(valid, ft[cfi], -1)
    n_in, n_wall = (tk == INFLOW).sum(1), (tk == WALL_BOUNDARY).sum(1)
    n_out, n_foil = (tk == OUTFLOW).sum(1), (tk == AIRFOIL).sum(1)
    n_norm = k - n_in - n_wall - n_out - n_foil
    ct = np.zeros(C, dtype=np.int32)
    wall = (n_wall > 0) & (((n_norm > 0) & (n_in == 0) & (n_out == 0)) | ((n_norm == 0) & (n_in > 0) & (n_out == 0))
                           | ((n_norm == 0) & (n_in == 0) & (n_out > 0)) | ((n_norm > 0) & (n_in > 0) & (n_out == 0))
                           | ((n_norm > 0) & (n_in == 0) & (n_out > 0)))
    foil = ~wall & (n_foil > 0) & (n_norm > 0) & (n_in == 0) & (n_out == 0)
    infl = ~wall & ~foil & (n_in > 0) & (n_norm > 0) & (n_wall == 0) & (n_out == 0)
    outf = ~wall & ~foil & ~infl & (n_out > 0) & (n_norm > 0) & (n_wall == 0) & (n_in == 0)
    ct[wall], ct[foil], ct[infl], ct[outf] = WALL_BOUNDARY, AIRFOIL, INFLOW, OUTFLOW
    # outward unit normals per (cell, face slot)
    unv = np.stack([-d[:, 1], d[:, 0]], 1)
    unv = unv / O._torch_norm2_f32(unv[:, 0], unv[:, 1])[:, None]
    unit_norm_v = np.zeros((C, 4, 2), dtype=np.float32)
    for j in range(4):
        f = cfi[:, j]
        ccv = centroid - fc[f]
        dot = unv[f, 0] * ccv[:, 0] + unv[f, 1] * ccv[:, 1]
        n = np.where((dot > 0)[:, None], unv[f] * np.float32(-1.0), unv[f])
        unit_norm_v[:, j, :] = np.where(valid[:, j:j + 1], n, np.float32(0.0))
    # neighbour_cell: first / second cell of a face in the scan c = 0..C-1, j = 0..k-1
    flat_f = cells_face[valid]                                                   # row-major over (c, j): the scan order
    flat_c = np.repeat(np.arange(C, dtype=np.int64), 4).reshape(C, 4)[valid]
    order = np.argsort(flat_f, kind="stable")
    sf, sc = flat_f[order], flat_c[order]
    first_pos = np.searchsorted(sf, np.arange(F), side="left")
    last_pos = np.searchsorted(sf, np.arange(F), side="right") - 1
    nb = np.stack([sc[first_pos], sc[last_pos]], 1)
    ev = centroid[nb[:, 0]] - centroid[nb[:, 1]]
    keep = (ev[:, 0] > 0) | ((ev[:, 0] == 0) & (ev[:, 1] > 0))
    neighbour_cell = np.where(keep[:, None], nb, nb[:, ::-1]).T.astype(np.int32)
    # cell_factor over the k nodes
    dist = np.stack([((pos[cni[:, j]] - centroid) ** 2).sum(1) for j in range(4)], 1).astype(np.float32)
    dist = np.where(valid, dist, np.float32(0.0))
    t3 = (dist[:, 0] + dist[:, 1]) + dist[:, 2]
    tot = np.where(quad, t3 + dist[:, 3], t3)
    cell_factor = (dist / tot[:, None]).astype(np.float32)
    # area: Heron (triangles), shoelace (quads)
    l1, l2, l3 = (face_length[cfi[:, j], 0] for j in range(3))
    p = np.float32(0.5) * ((l1 + l2) + l3)
    heron = np.sqrt(np.maximum(p * (p - l1) * (p - l2) * (p - l3), np.float32(0.0))).astype(np.float32)
    x, y = [pos[cni[:, j], 0] for j in range(4)], [pos[cni[:, j], 1] for j in range(4)]
    sh = (((x[0] * y[1] - x[1] * y[0]) + (x[1] * y[2] - x[2] * y[1])) + (x[2] * y[3] - x[3] * y[2])) + (x[3] * y[0] - x[0] * y[3])
    shoelace = (np.float32(0.5) * np.abs(sh)).astype(np.float32)
    cells_area = np.where(quad, shoelace, heron).astype(np.float32)[:, None]
    return dict(mesh_pos=pos, node_type=nt[:, None], cells_node=cn, cell_k=k, face=face, cells_face=cells_face, neighbour_cell=neighbour_cell,
                face_type=ft[:, None], cells_type=ct[:, None], face_length=face_length, unit_norm_v=unit_norm_v, centroid=centroid,
                cells_area=cells_area, cell_factor=cell_factor)
