"""Synthetic mesh + field generators shaped like the reference's datasets (there is no network, so no
cylinder_flow / airfoil / COMSOL files).  They emit exactly what the reference's converter consumes
(`extract_mesh_state`, Extract_mesh/parse_tfrecord_refactor.py:430-487,793-794):

    mesh_pos f32[N,2], cells i32[C,3], node_type i32[N], velocity f32[T,N,2], pressure f32[T,N,1]

Node types follow utils/utilities.py:6-15 (NORMAL 0, INFLOW 4, OUTFLOW 5, WALL_BOUNDARY 6) so the
reference's face/cell typing rules (parse_tfrecord_refactor.py:508-691) apply unchanged.
Everything is numpy/scipy on the host and deterministic in `seed`.
"""
from __future__ import annotations

import numpy as np

NORMAL, OBSTACLE, AIRFOIL, HANDLE, INFLOW, OUTFLOW, WALL_BOUNDARY = 0, 1, 2, 3, 4, 5, 6


def _compact(pos, cells):
    used = np.zeros(pos.shape[0], dtype=bool)
    used[cells.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    return pos[used], remap[cells].astype(np.int32), used


def _ccw(pos, cells):
    a, b, c = pos[cells[:, 0]], pos[cells[:, 1]], pos[cells[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    return area2


def _box_types(pos, xmin, xmax, ymin, ymax, inner_mask):
    t = np.zeros(pos.shape[0], dtype=np.int32)
    t[pos[:, 0] == xmin] = INFLOW
    t[pos[:, 0] == xmax] = OUTFLOW
    t[(pos[:, 1] == ymin) | (pos[:, 1] == ymax)] = WALL_BOUNDARY  # corners count as wall
    t[inner_mask] = WALL_BOUNDARY
    return t


def fields(pos, node_type, T, seed, ymin, ymax, umax=1.0):
    """velocity[T,N,2] = parabolic inflow profile x (1 + 0.05 N(0,1)); pressure[T,N,1] = N(0,1).
    No-slip on WALL_BOUNDARY nodes."""
    rng = np.random.default_rng(seed + 7919)
    N = pos.shape[0]
    s = (pos[:, 1] - ymin) / (ymax - ymin)
    prof = 4.0 * umax * s * (1.0 - s)
    vel = np.zeros((T, N, 2), dtype=np.float32)
    vel[:, :, 0] = prof[None, :] * (1.0 + 0.05 * rng.standard_normal((T, N)))
    vel[:, :, 1] = 0.05 * umax * rng.standard_normal((T, N))
    vel[:, node_type == WALL_BOUNDARY, :] = 0.0
    prs = rng.standard_normal((T, N, 1)).astype(np.float32)
    return vel.astype(np.float32), prs


def cylinder_mesh(seed=0, n_target=1900, T=3, xlim=(0.0, 1.6), ylim=(0.0, 0.41)):
    """Cylinder-in-channel, Delaunay, graded toward the cylinder: N~1.9k, C~3.6k, F~5.5k."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    xmin, xmax = xlim
    ymin, ymax = ylim
    cx, cy = 0.33 + 0.02 * (rng.random() - 0.5), 0.2 + 0.02 * (rng.random() - 0.5)
    r = 0.02 + 0.06 * rng.random()
    scale = np.sqrt(n_target / 1900.0)
    nx, ny = int(round(60 * scale)), int(round(16 * scale))
    ncirc = max(16, int(round(40 * scale * (0.6 + r / 0.08))))
    bx = np.linspace(xmin, xmax, nx + 1)
    by = np.linspace(ymin, ymax, ny + 1)
    pts = [np.stack([bx, np.full_like(bx, ymin)], 1), np.stack([bx, np.full_like(bx, ymax)], 1),
           np.stack([np.full_like(by[1:-1], xmin), by[1:-1]], 1),
           np.stack([np.full_like(by[1:-1], xmax), by[1:-1]], 1)]
    th = np.linspace(0, 2 * np.pi, ncirc, endpoint=False)
    circ = np.stack([cx + r * np.cos(th), cy + r * np.sin(th)], 1)
    pts.append(circ)
    n_b = sum(p.shape[0] for p in pts)
    # interior: rejection sampling with density ~ 1/(d^1.2) toward the cylinder + uniform floor
    need = max(n_target - n_b, 16)
    h = np.sqrt((xmax - xmin) * (ymax - ymin) / max(n_target, 1))
    out = []
    got = 0
    while got < need:
        cand = np.stack([xmin + (xmax - xmin) * rng.random(4 * need), ymin + (ymax - ymin) * rng.random(4 * need)], 1)
        d = np.hypot(cand[:, 0] - cx, cand[:, 1] - cy) - r
        wall = np.minimum.reduce([cand[:, 0] - xmin, xmax - cand[:, 0], cand[:, 1] - ymin, ymax - cand[:, 1]])
        ok = (d > 0.45 * h * (r / 0.05) ** 0.5) & (wall > 0.45 * h)
        w = 0.25 + 0.75 / (1.0 + (np.maximum(d, 0) / 0.12) ** 1.5)
        ok &= rng.random(cand.shape[0]) < w
        cand = cand[ok]
        out.append(cand)
        got += cand.shape[0]
    interior = np.concatenate(out)[:need]
    pos = np.concatenate(pts + [interior]).astype(np.float32).astype(np.float64)
    # remove near-duplicate points (would create degenerate triangles)
    key = np.round(pos / (0.2 * h)).astype(np.int64)
    _, first = np.unique(key[:, 0] * 1000003 + key[:, 1], return_index=True)
    keep = np.zeros(pos.shape[0], dtype=bool)
    keep[first] = True
    keep[:n_b] = True
    pos = pos[keep]
    tri = Delaunay(pos).simplices.astype(np.int32)
    cen = pos[tri].mean(1)
    inside = np.hypot(cen[:, 0] - cx, cen[:, 1] - cy) < r * 0.999
    tri = tri[~inside]
    a2 = _ccw(pos, tri)
    tri = tri[np.abs(a2) > 1e-12]
    pos, tri, used = _compact(pos, tri)
    pos32 = pos.astype(np.float32)
    on_c = np.abs(np.hypot(pos[:, 0] - cx, pos[:, 1] - cy) - r) < 1e-6
    nt = _box_types(pos32, np.float32(xmin), np.float32(xmax), np.float32(ymin), np.float32(ymax), on_c)
    vel, prs = fields(pos32, nt, T, seed, ymin, ymax)
    return dict(mesh_pos=pos32, cells=tri, node_type=nt, velocity=vel, pressure=prs,
                meta=dict(kind="cylinder", seed=seed, cx=cx, cy=cy, r=r))


def _naca0012(n):
    b = np.linspace(0, np.pi, n)
    x = 0.5 * (1 - np.cos(b))
    y = 0.6 * (0.2969 * np.sqrt(x) - 0.1260 * x - 0.3516 * x ** 2 + 0.2843 * x ** 3 - 0.1036 * x ** 4)
    up = np.stack([x, y], 1)
    lo = np.stack([x[1:-1][::-1], -y[1:-1][::-1]], 1)
    return np.concatenate([up, lo])


def _point_in_poly(pts, poly):
    x, y = pts[:, 0], pts[:, 1]
    inside = np.zeros(pts.shape[0], dtype=bool)
    n = poly.shape[0]
    j = n - 1
    for i in range(n):
        xi, yi, xj, yj = poly[i, 0], poly[i, 1], poly[j, 0], poly[j, 1]
        c = ((yi > y) != (yj > y)) & (x < (xj - xi) * (y - yi) / (yj - yi + 1e-300) + xi)
        inside ^= c
        j = i
    return inside


def _poly_distance(pts, poly):
    """distance of points to a closed polyline [M,2]"""
    d = np.full(pts.shape[0], np.inf)
    a = poly
    b = np.roll(poly, -1, axis=0)
    for k in range(poly.shape[0]):
        ab = b[k] - a[k]
        t = np.clip(((pts - a[k]) @ ab) / max(float(ab @ ab), 1e-300), 0.0, 1.0)
        d = np.minimum(d, np.hypot(*(pts - (a[k] + t[:, None] * ab)).T))
    return d


def channel_obstacle_mesh(shape="rect", seed=0, n_target=1900, T=3, xlim=(0.0, 1.6), ylim=(0.0, 0.41)):
    """Channel with a polygonal obstacle, Delaunay, graded toward the obstacle: the other two geometries of the reference's
    "hybrid" COMSOL dataset (cylinder / rect / NACA0012 — Extract_mesh/parse_comsol.py:411-418; all-triangle).
    shape = "rect" (axis-aligned box, random aspect) or "naca" (NACA0012 at a random angle of attack)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    xmin, xmax = xlim
    ymin, ymax = ylim
    cx, cy = 0.33 + 0.02 * (rng.random() - 0.5), 0.2 + 0.02 * (rng.random() - 0.5)
    scale = np.sqrt(n_target / 1900.0)
    nper = max(24, int(round(56 * scale)))
    if shape == "rect":
        w, hgt = 0.05 + 0.07 * rng.random(), 0.04 + 0.06 * rng.random()
        k = max(2, nper // 4)
        tx, ty = np.linspace(-0.5, 0.5, k + 1)[:-1], np.linspace(-0.5, 0.5, k + 1)[:-1]
        poly = np.concatenate([np.stack([tx * w, np.full(k, -0.5 * hgt)], 1), np.stack([np.full(k, 0.5 * w), ty * hgt], 1),
                               np.stack([-tx * w, np.full(k, 0.5 * hgt)], 1), np.stack([np.full(k, -0.5 * w), -ty * hgt], 1)])
        size = max(w, hgt)
    elif shape == "naca":
        chord = 0.16 + 0.08 * rng.random()
        aoa = np.deg2rad(-8.0 + 16.0 * rng.random())
        f = _naca0012(nper // 2 + 1)
        f = (f - np.array([0.5, 0.0])) * chord
        rot = np.array([[np.cos(aoa), np.sin(aoa)], [-np.sin(aoa), np.cos(aoa)]])
        poly = f @ rot.T
        size = chord
    else:
        raise ValueError(shape)
    poly = poly + np.array([cx, cy])
    nx, ny = int(round(60 * scale)), int(round(16 * scale))
    bx = np.linspace(xmin, xmax, nx + 1)
    by = np.linspace(ymin, ymax, ny + 1)
    pts = [np.stack([bx, np.full_like(bx, ymin)], 1), np.stack([bx, np.full_like(bx, ymax)], 1),
           np.stack([np.full_like(by[1:-1], xmin), by[1:-1]], 1), np.stack([np.full_like(by[1:-1], xmax), by[1:-1]], 1), poly]
    n_b = sum(p.shape[0] for p in pts)
    need = max(n_target - n_b, 16)
    h = np.sqrt((xmax - xmin) * (ymax - ymin) / max(n_target, 1))
    seg = float(np.hypot(*(np.roll(poly, -1, axis=0) - poly).T).max())
    out, got = [], 0
    while got < need:
        cand = np.stack([xmin + (xmax - xmin) * rng.random(4 * need), ymin + (ymax - ymin) * rng.random(4 * need)], 1)
        near = np.hypot(cand[:, 0] - cx, cand[:, 1] - cy) < size
        d = np.full(cand.shape[0], size)
        d[near] = _poly_distance(cand[near], poly)
        inside = np.zeros(cand.shape[0], dtype=bool)
        inside[near] = _point_in_poly(cand[near], poly)
        wall = np.minimum.reduce([cand[:, 0] - xmin, xmax - cand[:, 0], cand[:, 1] - ymin, ymax - cand[:, 1]])
        ok = (~inside) & (d > max(0.45 * h, 0.7 * seg)) & (wall > 0.45 * h)
        dist_c = np.maximum(np.hypot(cand[:, 0] - cx, cand[:, 1] - cy) - 0.5 * size, 0)
        ok &= rng.random(cand.shape[0]) < 0.25 + 0.75 / (1.0 + (dist_c / 0.12) ** 1.5)
        out.append(cand[ok])
        got += out[-1].shape[0]
    interior = np.concatenate(out)[:need]
    pos = np.concatenate(pts + [interior]).astype(np.float32).astype(np.float64)
    key = np.round(pos / (0.2 * h)).astype(np.int64)
    _, first = np.unique(key[:, 0] * 1000003 + key[:, 1], return_index=True)
    keep = np.zeros(pos.shape[0], dtype=bool)
    keep[first] = True
    keep[:n_b] = True
    pos = pos[keep]
    tri = Delaunay(pos).simplices.astype(np.int32)
    poly32 = pos[n_b - poly.shape[0]:n_b]
    inside = _point_in_poly(pos[tri].mean(1), poly32)
    on_poly = np.zeros(pos.shape[0], dtype=bool)
    on_poly[n_b - poly.shape[0]:n_b] = True
    inside |= on_poly[tri].all(1)  # slivers spanned by three obstacle nodes (concave trailing edge / chord-crossing triangles)
    tri = tri[~inside]
    tri = tri[np.abs(_ccw(pos, tri)) > 1e-12]
    pos, tri, used = _compact(pos, tri)
    pos32 = pos.astype(np.float32)
    nt = _box_types(pos32, np.float32(xmin), np.float32(xmax), np.float32(ymin), np.float32(ymax), on_poly[used])
    vel, prs = fields(pos32, nt, T, seed, ymin, ymax)
    return dict(mesh_pos=pos32, cells=tri, node_type=nt, velocity=vel, pressure=prs, meta=dict(kind=shape, seed=seed, cx=cx, cy=cy))


def airfoil_mesh(seed=0, n_target=5200, T=3, xlim=(-10.0, 20.0), ylim=(-10.0, 10.0)):
    """NACA0012 (chord 1) in a far-field box with strong grading: N~5.2k, C~10.2k, F~15.4k.
    Foil nodes are typed WALL_BOUNDARY (the incompressible branch has no AIRFOIL face rule)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    xmin, xmax = xlim
    ymin, ymax = ylim
    aoa = np.deg2rad(4.0 * (rng.random() - 0.5))
    foil = _naca0012(max(40, int(90 * np.sqrt(n_target / 5200.0))))
    rot = np.array([[np.cos(aoa), np.sin(aoa)], [-np.sin(aoa), np.cos(aoa)]])
    foil = (foil - np.array([0.25, 0.0])) @ rot.T + np.array([0.25, 0.0])
    nb = int(round(28 * np.sqrt(n_target / 5200.0)))
    bx = np.linspace(xmin, xmax, int(nb * 1.5) + 1)
    by = np.linspace(ymin, ymax, nb + 1)
    pts = [np.stack([bx, np.full_like(bx, ymin)], 1), np.stack([bx, np.full_like(bx, ymax)], 1),
           np.stack([np.full_like(by[1:-1], xmin), by[1:-1]], 1),
           np.stack([np.full_like(by[1:-1], xmax), by[1:-1]], 1), foil]
    n_b = sum(p.shape[0] for p in pts)
    need = max(n_target - n_b, 16)
    # radial sampling with log-uniform radius about the foil => strong grading
    out = []
    got = 0
    ctr = np.array([0.5, 0.0])
    while got < need:
        m = 4 * need
        rr = np.exp(rng.uniform(np.log(0.02), np.log(22.0), m))
        th = rng.uniform(0, 2 * np.pi, m)
        base = foil[rng.integers(0, foil.shape[0], m)]
        dirn = np.stack([np.cos(th), np.sin(th)], 1)
        cand = np.where((rr < 0.6)[:, None], base + rr[:, None] * dirn, ctr + rr[:, None] * dirn * np.array([1.3, 1.0]))
        ok = (cand[:, 0] > xmin + 0.3) & (cand[:, 0] < xmax - 0.3) & (cand[:, 1] > ymin + 0.3) & (cand[:, 1] < ymax - 0.3)
        ok &= ~_point_in_poly(cand, foil)
        dmin = np.min(np.hypot(cand[:, None, 0] - foil[None, ::4, 0], cand[:, None, 1] - foil[None, ::4, 1]), 1)
        ok &= dmin > 0.012
        cand = cand[ok]
        out.append(cand)
        got += cand.shape[0]
    interior = np.concatenate(out)[:need]
    pos = np.concatenate(pts + [interior]).astype(np.float32).astype(np.float64)
    # drop interior points that crowd each other (local spacing ~ 0.25 * distance to foil centre, min 0.01)
    d = np.hypot(pos[:, 0] - 0.5, pos[:, 1])
    cell = np.maximum(0.008, 0.02 * d)
    key = np.stack([np.round(pos[:, 0] / cell), np.round(pos[:, 1] / cell), np.round(np.log2(cell) * 2)], 1).astype(np.int64)
    _, first = np.unique(key[:, 0] * 1000003 + key[:, 1] * 1009 + key[:, 2], return_index=True)
    keep = np.zeros(pos.shape[0], dtype=bool)
    keep[first] = True
    keep[:n_b] = True
    pos = pos[keep]
    tri = Delaunay(pos).simplices.astype(np.int32)
    cen = pos[tri].mean(1)
    tri = tri[~_point_in_poly(cen, foil)]
    tri = tri[np.abs(_ccw(pos, tri)) > 1e-12]
    is_foil = np.zeros(pos.shape[0], dtype=bool)
    off = n_b - foil.shape[0]
    is_foil[off:n_b] = True
    pos, tri, used = _compact(pos, tri)
    is_foil = is_foil[used]
    pos32 = pos.astype(np.float32)
    nt = _box_types(pos32, np.float32(xmin), np.float32(xmax), np.float32(ymin), np.float32(ymax), is_foil)
    vel, prs = fields(pos32, nt, T, seed, ymin, ymax)
    # far-field: uniform stream rather than a parabola
    rng2 = np.random.default_rng(seed + 31)
    vel[:, :, 0] = (1.0 + 0.05 * rng2.standard_normal((T, pos32.shape[0]))).astype(np.float32)
    vel[:, nt == WALL_BOUNDARY, :] = 0.0
    return dict(mesh_pos=pos32, cells=tri, node_type=nt, velocity=vel, pressure=prs,
                meta=dict(kind="airfoil", seed=seed, aoa=float(aoa)))


def grid_mesh(nx=61, ny=31, seed=0, T=3, jitter=0.25, hole=True, xlim=(0.0, 1.6), ylim=(0.0, 0.41), shuffle=True):
    """Jittered structured triangulation with one rectangular hole (WALL_BOUNDARY) — the shape of the
    1M-cell sweep (nx=ny=708 -> C~1.0e6, F~1.5e6, N~0.5e6).  Vectorised, O(N)."""
    rng = np.random.default_rng(seed)
    xmin, xmax = xlim
    ymin, ymax = ylim
    xs = np.linspace(xmin, xmax, nx)
    ys = np.linspace(ymin, ymax, ny)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # [ny, nx]
    hx, hy = xs[1] - xs[0], ys[1] - ys[0]
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    if hole:
        i0, i1 = int(nx * 0.18), max(int(nx * 0.18) + 2, int(nx * 0.26))
        j0, j1 = int(ny * 0.40), max(int(ny * 0.40) + 2, int(ny * 0.60))
    else:
        i0 = i1 = j0 = j1 = -10
    on_hole = (ii >= i0) & (ii <= i1) & (jj >= j0) & (jj <= j1)
    border = (ii == 0) | (ii == nx - 1) | (jj == 0) | (jj == ny - 1)
    free = ~(border | on_hole)
    X = X + np.where(free, jitter * hx * (rng.random(X.shape) - 0.5) * 2 * 0.5, 0.0)
    Y = Y + np.where(free, jitter * hy * (rng.random(Y.shape) - 0.5) * 2 * 0.5, 0.0)
    pos = np.stack([X.reshape(-1), Y.reshape(-1)], 1)
    nid = (jj * nx + ii)
    q_i, q_j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    a = nid[q_j, q_i].reshape(-1)
    b = nid[q_j, q_i + 1].reshape(-1)
    c = nid[q_j + 1, q_i + 1].reshape(-1)
    d = nid[q_j + 1, q_i].reshape(-1)
    flip = rng.random(a.shape[0]) < 0.5
    t1 = np.where(flip[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t2 = np.where(flip[:, None], np.stack([b, c, d], 1), np.stack([a, c, d], 1))
    in_hole = ((q_i >= i0) & (q_i < i1) & (q_j >= j0) & (q_j < j1)).reshape(-1)
    tri = np.concatenate([t1[~in_hole], t2[~in_hole]]).astype(np.int32)
    perm = rng.permutation(tri.shape[0])  # shuffle=True: worst case, cell numbering carries no locality at all
    if shuffle:
        tri = tri[perm]
    pos, tri, used = _compact(pos, tri)
    pos32 = pos.astype(np.float32)
    nt = _box_types(pos32, np.float32(xmin), np.float32(xmax), np.float32(ymin), np.float32(ymax),
                    on_hole.reshape(-1)[used])
    vel, prs = fields(pos32, nt, T, seed, ymin, ymax)
    return dict(mesh_pos=pos32, cells=tri, node_type=nt, velocity=vel, pressure=prs,
                meta=dict(kind="grid", seed=seed, nx=nx, ny=ny))


def hybrid_mesh(nx=25, ny=17, seed=0, T=3, jitter=0.25, quad_fraction=0.5, hole=True, xlim=(0.0, 1.6), ylim=(0.0, 0.41)):
    """Jittered structured mesh whose grid quads are either kept as QUADRILATERAL cells (probability `quad_fraction`) or split into two
    triangles (random diagonal): a tri/quad hybrid in the shape BASELINE configs[3] names.  `cells` is int32 [C,4], 4th column -1 on
    triangles, in random order.  The reference cannot read such a mesh (triangles only); oracle/fvgn_oracle_poly.py defines the tables."""
    rng = np.random.default_rng(seed)
    xmin, xmax = xlim
    ymin, ymax = ylim
    xs, ys = np.linspace(xmin, xmax, nx), np.linspace(ymin, ymax, ny)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    hx, hy = xs[1] - xs[0], ys[1] - ys[0]
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    if hole:
        i0, i1 = int(nx * 0.18), max(int(nx * 0.18) + 2, int(nx * 0.26))
        j0, j1 = int(ny * 0.40), max(int(ny * 0.40) + 2, int(ny * 0.60))
    else:
        i0 = i1 = j0 = j1 = -10
    on_hole = (ii >= i0) & (ii <= i1) & (jj >= j0) & (jj <= j1)
    border = (ii == 0) | (ii == nx - 1) | (jj == 0) | (jj == ny - 1)
    free = ~(border | on_hole)
    X = X + np.where(free, jitter * hx * (rng.random(X.shape) - 0.5), 0.0)
    Y = Y + np.where(free, jitter * hy * (rng.random(Y.shape) - 0.5), 0.0)
    pos = np.stack([X.reshape(-1), Y.reshape(-1)], 1)
    nid = jj * nx + ii
    q_i, q_j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    a, b = nid[q_j, q_i].reshape(-1), nid[q_j, q_i + 1].reshape(-1)
    c, d = nid[q_j + 1, q_i + 1].reshape(-1), nid[q_j + 1, q_i].reshape(-1)
    in_hole = ((q_i >= i0) & (q_i < i1) & (q_j >= j0) & (q_j < j1)).reshape(-1)
    as_quad = (rng.random(a.shape[0]) < quad_fraction) & ~in_hole
    flip = rng.random(a.shape[0]) < 0.5
    minus = np.full(a.shape[0], -1)
    t1 = np.where(flip[:, None], np.stack([a, b, d, minus], 1), np.stack([a, b, c, minus], 1))
    t2 = np.where(flip[:, None], np.stack([b, c, d, minus], 1), np.stack([a, c, d, minus], 1))
    tri = ~as_quad & ~in_hole
    cells = np.concatenate([np.stack([a, b, c, d], 1)[as_quad], t1[tri], t2[tri]]).astype(np.int64)
    cells = cells[rng.permutation(cells.shape[0])]
    used = np.zeros(pos.shape[0], dtype=bool)
    used[cells[cells >= 0]] = True
    remap = np.cumsum(used) - 1
    cells = np.where(cells >= 0, remap[np.maximum(cells, 0)], -1).astype(np.int32)
    pos32 = pos[used].astype(np.float32)
    nt = _box_types(pos32, np.float32(xmin), np.float32(xmax), np.float32(ymin), np.float32(ymax), on_hole.reshape(-1)[used])
    vel, prs = fields(pos32, nt, T, seed, ymin, ymax)
    return dict(mesh_pos=pos32, cells=cells, node_type=nt, velocity=vel, pressure=prs,
                meta=dict(kind="hybrid", seed=seed, nx=nx, ny=ny, quads=int(as_quad.sum()), triangles=int(2 * tri.sum())))


def make(kind, seed=0, T=3, renumber=False, **kw):
    """renumber="cells" (or True) / "all": the mesh after renumber.renumber_mesh (cells / cells and nodes along a Z-order curve)"""
    if renumber:
        from .renumber import renumber_mesh

        return renumber_mesh(make(kind, seed=seed, T=T, **kw), nodes=(renumber == "all"))[0]
    if kind == "cyl":
        return cylinder_mesh(seed=seed, T=T, **kw)
    if kind in ("rect", "naca"):
        return channel_obstacle_mesh(shape=kind, seed=seed, T=T, **kw)
    if kind == "foil":
        return airfoil_mesh(seed=seed, T=T, **kw)
    if kind == "grid":
        return grid_mesh(seed=seed, T=T, **kw)
    if kind == "hybrid":
        return hybrid_mesh(seed=seed, T=T, **kw)
    if kind == "tiny":
        return grid_mesh(nx=9, ny=7, seed=seed, T=T, **kw)
    if kind == "1m":
        return grid_mesh(nx=708, ny=708, seed=seed, T=T, **kw)
    raise ValueError(kind)
