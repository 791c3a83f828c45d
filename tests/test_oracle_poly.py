"""CPU: the tri/quad generalisation of the oracle (oracle/fvgn_oracle_poly.py; test infrastructure for BASELINE configs[3]'s "hybrid tri/quad"
mesh, which the triangles-only reference cannot read).  What can be pinned is pinned: on all-triangle meshes the generalised tables and the
network over them reduce to the triangle oracle bit for bit (and that one is held to the reference); on hybrid meshes the tables satisfy the
geometric identities of a conforming polygonal mesh and the finite-volume step behaves (closed-surface flux cancellation, fp32 vs fp64)."""
import numpy as np
import pytest
import torch

import fvgn_b200.synthetic as S
from oracle import fvgn_oracle as O
from oracle import fvgn_oracle_poly as P
from tests import helpers as H


def _rollout(tab, m, steps, dtype=torch.float32, params=None):
    gn, ge, gc = O.build_graphs(tab, m["velocity"], m["pressure"])
    if dtype == torch.float64:
        for bag in (gn, ge, gc):
            for k, v in list(bag.__dict__.items()):
                if torch.is_tensor(v) and v.is_floating_point():
                    setattr(bag, k, v.double())
    gn, ge, gc, mi, mb = O.valid_datapreprocessing(gn, ge, gc, 0)
    orc = O.FVGNOracle(H.torch_reference_state_dict(0), params or H.params(), dtype=dtype).eval()
    ct, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    out = []
    with torch.no_grad():
        for _ in range(steps):
            uvp, nuv = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0)
            out.append((uvp.clone(), nuv.clone()))
            O.create_next_graph(gc, ge, mb, nuv, ct, Re, rmp)
    return out


@pytest.mark.parametrize("kind,seed,rule", [("tiny", 3, "B"), ("tiny", 4, "A"), ("cyl", 4, "B"), ("rect", 6, "B"), ("naca", 5, "B")])
def test_tables_reduce_to_the_triangle_oracle(kind, seed, rule):
    m = S.make(kind, seed=seed, **({} if kind == "tiny" else dict(n_target=400)))
    a = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], rule)
    b = P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], rule)  # [C,3] input is padded to [C,4]
    for k in ["face", "neighbour_cell", "face_type", "cells_type", "face_length", "centroid", "cells_area"]:
        assert np.array_equal(a[k], b[k]), k
    for k in ["cells_node", "cells_face", "cell_factor", "unit_norm_v"]:
        assert np.array_equal(a[k], b[k][:, :3]), k
    assert (b["cells_node"][:, 3] == -1).all() and (b["cells_face"][:, 3] == -1).all() and (b["cell_k"] == 3).all()


def test_network_over_padded_triangle_tables_is_bit_identical():
    m = S.make("tiny", seed=9, T=3)
    a = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    b = P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], "B")
    for (u1, v1), (u2, v2) in zip(_rollout(a, m, 4), _rollout(b, m, 4)):
        assert torch.equal(u1, u2) and torch.equal(v1, v2)
    p1 = H.params(mp_times=1)
    for (u1, v1), (u2, v2) in zip(_rollout(a, m, 2, params=p1), _rollout(b, m, 2, params=p1)):
        assert torch.equal(u1, u2) and torch.equal(v1, v2)


@pytest.mark.parametrize("seed,qf", [(0, 0.5), (1, 0.9), (2, 0.15)])
def test_hybrid_mesh_tables_are_a_conforming_polygonal_mesh(seed, qf):
    m = S.make("hybrid", seed=seed, quad_fraction=qf)
    t = P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], "B")
    C, F, N = t["cells_node"].shape[0], t["face"].shape[1], t["mesh_pos"].shape[0]
    k = t["cell_k"]
    assert (k == 4).sum() == m["meta"]["quads"] > 0 and (k == 3).sum() == m["meta"]["triangles"] > 0
    assert N - F + C == 0  # Euler characteristic of a planar mesh with one hole
    valid = np.arange(4)[None, :] < k[:, None]
    assert ((t["cells_face"] >= 0) == valid).all() and ((t["cells_node"] >= 0) == valid).all()
    # every face belongs to one (boundary) or two (interior) cells, and neighbour_cell names exactly them
    cnt = np.bincount(t["cells_face"][valid], minlength=F)
    assert set(np.unique(cnt)) <= {1, 2}
    nc = t["neighbour_cell"]
    assert ((nc[0] == nc[1]) == (cnt == 1)).all()
    for side in (0, 1):
        cf = t["cells_face"][nc[side]]
        assert (cf == np.arange(F)[:, None]).any(1).all()
    # closed polygons: sum_j n_j len_j = 0; areas add up to the domain minus the hole; interpolation weights sum to one
    cfi = np.where(valid, t["cells_face"], 0)
    closed = (t["unit_norm_v"] * np.where(valid, t["face_length"][cfi, 0], 0.0)[:, :, None]).sum(1)
    assert np.abs(closed).max() < 2e-6
    xs = np.unique(t["mesh_pos"][:, 0])
    ys = np.unique(t["mesh_pos"][:, 1])
    hole_nodes = t["mesh_pos"][(m["node_type"] == S.WALL_BOUNDARY) & (t["mesh_pos"][:, 1] > ys.min()) & (t["mesh_pos"][:, 1] < ys.max())]
    hole_area = np.ptp(hole_nodes[:, 0]) * np.ptp(hole_nodes[:, 1])
    assert abs(t["cells_area"].sum() - ((xs.max() - xs.min()) * (ys.max() - ys.min()) - hole_area)) < 1e-5
    np.testing.assert_allclose(t["cell_factor"].sum(1), 1.0, atol=1e-6)
    # outward normals: n . (face centre - centroid) > 0
    fc = (t["mesh_pos"][t["face"][0]] + t["mesh_pos"][t["face"][1]]) / 2
    out = ((fc[cfi] - t["centroid"][:, None, :]) * t["unit_norm_v"]).sum(2)
    assert (out[valid] > 0).all()


def test_hybrid_mesh_finite_volume_step():
    m = S.make("hybrid", seed=3, quad_fraction=0.6)
    t = P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], "B")
    # a uniform face state has no net flux through a closed cell: the Intergrator returns 0 (+ the diffusive term, also 0 here)
    F, C = t["face"].shape[1], t["cells_node"].shape[0]
    uv = torch.tensor([[1.0, 2.0]]).repeat(F, 1)
    cont, du = O.integrator(uv, torch.full((F, 1), 3.0), torch.zeros(F, 2), torch.as_tensor(t["unit_norm_v"]), 1.0, 1.0,
                            torch.as_tensor(t["face_length"]), torch.as_tensor(t["cells_face"]).long().T.contiguous(), loss_cont=1)
    assert du.abs().max() < 2e-5 and cont.abs().max() < 1e-5
    # the network over the hybrid tables: finite, and fp32 tracks fp64 at the reference's own noise level
    r32 = _rollout(t, m, 10)
    r64 = _rollout(t, m, 10, dtype=torch.float64)
    for (u32, v32), (u64, v64) in zip(r32, r64):
        assert torch.isfinite(v32).all() and H.rel_l2(v32, v64) < 1e-5 and H.rel_l2(u32, u64) < 1e-5
    # and it is not the answer of the same mesh with every quad cut into triangles (the quads really enter as 4-gons)
    tri_cells = []
    for row in m["cells"]:
        tri_cells += [[row[0], row[1], row[2]]] + ([[row[0], row[2], row[3]]] if row[3] >= 0 else [])
    tt = O.build_connectivity(m["mesh_pos"], np.asarray(tri_cells, dtype=np.int32), m["node_type"], "B")
    assert tt["cells_node"].shape[0] > C


def test_hybrid_mesh_600_step_rollout_fp32_tracks_fp64():
    """BASELINE configs[3] at the oracle level: 600 autoregressive steps on a tri/quad hybrid mesh; the fp32 oracle against its float64 copy
    with the rollout-error metric of rollout.py:446-482 (cumulative relative squared error)"""
    m = S.make("hybrid", seed=5, quad_fraction=0.5, nx=15, ny=9)
    t = P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], "B")
    r32 = _rollout(t, m, 600)
    r64 = _rollout(t, m, 600, dtype=torch.float64)
    uv32, uv64 = torch.stack([v for _, v in r32]).double(), torch.stack([v for _, v in r64])
    p32, p64 = torch.stack([u[:, 2:3] for u, _ in r32]).double(), torch.stack([u[:, 2:3] for u, _ in r64])
    assert torch.isfinite(uv32).all()
    err = O.rollout_relative_error(uv32, uv64, p32, p64)
    c_uv, c_p = float(err["uv_rmse"][-1]) ** 0.5, float(err["p_rmse"][-1]) ** 0.5
    assert c_uv < 1e-4 and c_p < 1e-4, (c_uv, c_p)
