"""CPU: renumber.py (Z-order renumbering of cells / nodes; SURVEY §8(f): the data format on the input side of the path).
A renumbered mesh is the same mesh; with CELL renumbering the oracle's rollout on it equals the original's after mapping back, up to
summation-order rounding (faces, orientations and sender / receiver order derive from node numbers); NODE renumbering gives another valid dataset."""
import numpy as np
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H


def _fv():
    import fvgn_b200 as fv
    return fv


def test_morton_order_properties():
    fv = _fv()
    from fvgn_b200 import renumber as R

    rng = np.random.default_rng(0)
    xy = rng.random((5000, 2))
    perm = R.morton_permutation(xy)
    assert sorted(perm.tolist()) == list(range(5000))
    inv = fv.inverse_permutation(perm)
    assert np.array_equal(perm[inv], np.arange(5000)) and np.array_equal(inv[perm], np.arange(5000))
    codes = R.morton_codes(xy)[perm]
    assert (np.diff(codes.astype(np.int64)) >= 0).all()
    # locality: consecutive points of the curve are close (mean step ~ 1/sqrt(n)), a random order is not (~0.52)
    step = np.linalg.norm(np.diff(xy[perm], axis=0), axis=1).mean()
    assert step < 0.05 < np.linalg.norm(np.diff(xy, axis=0), axis=1).mean()
    # a quadrant of the box is one contiguous quarter of the order
    pb = R.morton_permutation(xy, lo=(0.0, 0.0), hi=(1.0, 1.0))
    q = (xy[pb] < 0.5).all(1)
    assert q[: q.sum()].all()


def test_renumbered_mesh_is_the_same_mesh():
    fv = _fv()
    for kind, kw in (("grid", dict(nx=40, ny=23)), ("cyl", dict(n_target=600)), ("hybrid", dict(nx=21, ny=14))):
        m = fv.synthetic.make(kind, seed=3, **kw)
        m2, npm, cpm = fv.renumber_mesh(m, nodes=True)
        assert np.array_equal(m2["mesh_pos"], m["mesh_pos"][npm]) and np.array_equal(m2["node_type"], m["node_type"][npm])
        assert np.array_equal(m2["velocity"], m["velocity"][:, npm]) and np.array_equal(m2["pressure"], m["pressure"][:, npm])
        # cell c' of the new mesh is cell cpm[c'] of the old one, vertex by vertex (order within the cell kept: orientation unchanged)
        old = m["cells"][cpm]
        new = m2["cells"]
        assert np.array_equal(new >= 0, old >= 0)
        assert np.array_equal(np.where(new >= 0, npm[np.maximum(new, 0)], -1), old)
        assert m2["cells"].dtype == m["cells"].dtype
        again = fv.synthetic.make(kind, seed=3, renumber="all", **kw)
        assert np.array_equal(again["cells"], m2["cells"]) and np.array_equal(again["mesh_pos"], m2["mesh_pos"])
    # gather locality: neighbouring cell numbers now reference neighbouring node numbers
    m = fv.synthetic.make("grid", seed=1, nx=120, ny=90)
    m2 = fv.renumber_mesh(m)[0]
    assert np.array_equal(m2["mesh_pos"], m["mesh_pos"])  # default: cells only
    spread = lambda c: np.abs(np.diff(c.astype(np.int64).min(1))).mean()
    assert spread(m2["cells"]) * 20 < spread(m["cells"])


def _oracle_rollout(mm, sd, steps=3):
    tab = O.build_connectivity(mm["mesh_pos"], mm["cells"], mm["node_type"], "B")
    gn, ge, gc = O.build_graphs(tab, mm["velocity"], mm["pressure"])
    gn, ge, gc, mi, mb = O.valid_datapreprocessing(gn, ge, gc, 0)
    orc = O.FVGNOracle(sd, H.params()).eval()
    ct, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    uv = []
    with torch.no_grad():
        for _ in range(steps):
            uvp, nuv = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0)
            uv.append(nuv.double())
            O.create_next_graph(gc, ge, mb, nuv, ct, Re, rmp)
    return torch.stack(uv), tab


def test_oracle_rollout_is_equivariant_under_cell_renumbering():
    fv = _fv()
    sd = H.torch_reference_state_dict(0)
    for kind, kw in (("cyl", dict(n_target=500)), ("grid", dict(nx=30, ny=19))):
        m = fv.synthetic.make(kind, seed=7, T=3, **kw)
        m2, npm, cpm = fv.renumber_mesh(m)
        assert np.array_equal(npm, np.arange(npm.shape[0])) and not np.array_equal(cpm, np.arange(cpm.shape[0]))
        a, ta = _oracle_rollout(m, sd)
        b, tb = _oracle_rollout(m2, sd)
        inv = torch.as_tensor(fv.inverse_permutation(cpm))
        bb = b[:, inv]
        rel = ((a - bb) ** 2).sum(dim=(1, 2)).sqrt() / (a ** 2).sum(dim=(1, 2)).sqrt()
        print(kind, "cell renumbering: per-step rel-L2 vs the original numbering", rel.tolist())
        assert float(rel.max()) < 1e-6  # summation order inside the scatter-adds only (the first step is usually bit-equal)
        # the face tables do not move at all; the cell tables are permuted
        for k in ("face", "face_length", "face_type"):
            assert np.array_equal(ta[k], tb[k]), k
        assert np.array_equal(ta["cells_area"][cpm], tb["cells_area"]) and np.array_equal(ta["centroid"][cpm], tb["centroid"])


def test_node_renumbering_gives_another_valid_mesh():
    fv = _fv()
    m = fv.synthetic.make("cyl", seed=7, T=3, n_target=500)
    m2, npm, cpm = fv.renumber_mesh(m, nodes=True)
    ta = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    tb = O.build_connectivity(m2["mesh_pos"], m2["cells"], m2["node_type"], "B")
    for k in ("cells_area", "face_length"):  # the same geometry: same multisets
        np.testing.assert_allclose(np.sort(ta[k].reshape(-1)), np.sort(tb[k].reshape(-1)), rtol=1e-4)  # fp32 cross products in another vertex order
    assert np.array_equal(np.sort(ta["face_type"].reshape(-1)), np.sort(tb["face_type"].reshape(-1)))
    b, _ = _oracle_rollout(m2, H.torch_reference_state_dict(0), steps=2)
    assert torch.isfinite(b).all()
