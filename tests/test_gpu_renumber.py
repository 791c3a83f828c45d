"""GPU (B200): renumber.py on the kernels — a rollout on the cell-renumbered mesh is the original's under the permutation (within the
fp32 bound; the CSR orders, and with them the summation orders, move with the numbering), and the converter's --renumber groups load
through Data_Pool like any other."""
import numpy as np
import pytest
import torch

from tests import helpers as H
from tests.test_gpu_parity import fv, dev  # noqa: F401
from tests.test_gpu_poly import _gpu_tables, _rollout_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("math", ["fp32", "bf16x3"])
def test_rollout_on_the_cell_renumbered_mesh_is_the_permuted_rollout(fv, math):
    m = fv.synthetic.make("cyl", seed=21, T=3)
    m2, npm, cpm = fv.renumber_mesh(m)
    sd = H.torch_reference_state_dict(0)
    ua, pa, _, _ = _rollout_gpu(fv, _gpu_tables(fv, m), m, sd, math, 5)
    ub, pb, _, _ = _rollout_gpu(fv, _gpu_tables(fv, m2), m2, sd, math, 5)
    inv = torch.as_tensor(fv.inverse_permutation(cpm), device=ua.device)
    for a, b in ((ua, ub[:, inv]), (pa, pb)):  # next_uv is per cell; the face values keep their order (faces follow the node numbers)
        rel = ((a.double() - b.double()) ** 2).sum(dim=(1, 2)).sqrt() / (a.double() ** 2).sum(dim=(1, 2)).sqrt()
        assert float(rel.max()) < 1e-5, rel.tolist()


def test_converter_writes_renumbered_groups(fv, tmp_path):
    from fvgn_b200 import convert
    from fvgn_b200.dataset.mesh_store import MeshStore

    m = fv.synthetic.make("grid", seed=2, T=4, nx=30, ny=21)
    traj = {"mesh_pos": m["mesh_pos"][None], "cells": m["cells"][None], "node_type": m["node_type"].reshape(1, -1, 1),
            "velocity": m["velocity"], "pressure": m["pressure"]}
    w = MeshStore(str(tmp_path / "a.npz"), "w")
    g0 = convert.extract_mesh_state(traj, w, 0)
    g1 = convert.extract_mesh_state(traj, w, 1, renumber="cells")
    w.close()
    cpm = g1["renumber|cell_perm"]
    assert np.array_equal(g1["renumber|node_perm"], np.arange(m["mesh_pos"].shape[0]))
    assert np.array_equal(g1["face"], g0["face"]) and np.array_equal(g1["face_length"], g0["face_length"])  # faces follow the node numbers
    assert np.array_equal(g1["centroid"][0], g0["centroid"][0][cpm]) and np.array_equal(g1["cells_area"][0], g0["cells_area"][0][cpm])
    inv = fv.inverse_permutation(cpm)
    assert np.array_equal(inv[g0["neighbour_cell"][0]], g1["neighbour_cell"][0])
    r = MeshStore(str(tmp_path / "a.npz"), "r")
    assert "renumber|cell_perm" in r["1"] and "renumber|cell_perm" not in r["0"]
