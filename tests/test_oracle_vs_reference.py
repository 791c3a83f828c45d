"""CPU, build container only: the oracle against the UNMODIFIED reference executed live
(skipped where /root/reference is absent).  Random meshes beyond the committed fixtures."""
import copy
import io
import contextlib

import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from oracle import ref_loader as RL
from tests import helpers as H
import fvgn_b200.synthetic as S


@pytest.mark.parametrize("kind,seed,rule", [("tiny", 7, "A"), ("tiny", 8, "B"), ("cyl", 9, "B"), ("grid", 4, "B"),
                                            ("rect", 3, "B"), ("naca", 3, "B"), ("naca", 4, "A")])
def test_connectivity_live(ref, kind, seed, rule):
    m = S.make(kind, seed=seed)
    R = RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"], face_rule=rule)
    T = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], face_rule=rule)
    for k in ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type", "centroid", "cell_factor", "face_length", "unit_norm_v"]:
        assert np.array_equal(R[k][0], T[k]), k
    np.testing.assert_allclose(T["cells_area"], R["cells_area"][0], rtol=3e-7, atol=0)


def test_eval_and_rollout_live(ref):
    params = RL.default_params(train_traj_length=2)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.FVGN(device="cpu", model_dir=None, params=params).eval()
    sd = H.torch_reference_state_dict(0)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k  # rebuild-by-seed == reference constructor
    m = S.make("tiny", seed=21, T=4)
    R = RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"])
    dp = RL.ref_data_pool([R], params)
    gn, ge, gc, mi, mb = dp.require_one_mesh(start_epoch=0, rollout_index=0)
    T = {k: (v[0] if v.ndim > 0 and v.shape[0] == 1 else v) for k, v in R.items()}
    on, oe, oc = O.build_graphs(T, m["velocity"], m["pressure"])
    on, oe, oc, omi, omb = O.valid_datapreprocessing(on, oe, oc, 0)
    orc = O.FVGNOracle(model.state_dict(), params)
    ct, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    with torch.no_grad():
        for _ in range(3):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=1.0, mu=1e-3, dt=0.01, edge_one_hot=9, cell_one_hot=0, device="cpu")
            ouvp, onuv = orc(oc, oe, on, 1.0, 1e-3, 0.01, 9, 0)
            assert torch.equal(uvp, ouvp) and torch.equal(nuv, onuv)
            gc = ref.Data_Pool.create_next_graph(gc, ge, mb, nuv, ct, Re, rmp)
            oc = O.create_next_graph(oc, oe, omb, onuv, ct, Re, rmp)


def test_mp_times_1_eval_live(ref):
    """the reference's mp_times=1 variant (GN_blocks.py:130-138: edge halves scatter-added straight onto the face's two cells):
    oracle == reference bitwise in eval over 3 rollout steps"""
    params = RL.default_params(train_traj_length=2, mp_times=1, loss="global_mean_sum")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.FVGN(device="cpu", model_dir=None, params=params).eval()
    m = S.make("tiny", seed=22, T=4)
    R = RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"])
    dp = RL.ref_data_pool([R], params)
    gn, ge, gc, mi, mb = dp.require_one_mesh(start_epoch=0, rollout_index=0)
    T = {k: (v[0] if v.ndim > 0 and v.shape[0] == 1 else v) for k, v in R.items()}
    on, oe, oc = O.build_graphs(T, m["velocity"], m["pressure"])
    on, oe, oc, omi, omb = O.valid_datapreprocessing(on, oe, oc, 0)
    orc = O.FVGNOracle(model.state_dict(), params)
    ct, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    with torch.no_grad():
        for _ in range(3):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=1.0, mu=1e-3, dt=0.01, edge_one_hot=9, cell_one_hot=0, device="cpu")
            ouvp, onuv = orc(oc, oe, on, 1.0, 1e-3, 0.01, 9, 0)
            assert torch.equal(uvp, ouvp) and torch.equal(nuv, onuv)
            gc = ref.Data_Pool.create_next_graph(gc, ge, mb, nuv, ct, Re, rmp)
            oc = O.create_next_graph(oc, oe, omb, onuv, ct, Re, rmp)
    # and it is a different network from mp_times=2 on the same weights
    p2 = RL.default_params(train_traj_length=2, mp_times=2)
    on2, oe2, oc2 = O.build_graphs(T, m["velocity"], m["pressure"])
    on2, oe2, oc2, _, _ = O.valid_datapreprocessing(on2, oe2, oc2, 0)
    with torch.no_grad():
        _, nuv2 = O.FVGNOracle(model.state_dict(), p2)(oc2, oe2, on2, 1.0, 1e-3, 0.01, 9, 0)
        on1, oe1, oc1 = O.build_graphs(T, m["velocity"], m["pressure"])
        on1, oe1, oc1, _, _ = O.valid_datapreprocessing(on1, oe1, oc1, 0)
        _, nuv1 = O.FVGNOracle(model.state_dict(), params)(oc1, oe1, on1, 1.0, 1e-3, 0.01, 9, 0)
    assert not torch.allclose(nuv1, nuv2)
