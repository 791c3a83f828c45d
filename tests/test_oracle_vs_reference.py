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


@pytest.mark.parametrize("mp_times", [1, 2])
def test_training_loss_and_gradients_live(ref, mp_times):
    """training mode, live: the reference's FVGN.forward (train) + compute_FVM_loss + backward against torch autograd over the oracle on
    the same batch, noise and flips — loss and every parameter gradient, for the default network and the mp_times=1 variant"""
    from oracle import make_golden as MG

    params = RL.default_params(train_traj_length=2, mp_times=mp_times, loss="global_mean_sum")
    meshes = [S.make("tiny", seed=31 + i, T=4) for i in range(2)]
    Rs = [RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"]) for m in meshes]
    dp = RL.ref_data_pool(Rs, params, is_training=True)
    model = MG.new_model(ref, params, 0).train()
    gn, ge, gc, mi, mb, noise, flip = MG.train_batch(ref, dp, params, [(0, 1), (1, 0)], 321)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    res = model(graph=gc, graph_edge=ge, graph_node=gn, rho=1.0, mu=1e-3, dt=0.01, edge_one_hot=9, cell_one_hot=0, device="cpu")
    loss = ref.compute_FVM_loss(params=params, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=res[3],
                                target_face_uvp_normalized=res[4], predicted_delta_u=res[1], target_delta_u=res[2],
                                loss_function=torch.nn.MSELoss(reduction="mean"))
    loss[0].backward()
    want = {k: p.grad for k, p in model.named_parameters()}
    # oracle on the same inputs
    orc = O.FVGNOracle(sd0, params).train()
    for k, v in list(orc.sd.items()):
        if v.is_floating_point() and ("model." in k or k in ("_grid_feature_normalizer.weight", "_grid_feature_normalizer.bias")):
            orc.sd[k] = v.clone().requires_grad_(True)
    orc.bn["weight"], orc.bn["bias"] = orc.sd["_grid_feature_normalizer.weight"], orc.sd["_grid_feature_normalizer.bias"]
    bags = []
    for R, m, pair in zip(Rs, meshes, (1, 0)):
        T = {k: (v[0] if v.ndim > 0 and v.shape[0] == 1 else v) for k, v in R.items()}
        bags.append(O.build_graphs(T, m["velocity"], m["pressure"], training_pair=pair))
    on, oe, oc = [O.batch_graphs([b[k] for b in bags]) for k in range(3)]
    on, oe, oc, omi, omb = O.train_datapreprocessing(on, oe, oc, noise, flip)
    out = orc(oc, oe, on, 1.0, 1e-3, 0.01, 9, 0)
    oloss = O.compute_fvm_loss(params, oc.batch, oe.batch, omi, out[3], out[4], out[1], out[2], num_graphs=2)
    oloss[0].backward()
    assert abs(float(oloss[0].detach()) - float(loss[0].detach())) < 1e-6
    got = {k: v.grad for k, v in orc.sd.items() if v.requires_grad}
    assert set(got) == set(want)
    for k, w in want.items():
        assert H.rel_l2(got[k], w) < 1e-5, (k, H.rel_l2(got[k], w))


@pytest.mark.parametrize("seed", range(40, 52))
def test_connectivity_live_random_small_meshes(ref, seed):
    """a sweep of small random meshes of every generator (both face-typing rules): the oracle's tables against the live reference"""
    kind = ("cyl", "rect", "naca", "grid")[seed % 4]
    kw = dict(n_target=260 + 17 * (seed % 5)) if kind != "grid" else dict(nx=9 + seed % 7, ny=6 + seed % 5, jitter=0.3)
    m = S.make(kind, seed=seed, **kw)
    rule = "AB"[seed % 2]
    R = RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"], face_rule=rule)
    T = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], face_rule=rule)
    for k in ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type", "centroid", "cell_factor", "face_length", "unit_norm_v"]:
        assert np.array_equal(R[k][0], T[k]), (kind, seed, k)
    np.testing.assert_allclose(T["cells_area"], R["cells_area"][0], rtol=3e-7, atol=0)
