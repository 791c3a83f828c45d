"""CPU: the oracle (oracle/fvgn_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H

INT_TABLES = ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type"]
FP_TABLES = ["face_length", "unit_norm_v", "centroid", "cells_area", "cell_factor"]


@pytest.mark.parametrize("name", ["conn_tiny_A", "conn_tiny_B", "conn_cyl_A", "conn_cyl_B", "conn_foil_B"])
def test_connectivity_tables(golden, name):
    g = golden(name + ".npz")
    t = O.build_connectivity(g["in_mesh_pos"], g["in_cells"], g["in_node_type"], face_rule=name[-1])
    for k in INT_TABLES:
        assert t[k].dtype == np.int32 and np.array_equal(t[k], g[k]), k  # bit-exact
    for k in ("centroid", "cell_factor", "face_length", "unit_norm_v"):
        assert np.array_equal(t[k], g[k]), k  # bit-exact fp32 (reorder_face's sign tests depend on the centroid bits)
    # cells_area ends in torch.sqrt, whose vectorised CPU kernel (MKL VML) is 1-ulp, not correctly rounded
    np.testing.assert_allclose(t["cells_area"], g["cells_area"], rtol=3e-7, atol=0)


def test_weights_rebuild_matches_reference_init(golden):
    H.check_weights(H.torch_reference_state_dict(0), golden("fwd_tiny.npz"))


def _graphs(g, prefix="tab.", vel="velocity", prs="pressure", pair=None):
    return O.build_graphs(H.tables_from(g, prefix), g[vel], g[prs], training_pair=pair)


def test_eval_forward_intermediates_and_rollout(golden):
    g = golden("fwd_tiny.npz")
    sd = H.load_state(H.torch_reference_state_dict(0), g)
    orc = O.FVGNOracle(sd, H.params()).eval()
    gn, ge, gc = _graphs(g)
    gn, ge, gc, mi, mb = O.valid_datapreprocessing(gn, ge, gc, 0)
    assert np.array_equal(gc.x.numpy(), g["in_x"]) and np.array_equal(gc.edge_attr.numpy(), g["in_edge_attr"])
    assert np.array_equal(mb.numpy(), g["mask_face_boundary"])
    cell_type, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    with torch.no_grad():
        for step in range(3):
            cap = {} if step == 0 else None
            uvp, nuv = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0, capture=cap)
            if step == 0:
                for k in ["cell_in", "edge_in", "enc_x", "enc_e", "x_l0", "e_l0", "x_l1", "e_l1", "x_l14", "e_l14", "decoded", "face_area", "delta_u"]:
                    np.testing.assert_array_equal(cap[k].numpy(), g["cap." + k], err_msg=k)  # same torch ops => same bits
            np.testing.assert_array_equal(nuv.numpy(), g["roll_uv"][step])
            np.testing.assert_array_equal(uvp.numpy(), g["roll_uvp"][step])
            gc = O.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)


def test_cylinder_rollout_10(golden):
    g = golden("fwd_cyl.npz")
    sd = H.load_state(H.torch_reference_state_dict(0), g)
    orc = O.FVGNOracle(sd, H.params()).eval()
    gn, ge, gc = _graphs(g)
    gn, ge, gc, mi, mb = O.valid_datapreprocessing(gn, ge, gc, 0)
    cell_type, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    with torch.no_grad():
        for step in range(10):
            uvp, nuv = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0)
            if step == 0:
                assert H.rel_l2(nuv, g["uv_step1"]) < 1e-6 and H.rel_l2(uvp, g["uvp_step1"]) < 1e-6
            gc = O.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)
    assert H.rel_l2(nuv, g["uv_step10"]) < 1e-5 and H.rel_l2(uvp, g["uvp_step10"]) < 1e-5


@pytest.mark.parametrize("loss_name,tag", [("direct_mean_loss", "dm"), ("global_mean_sum", "gm")])
def test_training_step(golden, loss_name, tag):
    g = golden("train_tiny.npz")
    f = golden("fwd_tiny.npz")
    p = H.params(loss=loss_name)
    sd = H.load_state(H.torch_reference_state_dict(0), f)
    orc = O.FVGNOracle(sd, p).train()
    for k, v in list(orc.sd.items()):
        if v.is_floating_point() and ("model." in k or k in ("_grid_feature_normalizer.weight", "_grid_feature_normalizer.bias")):
            orc.sd[k] = v.clone().requires_grad_(True)
    orc.bn["weight"], orc.bn["bias"] = orc.sd["_grid_feature_normalizer.weight"], orc.sd["_grid_feature_normalizer.bias"]
    b0 = _graphs(f, "tab.", "velocity", "pressure", pair=1)
    b1 = _graphs(f, "tab1.", "velocity1", "pressure1", pair=0)
    gn, ge, gc = [O.batch_graphs([b0[k], b1[k]]) for k in range(3)]
    gn, ge, gc, mi, mb = O.train_datapreprocessing(gn, ge, gc, torch.as_tensor(g["noise"]), torch.as_tensor(g["flip"]))
    np.testing.assert_array_equal(gc.x.numpy(), g["pre_x"])
    np.testing.assert_array_equal(gc.edge_attr.numpy(), g["pre_edge_attr"])
    np.testing.assert_array_equal(gc.y.numpy(), g["pre_cell_y"])
    np.testing.assert_array_equal(ge.y.numpy(), g["pre_edge_y"])
    np.testing.assert_array_equal(mi.numpy(), g["mask_face_interior"])
    out = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0)
    loss = O.compute_fvm_loss(p, gc.batch, ge.batch, mi, out[3], out[4], out[1], out[2], num_graphs=2)
    loss[0].backward()
    if tag == "dm":
        np.testing.assert_array_equal(out[1].detach().numpy(), g["pred_delta_u"])
        np.testing.assert_array_equal(out[3].detach().numpy(), g["pred_edge"])
        np.testing.assert_array_equal(out[2].detach().numpy(), g["target_delta_u"])
        np.testing.assert_array_equal(out[4].detach().numpy(), g["target_face"])
    got = np.array([float(loss[0]), float(loss[2].mean()), float(loss[3].mean()), float(loss[4].mean())])
    np.testing.assert_allclose(got, g[tag + ".loss"], rtol=1e-6)
    keys = [str(k) for k in g[tag + ".gkeys"]]
    gn_ = np.array([float(orc.sd[k].grad.double().norm()) for k in keys])
    np.testing.assert_allclose(gn_, g[tag + ".gnorm"], rtol=2e-5, atol=1e-9)
    for k in g:
        if k.startswith(tag + ".grad."):
            name = k[len(tag) + 6:]
            assert H.rel_l2(orc.sd[name].grad, g[k]) < 1e-5, name


def test_mp_times_1_rollout_against_reference_fixture(golden):
    """the oracle's mp_times=1 CellBlock branch against 3 rollout steps of the unmodified reference run with mp_times=1 on fwd_tiny's mesh,
    weights and primed state (oracle/make_golden_mp1.py) — and the fixture is NOT what mp_times=2 gives"""
    g, g1 = golden("fwd_tiny.npz"), golden("fwd_tiny_mp1.npz")
    sd = H.load_state(H.torch_reference_state_dict(0), g)
    gn, ge, gc = _graphs(g)
    gn, ge, gc, mi, mb = O.valid_datapreprocessing(gn, ge, gc, 0)
    orc = O.FVGNOracle(sd, H.params(mp_times=1)).eval()
    ct, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    with torch.no_grad():
        for step in range(3):
            uvp, nuv = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0)
            assert H.rel_l2(nuv, g1["roll_uv"][step]) < 1e-6 and H.rel_l2(uvp, g1["roll_uvp"][step]) < 1e-6
            O.create_next_graph(gc, ge, mb, nuv, ct, Re, rmp)
    assert H.rel_l2(g["roll_uv"][0], g1["roll_uv"][0]) > 1e-3
