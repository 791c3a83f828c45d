"""GPU (B200): the reference's mp_times=1 CellBlock variant (GN_blocks.py:130-138 — the edge halves are scatter-added straight onto the
face's two cells) through the same fused kernels (cells_node_T = NULL: the per-cell aggregate enters the cell MLP as
it is).  The oracle's mp_times=1 branch is held to the live reference bitwise in tests/test_oracle_vs_reference.py; weights, normaliser
state and meshes are those of the committed fixtures (the variant has the same 283-entry state_dict).
Bounds: strict fp32 and the fp32-equivalent tensor-core mode 1e-5 per step, bf16 tiles 3e-2, gradients 1e-4 per tensor."""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H
from tests.test_gpu_parity import fv, dev, _graphs  # noqa: F401

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _model(fv, g, math):
    torch.manual_seed(0)
    m = fv.FVGN(device="cpu", params=H.params(mp_times=1))
    m.load_state_dict(H.load_state(m.state_dict(), g))
    return m.cuda().eval().set_math_mode(math)


def _oracle_rollout(g, steps):
    tab = H.tables_from(g, "tab.")
    on, oe, oc = O.build_graphs(tab, g["velocity"], g["pressure"])
    on, oe, oc, mi, mb = O.valid_datapreprocessing(on, oe, oc, 0)
    orc = O.FVGNOracle(H.load_state(H.torch_reference_state_dict(0), g), H.params(mp_times=1)).eval()
    ct, Re, rmp = oc.x[:, 0:1].clone(), oc.x[:, 1:2].clone(), oc.edge_attr[:, 2:5].clone()
    out = []
    with torch.no_grad():
        for _ in range(steps):
            uvp, nuv = orc(oc, oe, on, H.RHO, H.MU, H.DT, 9, 0)
            out.append((uvp, nuv))
            O.create_next_graph(oc, oe, mb, nuv, ct, Re, rmp)
    return out


def _cuda_rollout(fv, g, math, steps):
    model = _model(fv, g, math)
    gn, ge, gc = _graphs(fv, g)
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    state = fv.dataset.RolloutState(gc, ge)
    out = []
    with torch.no_grad():
        for _ in range(steps):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
            out.append((uvp.clone(), nuv.clone()))
            state.advance_(gc, nuv)
    return out, (model, gn, ge, gc)


@pytest.mark.parametrize("fixture,steps", [("fwd_tiny.npz", 3), ("fwd_cyl.npz", 10)])
def test_mp_times_1_rollout_all_math_modes(fv, golden, fixture, steps):
    g = golden(fixture)
    want = _oracle_rollout(g, steps)
    two = O.FVGNOracle(H.load_state(H.torch_reference_state_dict(0), g), H.params()).eval()  # mp_times=2 on the same inputs: a different answer
    ref_fixture = golden("fwd_tiny_mp1.npz") if fixture == "fwd_tiny.npz" else None  # the unmodified reference run with mp_times=1
    for math, tol in (("fp32", TOL), ("bf16x3", TOL), ("bf16", 3e-2)):
        got, _ = _cuda_rollout(fv, g, math, steps)
        for k, ((uvp, nuv), (ouvp, onuv)) in enumerate(zip(got, want)):
            assert H.rel_l2(nuv.cpu(), onuv) < tol * (k + 1), (math, k, H.rel_l2(nuv.cpu(), onuv))
            assert H.rel_l2(uvp.cpu(), ouvp) < tol * (k + 1), (math, k, H.rel_l2(uvp.cpu(), ouvp))
            if ref_fixture is not None:
                assert H.rel_l2(nuv.cpu(), ref_fixture["roll_uv"][k]) < 1.2 * tol * (k + 1), (math, k)
                assert H.rel_l2(uvp.cpu(), ref_fixture["roll_uvp"][k]) < 1.2 * tol * (k + 1), (math, k)
    on, oe, oc = O.build_graphs(H.tables_from(g, "tab."), g["velocity"], g["pressure"])
    on, oe, oc, _, _ = O.valid_datapreprocessing(on, oe, oc, 0)
    with torch.no_grad():
        _, nuv2 = two(oc, oe, on, H.RHO, H.MU, H.DT, 9, 0)
    assert H.rel_l2(nuv2, want[0][1]) > 1e-3  # the test would notice the mp_times=2 network


def test_mp_times_1_graphed_rollout_is_bit_identical(fv, golden):
    g = golden("fwd_cyl.npz")
    eager, (model, gn, ge, gc) = _cuda_rollout(fv, g, "bf16x3", 6)
    gn, ge, gc = _graphs(fv, g)
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    roll = fv.GraphedRollout(model, gn, ge, gc, H.RHO, H.MU, H.DT)
    for k in range(6):
        uvp, nuv = roll.step()
        assert torch.equal(uvp, eager[k][0]) and torch.equal(nuv, eager[k][1]), k


@pytest.mark.parametrize("math", ["fp32", "bf16x3"])
def test_mp_times_1_training_gradients(fv, golden, math):
    g = golden("train_tiny.npz")
    f = golden("fwd_tiny.npz")
    p = H.params(loss="global_mean_sum", mp_times=1)
    # oracle: torch-CPU autograd over the restated path
    orc = O.FVGNOracle(H.load_state(H.torch_reference_state_dict(0), f), p).train()
    for k, v in list(orc.sd.items()):
        if v.is_floating_point() and ("model." in k or k in ("_grid_feature_normalizer.weight", "_grid_feature_normalizer.bias")):
            orc.sd[k] = v.clone().requires_grad_(True)
    orc.bn["weight"], orc.bn["bias"] = orc.sd["_grid_feature_normalizer.weight"], orc.sd["_grid_feature_normalizer.bias"]
    b0 = O.build_graphs(H.tables_from(f, "tab."), f["velocity"], f["pressure"], training_pair=1)
    b1 = O.build_graphs(H.tables_from(f, "tab1."), f["velocity1"], f["pressure1"], training_pair=0)
    gn, ge, gc = [O.batch_graphs([b0[k], b1[k]]) for k in range(3)]
    gn, ge, gc, mi, mb = O.train_datapreprocessing(gn, ge, gc, torch.as_tensor(g["noise"]), torch.as_tensor(g["flip"]))
    out = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0)
    loss = O.compute_fvm_loss(p, gc.batch, ge.batch, mi, out[3], out[4], out[1], out[2], num_graphs=2)
    loss[0].backward()
    want = {k: v.grad for k, v in orc.sd.items() if v.requires_grad}
    # CUDA path
    model = _model(fv, f, math).train()
    gl = fv.dataset.collate([_graphs(fv, f, "tab.", "velocity", "pressure", pair=1), _graphs(fv, f, "tab1.", "velocity1", "pressure1", pair=0)])
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing(gl, cell_noise=dev(g["noise"]), flip_keep=dev(g["flip"]))
    out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    closs = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    closs[0].backward()
    assert abs(float(closs[0]) - float(loss[0])) < 2e-5
    named = dict(model.named_parameters())
    assert set(named) == set(want)
    worst = max(((k, H.rel_l2(named[k].grad.cpu(), w)) for k, w in want.items()), key=lambda kv: kv[1])
    assert worst[1] < 1e-4, worst


def test_mp_times_1_rejects_a_per_node_aggregate(fv):
    x = torch.zeros((8, 128), device="cuda")
    na = torch.zeros((5, 64), device="cuda")  # not per cell
    torch.manual_seed(0)
    m = fv.FVGN(device="cpu", params=H.params(mp_times=1)).cuda()
    with pytest.raises(ValueError):
        fv.ops.cell_block_fwd(m.model.processer_list[0].cb_module.net.mlp_ref(), x, na, None, 0)
