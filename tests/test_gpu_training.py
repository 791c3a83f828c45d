"""GPU (B200): the training path — forward 5-tuple, loss, and EVERY parameter gradient of the CUDA backward against
(a) torch-CPU autograd over the oracle and (b) the gradients the unmodified reference produced (tests/golden/train_tiny.npz).
Tolerance: relative L2 <= 1e-4 per parameter tensor (fp32 backward; summation order differs from torch's)."""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H
from tests.test_gpu_parity import fv, dev, _model, _graphs  # noqa: F401

pytestmark = pytest.mark.gpu
GTOL = 1e-4


def _oracle_grads(f, g, loss_name):
    p = H.params(loss=loss_name)
    sd = H.load_state(H.torch_reference_state_dict(0), f)
    orc = O.FVGNOracle(sd, p).train()
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
    return float(loss[0]), {k: v.grad for k, v in orc.sd.items() if v.requires_grad}


@pytest.mark.parametrize("loss_name,tag", [("direct_mean_loss", "dm"), ("global_mean_sum", "gm")])
def test_backward_matches_oracle_and_reference(fv, golden, loss_name, tag):
    g = golden("train_tiny.npz")
    f = golden("fwd_tiny.npz")
    want_loss, want = _oracle_grads(f, g, loss_name)
    model = _model(fv, f).train()
    gl = fv.dataset.collate([_graphs(fv, f, "tab.", "velocity", "pressure", pair=1), _graphs(fv, f, "tab1.", "velocity1", "pressure1", pair=0)])
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing(gl, cell_noise=dev(g["noise"]), flip_keep=dev(g["flip"]))
    p = H.params(loss=loss_name)
    out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                               target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    loss[0].backward()
    assert abs(float(loss[0]) - want_loss) < 2e-5 and abs(float(loss[0]) - g[tag + ".loss"][0]) < 2e-5
    named = dict(model.named_parameters())
    assert set(named) == set(want)
    worst = ("", 0.0)
    for k, w in want.items():
        got = named[k].grad
        assert got is not None, k
        err = H.rel_l2(got.cpu(), w)
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < GTOL, worst
    # the reference's own gradients (fixture): norms of all 266 tensors + a few full tensors
    keys = [str(k) for k in g[tag + ".gkeys"]]
    gnorm = np.array([float(named[k].grad.double().norm()) for k in keys])
    np.testing.assert_allclose(gnorm, g[tag + ".gnorm"], rtol=2e-4, atol=1e-9)
    for k in g:
        if k.startswith(tag + ".grad."):
            assert H.rel_l2(named[k[len(tag) + 6:]].grad.cpu(), g[k]) < GTOL, k


def test_backward_is_deterministic_and_optimizer_step_runs(fv, golden):
    g = golden("train_tiny.npz")
    f = golden("fwd_tiny.npz")
    p = H.params()
    grads = []
    for _ in range(2):
        model = _model(fv, f).train()
        gl = fv.dataset.collate([_graphs(fv, f, "tab.", "velocity", "pressure", pair=1), _graphs(fv, f, "tab1.", "velocity1", "pressure1", pair=0)])
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing(gl, cell_noise=dev(g["noise"]), flip_keep=dev(g["flip"]))
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        opt.zero_grad()
        out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        grads.append([q.grad.clone() for q in model.parameters()])
        opt.step()
    for a, b in zip(*grads):
        assert torch.equal(a, b)  # bitwise reproducible gradients (fixed-order reductions, no float atomics)


def test_mlp_rows_backward_shapes(fv):
    """ragged sizes: rows not a multiple of the tile, odd k_in, decoder-shaped output (n_out=5, no LayerNorm)"""
    for k_in, n_out, ln, rows in [(14, 128, True, 150), (2, 128, True, 65), (128, 5, False, 333), (37, 128, True, 64)]:
        torch.manual_seed(1)
        m = fv.build_mlp(k_in, 128, n_out, drop_out=False, lay_norm=ln)
        ref = torch.nn.Sequential(*[torch.nn.Linear(k_in, 128), torch.nn.SiLU(), torch.nn.Linear(128, 128), torch.nn.SiLU(), torch.nn.Linear(128, n_out)])
        seq = m[0] if ln else m
        for a, b in zip((ref[0], ref[2], ref[4]), (seq[0], seq[2], seq[4])):
            a.load_state_dict(b.state_dict())
        x = torch.randn(rows, k_in, generator=torch.Generator().manual_seed(3), requires_grad=True)
        gout = torch.randn(rows, n_out, generator=torch.Generator().manual_seed(4))
        y = ref(x.double() if False else x)
        if ln:
            lnm = torch.nn.LayerNorm(n_out)
            lnm.load_state_dict(m[1].state_dict())
            y = lnm(y)
        y.backward(gout)
        mc = m.cuda()
        grads, d_in = fv.ops.mlp_rows_bwd(mc.mlp_ref(), x.detach().cuda(), gout.cuda(), want_d_in=True)
        want = [ref[0].weight.grad, ref[0].bias.grad, ref[2].weight.grad, ref[2].bias.grad, ref[4].weight.grad, ref[4].bias.grad]
        if ln:
            want += [lnm.weight.grad, lnm.bias.grad]
        for gg, w in zip(grads, want):
            assert H.rel_l2(gg.cpu(), w) < GTOL, (k_in, n_out, gg.shape, H.rel_l2(gg.cpu(), w))
        assert H.rel_l2(d_in.cpu(), x.grad) < GTOL


def test_graphed_train_step_matches_eager_steps(fv, golden):
    """the whole training step replayed from one CUDA graph (fvgn_b200.GraphedTrainStep) == the eager loop of train.py:150-225
    (fixed noise / flip draws so that both see the same pre-processed batch)"""
    g = golden("train_tiny.npz")
    f = golden("fwd_tiny.npz")
    p = H.params(loss="global_mean_sum")
    pre = dict(cell_noise=dev(g["noise"]), flip_keep=dev(g["flip"]))

    def batch():
        return fv.dataset.collate([_graphs(fv, f, "tab.", "velocity", "pressure", pair=1), _graphs(fv, f, "tab1.", "velocity1", "pressure1", pair=0)])

    model = _model(fv, f).train().set_math_mode("bf16x3")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=True, fused=True)  # graphed: the fused multi-tensor AdamW; the eager twin below: the foreach one
    gn0, ge0, gc0 = batch()
    step = fv.GraphedTrainStep(model, opt, p, gn0, ge0, gc0, H.RHO, H.MU, H.DT, preprocess_kwargs=pre, warmup=2)  # 2 real steps (capture records, it does not run)
    replayed = [float(step.step()[0]) for _ in range(3)]                                                           # steps 3..5
    model2 = _model(fv, f).train().set_math_mode("bf16x3")
    opt2 = torch.optim.AdamW(model2.parameters(), lr=1e-3, capturable=True)
    eager = []
    for _ in range(5):
        gn, ge, gc = batch()
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc], **pre)
        opt2.zero_grad()
        out = model2(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        opt2.step()
        eager.append(float(loss[0]))
    np.testing.assert_allclose(replayed, eager[2:5], rtol=1e-5)
    for (k, a), (_, b) in zip(model.named_parameters(), model2.named_parameters()):
        assert H.rel_l2(a.detach().cpu(), b.detach().cpu()) < 1e-5, k
