"""GPU (B200): tri / quad hybrid meshes in the kernels (BASELINE configs[3]: "600-step autoregressive rollout on a hybrid tri/quad
COMSOL-style mesh").  The reference is triangles only, so for k = 4 the oracle is oracle/fvgn_oracle_poly.py
(PARITY UNPINNED for quads by construction); what is pinned, and tested here on the device as well:

  * on ALL-TRIANGLE meshes the [C,4] (-1 padded) entry points give the triangle entry points' tables and rollouts bit for bit — and
    those are held to the reference (tests/test_gpu_parity.py)
  * on hybrid meshes: connectivity / geometry tables bit-equal to the polygon oracle; one step and a 600-step rollout of the strict
    and the fp32-equivalent path within fp32 rel 1e-5 per step of the oracle's torch-CPU rollout over the same [4,C] tables; bf16 tiles
    within their stated bound; the pre-step (interpolation with four weights) bit-equal
  * TRAINING on hybrid meshes: loss and all 264 parameter gradients of the strict and the fp32-equivalent backward against torch autograd
    over the polygon oracle (mp_times 2 and 1)
"""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from oracle import fvgn_oracle_poly as P
from tests import helpers as H
from tests.test_gpu_parity import fv, dev  # noqa: F401
from tests.test_gpu_rollout600 import cumulative_relative_error

pytestmark = pytest.mark.gpu
INT_KEYS = ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type"]
F32_KEYS = ["centroid", "cell_factor", "face_length", "unit_norm_v"]


def _gpu_tables(fv, m, rule="B", pad=False):
    cells = np.asarray(m["cells"])
    if pad and cells.shape[1] == 3:
        cells = np.concatenate([cells, np.full((cells.shape[0], 1), -1, dtype=cells.dtype)], 1)
    return fv.ops.build_connectivity(dev(cells, torch.int32), dev(m["mesh_pos"]), dev(m["node_type"], torch.int32), face_rule=rule)


@pytest.mark.parametrize("seed,qf,rule", [(0, 0.5, "B"), (1, 0.9, "B"), (2, 0.15, "A"), (3, 1.0, "B")])
def test_hybrid_connectivity_bit_equal_to_the_polygon_oracle(fv, seed, qf, rule):
    m = fv.synthetic.make("hybrid", seed=seed, quad_fraction=qf, nx=40, ny=23)
    t = _gpu_tables(fv, m, rule)
    o = P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], rule)
    assert (o["cell_k"] == 4).any()
    for k in INT_KEYS:
        got = t[k].cpu().numpy()
        assert got.dtype == np.int32 and np.array_equal(got, o[k]), k
    for k in F32_KEYS:
        assert np.array_equal(t[k].cpu().numpy(), o[k]), k
    np.testing.assert_allclose(t["cells_area"].cpu().numpy(), o["cells_area"], rtol=3e-7, atol=0)
    assert np.array_equal(t["cell_k"].cpu().numpy(), o["cell_k"])
    assert t["mesh_pos"].shape[0] - t["face"].shape[1] + t["cells_node"].shape[0] == 0  # Euler: one hole


@pytest.mark.parametrize("kind,seed,rule", [("cyl", 11, "B"), ("naca", 4, "A"), ("grid", 5, "B")])
def test_padded_triangle_tables_equal_the_triangle_entry_points(fv, kind, seed, rule):
    m = fv.synthetic.make(kind, seed=seed, **({"nx": 60, "ny": 41} if kind == "grid" else {}))
    a = _gpu_tables(fv, m, rule)
    b = _gpu_tables(fv, m, rule, pad=True)
    for k in ["face", "neighbour_cell", "face_type", "cells_type", "face_length", "centroid", "cells_area"]:
        assert torch.equal(a[k], b[k]), k
    for k in ["cells_node", "cells_face", "cell_factor", "unit_norm_v"]:
        assert torch.equal(a[k], b[k][:, :3]), k
    assert (b["cells_node"][:, 3] == -1).all() and (b["cells_face"][:, 3] == -1).all() and (b["cell_factor"][:, 3] == 0).all()


def _rollout_gpu(fv, tab, m, sd, math, steps, graphed=False):
    model = fv.FVGN(device="cpu", params=H.params())
    model.load_state_dict(sd)
    model = model.cuda().eval().set_math_mode(math)
    gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device="cuda")
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    x0, ea0 = gc.x.clone(), gc.edge_attr.clone()
    uv, p = [], []
    with torch.no_grad():
        if graphed:
            roll = fv.GraphedRollout(model, gn, ge, gc, H.RHO, H.MU, H.DT)
            for _ in range(steps):
                uvp, nuv = roll.step()
                uv.append(nuv.clone()), p.append(uvp[:, 2:3].clone())
            roll.check()
        else:
            state = fv.dataset.RolloutState(gc, ge)
            for _ in range(steps):
                uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
                uv.append(nuv.clone()), p.append(uvp[:, 2:3].clone())
                state.advance_(gc, nuv)
    return torch.stack(uv), torch.stack(p), x0, ea0


def _rollout_oracle(tab, m, sd, steps):
    on, oe, oc = O.build_graphs(tab, m["velocity"], m["pressure"])
    on, oe, oc, mi, mb = O.valid_datapreprocessing(on, oe, oc, 0)
    x0, ea0 = oc.x.clone(), oc.edge_attr.clone()
    orc = O.FVGNOracle(sd, H.params()).eval()
    ct, Re, rmp = oc.x[:, 0:1].clone(), oc.x[:, 1:2].clone(), oc.edge_attr[:, 2:5].clone()
    uv, p = [], []
    with torch.no_grad():
        for _ in range(steps):
            uvp, nuv = orc(oc, oe, on, H.RHO, H.MU, H.DT, 9, 0)
            uv.append(nuv.double()), p.append(uvp[:, 2:3].double())
            O.create_next_graph(oc, oe, mb, nuv, ct, Re, rmp)
    return torch.stack(uv), torch.stack(p), x0, ea0


@pytest.mark.parametrize("math", ["fp32", "bf16x3"])
def test_padded_triangle_rollout_is_bit_identical_to_the_triangle_path(fv, math):
    m = fv.synthetic.make("cyl", seed=3, T=3, n_target=700)
    sd = H.torch_reference_state_dict(0)
    a = _gpu_tables(fv, m)
    b = _gpu_tables(fv, m, pad=True)
    ua, pa, xa, ea = _rollout_gpu(fv, a, m, sd, math, 6)
    ub, pb, xb, eb = _rollout_gpu(fv, b, m, sd, math, 6)
    assert torch.equal(xa, xb) and torch.equal(ea, eb)
    assert torch.equal(ua, ub) and torch.equal(pa, pb)


def test_hybrid_rollout_600_steps_vs_the_polygon_oracle(fv):
    """configs[3]: 600 autoregressive steps on a tri / quad hybrid mesh, rollout-error metric of rollout.py:446-482"""
    m = fv.synthetic.make("hybrid", seed=5, T=3, nx=34, ny=19, quad_fraction=0.5)
    sd = H.torch_reference_state_dict(0)
    otab = P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], "B")
    assert 0.2 < float((otab["cell_k"] == 4).mean()) < 0.8
    ouv, op, ox0, oea0 = _rollout_oracle(otab, m, sd, 600)
    assert torch.isfinite(ouv).all()
    tab = _gpu_tables(fv, m)
    for math in ("fp32", "bf16x3"):
        guv, gp, x0, ea0 = _rollout_gpu(fv, tab, m, sd, math, 600, graphed=(math == "bf16x3"))
        assert np.array_equal(x0.cpu().numpy(), ox0.numpy()) and np.array_equal(ea0.cpu().numpy(), oea0.numpy())  # 4-weight interpolation: same bits
        guv, gp = guv.double().cpu(), gp.double().cpu()
        c_uv, c_p = cumulative_relative_error(guv, ouv, gp, op)
        per_step = ((guv - ouv) ** 2).sum(dim=(1, 2)).sqrt() / (ouv ** 2).sum(dim=(1, 2)).sqrt()
        print("hybrid %s 600-step rollout: sqrt(cum rel err) uv %.2e p %.2e, step 1 rel-L2 %.2e, worst per-step %.2e"
              % (math, c_uv[-1].sqrt(), c_p[-1].sqrt(), per_step[0], per_step.max()))
        assert float(per_step[0]) < 1e-5
        assert float(c_uv[-1].sqrt()) < 1e-4 and float(c_p[-1].sqrt()) < 1e-4 and float(per_step.max()) < 1e-4
    ref_uv, ref_p = guv, gp
    buv, bp, _, _ = _rollout_gpu(fv, tab, m, sd, "bf16", 600)
    b_uv, b_p = cumulative_relative_error(buv.double().cpu(), ref_uv, bp.double().cpu(), ref_p)
    print("hybrid bf16 600-step rollout vs fp32-equivalent: sqrt(cum rel err) uv %.2e p %.2e" % (b_uv[-1].sqrt(), b_p[-1].sqrt()))
    assert float(b_uv[-1].sqrt()) < 3e-2 and float(b_p[-1].sqrt()) < 6e-2


def test_hybrid_batch_and_mp_times_1(fv):
    """two hybrid meshes batched block-diagonally (the -1 padding survives the PyG offsets) and the mp_times = 1 CellBlock variant"""
    ms = [fv.synthetic.make("hybrid", seed=s, T=3, nx=20 + s, ny=13, quad_fraction=0.6) for s in (1, 2)]
    sd = H.torch_reference_state_dict(0)
    for mp in (2, 1):
        p = H.params(mp_times=mp)
        lists, bags = [], []
        for m in ms:
            lists.append(fv.dataset.build_graphs(_gpu_tables(fv, m), m["velocity"], m["pressure"], device="cuda"))
            bags.append(O.build_graphs(P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], "B"), m["velocity"], m["pressure"]))
        gn, ge, gc = fv.dataset.collate(lists)
        assert int((gn.face[3] < 0).sum()) > 0 and int(gn.face[3].max()) >= lists[0][0].x.shape[0]
        gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
        on, oe, oc = [O.batch_graphs([b[k] for b in bags]) for k in range(3)]
        on, oe, oc, _, _ = O.valid_datapreprocessing(on, oe, oc, 0)
        orc = O.FVGNOracle(sd, p).eval()
        with torch.no_grad():
            ouvp, onuv = orc(oc, oe, on, H.RHO, H.MU, H.DT, 9, 0)
        for math in ("fp32", "bf16x3"):
            model = fv.FVGN(device="cpu", params=p)
            model.load_state_dict(sd)
            model = model.cuda().eval().set_math_mode(math)
            g2 = gc.clone()
            with torch.no_grad():
                uvp, nuv = model(graph=g2, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
            e1, e2 = H.rel_l2(uvp.cpu(), ouvp), H.rel_l2(nuv.cpu(), onuv)
            assert e1 < 1e-5 and e2 < 1e-5, (mp, math, e1, e2)


def _oracle_hybrid_grads(ms, sd, p, noise, flip):
    orc = O.FVGNOracle(sd, p).train()
    for k, v in list(orc.sd.items()):
        if v.is_floating_point() and ("model." in k or k in ("_grid_feature_normalizer.weight", "_grid_feature_normalizer.bias")):
            orc.sd[k] = v.clone().requires_grad_(True)
    orc.bn["weight"], orc.bn["bias"] = orc.sd["_grid_feature_normalizer.weight"], orc.sd["_grid_feature_normalizer.bias"]
    bags = [O.build_graphs(P.build_connectivity_poly(m["mesh_pos"], m["cells"], m["node_type"], "B"), m["velocity"], m["pressure"], training_pair=0)
            for m in ms]
    gn, ge, gc = [O.batch_graphs([b[k] for b in bags]) for k in range(3)]
    gn, ge, gc, mi, mb = O.train_datapreprocessing(gn, ge, gc, noise, flip)
    out = orc(gc, ge, gn, H.RHO, H.MU, H.DT, 9, 0)
    loss = O.compute_fvm_loss(p, gc.batch, ge.batch, mi, out[3], out[4], out[1], out[2], num_graphs=len(ms))
    loss[0].backward()
    return float(loss[0].detach()), {k: v.grad for k, v in orc.sd.items() if v.requires_grad}


@pytest.mark.parametrize("mp", [2, 1])
def test_training_step_on_hybrid_meshes_vs_oracle_autograd(fv, mp):
    """training on [4,C] tables (no reference counterpart: triangles only): loss and all 264 parameter gradients of the strict and the
    fp32-equivalent backward against torch autograd over the polygon oracle, on a batch of two tri / quad meshes"""
    ms = [fv.synthetic.make("hybrid", seed=s, T=3, nx=20 + s, ny=13, quad_fraction=0.5) for s in (1, 2)]
    sd = H.torch_reference_state_dict(0)
    p = H.params(loss="global_mean_sum", mp_times=mp)
    lists = [fv.dataset.build_graphs(_gpu_tables(fv, m), m["velocity"], m["pressure"], training_pair=0, device="cuda") for m in ms]
    C = sum(l[2].x.shape[0] for l in lists)
    F = sum(l[1].x.shape[0] for l in lists)
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(C, 2, generator=g) * 0.02
    flip = torch.rand(F, generator=g) < 0.5
    want_loss, want = _oracle_hybrid_grads(ms, sd, p, noise, flip)
    for math, tol in (("fp32", 1e-4), ("bf16x3", 2e-4)):
        model = fv.FVGN(device="cpu", params=p)
        model.load_state_dict(sd)
        model = model.cuda().train().set_math_mode(math)
        gl = fv.dataset.collate([[b.clone() for b in l] for l in lists])
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing(gl, cell_noise=noise.cuda(), flip_keep=flip.cuda())
        out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        got_loss = float(loss[0].detach())
        assert abs(got_loss - want_loss) < 5e-5 * max(1.0, abs(want_loss)), (math, got_loss, want_loss)
        named = dict(model.named_parameters())
        assert set(named) == set(want)
        worst = ("", 0.0)
        for k, w in want.items():
            err = H.rel_l2(named[k].grad.cpu(), w)
            if err > worst[1]:
                worst = (k, err)
        print(f"hybrid training step mp_times={mp} {math}: worst relative L2 over {len(want)} gradients: {worst}")
        assert worst[1] < tol, (math, worst)
    assert fv.ops.status(clear=True) == 0
