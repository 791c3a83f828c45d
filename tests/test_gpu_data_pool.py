"""GPU (B200): the dataset layer and the converters against the UNMODIFIED reference (oracle/_ref travels to the GPU box).

  * Data_Pool.require_minibatch_mesh / require_one_mesh / train_datapreprocessing / create_next_graph (this package, device kernels)
    == the reference's Data_Pool (torch CPU) on the same H5-layout groups: bit for bit (the pre / post step is FMA-free)
  * the loop body of the reference's rollout.py:368-425 and train.py:150-225 pointed at THIS package's Data_Pool + FVGN, next to the
    reference's own objects: outputs within fp32 rel 1e-5 per step; and the REFERENCE's own Data_Pool / Batch objects fed to fv.FVGN
  * convert.extract_mesh_state (device connectivity builder) == the reference's extract_mesh_state: every key of the H5 group
  * TFRecord -> store -> Data_Pool -> rollout end to end
"""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import ref_harness as RH
from oracle import ref_loader as RL
from tests import helpers as H
from tests.test_data_pool_cpu import _groups, write_synthetic_tfrecord
from tests.test_gpu_parity import fv, dev  # noqa: F401

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RL.available(), reason="reference copy (oracle/_ref) not present")]
TOL = 1e-5


def _pools(fv, groups, training, L=3):
    p = H.params(train_traj_length=L, dataset_size=len(groups))
    with RH._cpu_mode(True):
        ref = RL.ref_data_pool(groups, RL.default_params(train_traj_length=L, dataset_size=len(groups)), is_training=training, device="cpu")
    ref._set_status(is_training=training)
    mine = fv.dataset.Data_Pool(params=p, is_traning=training, device="cuda")
    mine.pool = {str(i): g for i, g in enumerate(groups)}
    return mine, ref


def _eq(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    assert torch.equal(a.cpu(), b.cpu()), what


@pytest.mark.parametrize("start", [0, 1])
def test_require_minibatch_mesh_bit_equal_to_the_reference(fv, start):
    groups, _ = _groups(n=3, T=4)
    mine, ref = _pools(fv, groups, False)
    got = mine.require_minibatch_mesh(start_epoch=start, batch_index=[0, 1, 2], is_training=False)
    with RH._cpu_mode(True), contextlib.redirect_stdout(io.StringIO()):
        want = ref.require_minibatch_mesh(start_epoch=start, batch_index=[0, 1, 2], is_training=False, mode="h5")
    for (a, b, nm) in zip(got[:3], want[:3], ("node", "edge", "cell")):
        for k in b.keys:
            vb = getattr(b, k)
            if torch.is_tensor(vb) and k not in ("_slices",):
                _eq(getattr(a, k), vb, f"{nm}.{k}")
    _eq(got[3], want[3], "mask_face_interior"), _eq(got[4], want[4], "mask_face_boundary")
    assert got[2].x.is_cuda and got[2].edge_attr.shape[1] == 5
    mine.reset()
    one = mine.require_one_mesh(start, 1)
    with RH._cpu_mode(True), contextlib.redirect_stdout(io.StringIO()):
        ref.reset()
        wone = ref.require_one_mesh(start, 1)
    _eq(one[2].x, wone[2].x, "one.x"), _eq(one[2].edge_attr, wone[2].edge_attr, "one.edge_attr"), _eq(one[1].y, wone[1].y, "one.edge.y")


def test_train_datapreprocessing_bit_equal_to_the_reference(fv):
    groups, _ = _groups(n=3, T=4)
    mine, ref = _pools(fv, groups, True)
    refL = RL.load()
    idx = [0, 4, 8]
    mine_list = fv.dataset.collate([mine.get(i) for i in idx])
    ref_list = [refL.Batch.from_data_list([ref.get(i)[k] for i in idx]) for k in range(3)]
    C, F = ref_list[2].x.shape[0], ref_list[1].x.shape[0]
    gen = torch.Generator().manual_seed(9)
    noise = torch.randn(C, 2, generator=gen) * 0.02
    flip = torch.rand(F, generator=gen) < 0.5
    # both sides consume the SAME draws (the reference: utils/noise.py:13-26 and Data_Pool.randbool, Load_mesh.py:442-446, 1046)
    mine.get_noise = lambda gl: list(gl) + [noise.cuda()]
    mine.randbool = lambda *size: flip.cuda()
    ref.randbool = lambda *size: flip.view(1, -1)
    saved = refL.Load_mesh.noise.get_v_noise_on_cell
    refL.Load_mesh.noise.get_v_noise_on_cell = lambda graph, noise_std, outchannel, device: noise.clone()
    try:
        with RH._cpu_mode(True):
            want = ref.train_datapreprocessing(ref_list, dual_edge=False)
    finally:
        refL.Load_mesh.noise.get_v_noise_on_cell = saved
    got = mine.train_datapreprocessing(mine_list, dual_edge=False)
    _eq(got[2].x, want[2].x, "cell.x"), _eq(got[2].edge_attr, want[2].edge_attr, "cell.edge_attr"), _eq(got[2].y, want[2].y, "cell.y")
    _eq(got[1].y, want[1].y, "edge.y"), _eq(got[3], want[3], "mask_interior"), _eq(got[4], want[4], "mask_boundary")
    assert got[2].x.is_cuda


def _ref_model(params):
    refL = RL.load()
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)
        return refL.FVGN(device="cpu", params=params)


def test_reference_rollout_loop_pointed_at_this_package(fv):
    """rollout.py:252-263, 368-425 with `datasets` / `model` = this package's Data_Pool / FVGN, next to the reference's own objects"""
    groups, _ = _groups(n=2, T=5)
    mine, ref = _pools(fv, groups, False, L=4)
    pr = RL.default_params(train_traj_length=4, dataset_size=2)
    ref_model = _ref_model(pr).eval()
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=H.params(train_traj_length=4, dataset_size=2))
    model.load_state_dict(ref_model.state_dict())
    model = model.cuda().eval()
    args = dict(rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    for mode in ("fp32", "bf16x3"):
        model.set_math_mode(mode)
        mine.reset()
        gn, ge, g_old, mi, mb = mine.require_minibatch_mesh(start_epoch=0, batch_index=[0, 1], is_training=False)
        with RH._cpu_mode(True), contextlib.redirect_stdout(io.StringIO()):
            ref.reset()
            rn, re_, r_old, rmi, rmb = ref.require_minibatch_mesh(start_epoch=0, batch_index=[0, 1], is_training=False, mode="h5")
        cell_type, Re, rmp = g_old.x[:, 0:1].clone(), g_old.x[:, 1:2].clone(), g_old.edge_attr[:, 2:5].clone()
        rct, rRe, rrmp = r_old.x[:, 0:1].clone(), r_old.x[:, 1:2].clone(), r_old.edge_attr[:, 2:5].clone()
        with torch.no_grad():
            for step in range(3):
                uvp, nuv = model(graph=g_old, graph_edge=ge, graph_node=gn, device="cuda", **args)
                with RH._cpu_mode(True):
                    ruvp, rnuv = ref_model(graph=r_old, graph_edge=re_, graph_node=rn, device="cpu", **args)
                e1, e2 = H.rel_l2(uvp.cpu(), ruvp), H.rel_l2(nuv.cpu(), rnuv)
                assert e1 < TOL * (step + 1) and e2 < TOL * (step + 1), (mode, step, e1, e2)
                g_old = mine.create_next_graph(g_old, ge, mb, nuv, cell_type, Re, rmp, dual_edge=False)
                with RH._cpu_mode(True):
                    r_old = ref.create_next_graph(r_old, re_, rmb, rnuv, rct, rRe, rrmp, dual_edge=False)
        assert fv.ops.status(clear=True) == 0


def test_the_references_own_data_pool_objects_feed_fv_FVGN(fv):
    """the reference's Data_Pool + its (shim) PyG Batch objects, built on the GPU by the reference's own code, go straight into fv.FVGN:
    eval forward and one training step (train.py:150-218) — against the reference model on the CPU"""
    groups, _ = _groups(n=2, T=4)
    pr = RL.default_params(train_traj_length=3, dataset_size=2, loss="global_mean_sum")
    refL = RL.load()
    with contextlib.redirect_stdout(io.StringIO()):
        dpc = RL.ref_data_pool(groups, pr, is_training=False, device="cuda")
        gn, ge, gc, mi, mb = dpc.require_minibatch_mesh(start_epoch=0, batch_index=[0, 1], is_training=False, mode="h5")
    assert gc.x.is_cuda and isinstance(gc, refL.Batch)
    ref_model = _ref_model(pr).eval()
    model = fv.FVGN(device="cpu", params=H.params(loss="global_mean_sum"))
    model.load_state_dict(ref_model.state_dict())
    model = model.cuda().eval().set_math_mode("fp32")
    args = dict(rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    with RH._cpu_mode(True), contextlib.redirect_stdout(io.StringIO()):
        dpr = RL.ref_data_pool(groups, pr, is_training=False, device="cpu")
        rn, re_, rc, _, _ = dpr.require_minibatch_mesh(start_epoch=0, batch_index=[0, 1], is_training=False, mode="h5")
        with torch.no_grad():
            ruvp, rnuv = ref_model(graph=rc, graph_edge=re_, graph_node=rn, device="cpu", **args)
    with torch.no_grad():
        uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, device="cuda", **args)
    assert H.rel_l2(uvp.cpu(), ruvp) < TOL and H.rel_l2(nuv.cpu(), rnuv) < TOL
    # one training step on the reference's training batch objects
    with contextlib.redirect_stdout(io.StringIO()):
        dpt = RL.ref_data_pool(groups, pr, is_training=True, device="cuda")
    dpt._set_status(is_training=True)
    idx = [0, 4]
    graph_list = [refL.Batch.from_data_list([dpt.get(i)[k] for i in idx]) for k in range(3)]
    torch.manual_seed(1)
    tn, te, tc, tmi, tmb = dpt.train_datapreprocessing(graph_list, dual_edge=False)  # the reference's own pre-step, on the GPU
    keep = {k: getattr(tc, k).clone() for k in ("x", "edge_attr", "y")}
    model.train()
    out = model(graph=tc, graph_edge=te, graph_node=tn, device="cuda", **args)
    loss = fv.compute_FVM_loss(params=pr, graph_cell=tc, graph_edge=te, mask_face_interior=tmi, predicted_edge_attr=out[3],
                               target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    loss[0].backward()
    ref_model.train()
    with RH._cpu_mode(True):
        cn, ce, cc = [b.clone().to("cpu") for b in (tn, te, tc)]
        for k, v in keep.items():
            setattr(cc, k, v.cpu())
        rout = ref_model(graph=cc, graph_edge=ce, graph_node=cn, device="cpu", **args)
        rloss = refL.compute_FVM_loss(params=pr, graph_cell=cc, graph_edge=ce, mask_face_interior=tmi.cpu(), predicted_edge_attr=rout[3],
                                      target_face_uvp_normalized=rout[4], predicted_delta_u=rout[1], target_delta_u=rout[2],
                                      loss_function=torch.nn.MSELoss())
        rloss[0].backward()
    assert abs(float(loss[0]) - float(rloss[0])) < 2e-5
    worst = max(H.rel_l2(a.grad.cpu(), b.grad) for a, b in zip(model.parameters(), ref_model.parameters()))
    assert worst < 1e-4, worst


@pytest.mark.parametrize("kind,rule_name", [("grid", "incompressible"), ("cyl", "incomp-flow")])
def test_convert_extract_mesh_state_matches_the_reference(fv, kind, rule_name):
    """the device converter's H5 group == the group the reference's extract_mesh_state (Python dict loops) writes for the same raw trajectory;
    the name selects the face-typing branch like the reference (:491: "compressible" in name)"""
    from fvgn_b200 import convert as CV

    m = fv.synthetic.make(kind, seed=3, T=3, **({"nx": 14, "ny": 9} if kind == "grid" else {}))
    raw = {"mesh_pos": m["mesh_pos"][None], "cells": m["cells"][None], "node_type": m["node_type"].reshape(1, -1, 1), "velocity": m["velocity"],
           "pressure": m["pressure"]}
    w = fv.dataset.MeshStore("unused", "w")
    got = CV.extract_mesh_state(raw, w, 0, {"name": rule_name, "rho": 1.0, "mu": 1e-3, "dt": 0.01})
    with RH._cpu_mode(True):
        want = RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"],
                                         face_rule="A" if "compressible" in rule_name else "B")
    assert sorted(got) == sorted(want)
    for k, v in want.items():
        g = np.asarray(got[k])
        assert g.shape == np.asarray(v).shape, (k, g.shape, np.asarray(v).shape)
        if np.issubdtype(np.asarray(v).dtype, np.integer):
            assert g.dtype == v.dtype and np.array_equal(g, v), k
        elif k in ("cells_area", "aoa", "mean_u", "charac_scale", "reynolds_num"):
            np.testing.assert_allclose(g, v, rtol=1e-5, atol=1e-7, err_msg=k)
        else:
            assert np.array_equal(g, v), k
    assert sorted(w["0"].keys()) == sorted(want)


def test_tfrecord_to_store_to_rollout_end_to_end(fv, tmp_path):
    from fvgn_b200 import convert as CV

    meshes = [fv.synthetic.make("cyl", seed=s, T=4) for s in range(2)]
    src, out = tmp_path / "tf", tmp_path / "h5"
    src.mkdir()
    write_synthetic_tfrecord(str(src), meshes, split="valid")
    counts = CV.convert_tfrecord(str(src), str(out), splits=("valid",), solving_params={"name": "incomp-flow", "rho": 1.0, "mu": 1e-3, "dt": 0.01})
    assert counts == {"valid": 2}
    p = H.params(train_traj_length=3, dataset_size=2, dataset_dir_h5=str(out))
    dp = fv.dataset.Data_Pool(params=p, is_traning=False, device="cuda")
    dp.load_mesh_to_cpu(mode="h5", split="valid")
    gn, ge, gc, mi, mb = dp.require_minibatch_mesh(start_epoch=0, batch_index=[0, 1], is_training=False)
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=p).cuda().eval()
    roll = fv.GraphedRollout(model, gn, ge, gc, H.RHO, H.MU, H.DT)
    uvp, nuv = roll.run(3, check=True)
    assert torch.isfinite(nuv).all() and nuv.shape == (gc.x.shape[0], 2)
