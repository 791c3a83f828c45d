"""GPU (B200) parity tests: every CUDA entry point of libfvgn_b200, called through the C-ABI, against the oracle
(oracle/fvgn_oracle.py) and against the committed fixtures of the unmodified reference (tests/golden).

Tolerances (stated by BASELINE.json north_star): integer tables bit-exact; fp32 path rel-L2 <= 1e-5 per step.
"""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = 1e-5  # fp32 relative L2 per step (north_star)


@pytest.fixture(scope="module")
def fv():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fvgn_b200

    fvgn_b200._lib.lib()  # raises if the extension is missing: no silent fallback
    return fvgn_b200


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)) if not torch.is_tensor(a) else a
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("name", ["conn_tiny_A", "conn_tiny_B", "conn_cyl_A", "conn_cyl_B", "conn_foil_B"])
def test_connectivity_bit_exact(fv, golden, name):
    g = golden(name + ".npz")
    t = fv.ops.build_connectivity(dev(g["in_cells"], torch.int32), dev(g["in_mesh_pos"]), dev(g["in_node_type"], torch.int32), face_rule=name[-1])
    for k in ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type"]:
        got = t[k].cpu().numpy()
        assert got.dtype == np.int32 and np.array_equal(got, g[k]), k
    for k in ["centroid", "cell_factor", "face_length", "unit_norm_v"]:
        assert np.array_equal(t[k].cpu().numpy(), g[k]), k  # bit-exact fp32
    np.testing.assert_allclose(t["cells_area"].cpu().numpy(), g["cells_area"], rtol=3e-7, atol=0)


def test_connectivity_large_vs_oracle(fv):
    m = fv.synthetic.make("grid", seed=5, nx=300, ny=200)  # ~119k cells
    t = fv.ops.build_connectivity(dev(m["cells"], torch.int32), dev(m["mesh_pos"]), dev(m["node_type"], torch.int32), face_rule="B")
    o = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    for k in ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type", "centroid", "cell_factor", "face_length", "unit_norm_v"]:
        assert np.array_equal(t[k].cpu().numpy(), o[k]), k
    np.testing.assert_allclose(t["cells_area"].cpu().numpy(), o["cells_area"], rtol=3e-7)
    # Euler characteristic of a planar mesh with one hole: N - F + C = 0
    assert t["mesh_pos"].shape[0] - t["face"].shape[1] + t["cells_node"].shape[0] == 0


@pytest.mark.parametrize("kind,seed,rule", [("rect", 3, "B"), ("naca", 3, "B"), ("naca", 4, "A"), ("cyl", 11, "A")])
def test_connectivity_hybrid_geometries_vs_oracle(fv, kind, seed, rule):
    """the three geometries of the reference's hybrid COMSOL set (cylinder / rect / NACA0012 in a channel, all-triangle); the oracle is
    held to the live reference on the same meshes in tests/test_oracle_vs_reference.py"""
    m = fv.synthetic.make(kind, seed=seed)
    t = fv.ops.build_connectivity(dev(m["cells"], torch.int32), dev(m["mesh_pos"]), dev(m["node_type"], torch.int32), face_rule=rule)
    o = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], rule)
    for k in ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type", "centroid", "cell_factor", "face_length", "unit_norm_v"]:
        assert np.array_equal(t[k].cpu().numpy(), o[k]), k
    np.testing.assert_allclose(t["cells_area"].cpu().numpy(), o["cells_area"], rtol=3e-7)
    assert t["mesh_pos"].shape[0] - t["face"].shape[1] + t["cells_node"].shape[0] == 0


def test_csr_is_stable_sort(fv):
    gen = torch.Generator().manual_seed(0)
    for n_items, n_dst in [(1, 1), (17, 5), (10000, 37), (200000, 70001)]:
        dst = torch.randint(0, n_dst, (n_items,), generator=gen, dtype=torch.int32)
        ptr, items = fv.ops.build_csr(dst.cuda(), n_dst)
        order = np.argsort(dst.numpy(), kind="stable")
        assert np.array_equal(items.cpu().numpy(), order.astype(np.int32))
        cnt = np.bincount(dst.numpy(), minlength=n_dst)
        assert np.array_equal(ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32))


# ------------------------------------------------------------------------------------------------ a6 pieces
def _tiny(golden):
    g = golden("fwd_tiny.npz")
    tab = H.tables_from(g)
    return g, tab


def test_node_aggregate_bit_exact(fv, golden):
    g, tab = _tiny(golden)
    F, N = tab["face"].shape[1], tab["mesh_pos"].shape[0]
    e = torch.randn(F, 128, generator=torch.Generator().manual_seed(1))
    ptr, items = fv.ops.build_csr(dev(tab["face"], torch.int32).reshape(-1), N)
    got = fv.ops.node_aggregate_fwd(e.cuda(), ptr, items, N).cpu()
    idx = torch.as_tensor(tab["face"]).long().reshape(-1)
    want = torch.zeros(N, 64).index_add_(0, idx, torch.cat(torch.chunk(e, 2, -1), 0))
    assert torch.equal(got, want)  # same summation order as the reference's scatter_add => same bits


def _mlp_module(fv, k_in, n_out=128, ln=True, seed=0):
    torch.manual_seed(seed)
    m = fv.build_mlp(k_in, 128, n_out, drop_out=False, lay_norm=ln)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))  # non-trivial biases / LN affine
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.cuda(), sd


@pytest.mark.parametrize("k_in,n_out,ln,rows", [(2, 128, True, 77), (14, 128, True, 150), (128, 5, False, 150), (128, 128, True, 1),
                                                (192, 128, True, 64), (384, 128, True, 65), (37, 128, True, 200)])
def test_mlp_rows(fv, k_in, n_out, ln, rows):
    m, sd = _mlp_module(fv, k_in, n_out, ln)
    x = torch.randn(rows, k_in, generator=torch.Generator().manual_seed(2))
    got = m(x.cuda()).cpu()
    want = O.mlp(sd, "", x.double() if False else x, layer_norm=ln)
    want64 = O.mlp({k: v.double() for k, v in sd.items()}, "", x.double(), layer_norm=ln)
    assert H.rel_l2(got, want64) < 2e-6, H.rel_l2(got, want64)
    assert H.rel_l2(got, want) < 2e-6


def test_cell_and_edge_block(fv, golden):
    g, tab = _tiny(golden)
    C, F, N = tab["cells_node"].shape[0], tab["face"].shape[1], tab["mesh_pos"].shape[0]
    gen = torch.Generator().manual_seed(3)
    x, e = torch.randn(C, 128, generator=gen), torch.randn(F, 128, generator=gen)
    cb, sdc = _mlp_module(fv, 192, seed=4)
    eb, sde = _mlp_module(fv, 384, seed=5)
    fn = torch.as_tensor(np.ascontiguousarray(tab["face"])).long()
    cnT = torch.as_tensor(tab["cells_node"]).long().T.contiguous()
    nc = torch.as_tensor(np.ascontiguousarray(tab["neighbour_cell"])).long()
    sd64 = lambda sd: {k: v.double() for k, v in sd.items()}
    xp_ref = O.cell_block(sd64({"net." + k: v for k, v in sdc.items()}), "", x.double(), e.double(), fn, cnT, N)
    ep_ref = O.edge_block(sd64({"net." + k: v for k, v in sde.items()}), "", xp_ref, e.double(), nc)
    ptr, items = fv.ops.build_csr(dev(tab["face"], torch.int32).reshape(-1), N)
    na = fv.ops.node_aggregate_fwd(e.cuda(), ptr, items, N)
    xp, xo = fv.ops.cell_block_fwd(cb.mlp_ref(), x.cuda(), na, cnT.to(torch.int32).cuda())
    assert H.rel_l2(xp.cpu(), xp_ref) < 2e-6 and H.rel_l2(xo.cpu(), x.double() + xp_ref) < 2e-6
    ep, eo = fv.ops.edge_block_fwd(eb.mlp_ref(), xp, e.cuda(), nc.to(torch.int32).cuda(), want_prime=True)
    assert H.rel_l2(ep.cpu(), ep_ref) < 3e-6 and H.rel_l2(eo.cpu(), e.double() + ep_ref) < 3e-6
    # module-level API (reference signatures)
    graph = fv.Data(x=x.cuda(), edge_attr=e.cuda(), edge_index=nc.cuda())
    gnode = fv.Data(x=torch.zeros(N, 1).cuda(), edge_index=fn.cuda(), face=cnT.cuda())
    out = fv.CellBlock(128, 128, custom_func=cb)(graph, gnode)
    assert torch.equal(out.x, xp)
    out2 = fv.EdgeBlock(custom_func=eb)(out)
    assert torch.equal(out2.edge_attr, ep)


# ------------------------------------------------------------------------------------------------ a10/a11/a12/a4
def test_integrator_and_face_area(fv, golden):
    g, tab = _tiny(golden)
    C, F = tab["cells_node"].shape[0], tab["face"].shape[1]
    gen = torch.Generator().manual_seed(6)
    dec = torch.randn(F, 5, generator=gen)
    fa = torch.randn(F, 1, generator=gen)
    unv = torch.as_tensor(tab["unit_norm_v"])
    cfT = torch.as_tensor(tab["cells_face"]).long().T.contiguous()
    cont_ref, du_ref = O.integrator(dec[:, 0:2], dec[:, 2:3], dec[:, 3:5], unv, 1.3, 1.0, fa, cfT, loss_cont=1)
    du, cont = fv.ops.integrator_fwd(dec.cuda(), unv.cuda(), fa.cuda(), cfT.to(torch.int32).cuda(), 1.3, 1.0, want_continuity=True)
    assert H.rel_l2(du.cpu(), du_ref) < 1e-6 and H.rel_l2(cont.cpu(), cont_ref) < 1e-6
    integ = fv.Intergrator(14, 2, 2, loss_cont=1)
    c2, du2 = integ(uv_face=dec[:, 0:2].cuda(), p_face=dec[:, 2:3].cuda(), flux_D=dec[:, 3:5].cuda(), unv=unv.cuda(), rho=1.3, rhs_coef=1.0,
                    face_area=fa.cuda(), cell_face=cfT.cuda())
    assert torch.equal(du2, du) and torch.equal(c2, cont)
    # face area + BatchNorm, training and eval
    ftl = torch.cat((torch.as_tensor(tab["face_type"]).float(), torch.as_tensor(tab["face_length"])), 1)
    area = torch.as_tensor(tab["cells_area"])
    nc = torch.as_tensor(np.ascontiguousarray(tab["neighbour_cell"])).long()
    raw = (ftl[:, 1:2] * (0.01 / ((area[nc[0]] + area[nc[1]]) / 2))).view(-1, 1)
    bn = torch.nn.BatchNorm1d(1)
    with torch.no_grad():
        bn.weight.fill_(1.7), bn.bias.fill_(-0.3)
    w, b = bn.weight.detach().clone().cuda(), bn.bias.detach().clone().cuda()
    rm, rv = torch.zeros(1).cuda(), torch.ones(1).cuda()
    bn.train()
    want = bn(raw).detach()
    got, _, _ = fv.ops.face_area_bn_fwd(ftl.cuda(), area.cuda(), nc.to(torch.int32).cuda(), 0.01, True, w, b, rm, rv)
    assert H.rel_l2(got.cpu(), want) < 2e-6
    assert H.rel_l2(rm.cpu(), bn.running_mean) < 1e-6 and H.rel_l2(rv.cpu(), bn.running_var) < 1e-6
    bn.eval()
    want = bn(raw).detach()
    got, _, _ = fv.ops.face_area_bn_fwd(ftl.cuda(), area.cuda(), nc.to(torch.int32).cuda(), 0.01, False, w, b, rm, rv)
    assert H.rel_l2(got.cpu(), want) < 2e-6


def test_normalizer(fv):
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(5000, 5, generator=gen) * 3 + 1
    types = torch.randint(0, 7, (5000, 1), generator=gen).float()
    feats = torch.cat((x, torch.nn.functional.one_hot(types.long().view(-1), 9).float()), 1)
    ref = O.NormState(14, 10)
    n = fv.Normalizer(14, max_accumulations=10).cuda()
    for _ in range(3):
        want = ref(feats, True)
        got = n(x.cuda(), True, type_col=types.cuda(), one_hot=9)
        assert H.rel_l2(got.cpu(), want) < 1e-6
    for k in ("acc_sum", "acc_sum_squared", "acc_count", "num_accumulations"):
        assert H.rel_l2(getattr(n, k).cpu(), getattr(ref, k)) < 1e-6, k
    assert H.rel_l2(n.inverse(got).cpu()[:, :14], ref.inverse(want)) < 1e-6
    # accumulation stops at max_accumulations (device-side test, normalization.py:40)
    n2 = fv.Normalizer(2, max_accumulations=2).cuda()
    y = torch.randn(100, 2, generator=gen).cuda()
    n2(y, True), n2(y, True), n2(y, True)
    assert float(n2.num_accumulations) == 2.0 and float(n2.acc_count) == 101.0


@pytest.mark.parametrize("loss_name", ["direct_mean_loss", "global_mean_sum", "global_mean"])
def test_loss(fv, golden, loss_name):
    g = golden("train_tiny.npz")
    p = H.params(loss=loss_name)
    t = lambda k: torch.as_tensor(g[k])
    want = O.compute_fvm_loss(p, t("cell_batch"), t("face_batch"), t("mask_face_interior"), t("pred_edge"), t("target_face"), t("pred_delta_u"),
                              t("target_delta_u"), num_graphs=2)
    gc = fv.Data(x=t("pred_delta_u").cuda(), batch=t("cell_batch").cuda())
    gc.num_graphs = 2
    ge = fv.Data(x=t("pred_edge").cuda(), batch=t("face_batch").cuda())
    got = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=t("mask_face_interior").cuda(),
                              predicted_edge_attr=t("pred_edge").cuda(), target_face_uvp_normalized=t("target_face").cuda(),
                              predicted_delta_u=t("pred_delta_u").cuda(), target_delta_u=t("target_delta_u").cuda())
    assert abs(float(got[0]) - float(want[0])) < 2e-6 * max(1.0, abs(float(want[0])))
    for i in (2, 3, 4):
        assert abs(float(got[i]) - float(want[i].mean())) < 2e-6 * max(1.0, abs(float(want[i].mean())))
    if loss_name != "global_mean":
        tag = "dm" if loss_name == "direct_mean_loss" else "gm"
        assert abs(float(got[0]) - g[tag + ".loss"][0]) < 1e-5  # the reference's own number


# ------------------------------------------------------------------------------------------------ full forward
def _model(fv, g, seed=0):
    torch.manual_seed(seed)
    m = fv.FVGN(device="cpu", params=H.params())
    H.check_weights({k: v for k, v in m.state_dict().items()}, g) if "wkeys" in g else None
    m.load_state_dict(H.load_state(m.state_dict(), g))
    return m.cuda().eval().set_math_mode("fp32")  # this file holds the STRICT FFMA path to the bound; tests/test_gpu_x3.py the default mode


def _graphs(fv, g, prefix="tab.", vel="velocity", prs="pressure", pair=None):
    return fv.dataset.build_graphs(H.tables_from(g, prefix), g[vel], g[prs], training_pair=pair, device="cuda")


def test_eval_forward_against_reference_fixture(fv, golden):
    g = golden("fwd_tiny.npz")
    model = _model(fv, g)
    gn, ge, gc = _graphs(fv, g)
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    assert np.array_equal(gc.x.cpu().numpy(), g["in_x"]) and np.array_equal(gc.edge_attr.cpu().numpy(), g["in_edge_attr"])  # fmad-free pre-step: same bits
    assert np.array_equal(mb.cpu().numpy(), g["mask_face_boundary"])
    cell_type, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    with torch.no_grad():
        for step in range(3):
            cap = {} if step == 0 else None
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0, capture=cap)
            if step == 0:
                for k in ["cell_in", "edge_in", "enc_x", "enc_e", "x_l0", "e_l0", "x_l1", "e_l1", "x_l14", "e_l14", "decoded", "face_area", "delta_u"]:
                    err = H.rel_l2(cap[k].cpu(), g["cap." + k])
                    assert err < TOL, (k, err)
            assert H.rel_l2(nuv.cpu(), g["roll_uv"][step]) < TOL * (step + 1)
            assert H.rel_l2(uvp.cpu(), g["roll_uvp"][step]) < TOL * (step + 1)
            gc = fv.dataset.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)


def test_cylinder_rollout_10_and_determinism(fv, golden):
    g = golden("fwd_cyl.npz")
    model = _model(fv, g)

    def run():
        gn, ge, gc = _graphs(fv, g)
        gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
        cell_type, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
        outs = []
        with torch.no_grad():
            for step in range(10):
                uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
                outs.append((uvp.clone(), nuv.clone()))
                gc = fv.dataset.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)
        return outs

    a, b = run(), run()
    assert H.rel_l2(a[0][1].cpu(), g["uv_step1"]) < TOL and H.rel_l2(a[0][0].cpu(), g["uvp_step1"]) < TOL
    assert H.rel_l2(a[9][1].cpu(), g["uv_step10"]) < 10 * TOL and H.rel_l2(a[9][0].cpu(), g["uvp_step10"]) < 10 * TOL
    for (u1, v1), (u2, v2) in zip(a, b):
        assert torch.equal(u1, u2) and torch.equal(v1, v2)  # run-to-run bitwise determinism (no float atomics)


def test_forward_vs_fp64_oracle_medium_mesh(fv):
    """~24k cells (several hundred CTAs, ragged tail tile) against the fp64 oracle as arbiter."""
    m = fv.synthetic.make("grid", seed=9, nx=111, ny=113, T=3)
    tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=H.params())
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval().set_math_mode("fp32")
    gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device="cuda")
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    with torch.no_grad():
        uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    on, oe, oc = O.build_graphs(tab, m["velocity"], m["pressure"])
    on, oe, oc, _, _ = O.valid_datapreprocessing(on, oe, oc, 0)
    for b in (on, oe, oc):
        for k, v in list(b.__dict__.items()):
            if torch.is_tensor(v) and v.is_floating_point():
                setattr(b, k, v.double())
    orc = O.FVGNOracle(sd, H.params(), dtype=torch.float64).eval()
    with torch.no_grad():
        ouvp, onuv = orc(oc, oe, on, H.RHO, H.MU, H.DT, 9, 0)
    assert H.rel_l2(uvp.cpu(), ouvp) < TOL and H.rel_l2(nuv.cpu(), onuv) < TOL, (H.rel_l2(uvp.cpu(), ouvp), H.rel_l2(nuv.cpu(), onuv))


def test_training_forward_tuple_and_loss(fv, golden):
    g = golden("train_tiny.npz")
    f = golden("fwd_tiny.npz")
    model = _model(fv, f).train()
    b0 = _graphs(fv, f, "tab.", "velocity", "pressure", pair=1)
    b1 = _graphs(fv, f, "tab1.", "velocity1", "pressure1", pair=0)
    gl = fv.dataset.collate([b0, b1])
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing(gl, cell_noise=dev(g["noise"]), flip_keep=dev(g["flip"]))
    assert np.array_equal(gc.x.cpu().numpy(), g["pre_x"]) and np.array_equal(gc.edge_attr.cpu().numpy(), g["pre_edge_attr"])
    assert np.array_equal(gc.y.cpu().numpy(), g["pre_cell_y"]) and np.array_equal(ge.y.cpu().numpy(), g["pre_edge_y"])
    assert np.array_equal(mi.cpu().numpy(), g["mask_face_interior"])
    with torch.no_grad():
        out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    assert H.rel_l2(out[1].cpu(), g["pred_delta_u"]) < TOL and H.rel_l2(out[3].cpu(), g["pred_edge"]) < TOL
    assert H.rel_l2(out[2].cpu(), g["target_delta_u"]) < 1e-6 and H.rel_l2(out[4].cpu(), g["target_face"]) < 1e-6
    sd = model.state_dict()
    for k in g:
        if k.startswith("after."):
            assert H.rel_l2(sd[k[6:]].float().cpu(), torch.as_tensor(g[k]).float()) < 2e-6, k
    for loss_name, tag in (("direct_mean_loss", "dm"), ("global_mean_sum", "gm")):
        p = H.params(loss=loss_name)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        assert abs(float(loss[0]) - g[tag + ".loss"][0]) < 2e-5


def test_non_contiguous_tables_rejected(fv, golden):
    g, tab = _tiny(golden)
    nc = torch.as_tensor(np.asfortranarray(tab["neighbour_cell"])).to(torch.int32).cuda()
    assert not nc.is_contiguous()
    F = nc.shape[1]
    with pytest.raises(ValueError, match="contiguous"):
        fv.ops.face_area_bn_fwd(torch.ones(F, 2).cuda(), torch.ones(tab["cells_node"].shape[0], 1).cuda(), nc, 0.01, False, torch.ones(1).cuda(),
                                torch.zeros(1).cuda(), torch.zeros(1).cuda(), torch.ones(1).cuda())


def test_no_cpu_fallback(fv):
    m, _ = _mlp_module(fv, 14)
    with pytest.raises(RuntimeError):
        m(torch.randn(4, 14))  # CPU tensor must be rejected, not silently computed
    with pytest.raises(RuntimeError):
        fv.ops.build_csr(torch.zeros(4, dtype=torch.int32), 2)
