"""GPU (B200): the fp32-equivalent tensor-core mode (FVGN_MATH_BF16X3_TC, csrc/mlp_tc3.cu — forward: two-term fp16 split of both
operands, three tcgen05 MMAs per product; backward: the three leading products of 3-way bf16 splits; fp32 epilogue) is held to the SAME
bound as the strict FFMA path: relative L2 <= 1e-5 per step against the oracle / the reference fixtures (north_star), per op <= 5e-6
against the fp64 oracle; the CTA-pair variants of the inference kernels (csrc/mlp_tc3_pair.cuh) to bit-identity with the single-CTA ones."""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H
from tests.test_gpu_parity import fv, dev, _mlp_module, _tiny, _model, _graphs, TOL  # noqa: F401

pytestmark = pytest.mark.gpu
X3 = 2


def _d(sd):
    return {k: v.double() for k, v in sd.items()}


@pytest.mark.parametrize("k_in,n_out,ln,rows", [(2, 128, True, 77), (14, 128, True, 150), (128, 5, False, 150), (128, 128, True, 1),
                                                (192, 128, True, 64), (384, 128, True, 65), (37, 128, True, 200), (384, 128, True, 40000)])
def test_x3_mlp_rows(fv, k_in, n_out, ln, rows):
    m, sd = _mlp_module(fv, k_in, n_out, ln)
    m.math_mode = "bf16x3"
    x = torch.randn(rows, k_in, generator=torch.Generator().manual_seed(2))
    got = m(x.cuda()).cpu()
    want64 = O.mlp(_d(sd), "", x.double(), layer_norm=ln)
    err = H.rel_l2(got, want64)
    print("x3 mlp_rows", k_in, n_out, rows, "rel-L2 vs fp64 %.2e" % err)
    assert err < 5e-6, err
    with torch.no_grad():  # re-pack follows weight updates
        for p_ in m.parameters():
            p_.mul_(0.5)
    sd2 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    assert H.rel_l2(m(x.cuda()).cpu(), O.mlp(_d(sd2), "", x.double(), layer_norm=ln)) < 5e-6


def test_x3_cell_and_edge_block(fv, golden):
    g, tab = _tiny(golden)
    C, F, N = tab["cells_node"].shape[0], tab["face"].shape[1], tab["mesh_pos"].shape[0]
    gen = torch.Generator().manual_seed(3)
    x, e = torch.randn(C, 128, generator=gen), torch.randn(F, 128, generator=gen)
    cb, sdc = _mlp_module(fv, 192, seed=4)
    eb, sde = _mlp_module(fv, 384, seed=5)
    fn = torch.as_tensor(np.ascontiguousarray(tab["face"])).long()
    cnT = torch.as_tensor(tab["cells_node"]).long().T.contiguous()
    nc = torch.as_tensor(np.ascontiguousarray(tab["neighbour_cell"])).long()
    xp_ref = O.cell_block(_d({"net." + k: v for k, v in sdc.items()}), "", x.double(), e.double(), fn, cnT, N)
    ep_ref = O.edge_block(_d({"net." + k: v for k, v in sde.items()}), "", xp_ref, e.double(), nc)
    ptr, items = fv.ops.build_csr(dev(tab["face"], torch.int32).reshape(-1), N)
    na = fv.ops.node_aggregate_fwd(e.cuda(), ptr, items, N)
    xp, xo = fv.ops.cell_block_fwd(cb.mlp_ref(), x.cuda(), na, cnT.to(torch.int32).cuda(), math_mode=X3)
    assert H.rel_l2(xp.cpu(), xp_ref) < 5e-6 and H.rel_l2(xo.cpu(), x.double() + xp_ref) < 5e-6
    ep, eo = fv.ops.edge_block_fwd(eb.mlp_ref(), xp, e.cuda(), nc.to(torch.int32).cuda(), math_mode=X3, want_prime=True)
    assert H.rel_l2(ep.cpu(), ep_ref) < 5e-6 and H.rel_l2(eo.cpu(), e.double() + ep_ref) < 5e-6
    # in-place residual (red.global.add) variant gives the same bits as the out-of-place one
    x2, e2 = x.cuda().clone(), e.cuda().clone()
    xp2, _ = fv.ops.cell_block_fwd(cb.mlp_ref(), x2, na, cnT.to(torch.int32).cuda(), math_mode=X3, x_out=x2)
    fv.ops.edge_block_fwd(eb.mlp_ref(), xp2, e2, nc.to(torch.int32).cuda(), math_mode=X3, e_out=e2)
    assert torch.equal(xp2, xp) and torch.equal(x2, xo) and torch.equal(e2, eo)


def _split_of(t):
    hi = (t * 8.0).half()
    lo = (t * 8.0 - hi.float()).half()
    return torch.stack([hi, lo], dim=1)


def test_x3_split_shadow_fast_path_is_bit_identical(fv, golden):
    """The inference fast path (latents in place, x' only as the split fp16 shadow moved by cp.async: fvgn_*_fwd_split) against the
    generic fp32-equivalent calls: same bits, and the shadow is exactly the two-term fp16 split of x'."""
    g, tab = _tiny(golden)
    C, F, N = tab["cells_node"].shape[0], tab["face"].shape[1], tab["mesh_pos"].shape[0]
    gen = torch.Generator().manual_seed(7)
    x0, e0 = torch.randn(C, 128, generator=gen).cuda(), torch.randn(F, 128, generator=gen).cuda()
    cb, _ = _mlp_module(fv, 192, seed=4)
    eb, _ = _mlp_module(fv, 384, seed=5)
    cnT = torch.as_tensor(tab["cells_node"]).to(torch.int32).T.contiguous().cuda()
    nc = torch.as_tensor(np.ascontiguousarray(tab["neighbour_cell"])).to(torch.int32).cuda()
    ptr, items = fv.ops.build_csr(dev(tab["face"], torch.int32).reshape(-1), N)
    x, e = x0.clone(), e0.clone()
    xr, er = x0.clone(), e0.clone()
    xps = torch.empty((C, 2, 128), dtype=torch.float16, device="cuda")
    for layer in range(3):
        na = fv.ops.node_aggregate_fwd(er, ptr, items, N)
        xp, xr = fv.ops.cell_block_fwd(cb.mlp_ref(), xr, na, cnT, math_mode=X3)
        _, er = fv.ops.edge_block_fwd(eb.mlp_ref(), xp, er, nc, math_mode=X3)
        na2 = fv.ops.node_aggregate_fwd(e, ptr, items, N)
        fv.ops.gn_block_fwd_inplace_split(cb.mlp_ref(), eb.mlp_ref(), x, e, na2, cnT, nc, xps)
        assert torch.equal(x, xr) and torch.equal(e, er), layer
        assert torch.equal(xps, _split_of(xp)), layer


@pytest.mark.parametrize("nx,ny", [(9, 7), (13, 11), (150, 130), (200, 150), (331, 160)])
def test_x3_pair_kernels_are_bit_identical(fv, nx, ny):
    """The CTA-pair kernels (tcgen05.mma.cta_group::2: each CTA of a pair holds half of every weight K-tile, csrc/mlp_tc3_pair.cuh) against the
    single-CTA kernels on the same inputs: the same MMAs on the same operands, same epilogue — same bits.  Sizes: one tile, an odd tile
    count below the SM count, several tiles per CTA (even and odd counts, a ragged last tile), > 2 tiles per CTA pair."""
    m = fv.synthetic.make("grid", seed=6, nx=nx, ny=ny, T=2)
    t = lambda a, dt: torch.as_tensor(a).to(dt).cuda()
    tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
    C, F, N = tab["cells_node"].shape[0], tab["face"].shape[1], m["mesh_pos"].shape[0]
    gen = torch.Generator().manual_seed(12)
    x0, e0 = torch.randn(C, 128, generator=gen).cuda(), torch.randn(F, 128, generator=gen).cuda()
    cb, _ = _mlp_module(fv, 192, seed=4)
    eb, _ = _mlp_module(fv, 384, seed=5)
    cnT, nc = tab["cells_node"].T.contiguous(), tab["neighbour_cell"].contiguous()
    ptr, items = fv.ops.build_csr(tab["face"].reshape(-1).contiguous(), N)
    L = fv._lib.lib()
    outs = {}
    prev = L.fvgn_x3_pair_enable(1)
    try:
        for on in (1, 0):
            L.fvgn_x3_pair_enable(on)
            x, e = x0.clone(), e0.clone()
            xps = torch.empty((C, 2, 128), dtype=torch.float16, device="cuda")
            for layer in range(3):
                fv.ops.gn_block_fwd_inplace_split(cb.mlp_ref(), eb.mlp_ref(), x, e, fv.ops.node_aggregate_fwd(e, ptr, items, N), cnT, nc, xps)
            torch.cuda.synchronize()
            outs[on] = (x, e, xps.clone())
    finally:
        L.fvgn_x3_pair_enable(prev)
    assert torch.isfinite(outs[1][0]).all() and torch.isfinite(outs[1][1]).all()
    for a, b in zip(outs[1], outs[0]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("nx,ny", [(13, 11), (150, 130), (200, 150)])
def test_x3_split_fast_path_mid_size_pairs(fv, nx, ny):
    """Same comparison on meshes that give a CTA several tiles (paired-tile schedule: two pairs + a single tile at ~60k cells, an odd
    tile count at ~39k), plus the strict-FFMA kernels as the arbiter for the values: 3 GnBlocks, in place."""
    m = fv.synthetic.make("grid", seed=5, nx=nx, ny=ny, T=2)
    t = lambda a, dt: torch.as_tensor(a).to(dt).cuda()
    tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
    C, F, N = tab["cells_node"].shape[0], tab["face"].shape[1], m["mesh_pos"].shape[0]
    gen = torch.Generator().manual_seed(11)
    x0, e0 = torch.randn(C, 128, generator=gen).cuda(), torch.randn(F, 128, generator=gen).cuda()
    cb, _ = _mlp_module(fv, 192, seed=4)
    eb, _ = _mlp_module(fv, 384, seed=5)
    cnT = tab["cells_node"].T.contiguous()
    nc = tab["neighbour_cell"].contiguous()
    ptr, items = fv.ops.build_csr(tab["face"].reshape(-1).contiguous(), N)
    xs, es = x0.clone(), e0.clone()   # split-shadow fast path
    xg, eg = x0.clone(), e0.clone()   # generic fp32-equivalent calls
    xf, ef = x0.clone(), e0.clone()   # strict FFMA
    xps = torch.empty((C, 2, 128), dtype=torch.float16, device="cuda")
    for layer in range(3):
        fv.ops.gn_block_fwd_inplace_split(cb.mlp_ref(), eb.mlp_ref(), xs, es, fv.ops.node_aggregate_fwd(es, ptr, items, N), cnT, nc, xps)
        xp, xg = fv.ops.cell_block_fwd(cb.mlp_ref(), xg, fv.ops.node_aggregate_fwd(eg, ptr, items, N), cnT, math_mode=X3)
        _, eg = fv.ops.edge_block_fwd(eb.mlp_ref(), xp, eg, nc, math_mode=X3)
        xpf, xf = fv.ops.cell_block_fwd(cb.mlp_ref(), xf, fv.ops.node_aggregate_fwd(ef, ptr, items, N), cnT, math_mode=0)
        _, ef = fv.ops.edge_block_fwd(eb.mlp_ref(), xpf, ef, nc, math_mode=0)
        assert torch.equal(xs, xg) and torch.equal(es, eg), layer
    ex, ee = H.rel_l2(xs.cpu(), xf.cpu().double()), H.rel_l2(es.cpu(), ef.cpu().double())
    print("x3 split path vs strict FFMA after 3 GnBlocks, C=%d F=%d: x %.2e e %.2e" % (C, F, ex, ee))
    assert ex < 5e-6 and ee < 5e-6


def test_x3_activation_overflow_is_loud(fv):
    """|activation| >= 8190 does not fit the scaled fp16 operand split: the row comes out non-finite (never a silently wrong number)."""
    m, _ = _mlp_module(fv, 14)
    m.math_mode = "bf16x3"
    x = torch.randn(300, 14, generator=torch.Generator().manual_seed(0)).cuda()
    x[7, 3] = 1.0e5
    y = m(x)
    assert not torch.isfinite(y[7]).all()
    keep = torch.ones(300, dtype=torch.bool, device="cuda")
    keep[7] = False
    assert torch.isfinite(y[keep]).all()


def test_x3_forward_against_reference_fixture_and_rollout(fv, golden):
    g = golden("fwd_tiny.npz")
    model = _model(fv, g).set_math_mode("bf16x3")
    gn, ge, gc = _graphs(fv, g)
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    cell_type, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    with torch.no_grad():
        for step in range(3):
            cap = {} if step == 0 else None
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0, capture=cap)
            if step == 0:
                for k in ["enc_x", "enc_e", "x_l0", "e_l0", "x_l1", "e_l1", "x_l14", "e_l14", "decoded", "delta_u"]:
                    err = H.rel_l2(cap[k].cpu(), g["cap." + k])
                    assert err < TOL, (k, err)
            e1, e2 = H.rel_l2(nuv.cpu(), g["roll_uv"][step]), H.rel_l2(uvp.cpu(), g["roll_uvp"][step])
            print("x3 step", step, "rel-L2 vs the reference fixture: next_uv %.2e uvp %.2e" % (e1, e2))
            assert e1 < TOL * (step + 1) and e2 < TOL * (step + 1)
            gc = fv.dataset.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)


def test_x3_cylinder_rollout_10_and_determinism(fv, golden):
    g = golden("fwd_cyl.npz")
    model = _model(fv, g).set_math_mode("bf16x3")

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
        assert torch.equal(u1, u2) and torch.equal(v1, v2)


def test_x3_training_step_gradients(fv, golden):
    """training with the fp32-equivalent tensor-core forward (the backward recomputes in FFMA): loss and gradients against the
    reference's own (fixture) within the training tolerances of tests/test_gpu_training.py"""
    g = golden("train_tiny.npz")
    f = golden("fwd_tiny.npz")
    model = _model(fv, f).train().set_math_mode("bf16x3")
    gl = fv.dataset.collate([_graphs(fv, f, "tab.", "velocity", "pressure", pair=1), _graphs(fv, f, "tab1.", "velocity1", "pressure1", pair=0)])
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing(gl, cell_noise=dev(g["noise"]), flip_keep=dev(g["flip"]))
    p = H.params(loss="global_mean_sum")
    out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                               target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    loss[0].backward()
    assert abs(float(loss[0]) - g["gm.loss"][0]) < 2e-5
    named = dict(model.named_parameters())
    keys = [str(k) for k in g["gm.gkeys"]]
    gnorm = np.array([float(named[k].grad.double().norm()) for k in keys])
    np.testing.assert_allclose(gnorm, g["gm.gnorm"], rtol=2e-4, atol=1e-9)
    for k in g:
        if k.startswith("gm.grad."):
            assert H.rel_l2(named[k[8:]].grad.cpu(), g[k]) < 1e-4, k


def test_bf16_tile_training_on_airfoil_meshes(fv):
    """BASELINE configs[2]: training with bf16 MLP tiles on airfoil-shaped meshes.  'bf16' training runs the bf16x3 kernels with the
    leading split product only (bf16 operands, fp32 accumulation) in forward, dgrad and wgrad.  Stated bound: loss within 1e-2
    (relative) and every large gradient tensor within 8e-2 relative L2 / cosine > 0.995 of the fp32-equivalent (bf16x3) training
    step on the same batch; a few AdamW steps reduce the loss."""
    meshes = [fv.synthetic.make("foil", seed=s, T=3, n_target=1500) for s in (0, 1)]
    lists = []
    for m in meshes:
        tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
        lists.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device="cuda"))
    p = H.params(loss="global_mean_sum")
    gen = torch.Generator(device="cuda").manual_seed(3)
    res = {}
    for mode in ("bf16x3", "bf16"):
        torch.manual_seed(0)
        model = fv.FVGN(device="cpu", params=p).cuda().train().set_math_mode(mode)
        gn, ge, gc = fv.dataset.collate([[b for b in l] for l in lists])
        C, F = gc.x.shape[0], ge.x.shape[0]
        gen.manual_seed(3)
        noise = torch.randn((C, 2), device="cuda", generator=gen) * 0.02
        flip = torch.rand(F, device="cuda", generator=gen) < 0.5
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc], cell_noise=noise, flip_keep=flip)
        out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        res[mode] = (float(loss[0].detach()), {k: v.grad.detach().double().flatten() for k, v in model.named_parameters()})
    l3, g3 = res["bf16x3"]
    l1, g1 = res["bf16"]
    assert abs(l1 - l3) < 1e-2 * abs(l3), (l1, l3)
    worst = 0.0
    for k, a in g3.items():
        if a.numel() < 1024:
            continue
        b = g1[k]
        rel = float((a - b).norm() / a.norm())
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        worst = max(worst, rel)
        assert rel < 8e-2 and cos > 0.995, (k, rel, cos)
    print("bf16-tile training vs fp32-equivalent: loss %.5f vs %.5f, worst weight-gradient rel-L2 %.2e" % (l1, l3, worst))
    # a few optimizer steps in bf16-tile mode
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=p).cuda().train().set_math_mode("bf16")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        gn, ge, gc = fv.dataset.collate([[b for b in l] for l in lists])
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc], cell_noise=noise, flip_keep=flip)
        opt.zero_grad()
        out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        opt.step()
        losses.append(float(loss[0].detach()))
    assert losses[-1] < losses[0], losses
