"""GPU (B200): the tcgen05/TMEM bf16 path of the fused MLP kernels against (a) the oracle with bf16-rounded GEMM operands
(same arithmetic model: bf16 operands, fp32 accumulate) and (b) the exact fp64 oracle with the stated bf16 bound.

Stated bounds: vs the bf16-operand oracle rel-L2 <= 2e-3 per op (differences: accumulation order, fast exp, and bf16
re-rounding of hidden activations at rounding boundaries); vs the exact oracle rel-L2 <= 3e-2 for the 15-layer forward
(SURVEY.md App. E measured 3e-3 for bf16 operands)."""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H
from tests.test_gpu_parity import fv, dev, _mlp_module, _tiny  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture()
def bf16_oracle():
    O.OPERAND_ROUND = torch.bfloat16
    yield
    O.OPERAND_ROUND = None


def _d(sd):
    return {k: v.double() for k, v in sd.items()}


@pytest.mark.parametrize("k_in,n_out,ln,rows", [(14, 128, True, 150), (2, 128, True, 77), (64, 128, True, 128), (128, 5, False, 300),
                                                (128, 128, True, 1), (192, 128, True, 129), (384, 128, True, 1000), (37, 128, True, 20000)])
def test_tc_mlp_rows(fv, bf16_oracle, k_in, n_out, ln, rows):
    m, sd = _mlp_module(fv, k_in, n_out, ln)
    m.math_mode = "bf16"
    x = torch.randn(rows, k_in, generator=torch.Generator().manual_seed(2))
    got = m(x.cuda()).cpu()
    want = O.mlp(_d(sd), "", x.double(), layer_norm=ln)
    err = H.rel_l2(got, want)
    assert err < 2e-3, err
    O.OPERAND_ROUND = None
    exact = O.mlp(_d(sd), "", x.double(), layer_norm=ln)
    assert H.rel_l2(got, exact) < 2e-2
    # repacking follows weight updates
    with torch.no_grad():
        for p_ in m.parameters():
            p_.mul_(0.5)
    sd2 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    O.OPERAND_ROUND = torch.bfloat16
    got2 = m(x.cuda()).cpu()
    assert H.rel_l2(got2, O.mlp(_d(sd2), "", x.double(), layer_norm=ln)) < 2e-3


def test_tc_cell_and_edge_block(fv, bf16_oracle, golden):
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
    ptr, items = fv.ops.build_csr(dev(tab["face"], torch.int32).reshape(-1), N)
    na = fv.ops.node_aggregate_fwd(e.cuda(), ptr, items, N)
    xp, xo = fv.ops.cell_block_fwd(cb.mlp_ref(), x.cuda(), na, cnT.to(torch.int32).cuda(), math_mode=1)
    assert H.rel_l2(xp.cpu(), xp_ref) < 2e-3, H.rel_l2(xp.cpu(), xp_ref)
    assert H.rel_l2(xo.cpu(), x.double() + xp_ref) < 2e-3
    ep_ref = O.edge_block(_d({"net." + k: v for k, v in sde.items()}), "", xp.cpu().double(), e.double(), nc)
    ep, eo = fv.ops.edge_block_fwd(eb.mlp_ref(), xp, e.cuda(), nc.to(torch.int32).cuda(), math_mode=1, want_prime=True)
    assert H.rel_l2(ep.cpu(), ep_ref) < 2e-3, H.rel_l2(ep.cpu(), ep_ref)
    assert H.rel_l2(eo.cpu(), e.double() + ep_ref) < 2e-3
    # in-place fast path with the bf16 x' shadow (fvgn_*_fwd_xbf16): same arithmetic => bit-identical to the two-call path above
    x2, e2 = x.cuda().clone(), e.cuda().clone()
    xb = torch.empty((C, 128), dtype=torch.bfloat16, device="cuda")
    fv.ops.gn_block_fwd_inplace_bf16(cb.mlp_ref(), eb.mlp_ref(), x2, e2, na, cnT.to(torch.int32).cuda(), nc.to(torch.int32).cuda(), xb)
    assert torch.equal(xb, xp.to(torch.bfloat16))
    assert torch.equal(x2, xo) and torch.equal(e2, eo)


@pytest.mark.parametrize("nx,ny", [(9, 7), (40, 33), (130, 97)])
def test_tc_gn_block_fast_path_matches_two_call_path(fv, nx, ny):
    """ragged sizes (last tile partial, fewer tiles than SMs / several tiles per SM): xbf16 fast path == generic tensor-core path"""
    m = fv.synthetic.make("grid", seed=3, nx=nx, ny=ny, T=3)
    tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    C, F, N = tab["cells_node"].shape[0], tab["face"].shape[1], tab["mesh_pos"].shape[0]
    gen = torch.Generator().manual_seed(11)
    x, e = torch.randn(C, 128, generator=gen).cuda(), torch.randn(F, 128, generator=gen).cuda()
    cb, _ = _mlp_module(fv, 192, seed=6)
    eb, _ = _mlp_module(fv, 384, seed=7)
    cnT = torch.as_tensor(tab["cells_node"]).T.contiguous().to(torch.int32).cuda()
    nc = torch.as_tensor(np.ascontiguousarray(tab["neighbour_cell"])).to(torch.int32).cuda()
    ptr, items = fv.ops.build_csr(dev(tab["face"], torch.int32).reshape(-1), N)
    na = fv.ops.node_aggregate_fwd(e, ptr, items, N)
    xp, xo = fv.ops.cell_block_fwd(cb.mlp_ref(), x, na, cnT, math_mode=1)
    _, eo = fv.ops.edge_block_fwd(eb.mlp_ref(), xp, e, nc, math_mode=1)
    x2, e2 = x.clone(), e.clone()
    xb = torch.empty((C, 128), dtype=torch.bfloat16, device="cuda")
    for _ in range(2):  # twice from the same inputs: also run-to-run determinism
        x2.copy_(x), e2.copy_(e)
        fv.ops.gn_block_fwd_inplace_bf16(cb.mlp_ref(), eb.mlp_ref(), x2, e2, na, cnT, nc, xb)
        assert torch.equal(xb, xp.to(torch.bfloat16)) and torch.equal(x2, xo) and torch.equal(e2, eo)


def test_tc_forward_bound_and_determinism(fv):
    m = fv.synthetic.make("grid", seed=9, nx=111, ny=113, T=3)
    tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=H.params())
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval().set_math_mode("bf16")

    def run():
        gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device="cuda")
        gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
        with torch.no_grad():
            return model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)

    uvp, nuv = run()
    uvp2, nuv2 = run()
    assert torch.equal(uvp, uvp2) and torch.equal(nuv, nuv2)  # deterministic
    on, oe, oc = O.build_graphs(tab, m["velocity"], m["pressure"])
    on, oe, oc, _, _ = O.valid_datapreprocessing(on, oe, oc, 0)
    for b in (on, oe, oc):
        for k, v in list(b.__dict__.items()):
            if torch.is_tensor(v) and v.is_floating_point():
                setattr(b, k, v.double())
    orc = O.FVGNOracle(sd, H.params(), dtype=torch.float64).eval()
    with torch.no_grad():
        ouvp, onuv = orc(oc, oe, on, H.RHO, H.MU, H.DT, 9, 0)
    e1, e2 = H.rel_l2(uvp.cpu(), ouvp), H.rel_l2(nuv.cpu(), onuv)
    print("bf16 tensor-core forward vs exact fp64 oracle: uvp", e1, "next_uv", e2)
    assert e1 < 3e-2 and e2 < 3e-2
