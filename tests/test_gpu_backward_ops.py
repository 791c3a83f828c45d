"""GPU (B200): the tensor-core backward of one build_mlp at the operator level — fvgn_mlp_rows_fwd (z_save) -> fvgn_ln_bwd (+ column-sum
partials) -> fvgn_mlp_dgrad_x3 -> fvgn_mlp_wgrad_x3 — against torch autograd over the same MLP in float64, at ragged and degenerate row
counts (1 row, one short of / one past a 64-row chunk and a 128-row tile, a count that spans many row ranges), narrow inputs / outputs
(encoder 14 -> 128, decoder 128 -> 5 without LayerNorm), with the three leading split products (default) and all six.
Bound: relative L2 <= 1e-4 per tensor (the training tolerance); plus the small kernels this round added, held to exact results:
fvgn_cell_mean (bit-equal to ((a + b) + c) / 3 in fp32) and fvgn_mlp_pack_bf16x3_many (byte-equal to the per-MLP pack)."""
import pytest
import torch

from tests import helpers as H
from tests.test_gpu_parity import fv  # noqa: F401

pytestmark = pytest.mark.gpu
GTOL = 1e-4


def _ref_mlp(m, k_in, n_out, ln):
    ref = torch.nn.Sequential(torch.nn.Linear(k_in, 128), torch.nn.SiLU(), torch.nn.Linear(128, 128), torch.nn.SiLU(), torch.nn.Linear(128, n_out))
    seq = m[0] if ln else m
    for a, b in zip((ref[0], ref[2], ref[4]), (seq[0], seq[2], seq[4])):
        a.load_state_dict(b.state_dict())
    lnm = None
    if ln:
        lnm = torch.nn.LayerNorm(n_out)
        lnm.load_state_dict(m[1].state_dict())
    return ref.double(), (lnm.double() if lnm is not None else None)


@pytest.mark.parametrize("products", [3, 6])
@pytest.mark.parametrize("k_in,n_out,ln,rows", [(128, 128, True, 1), (128, 128, True, 63), (128, 128, True, 65), (128, 128, True, 129),
                                                (14, 128, True, 150), (128, 5, False, 333), (384, 128, True, 1000), (128, 128, True, 70001)])
def test_x3_backward_chain_of_one_mlp(fv, monkeypatch, products, k_in, n_out, ln, rows):
    monkeypatch.setattr(fv.ops, "BWD_PRODUCTS", products)
    torch.manual_seed(7)
    m = fv.build_mlp(k_in, 128, n_out, drop_out=False, lay_norm=ln)
    with torch.no_grad():
        for p in m.parameters():  # LayerNorm weight / biases away from their 1 / 0 initial values
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    ref, lnm = _ref_mlp(m, k_in, n_out, ln)
    x = torch.randn(rows, k_in, generator=torch.Generator().manual_seed(3))
    gout = torch.randn(rows, n_out, generator=torch.Generator().manual_seed(4))
    xr = x.double().requires_grad_(True)
    y = ref(xr)
    if ln:
        y = lnm(y)
    y.backward(gout.double())
    mc = m.cuda()
    mc.math_mode = "bf16x3"
    mr = mc.mlp_ref()
    xc, gc = x.cuda(), gout.cuda()
    z = torch.empty((rows, 384), dtype=torch.float32, device="cuda")
    out = fv.ops.mlp_rows_fwd(mr, xc, 2, z_save=z)
    assert H.rel_l2(out.cpu(), y.detach()) < 1e-5
    if ln:
        dz, sums = fv.ops.ln_bwd(gc, z, mc[1].weight)
    else:
        dz, sums = gc, None
    da, d_in = fv.ops.mlp_dgrad_x3(mr, dz, z, want_dA=(k_in % 16 == 0))
    grads = fv.ops.mlp_wgrad_x3(mr, 0, gc, dz if ln else None, sums, z, da, x=xc)
    want = [ref[0].weight.grad, ref[0].bias.grad, ref[2].weight.grad, ref[2].bias.grad, ref[4].weight.grad, ref[4].bias.grad]
    if ln:
        want += [lnm.weight.grad, lnm.bias.grad]
    assert len(grads) == len(want)
    for i, (g, w) in enumerate(zip(grads, want)):
        assert H.rel_l2(g.cpu(), w) < GTOL, (i, tuple(g.shape), H.rel_l2(g.cpu(), w))
    if d_in is not None:
        assert H.rel_l2(d_in.cpu(), xr.grad) < GTOL
    # deterministic: a second run gives the same bits
    grads2 = fv.ops.mlp_wgrad_x3(mr, 0, gc, dz if ln else None, sums, z, da, x=xc)
    assert all(torch.equal(a, b) for a, b in zip(grads, grads2))


def test_cell_mean_is_bit_exact(fv):
    g = torch.Generator().manual_seed(0)
    N, Cn = 517, 1003
    na = torch.randn(N, 64, generator=g)
    cnT = torch.randint(0, N, (3, Cn), generator=g, dtype=torch.int32)
    got = fv.ops.cell_mean(na.cuda(), cnT.cuda())
    # GN_blocks.py:125-129 evaluated the way the reference's CPU path does: add, add, TRUE division (torch's CUDA kernel for `tensor / 3.0`
    # multiplies by the rounded reciprocal instead, which differs in the last bit for some inputs)
    want = ((na[cnT[0].long()] + na[cnT[1].long()]) + na[cnT[2].long()]) / 3.0
    assert torch.equal(got.cpu(), want)


def test_pack_many_equals_the_per_mlp_pack(fv):
    torch.manual_seed(1)
    mlps = [fv.build_mlp(k, 128, n, drop_out=False, lay_norm=ln).cuda() for k, n, ln in ((14, 128, True), (192, 128, True), (384, 128, True), (128, 5, False))]
    refs = [m.mlp_ref() for m in mlps]
    for r in refs:
        r.ensure_packed(2)  # allocates
    for r in refs:  # (bytes neither path writes — padding behind the header — must start equal)
        r.packed3.zero_()
        r.packed3_versions = None
        r.ensure_packed(2)
    single = [r.packed3.clone() for r in refs]
    for r in refs:
        r.packed3.zero_()
        r.packed3_versions = None
    fv.ops.ensure_packed_many(refs, 2)
    for a, r in zip(single, refs):
        assert torch.equal(a, r.packed3)
    # nothing stale: no launch
    l0 = fv.ops.LAUNCHES
    fv.ops.ensure_packed_many(refs, 2)
    assert fv.ops.LAUNCHES == l0
