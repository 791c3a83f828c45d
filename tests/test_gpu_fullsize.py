"""GPU (B200): BASELINE configs[4] at FULL size (one ~1.0M-cell mesh, 15 layers) through size-independent properties — the oracle
cannot run this size in seconds, so the checks are invariants of the tables, agreement between the independent arithmetic paths
(strict FFMA vs fp32-equivalent tcgen05 vs bf16 tiles), bitwise run-to-run determinism, and linearity of the finite-volume integrator.
Plus the degenerate sizes (0 and 1 rows) of every fused kernel."""
import numpy as np
import pytest
import torch

from tests import helpers as H
from tests.test_gpu_parity import fv, dev, _mlp_module  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(fv):
    m = fv.synthetic.make("grid", seed=11, nx=708, ny=708, T=3)
    t = lambda a, dt: torch.as_tensor(a).to(dt).cuda()
    tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
    return m, tab


def test_connectivity_invariants_at_1m_cells(fv, big):
    m, tab = big
    C, F, N = tab["cells_node"].shape[0], tab["face"].shape[1], tab["mesh_pos"].shape[0]
    assert C > 950_000
    face, cf, nc = tab["face"].long(), tab["cells_face"].long(), tab["neighbour_cell"].long()
    # faces: (max, min) node pairs, strictly increasing lexicographically (= torch.unique(dim=0) order)
    assert bool((face[0] > face[1]).all())
    key = face[0] * (N + 1) + face[1]
    assert bool((key[1:] > key[:-1]).all())
    # every cell's three faces are made of its own (sorted) nodes: (n1,n0), (n2,n1), (n2,n0)
    cn = tab["cells_node"].long()
    for j, (a, b) in enumerate(((1, 0), (2, 1), (2, 0))):
        f = cf[:, j]
        assert bool((face[0, f] == cn[:, a]).all()) and bool((face[1, f] == cn[:, b]).all())
    # each face is referenced once (boundary: self-loop in neighbour_cell) or twice (interior) by cells_face, and by those cells
    cnt = torch.bincount(cf.reshape(-1), minlength=F)
    boundary = nc[0] == nc[1]
    assert bool((cnt[boundary] == 1).all()) and bool((cnt[~boundary] == 2).all())
    owner_ok = torch.zeros(F, dtype=torch.long, device="cuda")
    for j in range(3):
        f = cf[:, j]
        c = torch.arange(C, device="cuda")
        owner_ok.index_add_(0, f, ((nc[0, f] == c) | (nc[1, f] == c)).long())
    assert bool((owner_ok == cnt).all())
    # Euler characteristic of a planar triangulation with one hole: N - F + C = 0
    assert N - F + C == 0
    # geometry: areas positive, unit normals, cell_factor rows sum to 1
    assert bool((tab["cells_area"] > 0).all())
    nrm = tab["unit_norm_v"].double().norm(dim=-1)
    assert float((nrm - 1).abs().max()) < 1e-5
    assert float((tab["cell_factor"].double().sum(1) - 1).abs().max()) < 1e-5


def test_forward_paths_agree_and_are_deterministic_at_1m_cells(fv, big):
    m, tab = big
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=H.params()).cuda().eval()

    def run(mode):
        gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device="cuda")
        gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
        model.set_math_mode(mode)
        with torch.no_grad():
            return model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)

    uvp32, uv32 = run("fp32")
    uvp3, uv3 = run("bf16x3")
    e1, e2 = H.rel_l2(uvp3.cpu(), uvp32.cpu()), H.rel_l2(uv3.cpu(), uv32.cpu())
    print("1M cells: fp32-equivalent tcgen05 vs strict FFMA rel-L2 uvp %.2e next_uv %.2e" % (e1, e2))
    assert e1 < 1e-5 and e2 < 1e-5
    uvp3b, uv3b = run("bf16x3")
    assert torch.equal(uvp3, uvp3b) and torch.equal(uv3, uv3b)  # bitwise run-to-run
    uvpb, uvb = run("bf16")
    uvpb2, uvb2 = run("bf16")
    assert torch.equal(uvpb, uvpb2) and torch.equal(uvb, uvb2)
    e1, e2 = H.rel_l2(uvpb.cpu(), uvp32.cpu()), H.rel_l2(uvb.cpu(), uv32.cpu())
    print("1M cells: bf16 tiles vs strict FFMA rel-L2 uvp %.2e next_uv %.2e" % (e1, e2))
    assert e1 < 3e-2 and e2 < 6e-2  # un-primed normalisers (std from one pseudo-sample) amplify the velocity update's relative error
    assert bool(torch.isfinite(uvp32).all()) and bool(torch.isfinite(uv32).all())


def test_integrator_linearity_and_flux_cancellation_at_1m_cells(fv, big):
    m, tab = big
    C, F = tab["cells_node"].shape[0], tab["face"].shape[1]
    gen = torch.Generator(device="cuda").manual_seed(4)
    cfT = tab["cells_face"].T.contiguous()
    unv = tab["unit_norm_v"].contiguous()
    area = torch.rand((F, 1), device="cuda", generator=gen) + 0.5
    d1 = torch.zeros((F, 5), device="cuda")
    d2 = torch.zeros((F, 5), device="cuda")
    d1[:, 2:5] = torch.randn((F, 3), device="cuda", generator=gen)  # pressure + diffusive flux columns: the integrator is linear in them
    d2[:, 2:5] = torch.randn((F, 3), device="cuda", generator=gen)
    a, _ = fv.ops.integrator_fwd(d1, unv, area, cfT, 1.3)
    b, _ = fv.ops.integrator_fwd(d2, unv, area, cfT, 1.3)
    ab, _ = fv.ops.integrator_fwd((2.0 * d1 - 0.5 * d2).contiguous(), unv, area, cfT, 1.3)
    assert H.rel_l2(ab.cpu(), (2.0 * a - 0.5 * b).cpu()) < 1e-5
    # a constant pressure field exerts no net force on the whole domain's interior: the sum over cells of p n a over interior faces
    # cancels pairwise (outward normals of the two neighbours are opposite) -> only boundary faces remain
    d3 = torch.zeros((F, 5), device="cuda")
    d3[:, 2] = 1.0
    du, _ = fv.ops.integrator_fwd(d3, unv, area, cfT, 1.0)
    nc = tab["neighbour_cell"].long()
    boundary = (nc[0] == nc[1])
    # reference sum over boundary faces only: -(1/rho) * sum_f p n_f a_f with the owner cell's outward normal
    cf = tab["cells_face"].long()
    tot = torch.zeros(2, dtype=torch.float64, device="cuda")
    for j in range(3):
        f = cf[:, j]
        sel = boundary[f]
        tot += -(unv[:, j, :].double()[sel] * area[f][sel].double()).sum(0)
    got = du.double().sum(0)
    assert float((got - tot).abs().max()) < 1e-6 * float(area.double().sum())


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_degenerate_row_counts(fv, mode):
    """0 rows (no launch, empty outputs) and 1 row through every fused kernel in every math mode"""
    cb, _ = _mlp_module(fv, 192, seed=1)
    eb, _ = _mlp_module(fv, 384, seed=2)
    enc, _ = _mlp_module(fv, 14, seed=3)
    for rows in (0, 1):
        x = torch.randn(rows, 14, device="cuda")
        enc.math_mode = {0: "fp32", 1: "bf16", 2: "bf16x3"}[mode]
        assert tuple(enc(x).shape) == (rows, 128)
    # one cell / three faces / three nodes
    xc = torch.randn(1, 128, device="cuda")
    e = torch.randn(3, 128, device="cuda")
    na = torch.randn(3, 64, device="cuda")
    cnT = torch.tensor([[0], [1], [2]], dtype=torch.int32, device="cuda")
    nc = torch.zeros((2, 3), dtype=torch.int32, device="cuda")
    xp, xo = fv.ops.cell_block_fwd(cb.mlp_ref(), xc, na, cnT, math_mode=mode)
    ep, eo = fv.ops.edge_block_fwd(eb.mlp_ref(), xp, e, nc, math_mode=mode, want_prime=True)
    assert bool(torch.isfinite(xo).all()) and bool(torch.isfinite(eo).all())
    xp0, _ = fv.ops.cell_block_fwd(cb.mlp_ref(), xc, na, cnT, math_mode=0)
    tol = {0: 1e-6, 1: 3e-2, 2: 1e-5}[mode]
    assert H.rel_l2(xp.cpu(), xp0.cpu()) <= tol
