"""GPU (B200): the C-side executor of the training step's layer loops (csrc/executor.cu: fvgn_processor_train_fwd / _bwd — two C calls
instead of ~13 ctypes calls per GnBlock, for the reference's real training regime of a NEW batch topology every step, train.py:119-160)
issues the same kernels in the same order on the same streams: forward outputs, loss and EVERY gradient are bit-identical to the
per-call path, in both tensor-core training modes, for mp_times 2 and 1, and in layer chunks (the data-parallel reduce groups)."""
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import helpers as H
from tests.test_gpu_parity import fv, dev  # noqa: F401

pytestmark = pytest.mark.gpu


def _step(fv, lists, p, mode, noise, flip, executor, chunk_cuts=None):
    from fvgn_b200 import autograd, parallel

    autograd.EXECUTOR = executor
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=p).cuda().train().set_math_mode(mode)
    gn, ge, gc = fv.dataset.collate([[b for b in l] for l in lists])
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc], cell_noise=noise, flip_keep=flip)
    out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
    loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                               target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    loss[0].backward()
    torch.cuda.synchronize()
    autograd.EXECUTOR = True
    return loss[0].detach().clone(), [o.detach().clone() for o in out[1:]], [q.grad.clone() for q in model.parameters()]


@pytest.mark.parametrize("mode,mp", [("bf16x3", 2), ("bf16", 2), ("bf16x3", 1)])
def test_executor_is_bit_identical_to_the_per_call_path(fv, mode, mp):
    meshes = [fv.synthetic.make("cyl", seed=s, T=3, n_target=900) for s in (1, 2, 3)]
    lists = []
    for m in meshes:
        tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
        lists.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device="cuda"))
    C = sum(l[2].x.shape[0] for l in lists)
    F = sum(l[1].x.shape[0] for l in lists)
    gen = torch.Generator().manual_seed(4)
    noise = (torch.randn(C, 2, generator=gen) * 0.02).cuda()
    flip = (torch.rand(F, generator=gen) < 0.5).cuda()
    p = H.params(loss="global_mean_sum", mp_times=mp)
    la, oa, ga = _step(fv, lists, p, mode, noise, flip, executor=False)
    lb, ob, gb = _step(fv, lists, p, mode, noise, flip, executor=True)
    assert torch.equal(la, lb)
    for a, b in zip(oa, ob):
        assert torch.equal(a, b)
    assert len(ga) == len(gb) and all(torch.equal(a, b) for a, b in zip(ga, gb))
    lc, oc, gc_ = _step(fv, lists, p, mode, noise, flip, executor=True)  # run-to-run deterministic
    assert all(torch.equal(a, b) for a, b in zip(gb, gc_))


def test_executor_backward_in_chunks_matches_one_call(fv):
    """the data-parallel reducer splits the backward into layer chunks (one C call per reduce group): same bits as one call"""
    from fvgn_b200 import ops, topology
    import ctypes as C

    torch.manual_seed(0)
    p = H.params()
    model = fv.FVGN(device="cpu", params=p).cuda().train().set_math_mode("bf16x3")
    m = fv.synthetic.make("cyl", seed=5, T=3, n_target=900)
    tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device="cuda")
    topo = topology.topology_of(gc, gn, ge)
    epd = model.model
    L = len(epd.processer_list)
    cms = [b.cb_module.net.mlp_ref() for b in epd.processer_list]
    ems = [b.eb_module.net.mlp_ref() for b in epd.processer_list]
    ops.ensure_packed_many(cms + ems, 2)
    ops.ensure_packed_bwd_many(cms + ems)
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device="cuda")
    X, E = f32(L + 1, topo.C, 128), f32(L + 1, topo.F, 128)
    X[0], E[0] = torch.randn(topo.C, 128, device="cuda"), torch.randn(topo.F, 128, device="cuda")
    XP, NA, ZC, ZE = f32(L, topo.C, 128), f32(L, topo.C, 64), f32(L, topo.C, 384), f32(L, topo.F, 384)
    tt = topo.train_tables(2)
    ops.processor_train_fwd(cms, ems, 2, tt, X, E, XP, NA, ZC, ZE, f32(topo.N, 64))
    lib = ops._lib.lib()
    sizes = [int(lib.fvgn_mlp_grad_floats(q.k_in)) for pair in zip(ems, cms) for q in pair]
    g_e0, g_x0 = torch.randn(topo.F, 128, device="cuda"), torch.randn(topo.C, 128, device="cuda")
    scratch = f32(int(lib.fvgn_processor_train_bwd_scratch_floats(topo.C, topo.F, topo.N)))
    side = torch.cuda.Stream()

    def run(bounds):
        G = torch.zeros(sum(sizes), dtype=torch.float32, device="cuda")
        slots, o = [], 0
        for n in sizes:
            slots.append(G[o:o + n])
            o += n
        ptrs = (C.c_void_p * (2 * L))(*[s.data_ptr() for s in slots])
        g_e, g_x = g_e0, g_x0
        bufs = [(f32(topo.F, 128), f32(topo.C, 128)) for _ in range(2)]
        for k, (hi, lo) in enumerate(zip(reversed(bounds[1:]), reversed(bounds[:-1]))):
            ops.processor_train_bwd(cms, ems, lo, hi, 2, tt, X, E, XP, NA, ZC, ZE, g_e, bufs[k & 1][0], g_x, bufs[k & 1][1], ptrs, scratch, side)
            g_e, g_x = bufs[k & 1]
        torch.cuda.synchronize()
        return G, g_e.clone(), g_x.clone()

    Ga, ea, xa = run([0, L])
    for bounds in ([0, 1, 3, 7, 11, L], [0, 14, L], list(range(L + 1))):
        Gb, eb, xb = run(bounds)
        assert torch.equal(Ga, Gb) and torch.equal(ea, eb) and torch.equal(xa, xb), bounds
    assert torch.isfinite(Ga).all() and float(Ga.abs().sum()) > 0
