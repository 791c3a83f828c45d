"""GPU (B200): single-mesh domain decomposition on the kernels (fvgn_b200.domain.KernelBackend).  The driver's GPU box has one GPU, so
all parts run on it one after the other with the in-process communicator — the kernels, tables and exchange plans are the ones a rank
per GPU runs over NCCL (`bench.py --gpus N`: `domain_decomposition` leg, N >= 2, which also checks the result against the single-GPU
rollout).  Bound: fp32 rel 1e-5 per step against the single-domain CUDA rollout (interface-node sums are re-associated: ~1e-7)."""
import pytest
import torch

from fvgn_b200 import domain as D
from oracle import fvgn_oracle as O
from tests import helpers as H
from tests.test_gpu_parity import fv, dev  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("math,n_parts", [("fp32", 2), ("bf16x3", 3), ("bf16x3", 8)])
def test_decomposed_rollout_on_the_kernels(fv, math, n_parts):
    m = fv.synthetic.make("grid", seed=3, nx=120, ny=70, T=3)  # ~16k cells
    tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    tab["mesh_pos"], tab["node_type"] = m["mesh_pos"], m["node_type"]
    model = fv.FVGN(device="cpu", params=H.params())
    model.load_state_dict(H.torch_reference_state_dict(0))
    model = model.cuda().eval().set_math_mode(math)
    gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device="cuda")
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    x0, ea0, bc = gc.x.clone(), gc.edge_attr.clone(), ge.y[:, 0, 0:2].clone()
    steps = 3
    state = fv.dataset.RolloutState(gc, ge)
    want = []
    with torch.no_grad():
        for _ in range(steps):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
            want.append((uvp.clone(), nuv.clone()))
            state.advance_(gc, nuv)
    parts, _ = D.decompose(tab, n_parts)
    be = D.KernelBackend(model)
    runs = [D.PartRollout(p, be, *D.initial_part_state(p, x0, ea0, bc), device="cuda") for p in parts]
    ms = dict(D.KernelBackend.state(model), dt=H.DT, rho=H.RHO)
    with torch.no_grad():
        outs = D.rollout_steps(runs, D.LocalComm(n_parts), ms, steps)
    C, F = x0.shape[0], ea0.shape[0]
    for k in range(steps):
        nuv = torch.zeros((C, 2), device="cuda")
        uvp = torch.zeros((F, 3), device="cuda")
        seen = torch.zeros(F, dtype=torch.long, device="cuda")
        for p, o in zip(parts, outs):
            nuv[torch.as_tensor(p.own).cuda()] = o[k][1]
            f = torch.as_tensor(p.faces).cuda()
            dup = seen[f] > 0
            assert torch.equal(uvp[f][dup], o[k][0][dup])  # a cut face is computed by two parts: same bits
            uvp[f] = o[k][0]
            seen[f] += 1
        e1, e2 = H.rel_l2(nuv.cpu(), want[k][1].cpu()), H.rel_l2(uvp.cpu(), want[k][0].cpu())
        assert e1 < 1e-5 * (k + 1) and e2 < 1e-5 * (k + 1), (math, n_parts, k, e1, e2)
    assert fv.ops.status(clear=True) == 0
    # the exchanges over peer memory (csrc/halo.cu: rows stored straight into the neighbour's receive buffer + sequence flags; here the
    # "peers" are the other parts' buffers on the same device): same bits as the torch index ops + copies above, eager and as a graph
    runs = [D.PartRollout(p, be, *D.initial_part_state(p, x0, ea0, bc), device="cuda") for p in parts]
    pcomm = D.LocalPeerComm(parts, runs, "cuda")
    with torch.no_grad():
        pouts = D.rollout_steps(runs, pcomm, ms, steps)
    pcomm.check()
    for o, g in zip(outs, pouts):
        for k in range(steps):
            assert torch.equal(o[k][0], g[k][0]) and torch.equal(o[k][1], g[k][1]), ("peer", math, n_parts, k)
    runs = [D.PartRollout(p, be, *D.initial_part_state(p, x0, ea0, bc), device="cuda") for p in parts]
    pcomm = D.LocalPeerComm(parts, runs, "cuda")
    gr = D.GraphedPartRollout(runs, pcomm, ms)
    for k in range(steps):
        got = gr.step()
        for o, g in zip(outs, got):
            assert torch.equal(o[k][0], g[0]) and torch.equal(o[k][1], g[1]), ("peer graph", math, n_parts, k)
    pcomm.check()
    # the same steps as ONE CUDA graph over all parts (D.GraphedPartRollout; one graph per rank with NCCL exchanges on a multi-GPU box): same bits
    runs = [D.PartRollout(p, be, *D.initial_part_state(p, x0, ea0, bc), device="cuda") for p in parts]
    gr = D.GraphedPartRollout(runs, D.LocalComm(n_parts), ms)
    for k in range(steps):
        got = gr.step()
        for o, g in zip(outs, got):
            assert torch.equal(o[k][0], g[0]) and torch.equal(o[k][1], g[1]), (math, n_parts, k)
