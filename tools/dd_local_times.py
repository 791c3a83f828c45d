"""Domain decomposition on ONE GPU (all parts in this process, LocalComm): time per step for 1 / 2 / 8 parts of the 1M-cell mesh next to
the plain single-domain forward — what the partitioned path itself costs, without any inter-GPU exchange.
usage: python tools/dd_local_times.py [workload=mesh1m] [parts=1,2,8]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
import fvgn_b200 as fv
from fvgn_b200 import domain as D

work = sys.argv[1] if len(sys.argv) > 1 else "mesh1m"
plist = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2,8").split(",")]
dev = torch.device("cuda", 0)
kind, kw, _, _ = bench.WORKLOADS[work]
m = fv.synthetic.make(kind, seed=4242, T=3, **kw)
t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
torch.manual_seed(0)
model = fv.FVGN(device="cpu", params=bench.params()).to(dev).eval().set_math_mode("bf16x3")
gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device=dev)
gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
x0, ea0, bc = gc.x.clone(), gc.edge_attr.clone(), ge.y[:, 0, 0:2].clone()
state = fv.dataset.RolloutState(gc, ge)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
with torch.no_grad():
    for rep in range(2):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=bench.RHO, mu=bench.MU, dt=bench.DT, edge_one_hot=9, cell_one_hot=0)
            state.advance_(gc, nuv)
        e1.record()
        torch.cuda.synchronize()
print("single-domain forward: %.2f ms/step" % (e0.elapsed_time(e1) / steps))
ms = dict(D.KernelBackend.state(model), dt=bench.DT, rho=bench.RHO)
for n in plist:
    parts, _ = D.decompose({k: v for k, v in tab.items()}, n)
    be = D.KernelBackend(model)
    runs = [D.PartRollout(p, be, *D.initial_part_state(p, x0, ea0, bc), device=dev) for p in parts]
    comm = D.LocalComm(n)
    with torch.no_grad():
        D.rollout_steps(runs, comm, ms, 1)
        torch.cuda.synchronize()
        e0.record()
        D.rollout_steps(runs, comm, ms, steps)
        e1.record()
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / steps
        # per part alone (no exchange partners needed for timing the kernels: time part 0's pieces)
        gr = D.GraphedPartRollout(runs, comm, ms)
        for _ in range(2):
            gr.step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            gr.step()
        e1.record()
        torch.cuda.synchronize()
    print("%d part(s) on one GPU: eager %.2f ms/step, one CUDA graph %.2f ms/step; cells per part %s, ghosts %s, faces %s"
          % (n, eager, e0.elapsed_time(e1) / steps, [p.C_own for p in parts][:3], [p.G for p in parts][:3], [p.F for p in parts][:3]))
    del runs, gr

if "--profile" in sys.argv:
    import collections

    from torch.profiler import ProfilerActivity, profile

    n = plist[-1]
    parts, _ = D.decompose({k: v for k, v in tab.items()}, n)
    be = D.KernelBackend(model)
    runs = [D.PartRollout(p, be, *D.initial_part_state(p, x0, ea0, bc), device=dev) for p in parts]
    comm = D.LocalComm(n)
    with torch.no_grad():
        D.rollout_steps(runs, comm, ms, 1)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            D.rollout_steps(runs, comm, ms, 1)
            torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            a = agg[e.name[:110]]
            a[0] += 1
            a[1] += e.device_time
    tot = sum(v[1] for v in agg.values())
    print("one step of %d parts: %d kernels, %.2f ms of GPU time" % (n, sum(v[0] for v in agg.values()), tot / 1e3))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print("%9.1f us %5.1f%%  x%-5d %s" % (v, 100 * v / tot, c, k))
