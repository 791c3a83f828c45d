"""Per-kernel GPU-time breakdown of one eager training step (cyl16, batch 16, default math mode) with the torch profiler (CUPTI).
usage: python tools/train_breakdown.py [math] [graph|fresh]   ("graph": profile one replay of training.GraphedTrainStep instead;
"fresh": the eager step with a new random batch topology every step — collate + CSR tables inside, fused AdamW — three steps profiled)"""
import collections
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench
import fvgn_b200 as fv

dev = torch.device("cuda", 0)
meshes = bench.make_meshes("cyl16", 0)
lists = []
for m in meshes:
    t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
    tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
    lists.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device=dev))
gn0, ge0, gc0 = fv.dataset.collate(lists)
p = bench.params()
p.loss = "global_mean_sum"
torch.manual_seed(0)
model = fv.FVGN(device="cpu", params=p).to(dev).train()
model.set_math_mode(sys.argv[1] if len(sys.argv) > 1 else "bf16x3")
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)


def tstep():
    gn, ge, gc = bench._shallow(gn0), bench._shallow(ge0), bench._shallow(gc0)
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc])
    opt.zero_grad(set_to_none=False)
    out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=bench.RHO, mu=bench.MU, dt=bench.DT, edge_one_hot=9, cell_one_hot=0)
    loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                               target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    loss[0].backward()
    opt.step()


if "graph" in sys.argv[2:]:
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=True, fused=True)  # one fused multi-tensor kernel per chunk (the foreach path divides by 0-dim step tensors one parameter at a time: 524 launches)
    gn_s, ge_s, gc_s = bench._shallow(gn0), bench._shallow(ge0), bench._shallow(gc0)
    gn_s.x = gn0.x.clone()
    graphed = fv.GraphedTrainStep(model, opt, p, gn_s, ge_s, gc_s, bench.RHO, bench.MU, bench.DT)
    tstep = graphed.step  # noqa: F811
NPROF = 1
if "fresh" in sys.argv[2:]:
    import random

    pool = list(lists)
    for i in range(16):
        m = fv.synthetic.make("cyl", seed=77000 + i, T=3)
        t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
        tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
        pool.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device=dev))
    rng = random.Random(0)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    NPROF = 3

    def tstep():  # noqa: F811
        bl = [[bench._shallow(b) for b in pool[i]] for i in rng.sample(range(len(pool)), 16)]
        gn, ge, gc = fv.dataset.collate(bl)
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc])
        opt.zero_grad(set_to_none=True)
        out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=bench.RHO, mu=bench.MU, dt=bench.DT, edge_one_hot=9, cell_one_hot=0)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        opt.step()

    for _ in range(30):
        tstep()
for _ in range(3):
    tstep()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NPROF):
        tstep()
    torch.cuda.synchronize()
iv = sorted((e.time_range.start, e.time_range.end) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA)
busy, cur_s, cur_e = 0.0, None, None
for a, b in iv:
    if cur_e is None or a > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
busy += (cur_e - cur_s) if cur_e is not None else 0.0
ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
gaps, end, last = [], None, None
for a, b, n in ev:
    if end is not None and a > end:
        gaps.append((a - end, last, n))
    if end is None or b > end:
        end, last = b, n
print("largest idle gaps of the GPU timeline (us, kernel before -> kernel after):")
for g, x, y in sorted(gaps, key=lambda t: -t[0])[:14]:
    print("  %7.1f  %s -> %s" % (g, x[:70], y[:70]))
print("gaps > 20 us: %d, total %.0f us; gaps <= 20 us: %d, total %.0f us (over %d step(s))"
      % (sum(1 for g in gaps if g[0] > 20), sum(g[0] for g in gaps if g[0] > 20), sum(1 for g in gaps if g[0] <= 20),
         sum(g[0] for g in gaps if g[0] <= 20), NPROF))
print("GPU timeline of the %d profiled step(s): span %.2f ms/step, some kernel running %.2f ms/step, idle %.2f ms/step"
      % (NPROF, (iv[-1][1] - iv[0][0]) / 1e3 / NPROF, busy / 1e3 / NPROF, ((iv[-1][1] - iv[0][0]) - busy) / 1e3 / NPROF))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[e.name[:90] if 'elementwise_kernel' not in e.name else e.name[:400]]
        a[0] += 1
        a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print("%d training step(s): %d kernels, %.2f ms of GPU time" % (NPROF, sum(v[0] for v in agg.values()), tot / 1e3))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:(60 if NPROF > 1 else 28)]:
    print("%8.1f us %5.1f%%  x%-4d %s" % (v, 100 * v / tot, c, k))
