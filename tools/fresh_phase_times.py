"""Host time (perf_counter, no synchronisation) and GPU time (CUDA events) per phase of the eager training step with a fresh batch
topology every step (train.py:119-160).  usage: python tools/fresh_phase_times.py [pool=32] [steps=20]"""
import random
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
import fvgn_b200 as fv

pool_n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
pool = []
for i in range(pool_n):
    m = fv.synthetic.make("cyl", seed=5000 + i, T=3)
    tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
    pool.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device=dev))
p = bench.params()
p.loss = "global_mean_sum"
torch.manual_seed(0)
model = fv.FVGN(device="cpu", params=p).to(dev).train().set_math_mode("bf16x3")
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
rng = random.Random(0)
PH = ["collate", "preprocess", "zero_grad", "forward", "loss", "backward", "optimizer"]
host = {k: 0.0 for k in PH}
gpu = {k: [] for k in PH}


def step(record):
    marks = []

    def mark(name):
        if record:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, time.perf_counter(), ev))

    mark("start")
    lists = [[bench._shallow(b) for b in pool[i]] for i in rng.sample(range(pool_n), 16)]
    gn, ge, gc = fv.dataset.collate(lists)
    mark("collate")
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc])
    mark("preprocess")
    opt.zero_grad(set_to_none=True)
    mark("zero_grad")
    out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=bench.RHO, mu=bench.MU, dt=bench.DT, edge_one_hot=9, cell_one_hot=0)
    mark("forward")
    loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                               target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    mark("loss")
    loss[0].backward()
    mark("backward")
    opt.step()
    mark("optimizer")
    return marks


for _ in range(8):
    step(False)
torch.cuda.synchronize()
all_marks = []
t0 = time.perf_counter()
for _ in range(steps):
    all_marks.append(step(True))
host_total = (time.perf_counter() - t0) / steps
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
for marks in all_marks:
    for (n0, t0_, e0), (n1, t1_, e1) in zip(marks[:-1], marks[1:]):
        host[n1] += (t1_ - t0_) / steps
        gpu[n1].append(e0.elapsed_time(e1))
print("executor:", fv.autograd.EXECUTOR if hasattr(fv, "autograd") else "?")
print(f"wall {wall * 1e3:.2f} ms/step ({16 / wall:.0f} graphs/s); host-side issue time {host_total * 1e3:.2f} ms/step")
print(f"{'phase':12s} {'host ms':>9s} {'gpu ms (between phase marks on the stream)':>45s}")
for k in PH:
    print(f"{k:12s} {host[k] * 1e3:9.3f} {sum(gpu[k]) / len(gpu[k]):45.3f}")
print(f"{'sum':12s} {sum(host.values()) * 1e3:9.3f} {sum(sum(v) / len(v) for v in gpu.values()):45.3f}")
