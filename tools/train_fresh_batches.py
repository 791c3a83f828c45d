"""Training step with a FRESH batch of meshes every step (the reference's DataLoader: 16 random (mesh, time step) samples per batch, so the
batch topology changes every step and nothing can be replayed from one CUDA graph): collate + topology tables + train_datapreprocessing +
forward + loss + backward + AdamW, eager.  Reports ms/step and graphs/s next to the fixed-batch CUDA-graph replay.
usage: python tools/train_fresh_batches.py [pool=64] [steps=20]"""
import random
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
import fvgn_b200 as fv

pool_n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
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


def step():
    lists = [[bench._shallow(b) for b in pool[i]] for i in rng.sample(range(pool_n), 16)]
    gn, ge, gc = fv.dataset.collate(lists)
    gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc])
    opt.zero_grad(set_to_none=True)
    out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=bench.RHO, mu=bench.MU, dt=bench.DT, edge_one_hot=9, cell_one_hot=0)
    loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                               target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
    loss[0].backward()
    opt.step()
    return loss[0]


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    l = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
print("fresh batch every step (eager, collate + topology inside): %.2f ms/step = %.0f graphs/s, loss %.4f" % (dt * 1e3, 16 / dt, float(l)))
# host-side share: the same loop without waiting for the GPU between steps is what was measured above; time the host alone
import cProfile
import pstats

pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
