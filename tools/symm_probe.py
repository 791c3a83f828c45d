"""probe: torch symmetric memory (peer buffers + signals) between the GPUs of one box.  torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
print(rank, "backend", symm_mem.get_backend(dev) if hasattr(symm_mem, "get_backend") else "?", flush=True)
buf = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok: world", hdl.world_size, "rank", hdl.rank, "multicast", hdl.has_multicast_support, flush=True)
peer = (rank + 1) % world
src = torch.full((1000, 128), float(rank + 1), device=dev)
idx = torch.arange(0, 1000, 2, device=dev)
buf.zero_()
hdl.barrier()
pv = hdl.get_buffer(peer, (500, 128), torch.float32, 0)
torch.index_select(src, 0, idx, out=pv)
hdl.put_signal(peer, 0)
hdl.wait_signal((rank - 1) % world, 0)
got = buf[: 500 * 128].view(500, 128)
torch.cuda.synchronize()
print(rank, "received from", (rank - 1) % world, "value", float(got[0, 0]), "ok", bool((got == float((rank - 1) % world + 1)).all()), flush=True)


def once():
    torch.index_select(src, 0, idx, out=pv)
    hdl.put_signal(peer, 0)
    hdl.wait_signal((rank - 1) % world, 0)


for _ in range(20):
    once()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
e0.record()
for _ in range(200):
    once()
e1.record()
torch.cuda.synchronize()
print(rank, "eager: %.1f us per exchange (256 KB)" % (e0.elapsed_time(e1) / 200 * 1e3), flush=True)
# the same inside a CUDA graph
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
try:
    with torch.cuda.stream(s):
        once()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.graph(g):
        for _ in range(31):
            once()
    torch.cuda.synchronize()
    dist.barrier()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(rank, "graph of 31 exchanges: %.1f us per exchange" % (e0.elapsed_time(e1) / 20 / 31 * 1e3), flush=True)
except Exception as e:
    print(rank, "graph capture failed:", type(e).__name__, e, flush=True)
# NCCL point-to-point for comparison
recv = torch.empty((500, 128), device=dev)


def nccl_once():
    ops = [dist.P2POp(dist.isend, src.index_select(0, idx), peer), dist.P2POp(dist.irecv, recv, (rank - 1) % world)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


for _ in range(10):
    nccl_once()
torch.cuda.synchronize()
dist.barrier()
e0.record()
for _ in range(200):
    nccl_once()
e1.record()
torch.cuda.synchronize()
print(rank, "NCCL batch_isend_irecv eager: %.1f us per exchange" % (e0.elapsed_time(e1) / 200 * 1e3), flush=True)
g2 = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g2):
        for _ in range(31):
            nccl_once()
    torch.cuda.synchronize()
    dist.barrier()
    for _ in range(3):
        g2.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g2.replay()
    e1.record()
    torch.cuda.synchronize()
    print(rank, "NCCL graph of 31 exchanges: %.1f us per exchange" % (e0.elapsed_time(e1) / 20 / 31 * 1e3), flush=True)
except Exception as e:
    print(rank, "NCCL graph capture failed:", type(e).__name__, e, flush=True)
torch.cuda.synchronize()
os._exit(0)  # (tearing the process group down after NCCL ops were captured into graphs hung on the box this was written on)
