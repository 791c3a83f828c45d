"""Wait-time accounting of the bf16x3 forward kernel (library built with FVGN_NVCC_EXTRA=-DFVGN_TC_TIMING_EXPERIMENTS=1).
usage: python tools/x3_prof.py [F] [C]   -> average clocks per 128-row tile, per role"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import fvgn_b200 as fv

F = int(sys.argv[1]) if len(sys.argv) > 1 else 86056
Cn = int(sys.argv[2]) if len(sys.argv) > 2 else 56291
lib = fv._lib.lib()
buf = (C.c_ulonglong * 32)()


def read(reset=1):
    lib.fvgn_dbg_x3_prof(buf, reset)
    return list(buf)


def show(tag, v, launches):
    tiles = v[7]
    per = lambda x: x / max(tiles, 1)
    print(f"== {tag}: {tiles // launches} tiles/launch")
    print("  MMA thread  per tile: wait DFREE %.0f | AFULL %.0f | WFULL(L1) %.0f | issue L1 %.0f | HREADY %.0f | WFULL(L23) %.0f | issue L23 %.0f" %
          tuple(per(v[i]) for i in range(7)))
    print("  epilogue w0 per tile: wait L1 %.0f | wait L2 %.0f | wait L3 %.0f | hidden work (2x) %.0f | final work %.0f" % tuple(per(v[8 + i]) for i in range(5)))
    print("  producer t0 per tile: gather issue %.0f | wait AEMPTY %.0f | split+store %.0f" % tuple(per(v[16 + i]) for i in range(3)))
    print("  loader      per tile: wait WEMPTY %.0f" % per(v[24]))


torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)
x = torch.randn(Cn, 128, generator=gen).cuda()
e = torch.randn(F, 128, generator=gen).cuda()
nc = torch.randint(0, Cn, (2, F), generator=gen, dtype=torch.int32).cuda()
nagg = torch.randn(Cn // 2, 64, generator=gen).cuda()
cnT = torch.randint(0, Cn // 2, (3, Cn), generator=gen, dtype=torch.int32).cuda()
eb = fv.build_mlp(384, 128, 128, drop_out=False, lay_norm=True).cuda()
cb = fv.build_mlp(192, 128, 128, drop_out=False, lay_norm=True).cuda()
for m in (eb, cb):
    m.math_mode = "bf16x3"
L = 10
for it in range(2):
    read()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(L):
        fv.ops.edge_block_fwd(eb.mlp_ref(), x, e, nc, math_mode=2, e_out=e)
    ev1.record()
    torch.cuda.synchronize()
    t = ev0.elapsed_time(ev1) / L
    if it:
        print("edge launch %.1f us" % (t * 1e3))
        show("EDGE", read(), L)
for it in range(2):
    read()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(L):
        fv.ops.cell_block_fwd(cb.mlp_ref(), x, nagg, cnT, math_mode=2, x_out=x)
    ev1.record()
    torch.cuda.synchronize()
    t = ev0.elapsed_time(ev1) / L
    if it:
        print("cell launch %.1f us" % (t * 1e3))
        show("CELL", read(), L)

# split-shadow (inference fast path) variants
xps = torch.empty((Cn, 2, 128), dtype=torch.float16, device="cuda").normal_()
cms = torch.empty((Cn, 2, 64), dtype=torch.float16, device="cuda").normal_()
import ctypes
P = lambda t: ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, call in (("EDGE split", lambda: lib.fvgn_edge_block_fwd_split(ctypes.byref(eb.mlp_ref().struct), P(xps), P(e), P(nc), F, S())),
                   ("CELL split", lambda: lib.fvgn_cell_block_fwd_split(ctypes.byref(cb.mlp_ref().struct), P(x), P(cms), Cn, P(xps), S())),
                   ("cell mean", lambda: lib.fvgn_cell_mean_split(P(nagg), P(cnT), Cn, P(cms), S()))):
    for it in range(2):
        e.normal_(), x.normal_()
        read()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(L):
            rc = call()
            assert rc == 0, lib.fvgn_last_error()
        ev1.record()
        torch.cuda.synchronize()
        if it:
            print("%s launch %.1f us" % (name, ev0.elapsed_time(ev1) / L * 1e3))
            show(name, read(), L)
