#!/usr/bin/env python
"""bench.py — FVGN hot-path benchmark (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cyl16|mesh1m|foil16] [--math fp32|bf16|bf16x3]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference's CPU torch path (oracle port) on the host cores

A "step" is one autoregressive rollout step of the hot path over one batch of synthetic meshes:
FVGN.forward (normalise -> encoder -> 15 GnBlocks -> decoder -> face-area BN -> flux integrator -> de-normalise)
followed by create_next_graph — i.e. one cell-update for every cell of the batch.
Metric: cell-updates/s = cells_in_batch x steps / seconds (whole job, all ranks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

RHO, MU, DT = 1.0, 1e-3, 0.01
WORKLOADS = {
    # name: (mesh kind, kwargs, graphs per rank, description)
    "cyl16": ("cyl", {}, 16, "16 synthetic cylinder_flow-shaped meshes (~1.9k nodes / ~3.5k cells each), batch 16, fp32 (BASELINE configs[1] shape)"),
    "foil16": ("foil", {"n_target": 6400}, 16, "16 synthetic airfoil-shaped meshes (~5k nodes each)"),
    "mesh1m": ("grid", {"nx": 708, "ny": 708}, 1, "one scaled synthetic mesh of ~1.0M cells (BASELINE configs[4] kernel roofline sweep)"),
    "mesh1m_ordered": ("grid", {"nx": 708, "ny": 708, "shuffle": False}, 1, "~1.0M cells, cells numbered in generation (row-major) order"),
    "mesh1m_renumbered": ("grid", {"nx": 708, "ny": 708, "renumber": "cells"}, 1, "mesh1m (random cell numbering) after fvgn_b200.renumber_mesh (cells along a Z-order curve; same results up to that permutation)"),
    "mesh1m_renumbered_all": ("grid", {"nx": 708, "ny": 708, "renumber": "all"}, 1, "mesh1m after renumber_mesh(nodes=True): cells, nodes and with them the faces along the Z-order curve (another valid numbering: edge orientations differ)"),
    "tiny": ("grid", {"nx": 25, "ny": 17}, 2, "debug"),
}


def params(**over):
    """the hot-path fields of the reference's utils/get_param.py:50-144 with their defaults (what train.py / rollout.py pass to FVGN)"""
    import types

    p = dict(edge_input_size=5, edge_one_hot=9, cell_input_size=2, cell_one_hot=0, dual_edge=False, message_passing_num=15,
             cell_output_size=2, edge_output_size=5, drop_out=False, mp_times=2, multihead=1, nn_act="SiLU", loss_cont=0,
             loss_mom=10, loss_face_uv=1, loss_face_p=1, loss="direct_mean_loss", cell_target_input_size=2,
             face_flux_target_input_size=3, dataset_size=1000, batch_size=16, cumulative_length=600, grid_feature_size=1,
             dataset_type="h5", n_epochs=1, train_traj_length=2, Noise_injection_factor=0.02, rho=RHO, mu=MU, dt=DT)
    p.update(over)
    return types.SimpleNamespace(**p)


def config_of(workload, C, F, N, graphs, world):
    """the `config` object — IDENTICAL in the native and the reference arm (same workload, same sizes, same step)"""
    return {"workload": workload, "description": WORKLOADS[workload][3], "cells_per_rank": int(C), "faces_per_rank": int(F),
            "nodes_per_rank": int(N), "graphs_per_rank": int(graphs), "message_passing_layers": 15, "hidden": 128,
            "step": "one autoregressive rollout step: FVGN.forward (eval) + create_next_graph over the whole batch",
            "l2": "GPU arms: flushed between timed steps (256 MiB write); CPU arm: not applicable",
            "parallelism": f"replicas x{world} (independent mesh batches, no collective)"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th = index, [], threading.Event(), None

    def _run_nvml(self):
        """in-process NVML polling every ~5 ms (the timed region of the default workload is ~50 ms: one nvidia-smi process per sample sees it
        once); returns False when NVML is not usable, and the nvidia-smi loop takes over"""
        try:
            import pynvml as N

            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), reasons_fn(h)  # probe once before committing to this path
        except Exception:
            return False
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self.stop.is_set():
            try:
                sm, r = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), int(reasons_fn(h))
                self.rows.append([str(sm), str(mx), ""] + ["Active" if r & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self.stop.wait(0.005)
        return True

    def _run(self):
        if self._run_nvml() and self.rows:
            return
        while True:  # at least one nvidia-smi sample even if the region has already ended
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            if self.stop.wait(0.2):
                break

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def make_meshes(workload, rank):
    import fvgn_b200 as fv

    kind, kw, n_graphs, _ = WORKLOADS[workload]
    return [fv.synthetic.make(kind, seed=1000 * rank + i, T=3, **kw) for i in range(n_graphs)]


def device_batch(meshes, device):
    """connectivity on device (a1) -> reference-layout graphs -> PyG-style batch -> valid_datapreprocessing (a3)."""
    import fvgn_b200 as fv

    lists = []
    for m in meshes:
        t = lambda a, dt: torch.as_tensor(a).to(dt).to(device)
        tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
        lists.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device=device))
    gn, ge, gc = fv.dataset.collate(lists)
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    return gn, ge, gc, mi, mb


def measure_forward(args, workload, math, steps, rank, local_rank, dev, dist):
    """One rollout benchmark: K steps of FVGN.forward (eval) + create_next_graph on `workload` in math mode `math`.
    Returns device-resident timing, end-to-end (host buffers) timing, per-kernel CUDA-event timings and clock samples."""
    import fvgn_b200 as fv
    from fvgn_b200 import ops

    meshes = make_meshes(workload, rank)
    gn, ge, gc, mi, mb = device_batch(meshes, dev)
    C, F, N = gc.x.shape[0], ge.x.shape[0], gn.x.shape[0]
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=params()).to(dev)
    model.set_math_mode(math)
    # prime the normalisers like the reference's first training rounds (2 accumulating forwards), then rollout mode
    model.train()
    with torch.no_grad():
        for _ in range(2):
            gtmp = gc.clone()
            gtmp.y = gc.y[:, 0, :]
            getmp = ge.clone()
            getmp.y = ge.y[:, 0, :]
            model(graph=gtmp, graph_edge=getmp, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0)
    model.eval()
    x_host0 = gc.x.cpu().pin_memory()
    ea_host0 = gc.edge_attr.cpu().pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # The step a user runs in a rollout loop: fv.GraphedRollout.step() = FVGN.forward (eval) + create_next_graph captured once into a
    # CUDA graph (same kernels, same bits as the eager calls; the ~57 Python-side launches per step no longer bound a cylinder-sized step)
    l0 = ops.LAUNCHES
    gr = fv.GraphedRollout(model, gn, ge, gc, RHO, MU, DT, edge_one_hot=9, cell_one_hot=0, warmup=2)
    launches_per_step = (ops.LAUNCHES - l0) // 3  # 2 eager warm-up steps + the captured one
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            gr.step()
        # ---------------- device-resident timing: K steps, L2 flushed between steps, CUDA events per step
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        with ClockSampler(local_rank) as clk:
            for k in range(steps):
                flush.zero_()
                ev[k][0].record()
                gr.step()
                ev[k][1].record()
            barrier()
        launches = launches_per_step * steps
        ms = sum(a.elapsed_time(b) for a, b in ev)
        # ---------------- end-to-end through the public API with HOST buffers: H2D inputs + D2H results every step
        x_host, ea_host = x_host0.clone().pin_memory(), ea_host0.clone().pin_memory()
        nuv_host = torch.empty((C, 2), dtype=torch.float32).pin_memory()
        uvp_host = torch.empty((F, 3), dtype=torch.float32).pin_memory()

        def e2e_step():
            gr.state.x0.copy_(x_host, non_blocking=True), gr.state.ea0.copy_(ea_host, non_blocking=True)
            uvp, nuv = gr.step()
            nuv_host.copy_(nuv, non_blocking=True), uvp_host.copy_(uvp, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller consumes the result on the host every step
            x_host[:, 2:4].copy_(nuv_host)              # host-side autoregression (next step's input comes from the host)

        for k in range(2):
            e2e_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(steps):
            e2e_step()
        t1.record()
        barrier()
        e2e_ms = t0.elapsed_time(t1)
        # ---------------- per-kernel CUDA-event timings (outside the timed regions): two eager steps of the same model
        ksteps = 2
        ops.TIMING = {"edge_block": [], "cell_block": [], "node_aggregate": [], "cell_mean": []}
        for k in range(ksteps):
            flush.zero_()
            gr._eager_step()
        torch.cuda.synchronize()
        kern = {k: [a.elapsed_time(b) for a, b in v] for k, v in ops.TIMING.items()}
        ops.TIMING = None
    del gr
    tms = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    cells = torch.tensor([float(C)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    del flush
    torch.cuda.empty_cache()
    return dict(ms=float(tms[0]), e2e_ms=float(tms[1]), total_cells=float(cells[0]), C=C, F=F, N=N, graphs=len(meshes), launches=int(launches),
                kern=kern, ksteps=ksteps, clocks=clk.summary(), meshes=meshes, steps=steps)


def rollout_config0(args, dev):
    """BASELINE configs[0]: 1-step forward + 10-step rollout, batch 1, one cylinder_flow-shaped mesh — the reference's own CPU-runnable
    case.  Eager loop vs the CUDA-graph rollout (fvgn_b200.GraphedRollout: zero host work per step)."""
    import fvgn_b200 as fv

    m = fv.synthetic.make("cyl", seed=0, T=3)
    gn, ge, gc, mi, mb = device_batch([m], dev)
    C = gc.x.shape[0]
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=params()).to(dev).eval().set_math_mode(args.math)
    state = fv.dataset.RolloutState(gc, ge)

    def eager(n):
        for _ in range(n):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0)
            state.advance_(gc, nuv)

    def timed(fn, reps=5):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)) * 1e3

    with torch.no_grad():
        eager(3)
        e1, e10 = timed(lambda: eager(1)), timed(lambda: eager(10))
        gn2, ge2, gc2, _, _ = device_batch([m], dev)
        roll = fv.GraphedRollout(model, gn2, ge2, gc2, RHO, MU, DT)
        roll.run(3)
        g1, g10 = timed(lambda: roll.run(1)), timed(lambda: roll.run(10))
    return {"workload": "1 cylinder_flow-shaped mesh, batch 1 (BASELINE configs[0])", "cells": C, "math": args.math,
            "eager_ms": {"1_step": e1, "10_step_rollout": e10}, "cuda_graph_ms": {"1_step": g1, "10_step_rollout": g10},
            "cell_updates_per_sec_cuda_graph": C * 10 / (g10 * 1e-3), "timing": "host wall clock around a synchronised call, median of 5"}


def measured_peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(pk)) if os.path.exists(pk) else {}


def roofline_of(m, math, workload):
    """roofline object of the dominant kernel (the fused EdgeBlock kernel) from a measure_forward() result"""
    peaks = measured_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    C, F = m["C"], m["F"]
    # Algorithmic bytes per launch, fp32 contract: e read + e written (2*512 B/face) + every x' row read once
    # (512 B/cell) + neighbour_cell (8 B/face).  (The bf16 x' shadow of the tensor-core path moves 256 B/cell instead.)
    edge_bytes = F * (2 * 512 + 8) + C * 512
    kern = m["kern"]
    edge_ms = float(np.mean(kern["edge_block"])) if kern["edge_block"] else None
    ach = edge_bytes / (edge_ms * 1e-3) / 1e9 if edge_ms else None
    shares = {k: ((sum(v) / m["ksteps"]) / (m["ms"] / m["steps"]) if v else 0.0) for k, v in kern.items()}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get(f"{workload}:{math}:edge_block")
        traffic = t["dram_bytes_per_launch"] if t else None
    name = {"fp32": "fused_mlp_simt_kernel<EDGE>", "bf16": "fused_mlp_tc_kernel<TC_EDGE, RED, XBF>",
            "bf16x3": "x3::fused_mlp_tc3_kernel<TC_EDGE, RED, SH>"}[math]
    if math == "bf16x3" and edge_ms:
        # the fp32-equivalent forward kernel runs 3 fp16 MMAs per algorithmic product (two-term fp16 split of both operands).  Algorithmic
        # GEMM FLOPs per face: 2*(384*128 + 2*128^2) = 163,840; executed on the tensor pipe: 3x that.
        tf_peak = peaks.get("bf16_tflops_sustained", 1375.0)
        alg = 163840.0 * F / (edge_ms * 1e-3) / 1e12
        return {"kernel": name, "workload": workload, "math": math, "bound": "tensor", "achieved": 3.0 * alg, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": 3.0 * alg / tf_peak, "traffic": traffic, "avg_launch_ms": edge_ms,
                "peak_source": ("measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1375 TFLOP/s"),
                "algorithmic_tflops_fp32_equivalent": alg, "executed_mma_flops_per_launch": int(3 * 163840 * F),
                "launches_timed": len(kern["edge_block"]), "share_of_step": shares,
                "note": "achieved = executed 16-bit MMA FLOPs (three two-term-split products per fp32-equivalent product) / CUDA-event launch "
                        "time; the algorithmic (fp32-equivalent) rate is a third of it: DESIGN.md §4.3"}
    return {"kernel": name, "workload": workload, "math": math, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
            "frac": (ach / hbm) if ach else None, "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": edge_ms,
            "algorithmic_bytes_per_launch": int(edge_bytes), "launches_timed": len(kern["edge_block"]), "share_of_step": shares,
            "note": {"fp32": "strict-fp32 mode is FFMA(compute)-bound (5.68 MFLOP/cell-update on the FP32 pipes), not HBM-bound: DESIGN.md §4.2",
                     "bf16x3": "fp32-equivalent mode is tensor/epilogue-bound (3 MMAs per product, all weights streamed from L2), not HBM-bound: "
                               "DESIGN.md §4.3", "bf16": ""}[math]}


DTYPES = {"fp32": "f32 (FFMA)", "bf16": "bf16 tiles, f32 accumulate/residual",
          "bf16x3": "f32-equivalent: two-term fp16 split of both operands on tcgen05 (3 MMAs/product; training backward: the 3 leading products "
                    "of 3-way bf16 splits, gradients held to 1e-4), f32 accumulate + f32 epilogue; held to the f32 parity bound (rel 1e-5/step)"}


def native(args, rank, world, local_rank):
    import fvgn_b200 as fv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:  # host-side setup (mesh generation, collation) of the ranks shares one box: no OpenMP oversubscription
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))
    fv._lib.lib()
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    m = measure_forward(args, args.workload, args.math, args.steps, rank, local_rank, dev, dist)
    meshes = m["meshes"]
    C, F, N = m["C"], m["F"], m["N"]
    result = None
    if rank == 0:
        result = {
            "metric": "cell_updates_per_sec", "value": m["total_cells"] * args.steps / (m["ms"] * 1e-3), "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": m["ms"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[args.math], "data": "synthetic",
            "config": config_of(args.workload, C, F, N, len(meshes), world),
            "math_mode": args.math,
            "implementation": "fv.GraphedRollout.step(): FVGN.forward (eval) + create_next_graph replayed from one CUDA graph of libfvgn_b200 kernels",
            "graphs_per_sec": len(meshes) * world * args.steps / (m["ms"] * 1e-3),
            "e2e": {"value": m["total_cells"] * args.steps / (m["e2e_ms"] * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": int(C * 16 + F * 20),
                    "d2h_bytes_per_step": int(C * 8 + F * 12), "ms_per_step": m["e2e_ms"] / args.steps,
                    "note": "pinned host inputs copied H2D and results D2H + host sync every step"},
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
            "roofline": roofline_of(m, args.math, args.workload),
        }
    # ---- kernel roofline sweep (BASELINE configs[4]): the tcgen05 kernels on the 1M-cell mesh, with its own roofline object under
    # `sweep.roofline` (the line's `roofline` describes the kernel that produced the line's `value`)
    if args.sweep != "none" and tuple(args.sweep.split(":")) != (args.workload, args.math):
        sw_work, sw_math = args.sweep.split(":")
        sm = measure_forward(args, sw_work, sw_math, min(args.steps, 5), rank, local_rank, dev, dist)
        if rank == 0:
            sweep_roof = roofline_of(sm, sw_math, sw_work)
            result["sweep"] = {"workload": sw_work, "math": sw_math, "dtype": DTYPES[sw_math], "cells_per_rank": sm["C"], "faces_per_rank": sm["F"],
                               "roofline": sweep_roof,
                               "value": sm["total_cells"] * sm["steps"] / (sm["ms"] * 1e-3), "unit": "cell-updates/s",
                               "ms_per_step": sm["ms"] / sm["steps"], "steps": sm["steps"],
                               "e2e": {"value": sm["total_cells"] * sm["steps"] / (sm["e2e_ms"] * 1e-3), "unit": "cell-updates/s"},
                               "whole_step_roofline": {"bytes_per_cell_update": 64200, "peak": measured_peaks().get("hbm_gbs", 6650.0), "unit": "GB/s",
                                                       "frac": 64200 * sm["total_cells"] * sm["steps"] / (sm["ms"] * 1e-3) / 1e9 / world
                                                       / measured_peaks().get("hbm_gbs", 6650.0)},
                               "clocks": sm["clocks"], "gpu_launches": sm["launches"]}
        del sm
        for suffix in ("_renumbered", "_renumbered_all"):
            # the same mesh after the converter's renumbering (fvgn_b200.renumber_mesh; "cells": same results up to that permutation, "all":
            # nodes and faces too — another valid numbering): what the random cell numbering of the sweep mesh costs in gather traffic
            if sw_work + suffix not in WORKLOADS or args.no_renumbered_leg:
                continue
            rm = measure_forward(args, sw_work + suffix, sw_math, min(args.steps, 5), rank, local_rank, dev, dist)
            if rank == 0:
                rr = roofline_of(rm, sw_math, sw_work + suffix)
                v = rm["total_cells"] * rm["steps"] / (rm["ms"] * 1e-3)
                result["sweep"][suffix[1:]] = {
                    "workload": sw_work + suffix, "description": WORKLOADS[sw_work + suffix][3], "value": v, "unit": "cell-updates/s",
                    "ms_per_step": rm["ms"] / rm["steps"], "edge_block_frac": rr["frac"], "share_of_step": rr.get("share_of_step"),
                    "whole_step_roofline_frac": 64200 * v / 1e9 / world / measured_peaks().get("hbm_gbs", 6650.0)}
            del rm
    if args.math != "fp32" and not args.no_strict_leg:
        fm = measure_forward(args, args.workload, "fp32", 3, rank, local_rank, dev, dist)
        if rank == 0:
            result["strict_fp32_ffma"] = {"value": fm["total_cells"] * fm["steps"] / (fm["ms"] * 1e-3), "unit": "cell-updates/s",
                                          "ms_per_step": fm["ms"] / fm["steps"], "steps": fm["steps"], "dtype": DTYPES["fp32"],
                                          "note": "the same workload on the strict FFMA kernels (bit-level evaluation order of the reference)"}
        del fm
    if rank == 0 and not args.no_config0:
        result["rollout_config0"] = rollout_config0(args, dev)
    train = None
    if not args.no_training:
        torch.cuda.empty_cache()
        train = training_measure(args, rank, world, dev, meshes, dist)
    if not args.no_foil_leg and (args.workload, args.math) != ("foil16", "bf16"):
        # BASELINE configs[2] on every N: airfoil-shaped meshes (~5k nodes, far field), bf16 MLP tiles, data-parallel training
        import copy

        a2 = copy.copy(args)
        a2.workload, a2.math = "foil16", "bf16"
        torch.cuda.empty_cache()
        fm = measure_forward(a2, "foil16", "bf16", min(args.steps, 10), rank, local_rank, dev, dist)
        ft = training_measure(a2, rank, world, dev, fm["meshes"], dist, fresh=False) if not args.no_training else None
        if rank == 0:
            result["foil16_bf16"] = {
                "workload": "foil16", "description": WORKLOADS["foil16"][3] + " (BASELINE configs[2])", "math": "bf16", "dtype": DTYPES["bf16"],
                "cells_per_rank": fm["C"], "faces_per_rank": fm["F"], "nodes_per_rank": fm["N"], "graphs_per_rank": fm["graphs"],
                "rollout": {"value": fm["total_cells"] * fm["steps"] / (fm["ms"] * 1e-3), "unit": "cell-updates/s", "ms_per_step": fm["ms"] / fm["steps"],
                            "steps": fm["steps"], "e2e": {"value": fm["total_cells"] * fm["steps"] / (fm["e2e_ms"] * 1e-3), "unit": "cell-updates/s"},
                            "gpu_launches": fm["launches"], "clocks": fm["clocks"]},
                "training": ft}
        del fm
    if rank == 0:
        if train is not None:
            result["training"] = train
        if not args.no_torch_eager and world == 1:
            te = torch_eager_cuda(args, dev, args.steps)
            result["torch_eager_cuda"] = te
            if "unavailable" not in te:  # the speed-ups that matter: over the reference's own PyTorch path on the same GPU
                sp = {f"rollout_vs_{k}": result["value"] / v["value"] for k, v in te["rollout"].items()}
                if train is not None:
                    sp.update({f"training_vs_{k}": train["value"] / v["value"] for k, v in te["training"].items()})
                    sp.update({f"training_fresh_batches_vs_{k}": train["fresh_batches"]["value"] / v["value"] for k, v in te["training"].items()})
                result["speedup_over_torch_eager_cuda"] = sp
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only: at N > 1 the other ranks' host threads share the cores
            result["cpu_baseline"] = cpu_baseline(args, budget_s=args.cpu_seconds)
    if dist is not None and not args.no_dd_leg:
        torch.cuda.empty_cache()
        dd = domain_decomposition_leg(args, rank, world, dev, dist)
        if rank == 0:
            result["domain_decomposition"] = dd
            if not dd["pass"]:
                print(f"[bench] DOMAIN-DECOMPOSITION PARITY CHECK FAILED: {dd}", file=sys.stderr)
    if dist is not None and (args.check or world == 2):
        chk = dp_parity_check(rank, world, dev, dist)
        if rank == 0:
            result["dp_check"] = chk
            if not chk["pass"]:
                print(f"[bench] DATA-PARALLEL PARITY CHECK FAILED: {chk}", file=sys.stderr)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and result.get("dp_check", {}).get("pass") is False and args.check:
        print(json.dumps(result))
        sys.exit(3)
    return result


# ------------------------------------------------------------------------------------------------ training step
def _shallow(bag):
    """new attribute bag over the same tensors (index tables keep their identity => the topology cache hits)"""
    out = bag.__class__()
    out.__dict__.update(bag.__dict__)
    return out


def training_measure(args, rank, world, dev, meshes, dist, fresh=True):
    """training graphs/s: per step = H2D of the batch's node fields (e2e leg only) + train_datapreprocessing on device (a2) +
    FVGN.forward (training) + compute_FVM_loss + backward (all in libfvgn_b200 kernels) + flat-bucket gradient all-reduce
    (N>1, NCCL) + AdamW (train.py:150-225).  Strict fp32."""
    import fvgn_b200 as fv
    from fvgn_b200 import ops, parallel

    lists = []
    for m in meshes:
        t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
        tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
        lists.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device=dev))
    gn0, ge0, gc0 = fv.dataset.collate(lists)
    C, F = gc0.x.shape[0], ge0.x.shape[0]
    p = params()
    p.loss = "global_mean_sum"  # the reference's default (utils/get_param.py:68)
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=p).to(dev).train()
    train_math = args.math  # "bf16": the bf16x3 kernels with the leading split product only (bf16 tiles, fp32 accumulation)
    model.set_math_mode(train_math)  # forward arithmetic; the backward kernels are FFMA
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)  # train.py:45, get_param.py:80
    bucket = parallel.GradBucket(model.parameters())
    reducer = None
    if world > 1 and train_math != "fp32" and not args.no_overlap:
        # tensor-core training path: the gradient all-reduce is issued group by group from inside the backward (overlapped with it)
        reducer = parallel.OverlappedGradReducer(model, n_groups=args.dp_groups).attach(model)
    dp_sync = (lambda: None) if reducer is not None else bucket.all_reduce_mean_
    node_host = gn0.x.cpu().pin_memory()
    node_dev = gn0.x
    loss_host = torch.empty(5, dtype=torch.float32).pin_memory()

    def tstep(h2d):
        gn, ge, gc = _shallow(gn0), _shallow(ge0), _shallow(gc0)
        gn.x = node_host.to(dev, non_blocking=True) if h2d else node_dev
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc])
        opt.zero_grad(set_to_none=False)
        out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0)
        loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                   target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        dp_sync()
        opt.step()
        if h2d:
            loss_host.copy_(torch.stack([loss[0].detach().reshape(()), loss[2].mean(), loss[3].mean(), loss[4].mean(), loss[0].detach().reshape(())]),
                            non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps = max(2, min(args.steps, args.train_steps))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    graphed = None
    if args.train_graph:
        # the same step captured into one CUDA graph (fvgn_b200.GraphedTrainStep): no Python / launch overhead per step
        try:
            opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=True, fused=True)  # one fused multi-tensor kernel per chunk (the foreach path divides by 0-dim step tensors one parameter at a time: 524 launches)
            gn_s, ge_s, gc_s = _shallow(gn0), _shallow(ge0), _shallow(gc0)
            gn_s.x = node_dev.clone()
            graphed = fv.GraphedTrainStep(model, opt, p, gn_s, ge_s, gc_s, RHO, MU, DT, all_reduce=dp_sync if world > 1 else None)

            def tstep(h2d):  # noqa: F811
                if h2d:
                    graphed.load(node_host)
                loss4 = graphed.step()
                if h2d:
                    loss_host[:4].copy_(loss4, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                return (loss4[0:1],)
        except Exception as e:  # pragma: no cover
            print(f"[bench] CUDA-graph training step unavailable ({type(e).__name__}: {e}); eager step", file=sys.stderr)
            graphed = None
    for _ in range(3):
        tstep(False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    l0 = ops.LAUNCHES
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        tstep(False)
        ev[k][1].record()
    barrier()
    launches = ops.LAUNCHES - l0 if graphed is None else graphed.launches_per_step * steps
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tstep(True)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        loss = tstep(True)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    tms = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    g = len(meshes) * world
    fresh_ms, fsteps = None, 0
    if fresh:
        # ---------------- the reference's DataLoader regime (train.py:119-160): every batch is 16 random (mesh, time step) samples, so the
        # batch TOPOLOGY changes every step and one captured graph cannot be replayed: collate + CSR tables + the eager step, per step
        import random

        kind, kw, _, _ = WORKLOADS[args.workload]
        pool = list(lists)
        for i in range(len(meshes)):
            m = fv.synthetic.make(kind, seed=77000 + 1000 * rank + i, T=3, **kw)
            t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
            tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
            pool.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device=dev))
        rng = random.Random(rank)
        opt_e = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)

        def fresh_step():
            bl = [[_shallow(b) for b in pool[i]] for i in rng.sample(range(len(pool)), len(meshes))]
            gn, ge, gc = fv.dataset.collate(bl)
            gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc])
            opt_e.zero_grad(set_to_none=True)
            out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0)
            lo = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                     target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
            lo[0].backward()
            dp_sync()
            opt_e.step()

        fsteps = 2 * steps
        for _ in range(40):  # the batch sizes vary from step to step: let the caching allocator (two streams) see the range of block sizes first
            fresh_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(fsteps):
            fresh_step()
        t1.record()
        barrier()
        fresh = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(fresh, op=dist.ReduceOp.MAX)
        fresh_ms = float(fresh[0])
    return {"metric": "training_graphs_per_sec", "value": g * steps / (ms * 1e-3), "unit": "graphs/s", "ms_per_step": ms / steps, "steps": steps,
            "global_batch": g, "cells_per_rank": C, "faces_per_rank": F, "dtype": DTYPES[train_math] + (" (forward and backward GEMMs on tcgen05)" if train_math != "fp32" else " forward and backward"), "loss": p.loss, "optimizer": "AdamW lr 1e-3",
            "cell_updates_per_sec": C * world * steps / (ms * 1e-3),
            "e2e": {"value": g * steps / (e2e_ms * 1e-3), "unit": "graphs/s", "ms_per_step": e2e_ms / steps,
                    "h2d_bytes_per_step": int(node_host.numel() * 4), "d2h_bytes_per_step": 20,
                    "note": "pinned node fields H2D + loss D2H + host sync every step"},
            "fresh_batches": None if not fresh else {
                "value": g * fsteps / (fresh_ms * 1e-3), "unit": "graphs/s", "ms_per_step": fresh_ms / fsteps, "steps": fsteps, "cuda_graph": False,
                "note": f"every step a new random batch of {len(meshes)} meshes out of {2 * len(meshes)} (new topology: collate + CSR tables + "
                        "eager step inside the timed region)"},
            "gpu_launches": int(launches), "final_loss": float(loss[0]),
            "parallelism": (f"dp{world} over mesh graphs; " + (f"gradient all-reduce (8.84 MB, AVG) in {args.dp_groups} groups issued from inside the backward "
                            "(overlapped with it; parallel.OverlappedGradReducer)" if reducer is not None else "one flat fp32 gradient all-reduce (8.84 MB) per step"))
            if world > 1 else "single GPU",
            "step": "train_datapreprocessing + FVGN.forward(train) + compute_FVM_loss + backward + grad all-reduce + AdamW",
            "cuda_graph": graphed is not None}


# ------------------------------------------------------------------------------------------------ single-mesh domain decomposition
def domain_decomposition_leg(args, rank, world, dev, dist, steps=5):
    """ONE mesh (`--dd-workload`, default the 1M-cell mesh of configs[4]) partitioned over the N GPUs of the box; every
    rank runs the rollout step on its part (fvgn_b200.domain: two NCCL exchanges per GnBlock + one per step).  Checked against the
    single-GPU rollout of the same mesh (rank 0 runs it and broadcasts next_uv) and timed next to it."""
    import fvgn_b200 as fv
    from fvgn_b200 import domain as D

    kind, kw, _, _ = WORKLOADS[args.dd_workload]
    m = fv.synthetic.make(kind, seed=4242, T=3, **kw)
    t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
    tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
    torch.manual_seed(0)
    model = fv.FVGN(device="cpu", params=params()).to(dev).eval().set_math_mode("bf16x3")
    gn, ge, gc = fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], device=dev)
    gn, ge, gc, mi, mb = fv.dataset.valid_datapreprocessing(gn, ge, gc, 0)
    x0, ea0, bc = gc.x.clone(), gc.edge_attr.clone(), ge.y[:, 0, 0:2].clone()
    C, F = x0.shape[0], ea0.shape[0]
    # single-GPU rollout (every rank runs it: the comparison needs no broadcast and all ranks stay in step)
    state = fv.dataset.RolloutState(gc, ge)
    ref = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0)  # warm-up (tables, packs)
        gc.x, gc.edge_attr = x0.clone(), ea0.clone()
        state = fv.dataset.RolloutState(gc, ge)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0)
            ref.append(nuv.clone())
            state.advance_(gc, nuv)
        e1.record()
        torch.cuda.synchronize()
    single_ms = e0.elapsed_time(e1) / steps
    parts, _ = D.decompose({k: v for k, v in tab.items()}, world)
    part = parts[rank]
    run = D.PartRollout(part, D.KernelBackend(model), *D.initial_part_state(part, x0, ea0, bc), device=dev)
    ms = dict(D.KernelBackend.state(model), dt=DT, rho=RHO)
    # exchanges: rows stored straight into the neighbours' receive buffers over NVLink peer memory + sequence flags (csrc/halo.cu,
    # D.SymmPeerComm); `--dd-comm nccl`, or a box where the symmetric allocation is unavailable: NCCL point-to-point batches (D.DistComm)
    exch = "peer memory (fvgn_exchange_push / _wait_*: stores into the neighbour's buffer over NVLink + sequence flags; 2 launches per exchange)"
    ok = torch.ones(1, device=dev)
    comm = None
    if args.dd_comm == "peer":
        try:
            comm = D.SymmPeerComm(parts, run, dev)
        except Exception as e:  # pragma: no cover
            print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); NCCL point-to-point exchanges", file=sys.stderr)
            ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if args.dd_comm != "peer" or float(ok[0]) == 0:
        comm = D.DistComm()
        exch = "NCCL point-to-point batches (torch.distributed.batch_isend_irecv)"
    with torch.no_grad():
        D.rollout_steps([run], comm, ms, 1)  # warm-up step (NCCL connections, allocator)
        run = D.PartRollout(part, D.KernelBackend(model), *D.initial_part_state(part, x0, ea0, bc), device=dev)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        outs = D.rollout_steps([run], comm, ms, steps)[0]
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
    dd_ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(dd_ms, op=dist.ReduceOp.MAX)
    eager_ms = float(dd_ms[0])
    # the same step as ONE CUDA graph per rank (kernels + the NCCL point-to-point exchanges): D.GraphedPartRollout
    graph_note = None
    ok = torch.ones(1, device=dev)
    try:
        run_g = D.PartRollout(part, D.KernelBackend(model), *D.initial_part_state(part, x0, ea0, bc), device=dev)
        gr = D.GraphedPartRollout(run_g, comm, ms)
    except Exception as e:  # pragma: no cover
        graph_note = f"capture failed ({type(e).__name__}: {e}); eager steps reported"
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks replay, or none does (a replay with a missing peer would hang)
    if float(ok[0]) > 0:
        with torch.no_grad():
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            outs = []
            for _ in range(steps):
                uvp_g, nuv_g = gr.step()
                outs.append((None, nuv_g.clone()))
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
        dd_ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(dd_ms, op=dist.ReduceOp.MAX)
    elif graph_note is None:
        graph_note = "capture failed on another rank; eager steps reported"
    own = torch.as_tensor(part.own).to(dev)
    num = torch.stack([((outs[k][1].double() - ref[k][own].double()) ** 2).sum() for k in range(steps)])
    den = torch.stack([(ref[k][own].double() ** 2).sum() for k in range(steps)])
    both = torch.stack([num, den])
    dist.all_reduce(both)
    rel = torch.sqrt(both[0] / both[1])
    ghosts = torch.tensor([float(part.G), float(sum(ix.size for ix in part.node_share.values()))], dtype=torch.float64, device=dev)
    dist.all_reduce(ghosts, op=dist.ReduceOp.MAX)
    return {"workload": args.dd_workload, "cells": int(C), "faces": int(F), "parts": world, "cells_per_part": int(part.C_own), "math": "bf16x3 (the in-place split-shadow kernels of the single-GPU fast path)",
            "max_ghost_cells_per_part": int(ghosts[0]), "max_interface_nodes_per_part": int(ghosts[1]), "steps": steps,
            "ms_per_step": float(dd_ms[0]), "value": C / (float(dd_ms[0]) * 1e-3), "unit": "cell-updates/s", "scaling": "strong (one mesh over N GPUs)",
            "cuda_graph": graph_note is None, "eager_ms_per_step": eager_ms, "note": graph_note,
            "single_gpu_eager_ms_per_step": single_ms, "speedup_over_single_gpu_eager": single_ms / float(dd_ms[0]),
            "rel_l2_next_uv_vs_single_gpu_per_step": [float(v) for v in rel], "tolerance_per_step": 1e-5,
            "pass": bool(all(float(v) < 1e-5 * (k + 1) for k, v in enumerate(rel))),
            "exchanges_per_step": "2 per GnBlock (interface-node partial sums; ghost-cell x') + 1 (ghost next_uv)", "exchange": exch}


# ------------------------------------------------------------------------------------------------ data-parallel parity check
def dp_parity_check(rank, world, dev, dist, graphs_per_rank=4):
    """DP-N == single process (SURVEY §4 'distributed', §5): every rank trains one step on ITS `graphs_per_rank` meshes with the
    statistic / loss / gradient collectives of fvgn_b200.parallel over NCCL, and — redundantly, collectives disabled — one step on the
    WHOLE batch of N x graphs_per_rank meshes.  Strict fp32 kernels (deterministic), same weights, same noise / flip draws.
    Compared on every rank: loss, all 266 gradients, the four normalisers' buffers, the face-area BatchNorm running statistics, for
    both loss flavours the reference trains with (global_mean_sum = get_param.py default; direct_mean_loss = run_train.sh:17, whose
    single log over the batch means needs the partial sums reduced before the log)."""
    import fvgn_b200 as fv
    from fvgn_b200 import parallel

    n = graphs_per_rank
    lists, sizes = [], []
    for r in range(world):
        for i in range(n):
            m = fv.synthetic.make("cyl", seed=5000 + 1000 * r + i, T=3)
            t = lambda a, dt: torch.as_tensor(a).to(dt).to(dev)
            tab = fv.ops.build_connectivity(t(m["cells"], torch.int32), t(m["mesh_pos"], torch.float32), t(m["node_type"], torch.int32), "B")
            lists.append(fv.dataset.build_graphs(tab, m["velocity"], m["pressure"], training_pair=0, device=dev))
            sizes.append((int(tab["cells_node"].shape[0]), int(tab["face"].shape[1])))
    Ct, Ft = sum(c for c, _ in sizes), sum(f for _, f in sizes)
    gen = torch.Generator().manual_seed(123)
    noise = (torch.randn((Ct, 2), generator=gen) * 0.02).to(dev)
    flip = (torch.rand(Ft, generator=gen) < 0.5).to(dev)
    c0 = sum(c for c, _ in sizes[:rank * n]); c1 = c0 + sum(c for c, _ in sizes[rank * n:(rank + 1) * n])
    f0 = sum(f for _, f in sizes[:rank * n]); f1 = f0 + sum(f for _, f in sizes[rank * n:(rank + 1) * n])
    out = {"graphs_per_rank": n, "world": world, "cells_total": Ct, "losses": {}}

    def one_step(model, p, sel, cn, fk):
        gn, ge, gc = fv.dataset.collate([[_shallow(b) for b in lists[i]] for i in sel])
        gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing([gn, ge, gc], cell_noise=cn, flip_keep=fk)
        o = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0)
        lo = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=o[3],
                                 target_face_uvp_normalized=o[4], predicted_delta_u=o[1], target_delta_u=o[2])
        lo[0].backward()
        return lo[0].detach()

    def buffers(model):
        b = {f"{nm}.{k}": getattr(getattr(model, nm), k).detach().double().reshape(-1) for nm in parallel.NORMALIZERS
             for k in ("acc_sum", "acc_sum_squared", "acc_count", "num_accumulations")}
        bn = model._grid_feature_normalizer
        b["bn.running_mean"], b["bn.running_var"] = bn.running_mean.double(), bn.running_var.double()
        return b

    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))
    ok = True
    for loss_name, math in (("global_mean_sum", "fp32"), ("direct_mean_loss", "fp32"), ("global_mean_sum", "bf16x3")):
        p = params(loss=loss_name)
        torch.manual_seed(0)
        ref = fv.FVGN(device="cpu", params=p).to(dev).train().set_math_mode(math)
        with parallel.single_process():
            l_ref = one_step(ref, p, range(world * n), noise, flip)
        torch.manual_seed(0)
        dp = fv.FVGN(device="cpu", params=p).to(dev).train().set_math_mode(math)
        if math != "fp32":  # tensor-core path: the all-reduce runs in groups from inside the backward, .grad are views of the reduced buffer
            parallel.OverlappedGradReducer(dp, n_groups=4).attach(dp)
        l_dp = one_step(dp, p, range(rank * n, (rank + 1) * n), noise[c0:c1], flip[f0:f1])
        if math == "fp32":
            parallel.GradBucket(dp.parameters()).all_reduce_mean_()
        l_mean = l_dp.clone().reshape(1)
        dist.all_reduce(l_mean)
        l_mean /= world
        g_err = max(rel(a.grad, b.grad) for a, b in zip(dp.parameters(), ref.parameters()))
        b_dp, b_ref = buffers(dp), buffers(ref)
        b_err = max(rel(b_dp[k], b_ref[k]) for k in b_ref)
        l_err = abs(float(l_mean) - float(l_ref))
        errs = torch.tensor([g_err, b_err, l_err], dtype=torch.float64, device=dev)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        out["losses"][f"{loss_name}:{math}"] = {"gradient_all_reduce": "flat bucket after the backward" if math == "fp32" else "4 groups overlapped with the backward",
                                    "loss_single_process": float(l_ref), "loss_dp_mean_over_ranks": float(l_mean), "loss_abs_err": float(errs[2]),
                                    "worst_gradient_rel_l2": float(errs[0]), "worst_buffer_rel_l2": float(errs[1])}
        ok = ok and float(errs[0]) < 1e-5 and float(errs[1]) < 1e-6 and float(errs[2]) < 1e-5
    out["tolerance"] = {"gradients_rel_l2": 1e-5, "buffers_rel_l2": 1e-6, "loss_abs": 1e-5}
    out["pass"] = bool(ok)
    return out


# ------------------------------------------------------------------------------------------------ reference arms
# The reference's own implementation of the path (the UNMODIFIED FVGN-src/main modules: /root/reference in the build container, the
# byte-for-byte copy oracle/make_ref.py places under the git-ignored oracle/_ref/ on the GPU box) driven through its own Data_Pool /
# FVGN / compute_FVM_loss objects by oracle/ref_harness.py — on the host cores (`--impl reference`, `cpu_baseline`: kind "reference")
# and on the GPU (`torch_eager_cuda`: what a user of the reference runs on a B200 today).  Where the copy is missing the CPU arm falls
# back to the oracle port (oracle/fvgn_oracle.py, kind "port").  These legs are the only place bench.py executes anything under oracle/.
def _ref_harness(workload, n_graphs, device, loss="global_mean_sum"):
    import fvgn_b200 as fv
    from oracle import ref_harness as RH

    if not RH.available():
        return None
    kind, kw, _, _ = WORKLOADS[workload]
    meshes = [fv.synthetic.make(kind, seed=i, T=3, **kw) for i in range(n_graphs)]  # rank 0's meshes of the native arm
    return RH.RefHarness(meshes, params(), device=device, loss=loss)


def cpu_setup_port(workload, n_graphs):
    """fallback CPU arm (kind "port"): the reference's torch-CPU path restated (oracle/fvgn_oracle.py, pinned against the reference)"""
    import fvgn_b200 as fv
    from oracle import fvgn_oracle as O
    from tests import helpers as H

    kind, kw, _, _ = WORKLOADS[workload]
    bags = []
    for i in range(n_graphs):
        m = fv.synthetic.make(kind, seed=i, T=3, **kw)
        tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
        bags.append(O.build_graphs(tab, m["velocity"], m["pressure"]))
    gn, ge, gc = [O.batch_graphs([b[k] for b in bags]) for k in range(3)]
    gn, ge, gc, mi, mb = O.valid_datapreprocessing(gn, ge, gc, 0)
    orc = O.FVGNOracle(H.torch_reference_state_dict(0), params()).eval()
    ct, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()

    def step(gc=gc):
        with torch.no_grad():
            uvp, nuv = orc(gc, ge, gn, RHO, MU, DT, 9, 0)
            O.create_next_graph(gc, ge, mb, nuv, ct, Re, rmp)
        return nuv

    return step, gc.x.shape[0], ge.x.shape[0], gn.x.shape[0]


def _cpu_arm(workload, n_graphs):
    """(kind, rollout_step, train_step or None, C, F, N, description)"""
    h = _ref_harness(workload, n_graphs, "cpu")
    if h is not None:
        return ("reference", h.rollout_step, h.train_step, h.C, h.F, int(h.gn.x.shape[0]),
                "the UNMODIFIED reference (FVGN-src/main via oracle/_ref + shims for torch_geometric / torch_scatter)")
    step, C, F, N = cpu_setup_port(workload, n_graphs)
    return ("port", step, None, C, F, N, "oracle/fvgn_oracle.py (restatement of the reference's torch path, pinned against it)")


def cpu_baseline(args, budget_s=15.0):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_graphs = WORKLOADS[args.workload][2]  # the whole batch of one rank: same workload as the GPU arm, bounded in the number of steps
    kind, step, tstep, C, F, N, what = _cpu_arm(args.workload, n_graphs)
    step()
    t0 = time.perf_counter()
    step()
    one = time.perf_counter() - t0
    n = int(max(2, min(50, budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = time.perf_counter() - t0
    out = {"value": C * n / dt, "unit": "cell-updates/s", "cores": torch.get_num_threads(), "kind": kind,
           "sample": f"{n} rollout steps of {n_graphs} {WORKLOADS[args.workload][0]} mesh(es) ({C} cells) on the host CPU, torch {torch.__version__}, "
                     f"{threads} threads; {what}",
           "ms_per_step": dt / n * 1e3}
    if tstep is not None:  # the batch-16 training step of train.py:150-225 on the host cores (the metric's "training graphs/sec vs CPU torch")
        tstep()
        t0 = time.perf_counter()
        tstep()
        one = time.perf_counter() - t0
        nt = int(max(2, min(20, budget_s / max(one, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(nt):
            tstep()
        dtt = time.perf_counter() - t0
        out["training"] = {"value": n_graphs * nt / dtt, "unit": "graphs/s", "ms_per_step": dtt / nt * 1e3, "cores": torch.get_num_threads(),
                           "kind": kind, "sample": f"{nt} training steps (train_datapreprocessing + FVGN.forward + compute_FVM_loss + backward + AdamW, "
                                                   f"train.py:150-225) on one batch of {n_graphs} meshes ({C} cells), loss global_mean_sum"}
    return out


def torch_eager_cuda(args, dev, steps):
    """The reference's own PyTorch path on the B200 ("the bar is torch-eager on B200"): the unmodified reference modules
    on `dev`, same 16 meshes / same initial weights, rollout step and batch-16 training step, in strict fp32 matmuls and with TF32 allowed
    (what the authors' torch 1.9 did by default).  CUDA-event timed, L2 flushed between steps; outside the native timed regions."""
    try:
        h = _ref_harness(args.workload, WORKLOADS[args.workload][2], dev)
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"{type(e).__name__}: {e}"}
    if h is None:
        return {"unavailable": "oracle/_ref (copy of the reference made by oracle/make_ref.py in the build container) is not present"}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    out = {"impl": "UNMODIFIED reference modules (oracle/_ref) through their own Data_Pool / FVGN / compute_FVM_loss API on cuda; torch_scatter / "
                   "torch_geometric replaced by index_add_ shims (oracle/shims); eager, one CUDA stream",
           "torch": torch.__version__, "cells": h.C, "faces": h.F, "graphs": h.n_graphs, "rollout": {}, "training": {}}

    def timed(fn, n):
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        torch.cuda.synchronize()
        for k in range(n):
            flush.zero_()
            ev[k][0].record()
            fn()
            ev[k][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / n

    try:
        for leg, fn, per, unit in (("rollout", h.rollout_step, h.C, "cell-updates/s"), ("training", h.train_step, h.n_graphs, "graphs/s")):
            for name, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                if leg == "rollout":
                    h.reset_rollout()
                n = steps if leg == "rollout" else max(2, min(steps, args.train_steps))
                ms = timed(fn, n)
                out[leg][name] = {"value": per / (ms * 1e-3), "unit": unit, "ms_per_step": ms, "steps": n}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    del h, flush
    torch.cuda.empty_cache()
    return out


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, all host threads, on the native arm's config."""
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_graphs = WORKLOADS[args.workload][2]  # the whole batch of one rank (the GPU arm's config), bounded in the number of steps
    kind, step, _, C, F, N, what = _cpu_arm(args.workload, n_graphs)
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = C * args.steps / dt
    sample = (f"{args.steps} rollout steps of {n_graphs} {WORKLOADS[args.workload][0]} mesh(es) ({C} cells): bounded sample of workload {args.workload}; "
              f"{what}; torch {torch.__version__}, {threads} threads")
    return {"impl": "reference", "metric": "cell_updates_per_sec", "value": v, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(args.workload, C, F, N, n_graphs, world),
            "cpu_baseline": {"value": v, "unit": "cell-updates/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cyl16", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="bf16x3", choices=["fp32", "bf16", "bf16x3"],
                    help="bf16x3 (default): fp32-equivalent tensor-core mode; fp32: strict FFMA; bf16: bf16 tiles")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one flat gradient all-reduce after the backward instead of the overlapped groups")
    ap.add_argument("--dp-groups", type=int, default=4, help="N > 1: number of gradient all-reduce groups issued from inside the backward")
    ap.add_argument("--check", action="store_true", help="N > 1: run the data-parallel parity check (DP-N step == single-process batch; always on at N = 2) and exit 3 if it fails")
    ap.add_argument("--no-torch-eager", action="store_true", help="skip the torch-eager CUDA leg (the unmodified reference on the GPU)")
    ap.add_argument("--no-training", action="store_true", help="skip the training graphs/s leg")
    ap.add_argument("--no-strict-leg", action="store_true", help="skip the strict-FFMA comparison leg")
    ap.add_argument("--no-dd-leg", action="store_true", help="N > 1: skip the single-mesh domain-decomposition leg")
    ap.add_argument("--dd-comm", default="peer", choices=["peer", "nccl"], help="N > 1: exchanges of the domain-decomposition leg over peer memory (csrc/halo.cu) or NCCL point-to-point")
    ap.add_argument("--dd-workload", default="mesh1m", choices=sorted(WORKLOADS), help="the ONE mesh the domain-decomposition leg partitions over the N GPUs")
    ap.add_argument("--no-foil-leg", action="store_true", help="skip the foil16 / bf16 leg (BASELINE configs[2])")
    ap.add_argument("--no-config0", action="store_true", help="skip the batch-1 rollout latency leg (configs[0])")
    ap.add_argument("--train-graph", type=int, default=1, help="1: time the training step replayed from one CUDA graph; 0: eager")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--no-renumbered-leg", action="store_true", help="skip the sweep mesh after fvgn_b200.renumber_mesh")
    ap.add_argument("--sweep", default="mesh1m:bf16", help="<workload>:<math> of the kernel roofline sweep leg, or 'none'")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        res = reference_arm(args, rank, world)
    else:
        res = native(args, rank, world, local_rank)
    if rank == 0 and res is not None:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
