#!/bin/bash
# Timing-experiment sweep of the tensor-core kernels (library built with FVGN_NVCC_EXTRA=-DFVGN_TC_TIMING_EXPERIMENTS=1).
# usage: tools/dbg_sweep.sh <workload> <math> <flag> [<flag> ...]   -> gpurun_out/dbgx_<flag>.log ; results are garbage by design
w=$1; m=$2; shift 2
mkdir -p gpurun_out
for f in "$@"; do
  FVGN_TC_DBG=$f timeout 120 python bench.py --workload $w --math $m --no-training --no-cpu-baseline --sweep none --steps 5 --warmup 3 > gpurun_out/dbgx_$f.log 2>&1
  python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f"gpurun_out/dbgx_{f}.log") if x.startswith("{")][-1]
    d = json.loads(l); r = d["roofline"]; sh = r["share_of_step"]; ms = d["ms_per_step"]
    print(f"dbg={f:>4} step {ms:.3f} ms  edge {r['avg_launch_ms']*1e3:.1f} us  cell {sh['cell_block']*ms/15*1e3:.1f} us  node {sh['node_aggregate']*ms/15*1e3:.1f} us")
except Exception as e:
    print(f"dbg={f} failed: {e}")
PY
done
