#!/bin/bash
# A/B of the CTA-pair inference kernels (opt-in: FVGN_X3_PAIR=1) against the single-CTA kernels.
#   tools/ab_pair.sh base pair [VARIANT ...]   base = product library, single-CTA kernels; pair = product library, pair kernels;
#   VARIANT = libfvgn_b200_<VARIANT>.so (build.py --variant VARIANT -D...), pair kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  unset FVGN_B200_LIB
  export FVGN_X3_PAIR=1
  if [ "$v" = base ]; then export FVGN_X3_PAIR=0; elif [ "$v" != pair ]; then export FVGN_B200_LIB=$PWD/finite-volume-graph-network_b200/libfvgn_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_x3.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pair_${v}_tests.log
  timeout 300 python bench.py --steps 30 --warmup 5 --no-training --no-cpu-baseline --no-torch-eager --no-strict-leg --no-config0 --no-foil-leg \
      --sweep mesh1m:bf16x3 > gpurun_out/pair_$v.json 2> gpurun_out/pair_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/pair_{v}.json").read().strip().splitlines()[-1])
    r = d["roofline"]; sh = r["share_of_step"]; ms = d["ms_per_step"]
    print(f"{v:8s} cyl16 {ms:.3f} ms/step {d['value']/1e6:.2f} M/s e2e {d['e2e']['value']/1e6:.2f} | edge {r['avg_launch_ms']*1e3:.1f} us cell {sh['cell_block']*ms/15*1e3:.1f} us | "
          f"mesh1m x3 {d['sweep']['ms_per_step']:.2f} ms {d['sweep']['value']/1e6:.2f} M/s edge {d['sweep']['roofline']['avg_launch_ms']*1e3:.0f} us | "
          + open(f"gpurun_out/pair_{v}_tests.log").read().strip().splitlines()[-1])
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/pair_{v}.err").read()[-800:], open(f"gpurun_out/pair_{v}_tests.log").read()[-600:])
PY
done
