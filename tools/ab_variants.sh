#!/bin/bash
# A/B measurement of experimental kernel builds (finite-volume-graph-network_b200/build.py --variant NAME -D...): for every
# libfvgn_b200_<NAME>.so given (and the product build), the x3 parity tests and the rollout legs of bench.py.
#   tools/ab_variants.sh base pipe rcp2 both      -> gpurun_out/ab_<NAME>.json / ab_<NAME>_tests.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset FVGN_B200_LIB; else export FVGN_B200_LIB=$PWD/finite-volume-graph-network_b200/libfvgn_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_x3.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/ab_${v}_tests.log
  timeout 300 python bench.py --steps 30 --warmup 5 --no-training --no-cpu-baseline --no-torch-eager --no-strict-leg --no-config0 \
      --sweep mesh1m:bf16x3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{v}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{v:8s} cyl16 {d['ms_per_step']:.3f} ms/step  {d['value']/1e6:.2f} M/s | edge {r['avg_launch_ms']*1e3:.1f} us shares {r['share_of_step']} | "
          f"mesh1m x3 {d['sweep']['ms_per_step']:.2f} ms {d['sweep']['value']/1e6:.2f} M/s edge {d['sweep']['roofline']['avg_launch_ms']*1e3:.0f} us | "
          + open(f"gpurun_out/ab_{v}_tests.log").read().strip().splitlines()[-1])
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/ab_{v}.err").read()[-500:])
PY
done
