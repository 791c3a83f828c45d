#!/bin/bash
# quick A/B of the rollout legs: tools/ab_quick.sh "<ENV=VAL ...>" ...   (one bench run per argument; x3 parity tests first)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python -m pytest tests/test_gpu_x3.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2 > gpurun_out/abq_${i}_tests.log
  env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-training --no-cpu-baseline --no-torch-eager --no-strict-leg --no-config0 --no-foil-leg \
      --sweep mesh1m:bf16x3 > gpurun_out/abq_$i.json 2> gpurun_out/abq_$i.err
  python - "$i" "$envs" <<'PY'
import json, sys
i, envs = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/abq_{i}.json").read().strip().splitlines()[-1])
    r = d["roofline"]; sh = r["share_of_step"]; ms = d["ms_per_step"]
    print(f"[{envs}] cyl16 {ms:.3f} ms/step {d['value']/1e6:.2f} M/s e2e {d['e2e']['value']/1e6:.2f} | edge {r['avg_launch_ms']*1e3:.1f} us cell {sh['cell_block']*ms/15*1e3:.1f} us | "
          f"mesh1m x3 {d['sweep']['ms_per_step']:.2f} ms {d['sweep']['value']/1e6:.2f} M/s edge {d['sweep']['roofline']['avg_launch_ms']*1e3:.0f} us | "
          + open(f"gpurun_out/abq_{i}_tests.log").read().strip().splitlines()[-1])
except Exception as e:
    print(envs, "FAILED", e, open(f"gpurun_out/abq_{i}.err").read()[-800:], open(f"gpurun_out/abq_{i}_tests.log").read()[-600:])
PY
done
