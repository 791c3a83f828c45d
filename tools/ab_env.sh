#!/bin/bash
# A/B of run-time switches: tools/ab_env.sh [--train] LABEL:VAR=VAL[,VAR=VAL...] ...   (e.g. on:FVGN_PDL=1 off:FVGN_PDL=0)
# per label: the x3 / rollout parity tests and the rollout (+ training with --train) legs of bench.py -> gpurun_out/env_<LABEL>.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TRAIN="--no-training"
if [ "$1" = "--train" ]; then TRAIN=""; shift; fi
for spec in "$@"; do
  label=${spec%%:*}
  envs=${spec#*:}
  (
    IFS=',' read -ra kv <<< "$envs"
    for e in "${kv[@]}"; do export "$e"; done
    timeout 400 python -m pytest tests/test_gpu_x3.py tests/test_gpu_round2.py tests/test_gpu_rollout600.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/env_${label}_tests.log
    timeout 400 python bench.py --steps 30 --warmup 5 $TRAIN --no-cpu-baseline --no-torch-eager --no-strict-leg --no-config0 --no-foil-leg \
        --sweep mesh1m:bf16x3 > gpurun_out/env_$label.json 2> gpurun_out/env_$label.err
  )
  python - "$label" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/env_{v}.json").read().strip().splitlines()[-1])
    r = d["roofline"]; ms = d["ms_per_step"]
    tr = d.get("training") or {}
    print(f"{v:8s} cyl16 {ms:.3f} ms/step {d['value']/1e6:.2f} M/s e2e {d['e2e']['value']/1e6:.2f} | edge {r['avg_launch_ms']*1e3:.1f} us | "
          f"mesh1m x3 {d['sweep']['ms_per_step']:.2f} ms | train {tr.get('ms_per_step')} ms fresh {(tr.get('fresh_batches') or {}).get('ms_per_step')} | "
          + open(f"gpurun_out/env_{v}_tests.log").read().strip().splitlines()[-1])
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/env_{v}.err").read()[-800:], open(f"gpurun_out/env_{v}_tests.log").read()[-600:])
PY
done
