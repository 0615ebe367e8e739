#!/bin/bash
# 2-GPU A/B: ghost-load policy (ld.relaxed.sys vs weak ld.global, both L1::no_allocate) x schedule
N=2
mkdir -p gpurun_out
T0=$(date +%s)
run() { # label [bench args...]
  local label=$1; shift
  env "${ENVV[@]}" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) \
     bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-e2e "$@" 2>> gpurun_out/bench_err.txt | tail -1 > gpurun_out/r2_scale_${N}_${label}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${N}_${label}.json"))
    print("${label}", {k: d.get(k) for k in ("value","ms_per_step","n_gpus")}, d.get("halo",{}).get("parity_dev"), d.get("breakdown_us"))
except Exception as e:
    print("${label} FAILED", e)
PY
  echo "[t=$(( $(date +%s)-T0 ))s]"
}
ENVV=(A=1); run xsplit_auto --grid 2 1 1 1
echo "== done"
