#!/bin/bash
# lean 2-GPU check of the new arrival protocol (charged 2x): self-mode sanity on GPU 0, then the N=2 bench (fused, streams)
N=2
mkdir -p gpurun_out
T0=$(date +%s)
timeout 100 python tools/fused_self.py "0 0 0 1" 200 2>> gpurun_out/bench_err.txt | tail -1
timeout 120 python -m pytest tests/test_gpu_dirac.py -m gpu -x -q -k "partitioned" 2>&1 | tail -2
echo "[t=$(( $(date +%s)-T0 ))s]"
run() { # label [bench args...]
  local label=$1; shift
  env "${ENVV[@]}" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) \
     bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline "$@" 2>> gpurun_out/bench_err.txt | tail -1 > gpurun_out/r2_scale_${N}_${label}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${N}_${label}.json"))
    print("${label}", {k: d.get(k) for k in ("value","ms_per_step","n_gpus")}, d.get("halo",{}).get("parity_dev"), d.get("breakdown_us"), (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("${label} FAILED", e)
PY
  echo "[t=$(( $(date +%s)-T0 ))s]"
}
ENVV=(A=1); run fused
ENVV=(B200_HALO_SCHEDULE=streams); run streams --no-e2e
ENVV=(A=1); run xsplit --grid 2 1 1 1 --no-e2e
echo "== done"
