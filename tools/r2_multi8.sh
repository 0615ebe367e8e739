#!/bin/bash
# Round-2 8-GPU pack (gpurun --gpus 8; charged 8x, so keep it lean): parity at 8 ranks (fused schedule, 8^4 local),
# weak-scaling bench with its in-line global-oracle check, the x-including grid, fp64 / half points, and BASELINE config 5
# (48^3 x 96 clover CG, mixed precision) through bench.py --op cg.
N=8
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "eight" 2>&1 | tail -4 | tee gpurun_out/r2_pytest_multi_8gpu.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
run() { # label [bench args...]
  local label=$1; shift
  env "${ENVV[@]}" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) \
     bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline "$@" 2>> gpurun_out/bench_err.txt | tail -1 > gpurun_out/r2_scale_${N}_${label}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${N}_${label}.json"))
    print("${label}", {k: d.get(k) for k in ("value","ms_per_step","n_gpus")}, d.get("halo",{}).get("parity_dev"), d.get("breakdown_us"), (d.get("e2e") or {}).get("ms_per_step"), d.get("cg"))
except Exception as e:
    print("${label} FAILED", e)
PY
  echo "[t=$(( $(date +%s)-T0 ))s]"
}
ENVV=(A=1); run fused
ENVV=(B200_HALO_SCHEDULE=streams); run streams --no-e2e
ENVV=(A=1); run cg --op cg --steps 1 --warmup 1
ENVV=(A=1); run xgrid --grid 2 2 2 1 --no-e2e
ENVV=(A=1); run double --prec double --recon 18 --no-e2e
ENVV=(A=1); run half --prec half --recon 12 --no-e2e
echo "== done"
