#!/bin/bash
# Round-2 multi-GPU pack (run with gpurun --gpus N): parity of the fused single-launch schedule against the global oracle,
# then the weak-scaling bench at N with its in-line parity check, x-split grid, and the round-1 two-stream schedule for comparison.
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_multi_${N}gpu.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
run() { # label [bench args / env...]
  local label=$1; shift
  env "${ENVV[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline "$@" 2>> gpurun_out/bench_err.txt | tail -1 > gpurun_out/r2_scale_${N}_${label}.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_scale_${N}_${label}.json"))
print("${label}", {k: d.get(k) for k in ("value","ms_per_step","n_gpus")}, d.get("halo",{}).get("parity_dev"), d.get("breakdown_us"), (d.get("e2e") or {}).get("ms_per_step"))
PY
}
ENVV=(A=1); run fused
ENVV=(B200_HALO_SCHEDULE=streams); run streams --no-e2e
if [ "$N" == "2" ]; then ENVV=(A=1); run xsplit --grid 2 1 1 1 --no-e2e; fi
if [ "$N" == "8" ]; then ENVV=(A=1); run xsplit --grid 2 2 2 1 --no-e2e; fi
ENVV=(A=1); run double --prec double --recon 18 --no-e2e
ENVV=(A=1); run half --prec half --recon 12 --no-e2e
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
