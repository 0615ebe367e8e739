#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
for v in "" _i2f1 _i2f2; do
  echo "== tune lib$v"; B200_LIB=$PWD/quda_b200/libquda_b200$v.so timeout 300 python tools/tune.py "lib$v" half:12,half:8,half:18 2>&1 | grep tune
done
echo "== clover-PC Dslash (config 3 and friends)"
for cfg in "half 8" "half 12" "single 12" "double 18"; do set -- $cfg; timeout 200 python bench.py --op clover_pc --prec $1 --recon $2 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_cloverpc_$1_r$2.json | cut -c1-330; done
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 300 python -m pytest tests/test_gpu_dirac.py tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -3
echo "[t=$(( $(date +%s)-T0 ))s]"
