#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, tile sweeps, ncu launch list + full capture of the top kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [quick]'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/smoke.txt
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 5 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_single_r12.json
if [ "$1" != "quick" ]; then
  for cfg in "single 12" "double 18" "half 8" "single 18" "single 8" "double 12" "half 12"; do
    set -- $cfg
    echo "== sweep $1 r$2"; timeout 400 python bench.py --sweep --prec $1 --recon $2 2>&1 | tail -10
  done
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dslash -c 40 --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "== ncu full"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dslash_interior -s 5 -c 2 \
     -o gpurun_out/prof_single_r12 -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
echo "== done"
