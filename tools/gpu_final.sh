#!/bin/bash
# Final single-GPU check of the round: full GPU parity tier, smoke, default bench line, ncu evidence of the final kernels.
mkdir -p gpurun_out
T0=$(date +%s)
echo "== pytest -m gpu"; timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_final.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_final.txt
echo "== bench"; timeout 300 python bench.py 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_final.json | cut -c1-600
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dslash --csv \
   --log-file gpurun_out/launches_final.csv python tools/prof_target.py single 12 8 > gpurun_out/ncu_list.log 2>&1
for cfg in "single 12" "half 12"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none -k regex:dslash_interior -s 3 -c 1 \
     -o gpurun_out/final_$1_r$2 -f python tools/prof_target.py $1 $2 5 > gpurun_out/ncu_full_$1_r$2.log 2>&1
  ncu -i gpurun_out/final_$1_r$2.ncu-rep --page raw --csv > gpurun_out/final_$1_r$2.raw.csv 2>/dev/null
  rm -f gpurun_out/final_$1_r$2.ncu-rep
done
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
