#!/bin/bash
# Round-2 final single-GPU validation: full GPU test tier, smoke(), the default bench line (+ reference arm), ncu launch list and
# one --set full capture of the dominant kernel with rotating buffers.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_final_pytest_gpu.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/r2_final_smoke.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 400 python bench.py 2> gpurun_out/bench_final_err.txt | tee gpurun_out/r2_final_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_final_err.txt | tee gpurun_out/r2_final_reference_arm.json | cut -c1-300
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 8 --csv --log-file gpurun_out/r2_final_launches_single_r12.csv python tools/prof_target.py single 12 12 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dslash_interior -s 5 -c 1 -o gpurun_out/r2_final_prof -f python tools/prof_target.py single 12 8 > gpurun_out/ncu_r2_final.log 2>&1
ncu -i gpurun_out/r2_final_prof.ncu-rep --page raw --csv > gpurun_out/r2_final_single_r12.raw.csv 2>/dev/null
rm -f gpurun_out/r2_final_prof.ncu-rep
for cfg in "half 12" "half 8"; do set -- $cfg
  timeout 200 ncu --set full --clock-control none -k regex:dslash_interior -s 5 -c 1 -o gpurun_out/r2_final_prof_$1_$2 -f python tools/prof_target.py $1 $2 8 > /dev/null 2>&1
  ncu -i gpurun_out/r2_final_prof_$1_$2.ncu-rep --page raw --csv > gpurun_out/r2_final_$1_r$2.raw.csv 2>/dev/null
  rm -f gpurun_out/r2_final_prof_$1_$2.ncu-rep
done
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
