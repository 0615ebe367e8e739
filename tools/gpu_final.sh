#!/bin/bash
# Final single-GPU check of the round: full GPU parity tier, smoke, default bench line, multi-RHS sweep, ncu evidence.
mkdir -p gpurun_out
T0=$(date +%s)
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_final.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_final.txt
echo "== bench"; timeout 300 python bench.py 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_final.json | cut -c1-400
echo "[t=$(( $(date +%s)-T0 ))s]"
: > gpurun_out/mrhs_sweep.jsonl
run() { # prec recon nsrc batch [tile]
  B200_MRHS_BATCH=$4 timeout 120 python bench.py --prec $1 --recon $2 --nsrc $3 --steps 100 --no-cpu-baseline --no-e2e ${5:+--tile $5 $6 $7 $8} 2>> gpurun_out/bench_err.txt \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'prec':'$1','recon':$2,'nsrc':$3,'batch':$4,'tile':'${5:-def} ${6:-} ${7:-} ${8:-}','ms':d['ms_per_step'],'us_per_rhs':d['ms_per_rhs']*1e3,'gflops':d['value'],'frac':d['roofline']['frac'],'sustained_ms':d['sustained']['ms_per_step']}))" >> gpurun_out/mrhs_sweep.jsonl
}
run single 12 8 4
run single 12 8 2
run single 12 8 4 16 2 1 1
run single 12 16 4
run single 18 8 4
run single 8 8 4
run half 12 8 4
run half 8 8 4
run double 18 8 2
run double 18 8 2 16 2 2 1
cat gpurun_out/mrhs_sweep.jsonl
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dslash --csv \
   --log-file gpurun_out/launches_final.csv python tools/prof_target.py single 12 8 > gpurun_out/ncu_list.log 2>&1
for cfg in "single 12 1 dslash_interior" "half 12 1 dslash_interior" "single 12 8 dslash_mrhs"; do
  set -- $cfg
  timeout 200 ncu --set full --clock-control none -k regex:$4 -s 3 -c 1 \
     -o gpurun_out/final_$1_r$2_n$3 -f python tools/prof_target.py $1 $2 5 $3 > gpurun_out/ncu_full_$1_r$2_n$3.log 2>&1
  ncu -i gpurun_out/final_$1_r$2_n$3.ncu-rep --page raw --csv > gpurun_out/final_$1_r$2_n$3.raw.csv 2>/dev/null
  rm -f gpurun_out/final_$1_r$2_n$3.ncu-rep
done
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
