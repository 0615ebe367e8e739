#!/bin/bash
# 1-GPU study of the partitioned schedules in "self" mode + the TMA variant with the rolled producer loop
mkdir -p gpurun_out
T0=$(date +%s)
: > gpurun_out/r2_fused_self.jsonl
for m in "0 0 0 1" "0 1 1 1" "1 0 0 0"; do
  for sch in fused streams; do
    B200_HALO_SCHEDULE=$sch timeout 120 python tools/fused_self.py "$m" 200 2>> gpurun_out/bench_err.txt | tail -1 | tee -a gpurun_out/r2_fused_self.jsonl
  done
done
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dslash_fused -s 3 -c 1 -o gpurun_out/r2_fused -f python tools/fused_self.py "0 1 1 1" 6 > gpurun_out/ncu_r2_fused.log 2>&1
ncu -i gpurun_out/r2_fused.ncu-rep --page raw --csv > gpurun_out/r2_fused_single_r12.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fused.ncu-rep --page source --csv > gpurun_out/r2_fused_single_r12.source.csv 2>/dev/null
rm -f gpurun_out/r2_fused.ncu-rep
echo "[t=$(( $(date +%s)-T0 ))s]"
B200_TMA=1 timeout 120 python bench.py --steps 200 --no-cpu-baseline --no-e2e --no-mrhs --no-extra 2>> gpurun_out/bench_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma rolled producer', d['ms_per_step']*1e3, d['sustained']['ms_per_step']*1e3)"
timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-e2e --no-mrhs 2>> gpurun_out/bench_err.txt | tee gpurun_out/r2_bench_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', d['ms_per_step']*1e3, [(o.get('prec'),o.get('recon'),o.get('op'),o.get('ms_per_step'),o.get('frac'),o.get('error')) for o in d.get('other_configs',[])])"
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
