#!/bin/bash
# Lean single-GPU call: parity tests, variant/tile tuning in one process per library build, bench, ncu evidence.
mkdir -p gpurun_out
T0=$(date +%s)
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
B200_CG_VERBOSE=1 timeout 300 python -m pytest tests/test_gpu_dirac.py -m gpu -q -k mixed 2>&1 | grep "\[cg\]" | head -40
for v in "" _r255 _r170 _r102 _r85 _l2; do
  echo "== tune lib$v"; B200_LIB=$PWD/quda_b200/libquda_b200$v.so timeout 400 python tools/tune.py "lib$v" 2>&1 | grep tune
done
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_single_r12.json
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dslash --csv \
   --log-file gpurun_out/launches_single_r12.csv python tools/prof_target.py single 12 8 > gpurun_out/ncu_list.log 2>&1
echo "== ncu full (single r12 with source, double r18, half r8)"
for cfg in "single 12" "double 18" "half 8"; do
  set -- $cfg
  SRC=""; [ "$1" = "single" ] && SRC="--import-source on"
  timeout 400 ncu --set full --clock-control none $SRC -k regex:dslash_interior -s 3 -c 1 \
     -o gpurun_out/prof_$1_r$2 -f python tools/prof_target.py $1 $2 5 > gpurun_out/ncu_full_$1_r$2.log 2>&1
  tail -1 gpurun_out/ncu_full_$1_r$2.log
  ncu -i gpurun_out/prof_$1_r$2.ncu-rep --page raw --csv > gpurun_out/prof_$1_r$2.raw.csv 2>/dev/null
  [ "$1" = "single" ] && ncu -i gpurun_out/prof_$1_r$2.ncu-rep --page source --csv > gpurun_out/prof_$1_r$2.source.csv 2>/dev/null
  ls -la gpurun_out/prof_$1_r$2.ncu-rep
  # the merge-back limit is 64 MiB in total: keep the report only if it is small
  [ $(stat -c %s gpurun_out/prof_$1_r$2.ncu-rep) -gt 15000000 ] && rm -f gpurun_out/prof_$1_r$2.ncu-rep
done
du -sh gpurun_out
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== done"
