#!/bin/bash
# Round-2 experiment pack, FIRST call of round 2 (kept as the record of how profiles/r02_march_* were produced; the
# B200_MARCH_T launch it sweeps was removed afterwards because it did not help -- DESIGN.md section 6.2).
# Round-2 experiment pack (single GPU, ~3 min): everything that was written in round 1 after the GPU budget ran out.
#   1. experimental GPU tests (time-marching launch)                      B200_EXPERIMENTAL=1
#   2. time-marching sweep: does walking t slices with a fixed (x,y,z) tile cut the L2->L1 spinor traffic?
#   3. ncu of the best marching shape (xbar bytes = l1tex__m_xbar2l1tex_read_bytes.sum, the quantity to drive down)
# Multi-GPU items (run with gpurun --gpus 2/8): tools/r2_multi.sh
mkdir -p gpurun_out
T0=$(date +%s)
B200_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_wilson.py -m gpu -q -k marching 2>&1 | tail -4 | tee gpurun_out/r2_pytest_marching.txt
: > gpurun_out/r2_march_sweep.jsonl
run() { # march prec recon tile...
  local m=$1 prec=$2 recon=$3; shift 3
  B200_MARCH_T=$m timeout 100 python bench.py --prec $prec --recon $recon --steps 100 --no-cpu-baseline --no-e2e --no-mrhs --tile $1 $2 $3 $4 2>> gpurun_out/bench_err.txt \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'march':$m,'prec':'$prec','recon':$recon,'tile':'$*','us':d['ms_per_step']*1e3,'frac':d['roofline']['frac'],'sustained_us':d['sustained']['ms_per_step']*1e3}))" >> gpurun_out/r2_march_sweep.jsonl
}
for m in 0 4 8 16 32; do
  run $m single 12 16 2 2 1
  run $m single 12 16 4 2 1
  run $m single 12 16 4 1 1
done
run 8 double 12 16 2 1 1
run 8 double 12 16 2 2 1
run 8 half 12 16 8 1 1
# half precision: conversion-split variants (build them first with tools/r2_build_variants.sh)
for v in "" _i2f3 _i2f2 _nodef; do
  lib=quda_b200/libquda_b200$v.so
  [ -f $lib ] || continue
  for rc in 12 8; do
    B200_LIB=$lib timeout 100 python bench.py --prec half --recon $rc --steps 100 --no-cpu-baseline --no-e2e --no-mrhs 2>> gpurun_out/bench_err.txt \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$lib','prec':'half','recon':$rc,'us':d['ms_per_step']*1e3,'sustained_us':d['sustained']['ms_per_step']*1e3}))" >> gpurun_out/r2_march_sweep.jsonl
  done
  B200_LIB=$lib timeout 100 python -m pytest tests/test_gpu_wilson.py -m gpu -q -k "half or prec2 or 2-" 2>&1 | tail -1
done
cat gpurun_out/r2_march_sweep.jsonl
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 150 env B200_MARCH_T=8 PROF_TILE="16 4 2 1" ncu --set full --clock-control none -k regex:dslash_march -s 2 -c 1 \
   -o gpurun_out/r2_march -f python tools/prof_target.py single 12 4 > gpurun_out/ncu_r2_march.log 2>&1
ncu -i gpurun_out/r2_march.ncu-rep --page raw --csv > gpurun_out/r2_march_single_r12.raw.csv 2>/dev/null
rm -f gpurun_out/r2_march.ncu-rep
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
