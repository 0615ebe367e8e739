#!/bin/bash
# Multi-RHS experiment: CTA flavour (sources in threadIdx.y, links shared through L1) vs thread flavour (links in registers)
mkdir -p gpurun_out
T0=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "multi_rhs" 2>&1 | tail -6 | tee gpurun_out/pytest_mrhs.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
: > gpurun_out/mrhs_cta_sweep.jsonl
run() { # lib mode prec recon nsrc cta_sources l1 tile...
  local lib=$1 mode=$2 prec=$3 recon=$4 nsrc=$5 cs=$6 l1=$7; shift 7
  B200_LIB=$lib B200_MRHS_MODE=$mode B200_MRHS_CTA_SOURCES=$cs B200_MRHS_L1=$l1 timeout 100 python bench.py --prec $prec --recon $recon --nsrc $nsrc \
     --steps 60 --no-cpu-baseline --no-e2e ${1:+--tile $1 $2 $3 $4} 2>> gpurun_out/bench_err.txt \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$lib'.split('/')[-1],'mode':'$mode','prec':'$prec','recon':$recon,'nsrc':$nsrc,'cta_sources':$cs,'l1':$l1,'cfg':'$B200_MRHS_CTA_CFG','tile':'$*','us_per_rhs':d['ms_per_rhs']*1e3,'gflops':d['value'],'sustained_us_per_rhs':d['sustained']['ms_per_step']/$nsrc*1e3}))" >> gpurun_out/mrhs_cta_sweep.jsonl
}
D=quda_b200/libquda_b200.so
for cfg in 1 2; do
  export B200_MRHS_CTA_CFG=$cfg
  run $D cta single 12 8 0 1 16 2 1 1
  run $D cta single 12 8 0 1 16 2 2 1
  run $D cta single 12 8 0 1 16 4 1 1
  run $D cta single 12 8 0 1 16 1 1 1
  run $D cta half 12 8 0 1 16 2 1 1
  run $D cta half 12 8 0 1 16 4 1 1
  run $D cta double 18 8 0 1 16 2 1 1
done
export B200_MRHS_CTA_CFG=1
run $D cta single 8 8 0 1 16 2 1 1
run $D cta single 18 8 0 1 16 2 1 1
run $D cta half 8 8 0 1 16 2 1 1
cat gpurun_out/mrhs_cta_sweep.jsonl
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 150 env B200_MRHS_MODE=cta B200_MRHS_CTA_CFG=1 PROF_TILE="16 2 1 1" ncu --set full --clock-control none -k regex:dslash_mrhs_cta -s 1 -c 1 \
   -o gpurun_out/mrhs_cta -f python tools/prof_target.py single 12 3 8 > gpurun_out/ncu_mrhs_cta.log 2>&1
ncu -i gpurun_out/mrhs_cta.ncu-rep --page raw --csv > gpurun_out/mrhs_cta_single_r12_n8.raw.csv 2>/dev/null
rm -f gpurun_out/mrhs_cta.ncu-rep
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
