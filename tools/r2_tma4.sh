#!/bin/bash
# Round-2 GPU pack 4: converged producer warp (elect.sync), new host layer (device-scalar CG)
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_tma.py tests/test_gpu_dirac.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r2_pytest_tma4.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
: > gpurun_out/r2_tma4_sweep.jsonl
run() { # label prec recon [env...]
  local label=$1 prec=$2 recon=$3; shift 3
  env "$@" timeout 120 python bench.py --prec $prec --recon $recon --steps 200 --no-cpu-baseline --no-e2e --no-mrhs 2>> gpurun_out/bench_err.txt \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'label':'$label','prec':'$prec','recon':$recon,'env':'$*','us':d['ms_per_step']*1e3,'frac':d['roofline']['frac'],'sustained_us':d['sustained']['ms_per_step']*1e3,'clocks':d['clocks']}))" >> gpurun_out/r2_tma4_sweep.jsonl
  tail -1 gpurun_out/r2_tma4_sweep.jsonl
}
run gather single 12 B200_TMA=0
run tma single 12 B200_TMA=1
run tma_pf2 single 12 B200_TMA=1 B200_TMA_L2PF=2
run tma_l2 single 12 B200_TMA=1 B200_TMA_LINKS=2
run tma_82 single 12 B200_TMA=1 "B200_TMA_TILE=8 2"
run tma_28 single 12 B200_TMA=1 "B200_TMA_TILE=2 8"
run tma_g296 single 12 B200_TMA=1 B200_TMA_GRID=296
run tma_reg single 12 B200_TMA=1 B200_TMA_LINKS=-1 B200_TMA_PREFETCH=2 "B200_TMA_RINGS=4 2"
run gather single 8 B200_TMA=0
run tma    single 8 B200_TMA=1
run tma    single 18 B200_TMA=1
run gather double 12 B200_TMA=0
run tma    double 12 B200_TMA=1
echo "[t=$(( $(date +%s)-T0 ))s]"
timeout 200 env B200_TMA=1 ncu --set full --clock-control none --import-source on -k regex:dslash_tma -s 2 -c 1 \
   -o gpurun_out/r2_tma4 -f python tools/prof_target.py single 12 4 > gpurun_out/ncu_r2_tma4.log 2>&1
ncu -i gpurun_out/r2_tma4.ncu-rep --page raw --csv > gpurun_out/r2_tma4_single_r12.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_tma4.ncu-rep --page source --csv > gpurun_out/r2_tma4_single_r12.source.csv 2>/dev/null
rm -f gpurun_out/r2_tma4.ncu-rep
echo "[t=$(( $(date +%s)-T0 ))s]"; echo "== done"
