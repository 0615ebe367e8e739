#!/bin/bash
# Scaling call (gpurun --gpus 8): multi-GPU parity tests + bench lines at N = 1, 2, 4, 8 + reference arm.
N=${1:-8}
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== pytest multi (2- and 4-GPU cases)"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "four or eight or (p2p and grid_dims0)" 2>&1 | tail -5 | tee gpurun_out/pytest_multi_$N.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
for n in 1 2 4 8; do
  [ $n -gt $N ] && continue
  echo "== bench --gpus $n"
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 2> gpurun_out/bench_err_1.txt | tee gpurun_out/scale_1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --breakdown --no-cpu-baseline 2> gpurun_out/bench_err_$n.txt | tee gpurun_out/scale_$n.json
    grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_err_$n.txt | tail -3
  fi
  echo "[t=$(( $(date +%s)-T0 ))s]"
done
echo "== reference arm under torchrun (rank 0 only)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
   bench.py --impl reference --gpus 2 --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/reference_arm.json
echo "== other precisions at N=1"
for cfg in "double 18" "half 12" "half 8" "single 18"; do set -- $cfg; timeout 200 python bench.py --prec $1 --recon $2 --no-e2e --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_$1_r$2.json | cut -c1-400; done
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== done"
