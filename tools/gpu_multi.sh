#!/bin/bash
# Multi-GPU call (gpurun --gpus N): NVLink peer-write halo tests against the global oracle + weak-scaling bench lines.
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_multi_$N.txt
echo "[t=$(( $(date +%s)-T0 ))s]"
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  echo "== bench --gpus $n"
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --no-cpu-baseline 2> gpurun_out/bench_err_1.txt | tee gpurun_out/scale_1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --breakdown --no-cpu-baseline 2> gpurun_out/bench_err_$n.txt | tee gpurun_out/scale_$n.json
    tail -3 gpurun_out/bench_err_$n.txt
  fi
done
echo "[t=$(( $(date +%s)-T0 ))s]"
echo "== done"
