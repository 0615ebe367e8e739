#!/bin/bash
mkdir -p gpurun_out
for v in "" _pf600 _pf1500 _pf4000; do
  echo "== tune lib$v"; B200_LIB=$PWD/quda_b200/libquda_b200$v.so timeout 300 python tools/tune.py "lib$v" single:12,double:18,half:12,single:18,double:12 2>&1 | grep tune
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --breakdown --no-cpu-baseline 2>gpurun_out/bench_err_2.txt | tee gpurun_out/scale_2.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_err_2.txt | tail -5
