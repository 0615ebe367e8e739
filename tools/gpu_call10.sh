#!/bin/bash
mkdir -p gpurun_out
for v in _nopre "" _pre170 _pre255; do
  echo "== tune lib$v"; B200_LIB=$PWD/quda_b200/libquda_b200$v.so timeout 300 python tools/tune.py "lib$v" single:12,half:12,half:8,single:8 2>&1 | grep tune
done
timeout 300 python -m pytest tests/test_gpu_wilson.py tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -3
