#!/bin/bash
# last seconds of the round's GPU budget: the GPU parity tier (minus the 32^4 cases) with the no-lineinfo build of the
# final sources (identical code generation, 38 MB instead of 133 MB to push)
mkdir -p gpurun_out
export B200_LIB=$PWD/quda_b200/libquda_b200_nl.so
timeout 100 python -m pytest tests -m gpu -q --ignore tests/test_gpu_fullsize.py -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_tiny.txt
