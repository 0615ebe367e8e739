#!/bin/bash
# last seconds of the round's GPU budget: the 32^4 parity / property tests with the no-lineinfo build of the final sources
mkdir -p gpurun_out
export B200_LIB=$PWD/quda_b200/libquda_b200_nl.so
timeout 52 python -m pytest -m gpu -v -p no:cacheprovider \
  "tests/test_gpu_fullsize.py::test_config2_fp32_recon12_vs_oracle_and_properties" \
  "tests/test_gpu_fullsize.py::test_config3_clover_pc_half_recon8_vs_oracle" \
  "tests/test_gpu_fullsize.py::test_fp64_recon18_32cubed_vs_oracle" 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed|Error|assert" | tee gpurun_out/pytest_gpu_fullsize.txt
