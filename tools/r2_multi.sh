#!/bin/bash
# Round-2 multi-GPU items (gpurun --gpus N -- 'bash tools/r2_multi.sh N'): NVLink mailbox all-reduce in the CG, and the
# BASELINE config-5 solve (48^3 x 96 Wilson-clover CG, double / single mixed) through bench.py --op cg
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
B200_ALLREDUCE=nvlink timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "cg" 2>&1 | tail -4 | tee gpurun_out/r2_pytest_cg_nvlink_$N.txt
for mode in callback nvlink; do
  B200_ALLREDUCE=$mode run bench.py --gpus $N --op cg 2> gpurun_out/r2_cg_err_${mode}_$N.txt | tee gpurun_out/r2_cg_${mode}_$N.json | cut -c1-600
done
