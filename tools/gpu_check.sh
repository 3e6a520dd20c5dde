#!/bin/bash
# GPU-box check used while iterating on the conv kernel: correctness first (fail fast), then timings.
#   bash tools/gpu_check.sh "<exp flags...>" <logname>
set -e
EXPS=${1:-0}
LOG=${2:-tcm}
timeout 120 python tools/tc_check.py > gpurun_out/tcc.log 2>&1 || { tail -20 gpurun_out/tcc.log; exit 1; }
tail -1 gpurun_out/tcc.log
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1 || { tail -30 gpurun_out/t_gpu.log | cut -c1-200; exit 1; }
tail -1 gpurun_out/t_gpu.log
for e in $EXPS; do
  echo "EXP=$e"
  for a in "500000 27 32 32" "317485 27 64 64" "150000 27 96 96" "60000 27 128 128" "20000 27 160 160" "7000 27 192 192" "317485 8 32 64" "317485 8 64 32"; do SCN_B200_TC_EXP=$e timeout 30 python tools/tc_profile.py $a 3; done
done > gpurun_out/$LOG.log 2>&1
cat gpurun_out/$LOG.log
# the plain-C caller of the C ABI (no Python in the process)
gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include tools/abi_harness.c -Lsparseeventid_b200/lib -lscn_b200 \
    -L/usr/local/cuda/lib64 -lcudart -lm -o /tmp/abi_harness && LD_LIBRARY_PATH=sparseeventid_b200/lib timeout 60 /tmp/abi_harness
