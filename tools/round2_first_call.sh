#!/bin/bash
# First GPU call of the next round: validate what was written without a GPU, then time it.  Fail fast, short timeouts.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh > gpurun_out/first_call.log 2>&1; tail -60 gpurun_out/first_call.log'
set -u
mkdir -p gpurun_out
step() { echo; echo "=== $1"; shift; "$@"; local rc=$?; echo "--- rc=$rc"; return $rc; }

# 1. the default product path is intact (every change of the last session was meant to leave it untouched)
step "default gpu suite" timeout 400 python -m pytest tests -m gpu -x -q || exit 1
step "smoke" timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
# 2. plain-C caller of the C ABI
step "abi harness" bash -c 'gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include tools/abi_harness.c -Lsparseeventid_b200/lib -lscn_b200 -L/usr/local/cuda/lib64 -lcudart -lm -o /tmp/abi_harness && LD_LIBRARY_PATH=sparseeventid_b200/lib timeout 60 /tmp/abi_harness'
# 3. the experimental stage-list convolution: builder first, one conv, then everything
step "stage lists: builder" env SCN_B200_STAGE_LISTS_TESTS=1 timeout 120 python -m pytest tests/test_gpu_stage_lists.py -m gpu -x -q -k builder || exit 0
step "stage lists: one conv per kernel instantiation" bash -c 'for a in "20000 27 32 32" "20000 27 64 64" "20000 27 96 96" "8000 27 128 128" "8000 27 192 192"; do timeout 60 python tools/tc_profile.py $a 3 lists || exit 1; done' || exit 0
step "stage lists: acceptance tests" env SCN_B200_STAGE_LISTS_TESTS=1 timeout 400 python -m pytest tests/test_gpu_stage_lists.py -m gpu -x -q || exit 0
# 4. timings: per shape, then the whole step
step "per-shape timings (bench level sizes)" bash -c 'for a in "495518 27 32 32" "317485 27 64 64" "154605 27 96 96" "59700 27 128 128" "20727 27 160 160" "7332 27 192 192"; do timeout 60 python tools/tc_profile.py $a 5 lists; done'
step "bench default" bash -c 'timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json && tail -c 1500 gpurun_out/bench_default.json'
step "bench stage lists" bash -c 'SCN_B200_STAGE_LISTS=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_stage_lists.json && tail -c 1500 gpurun_out/bench_stage_lists.json'
# 5. weight re-streaming: tiles per group forced up (single TMEM buffer) at the levels where weights dominate L2 traffic
step "T sweep" bash -c 'for t in 2 3 4 5; do echo "SCN_B200_TC_T=$t"; for a in "154605 27 96 96" "59700 27 128 128" "20727 27 160 160" "7332 27 192 192"; do SCN_B200_TC_T=$t timeout 60 python tools/tc_profile.py $a 5; done; done'
step "grid sweep (deep levels: fewer CTAs x more tiles per group)" bash -c 'for g in 37 74 111; do for t in 2 3; do echo "SCN_B200_TC_GRID=$g SCN_B200_TC_T=$t"; for a in "20727 27 160 160" "7332 27 192 192"; do SCN_B200_TC_GRID=$g SCN_B200_TC_T=$t timeout 60 python tools/tc_profile.py $a 5; done; done; done'
