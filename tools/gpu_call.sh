#!/bin/bash
# Generic GPU call: tests first (fail fast), then whatever the arguments say.  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
step() { echo; echo "=== $1"; shift; "$@"; local rc=$?; echo "--- rc=$rc"; return $rc; }
for s in "$@"; do
  case "$s" in
    tests) step "gpu suite" timeout 700 python -m pytest tests -m gpu -x -q || exit 1 ;;
    tests_nf) step "gpu suite (non-fatal)" timeout 700 python -m pytest tests -m gpu -q ;;
    timeline2) for a in "495518 27 32 32" "317485 27 64 64"; do step "timeline $a" env SCN_B200_LIB=sparseeventid_b200/lib/libscn_b200_dbg.so timeout 120 python tools/tc_timeline.py $a; done ;;
    timeline) for a in "495518 27 32 32" "317485 27 64 64" "317485 27 64 64 1" "154605 27 96 96" "7332 27 192 192"; do step "timeline $a" env SCN_B200_LIB=sparseeventid_b200/lib/libscn_b200_dbg.so timeout 120 python tools/tc_timeline.py $a; done ;;
    convtests) step "conv kernel tests" timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q || exit 1 ;;
    smoke) step "smoke" timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1 ;;
    sweep) for i in 0 1 2 3 4 5; do step "tc sweep shape $i" timeout 200 python tools/tc_sweep.py full $i; done ;;
    sweepq) step "tc sweep quick" timeout 300 python tools/tc_sweep.py quick ;;
    bench) step "bench" bash -c 'timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json && tail -c 2500 gpurun_out/bench_n1.json' ;;
    benchfull) step "bench full" bash -c 'timeout 600 python bench.py > gpurun_out/bench_n1.json && tail -c 2500 gpurun_out/bench_n1.json' ;;
    wgsweep) step "wgrad kernel tests" timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k wgrad || exit 1
        step "wgrad sweep (default)" timeout 200 python tools/wgrad_sweep.py ;;
    ncu_wgrad) for i in 4 1; do
        step "ncu k_wgrad_tc shape $i" bash -c "timeout 120 python tools/wgrad_sweep.py $i > gpurun_out/plain_wg$i.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_tc -s 3 -c 1 -f -o gpurun_out/wgrad_tc_shape$i python tools/wgrad_sweep.py $i > gpurun_out/ncu_wg$i.log 2>&1; tail -2 gpurun_out/ncu_wg$i.log"
      done ;;
    ncu_conv6) for a in "495518 27 32 32" "317485 27 64 64" "154605 27 96 96" "59700 27 128 128" "20727 27 160 160" "7332 27 192 192"; do
        tag=$(echo $a | tr ' ' '_')
        step "ncu k_conv_tc $a" bash -c "TC_PROFILE_CHECK=0 timeout 120 python tools/tc_profile.py $a 3 > gpurun_out/plain_$tag.log 2>&1 && TC_PROFILE_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -f -o gpurun_out/conv_tc_$tag python tools/tc_profile.py $a 3 > gpurun_out/ncu_$tag.log 2>&1; tail -2 gpurun_out/ncu_$tag.log"
      done ;;
    ncu_conv) for a in "495518 27 32 32" "317485 27 64 64"; do
        tag=$(echo $a | tr ' ' '_')
        step "ncu k_conv_tc $a" bash -c "TC_PROFILE_CHECK=0 timeout 120 python tools/tc_profile.py $a 3 > gpurun_out/plain_$tag.log 2>&1 && TC_PROFILE_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -f -o gpurun_out/conv_tc_$tag python tools/tc_profile.py $a 3 > gpurun_out/ncu_$tag.log 2>&1; tail -3 gpurun_out/ncu_$tag.log"
      done ;;
    *) step "$s" bash -c "$s" ;;
  esac
done
