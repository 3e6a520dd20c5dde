#!/bin/bash
# Generic GPU call: tests first (fail fast), then whatever the arguments say.  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
step() { echo; echo "=== $1"; shift; "$@"; local rc=$?; echo "--- rc=$rc"; return $rc; }
for s in "$@"; do
  case "$s" in
    tests) step "gpu suite" timeout 500 python -m pytest tests -m gpu -x -q || exit 1 ;;
    convtests) step "conv kernel tests" timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q || exit 1 ;;
    smoke) step "smoke" timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1 ;;
    sweep) for i in 0 1 2 3 4 5; do step "tc sweep shape $i" timeout 200 python tools/tc_sweep.py full $i; done ;;
    sweepq) step "tc sweep quick" timeout 300 python tools/tc_sweep.py quick ;;
    bench) step "bench" bash -c 'timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json && tail -c 2500 gpurun_out/bench_n1.json' ;;
    benchfull) step "bench full" bash -c 'timeout 600 python bench.py > gpurun_out/bench_n1.json && tail -c 2500 gpurun_out/bench_n1.json' ;;
    *) step "$s" bash -c "$s" ;;
  esac
done
