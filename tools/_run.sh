timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1 || { tail -40 gpurun_out/t_gpu.log | cut -c1-220; exit 1; }
tail -2 gpurun_out/t_gpu.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.log 2> gpurun_out/bench11.err; tail -1 gpurun_out/bench11.log | cut -c1-330; tail -3 gpurun_out/bench11.err


python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-prefetch 2>/dev/null | tail -1 | cut -c1-200
