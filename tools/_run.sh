timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1 || { tail -40 gpurun_out/t_gpu.log | cut -c1-220; exit 1; }
tail -2 gpurun_out/t_gpu.log
timeout 300 python tools/conv_sweep.py --out gpurun_out/conv_sweep.json > gpurun_out/conv_sweep.log 2>&1; tail -45 gpurun_out/conv_sweep.log | cut -c1-200
timeout 200 python tools/fullres_inference.py --batch 128 --depth 9 > gpurun_out/fullres.log 2>&1; tail -3 gpurun_out/fullres.log | cut -c1-600
timeout 200 python bench.py --dataset dune2d --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dune2d.log 2> gpurun_out/bench_dune2d.err; tail -1 gpurun_out/bench_dune2d.log | cut -c1-500; tail -3 gpurun_out/bench_dune2d.err
