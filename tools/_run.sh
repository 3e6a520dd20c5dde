timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1 || { tail -40 gpurun_out/t_gpu.log | cut -c1-220; exit 1; }
tail -2 gpurun_out/t_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.log 2> gpurun_out/bench8.err; tail -1 gpurun_out/bench8.log | cut -c1-400
python tools/step_profile.py > gpurun_out/step_profile2.txt 2>&1; head -4 gpurun_out/step_profile2.txt | cut -c1-200
