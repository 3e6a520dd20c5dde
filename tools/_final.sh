python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_r01c_final.log 2> gpurun_out/bench_r01c_final.err; tail -1 gpurun_out/bench_r01c_final.log | cut -c1-300; tail -2 gpurun_out/bench_r01c_final.err
cp gpurun_out/bench_breakdown_n1.json gpurun_out/r01c_breakdown_n1.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_tc -s 60 -c 1 -o gpurun_out/r01c_wgrad_tc_full -f python tools/wgrad_check.py 11 > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
