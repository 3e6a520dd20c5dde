python tools/bn_profile.py > gpurun_out/bn1.log 2>&1; cat gpurun_out/bn1.log
timeout 120 ncu --set full --clock-control none -k regex:k_bn_bwd_apply_rows -s 3 -c 1 -o gpurun_out/bn_bwd_apply -f python tools/bn_profile.py 1 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none -k regex:k_wgrad_tc -s 12 -c 1 -o gpurun_out/r01b_wgrad_tc_full -f python tools/wgrad_check.py 11 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
