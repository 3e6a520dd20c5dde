timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 2 -c 1 -o gpurun_out/tc_new2 -f python tools/tc_profile.py 317485 27 64 64 3 > gpurun_out/ncu_new.log 2>&1
tail -2 gpurun_out/ncu_new.log
