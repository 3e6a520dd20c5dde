python tools/tc_timeline.py 150000 27 96 96 > gpurun_out/tl_96.txt 2>&1
python tools/tc_timeline.py 317485 27 64 64 > gpurun_out/tl_64.txt 2>&1
tail -n 2 gpurun_out/tl_96.txt gpurun_out/tl_64.txt
