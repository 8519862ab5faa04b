#!/bin/bash
python profiles/_fwd_once.py > gpurun_out/fwd_once_plain.txt 2>&1 || { tail -3 gpurun_out/fwd_once_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 14 -c 14 -o gpurun_out/r02_conv_tc_v2 python profiles/_fwd_once.py > gpurun_out/ncu_conv2.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-record > gpurun_out/b_plain2.json 2> gpurun_out/b_plain2.err || { tail -3 gpurun_out/b_plain2.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-record > gpurun_out/ncu_launch2.log 2>&1
tail -2 gpurun_out/ncu_conv2.log; ls -la gpurun_out/r02_conv_tc_v2.ncu-rep gpurun_out/r02_launches_v2.csv
