#!/bin/bash
# ncu --set full of g_a.0 (first conv launch of the second forward) and g_s.4 (13th) with the pipelined GDN epilogue
python profiles/_fwd_once.py > gpurun_out/fwd_once_plain.txt 2>&1 || { tail -3 gpurun_out/fwd_once_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 14 -c 1 -f -o gpurun_out/r02_ga0_pipe python profiles/_fwd_once.py > gpurun_out/ncu_ga0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 26 -c 1 -f -o gpurun_out/r02_gs4_pipe python profiles/_fwd_once.py > gpurun_out/ncu_gs4.log 2>&1
tail -1 gpurun_out/ncu_ga0.log gpurun_out/ncu_gs4.log; ls -la gpurun_out/r02_ga0_pipe.ncu-rep gpurun_out/r02_gs4_pipe.ncu-rep
