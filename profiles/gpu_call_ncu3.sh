#!/bin/bash
# one ncu --set full capture of the reconstruction layer (g_s.6, the 14th conv launch of the second forward)
python profiles/_fwd_once.py > gpurun_out/fwd_once_plain.txt 2>&1 || { tail -3 gpurun_out/fwd_once_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 27 -c 1 -f -o gpurun_out/r02_scatter_v3 python profiles/_fwd_once.py > gpurun_out/ncu_scatter3.log 2>&1
tail -2 gpurun_out/ncu_scatter3.log; ls -la gpurun_out/r02_scatter_v3.ncu-rep
