#!/bin/bash
# phase-inner tile order of the transposed convolutions: parity tests with it forced on (small shapes) and with the default rule,
# per-layer times off / on, DRAM bytes of the 14 conv launches with the default rule
MMC_TC_PHASE_INNER=1 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py tests/test_gpu_models_video.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/t_phase.txt
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py tests/test_gpu_fullsize_parity.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/t_phase.txt
cat gpurun_out/t_phase.txt
for p in 0 1; do MMC_TC_PHASE_INNER=$p TAG="phase_inner=$p" python profiles/probe_layers.py 2>&1 | tail -1; done > gpurun_out/probe_phase.txt
TAG="phase_inner=default" python profiles/probe_layers.py 2>&1 | tail -1 >> gpurun_out/probe_phase.txt
cat gpurun_out/probe_phase.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_tc -s 14 -c 14 --csv --log-file gpurun_out/r02_ncu_conv_phase_inner.csv python profiles/_fwd_once.py > gpurun_out/ncu_phase.log 2>&1
cut -d, -f5,12- gpurun_out/r02_ncu_conv_phase_inner.csv | tail -15
