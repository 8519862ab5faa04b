#!/bin/bash
# software-pipelined GDN epilogue (default; MMC_TC_EPI_PIPE=0 = the serial chain): parity tests, per-layer times off / on
timeout 400 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py tests/test_gpu_models_mm.py tests/test_gpu_models_video.py tests/test_gpu_fullsize_parity.py -q -m gpu 2>&1 | tail -4 > gpurun_out/t_pipe.txt
cat gpurun_out/t_pipe.txt
for p in 0 1; do MMC_TC_EPI_PIPE=$p TAG="epi_pipe=$p" timeout 120 python profiles/probe_layers.py 2>&1 | tail -1; done > gpurun_out/probe_epi_pipe.txt
cat gpurun_out/probe_epi_pipe.txt
