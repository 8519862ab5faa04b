#!/bin/bash
# programmatic dependent launch between the conv layers: parity tests, per-layer probe and pipeline probe with it on / off
MMC_TC_PDL=1 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/t_pdl.txt
cat gpurun_out/t_pdl.txt
for p in 0 1; do export MMC_TC_PDL=$p; MMC_TC_PDL=$p TAG="pdl=$p" python profiles/probe_layers.py 2>&1 | tail -1; MMC_TC_PDL=$p python profiles/probe_pipe.py 8,0 16,0 2>&1 | tail -2; done > gpurun_out/probe_pdl.txt
cat gpurun_out/probe_pdl.txt
