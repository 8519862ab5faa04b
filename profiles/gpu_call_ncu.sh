#!/bin/bash
python profiles/_fusion_ncu.py > gpurun_out/fusion_ncu_plain.txt 2>&1 || { tail -5 gpurun_out/fusion_ncu_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"maxpool|upsample_bwd|sigmoid_gate_bwd|layernorm_bwd|gelu_bwd|window_attention|rans_lane" \
    --launch-skip 13 -c 13 -o gpurun_out/r02_fusion_coder python profiles/_fusion_ncu.py > gpurun_out/fusion_ncu.log 2>&1
tail -3 gpurun_out/fusion_ncu_plain.txt; tail -3 gpurun_out/fusion_ncu.log; ls -la gpurun_out/r02_fusion_coder.ncu-rep
