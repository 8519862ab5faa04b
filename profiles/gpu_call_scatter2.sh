#!/bin/bash
# col2im epilogue teams: per-layer times with 2 / 3 / 4 teams, then the col2im parity tests with the default and the other team counts
for t in 2 3 4; do MMC_TC_SCATTER_TEAMS=$t TAG="scatter_teams=$t" python profiles/probe_layers.py 2>&1 | tail -1; done > gpurun_out/probe_scatter_teams.txt
cat gpurun_out/probe_scatter_teams.txt
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/t_scatter.txt
MMC_TC_SCATTER_TEAMS=2 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/t_scatter.txt
MMC_TC_SCATTER_TEAMS=4 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/t_scatter.txt
cat gpurun_out/t_scatter.txt
