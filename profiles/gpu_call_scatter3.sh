#!/bin/bash
# where is the reconstruction layer bound? main loop switched off piecewise (MMC_TC_DEBUG: 1 = no TMA after priming, 2 = no MMAs)
for t in 2 3; do for d in 0 1 2; do MMC_TC_SCATTER_TEAMS=$t MMC_TC_DEBUG=$d TAG="scatter_teams=$t debug=$d" python profiles/probe_layers.py 2>&1 | tail -1; done; done > gpurun_out/probe_scatter_debug.txt
cat gpurun_out/probe_scatter_debug.txt
