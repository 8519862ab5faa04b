#!/bin/bash
# fresh-process stress of the round-end default path: N runs of two batch-64 forwards each, every run must end with the same bpp
N=${1:-12}
for i in $(seq 1 $N); do
  if [ $((i % 3)) -eq 0 ]; then export CUDA_LAUNCH_BLOCKING=1; else unset CUDA_LAUNCH_BLOCKING; fi
  timeout 120 python profiles/_fwd_once.py 2>&1 | tail -1
done > gpurun_out/stress_default.txt
sort gpurun_out/stress_default.txt | uniq -c
