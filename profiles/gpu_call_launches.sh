#!/bin/bash
# launch list of the headline bench command with the round-end kernels (ncu per-launch times are cold-cache and serialised: shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-record > gpurun_out/b_plain3.json 2> gpurun_out/b_plain3.err || { tail -3 gpurun_out/b_plain3.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-record > gpurun_out/ncu_launch3.log 2>&1
ls -la gpurun_out/r02_launches_v3.csv; tail -1 gpurun_out/b_plain3.json | cut -c1-200
