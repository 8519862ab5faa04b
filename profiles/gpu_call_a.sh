#!/bin/bash
# one GPU-box call: device coder tests, compress bench, full-model fault probe (timing per variant, then a fresh-process loop)
python -m pytest tests/test_gpu_rans_device.py -x -q 2>&1 | tail -30 > gpurun_out/t_rans.txt
python bench.py --workload mbt-mean-compress --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_compress.json 2> gpurun_out/b_compress.err
export LD_LIBRARY_PATH=$PWD/$(ls -d 165-*_b200)/mmcodec
for v in "X=1" "MMC_TC_GROUPED=1" "MMC_TC_TEAMS=2" "MMC_TC_GROUPED=1 MMC_TC_TEAMS=2"; do
  echo "== [$v]"; s=$(date +%s%N); env $v PROBE_SYNC_EACH=1 profiles/bin/probe_fault 64 2; echo "rc $? wall $(( ($(date +%s%N) - s) / 1000000 )) ms"
done > gpurun_out/probe_timing.txt 2>&1
bash profiles/fault_loop.sh ${1:-170} grouped teams both > gpurun_out/fault_loop_3.txt 2>&1
tail -12 gpurun_out/t_rans.txt; tail -3 gpurun_out/b_compress.err; tail -6 gpurun_out/fault_loop_3.txt
