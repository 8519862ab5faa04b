#!/bin/bash
export LD_LIBRARY_PATH=$PWD/$(ls -d 165-*_b200)/mmcodec
for v in "X=1" "MMC_TC_PAIR=2" "MMC_TC_PAIR_MINKB=16" "MMC_TC_PAIR_MINKB=12" "MMC_TC_PAIR=2 MMC_TC_LATE_RELEASE=0"; do
  echo "== [$v]"; env $v PROBE_SYNC_EACH=1 profiles/bin/probe_fault 64 3 | grep "^rep 2" | tr '\n' ' '; echo
done > gpurun_out/probe_pair.txt 2>&1
cat gpurun_out/probe_pair.txt
