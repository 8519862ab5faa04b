#!/bin/bash
# one GPU-box call: device coder tests + compress bench, then the training-step and headline records after the fusion-layer kernels
python -m pytest tests/test_gpu_rans_device.py -x -q 2>&1 | tail -30 > gpurun_out/t_rans.txt
python bench.py --workload mbt-mean-compress --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_compress.json 2> gpurun_out/b_compress.err
python bench.py --workload mm-train --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b_mmtrain.json 2> gpurun_out/b_mmtrain.err
python bench.py --workload master-train --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b_mastertrain.json 2> gpurun_out/b_mastertrain.err
python bench.py --workload mm-forward --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b_mmforward.json 2> gpurun_out/b_mmforward.err
tail -12 gpurun_out/t_rans.txt; tail -2 gpurun_out/b_*.err
