#!/bin/bash
# 2-GPU records with the round-end kernels: headline workload and ssf2020 (one GOP per GPU)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-train-record > gpurun_out/b_n2_final.json 2> gpurun_out/b_n2_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload ssf2020 > gpurun_out/b_ssf_n2_final.json 2> gpurun_out/b_ssf_n2_final.err
python bench.py --workload ssf2020 > gpurun_out/b_ssf_n1_final.json 2> gpurun_out/b_ssf_n1_final.err
tail -2 gpurun_out/b_n2_final.err gpurun_out/b_ssf_n2_final.err
python - <<'PY'
import json
for f in ('b_n2_final','b_ssf_n2_final','b_ssf_n1_final'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'])
    except Exception as e:
        print(f, 'ERR', e)
PY
