#!/bin/bash
# reconstruction-layer gather rewrite: col2im / conv_tc parity tests, model parity tests, headline bench without the side records
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/t_scatter.txt
python bench.py --no-cpu-baseline --no-train-record > gpurun_out/b_scatter.json 2> gpurun_out/b_scatter.err
tail -4 gpurun_out/t_scatter.txt; tail -2 gpurun_out/b_scatter.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_scatter.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
print(d['roofline']['layer_ms'])
PY
