#!/bin/bash
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py tests/test_gpu_fullsize_parity.py tests/test_gpu_models_mm.py -x -q 2>&1 | tail -4 > gpurun_out/t_conv.txt
python bench.py --no-cpu-baseline --no-train-record > gpurun_out/b_default2.json 2> gpurun_out/b_default2.err
tail -3 gpurun_out/t_conv.txt; tail -2 gpurun_out/b_default2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_default2.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])
print({k: round(v,3) for k,v in d['roofline']['layer_ms'].items()})
PY
