#!/bin/bash
# full GPU suite, smoke, headline bench (what the driver runs at round end)
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 > gpurun_out/t_all_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err
tail -6 gpurun_out/t_all_gpu.txt; tail -2 gpurun_out/smoke.txt; tail -2 gpurun_out/b_default.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_default.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'train', d.get('train_step',{}).get('value'), d.get('retried_after_fault'))
PY
