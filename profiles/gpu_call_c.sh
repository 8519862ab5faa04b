#!/bin/bash
python -m pytest tests/test_gpu_rans_device.py -x -q 2>&1 | tail -30 > gpurun_out/t_rans.txt
python bench.py --workload mbt-mean-compress --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_compress.json 2> gpurun_out/b_compress.err
tail -5 gpurun_out/t_rans.txt; tail -2 gpurun_out/b_compress.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_compress.json').read().strip().splitlines()[-1])
print({k: (d[k] if not isinstance(d[k], dict) else {a: d[k][a] for a in ('value','ms_per_step','coder_ms_per_step','rate_overhead','lanes_y') if a in d[k]}) for k in ('value','e2e','e2e_sync_compress','e2e_device_coder') if k in d})
PY
