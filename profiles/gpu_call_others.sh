#!/bin/bash
# for-the-record lines of other workloads with the round-end kernels
for w in factorized mbt-mean-symbols mm-forward; do
  timeout 150 python bench.py --workload $w --no-cpu-baseline > gpurun_out/b_final_$w.json 2> gpurun_out/b_final_$w.err || tail -2 gpurun_out/b_final_$w.err
done
python - <<'PY'
import json
for w in ("factorized","mbt-mean-symbols","mm-forward"):
    try:
        d=json.loads(open(f"gpurun_out/b_final_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(w, "ERR", e)
PY
