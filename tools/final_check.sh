#!/bin/bash
# One-call GPU verification of a build (what the driver runs at round end, condensed): GPU tests, smoke, both bench arms.
#   gpurun --timeout 1500 -- bash tools/final_check.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/final_ref.log", "gpurun_out/final_bench.log"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d.get("impl"), round(d["value"]), "%.1f ms" % d["ms_per_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None,
                  "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
PY
