#!/bin/bash
# round 2, after the measurement-scan rework: full GPU suite, Shor workload, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_m.log
timeout 600 python bench.py --workload shor > gpurun_out/bench_shor_m.json 2> gpurun_out/bench_shor_m.err; echo "shor rc=$?"
timeout 900 python bench.py > gpurun_out/bench_default_m.json 2> gpurun_out/bench_default_m.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_shor_m.json", "gpurun_out/bench_default_m.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "parity", d.get("parity", {}).get("ok"),
              "frac", round(d["roofline"]["frac"], 3), d["config"].get("measure_state_ms"), "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
