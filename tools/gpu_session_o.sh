#!/bin/bash
# round 2, final: full GPU suite, measurement evidence, Shor workload
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_pytest_gpu_final.log
timeout 300 python tools/run_measure.py > $O/r02_measure_n30.log 2>&1; tail -4 $O/r02_measure_n30.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_ncu_launches_measure_n30.csv python tools/run_measure.py > /dev/null 2>&1; echo "launch list rc=$?"
timeout 400 python bench.py --workload shor > $O/r02_bench_shor_final.json 2> $O/r02_bench_shor_final.err; echo "shor rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_shor_final.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["config"]["measure_state_ms"], d["parity"])
PY
