#!/bin/bash
# round 2, final evidence for the measurement scan + Shor workload
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/run_measure.py > $O/r02_measure_n30.log 2>&1; tail -4 $O/r02_measure_n30.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_ncu_launches_measure_n30.csv python tools/run_measure.py > /dev/null 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_chunk|k_exact|k_classify|k_super" -c 10 -o $O/r02_ncu_measure_n30 -f python tools/run_measure.py > $O/ncu_measure.log 2>&1; echo "ncu rc=$?"
timeout 400 python bench.py --workload shor > $O/r02_bench_shor_final.json 2> $O/r02_bench_shor_final.err; echo "shor rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_shor_final.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["config"]["measure_state_ms"], d["parity"], [(c.get("name"), c.get("gpu_ms_per_find_period")) for c in d.get("find_period", [])])
PY
