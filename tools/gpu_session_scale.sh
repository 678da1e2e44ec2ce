#!/bin/bash
# the driver's scaling sequence at one N: bench.py under torchrun (or plain for N = 1)
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $RUN --master-port 29630 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_scale_n${N}.json 2> $O/r02_scale_n${N}.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$O/r02_scale_n${N}.json") if l.startswith('{"metric')][-1])
ns = d.get("north_star", {})
print("N", d["n_gpus"], "n", d["config"]["qubits"], "ms", round(d["ms_per_step"], 2), "value", round(d["value"], 1), "parity", d["parity"]["ok"], d["parity"]["max_rel_err"],
      "| roofline", d["roofline"]["kernel"], round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "| e2e ms", round(d["e2e"]["ms_per_step"], 1), "h2d GB/s", round(d["e2e"]["h2d_GBps_per_gpu"], 1),
      "| north_star n", ns.get("qubits"), "ms", round(ns.get("ms_per_qft", 0), 1), "parity", ns.get("parity", {}).get("ok"))
print({k: (v["launches"], v["ms"], v["GBps"]) for k, v in d["kernels"].items()})
PY
tail -2 $O/r02_scale_n${N}.err
