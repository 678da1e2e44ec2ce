for lib in "" build/alt/libqcs.so; do
  echo "== lib: ${lib:-default}"
  QCS_LIB_PATH=$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m.csv python tools/run_measure.py > gpurun_out/measure_ncu.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/m.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows:
    n=r[4][:40]
    if any(k in n for k in ("chunk_maps","chunk_sums")): print(n, r[-1], r[-2])
PY
done
