set -x
mkdir -p gpurun_out
(timeout 90 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/c61_tests.log; cat gpurun_out/c61_tests.log
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c61_lossless_launches.csv python profiles/runs/r2bj_lossless_launches.py > gpurun_out/c61.log 2>&1
tail -3 gpurun_out/c61.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c61_lossless_launches.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value'); u=hdr.index('Metric Unit')
for r in rows[-9:]: print(r[k][:60], r[v], r[u])
PY
