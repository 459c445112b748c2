set -x
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c59_lossless_launches.csv python profiles/runs/r2bj_lossless_launches.py > gpurun_out/c59.log 2>&1
tail -3 gpurun_out/c59.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c59_lossless_launches.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value'); u=hdr.index('Metric Unit')
for r in rows[1:]: print(r[k][:60], r[v], r[u])
PY
