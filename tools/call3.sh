set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -q -x -k "resample or slab" 2>&1 | tail -5
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --csv --log-file gpurun_out/overfetch.csv tools/micro/overfetch > gpurun_out/overfetch.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/overfetch.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; break
ki,mi,vi=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value')
cur=None
for r in rows[i+1:]:
    if len(r)>vi: print(r[0], r[ki][:40], r[mi], r[vi])
PY
