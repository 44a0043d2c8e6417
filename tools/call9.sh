ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'march3|prefilter|taps|zwin|select_hist|normalize' -c 14 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 1 --warmup 0 --no-variant --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_f.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_r01d.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; break
ki,mi,vi,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
d=collections.OrderedDict()
for r in rows[i+1:]:
    if len(r)>vi: d.setdefault((r[ii],r[ki][:50]),{})[r[mi]]=r[vi]
for k,m in d.items(): print(k, m)
PY
