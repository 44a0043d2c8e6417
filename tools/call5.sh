ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 1 --no-variant --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_d.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_r01c.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; break
ki,mi,vi,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
d=collections.OrderedDict()
for r in rows[i+1:]:
    if len(r)>vi:
        d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=float(r[vi].replace(',',''))
agg=collections.OrderedDict()
for (id_,k),m in d.items():
    a=agg.setdefault(k,[0,0,0,0]); a[0]+=1; a[1]+=m.get('gpu__time_duration.sum',0); a[2]+=m.get('dram__bytes_read.sum',0); a[3]+=m.get('dram__bytes_write.sum',0)
for k,a in agg.items(): print(f'{k:60s} n={a[0]:4d} t={a[1]/1e3:9.1f}us rd={a[2]/1e6:9.1f}MB wr={a[3]/1e6:9.1f}MB')
PY
