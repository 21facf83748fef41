export PYTHONPATH=.
echo "=== C5 launch list (none)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file /tmp/c5.csv python scripts/c5_probe.py none > /tmp/c5.log 2>&1; tail -1 /tmp/c5.log | cut -c1-600
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('/tmp/c5.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    a=agg.setdefault(r[ki].split('(')[0],[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k_,a in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print(f"{k_:50s} n={a[0]:4d} us/launch={a[1]/a[0]:10.1f} share={a[1]/tot:.3f}")
PY
echo "=== C5 launch list (zstd)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pq -c 60 --csv --log-file /tmp/c5z.csv python scripts/c5_probe.py zstd > /tmp/c5z.log 2>&1; tail -1 /tmp/c5z.log | cut -c1-300
grep -E "k_pq_zstd|k_pq_levels|k_pq_expand" /tmp/c5z.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -12
