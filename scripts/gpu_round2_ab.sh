# emit kernel A/B on one box: same library, only emit.cu differs (PAIMON_GPU_LIB); then a launch list of the parquet step
export PYTHONPATH=.
for rep in 1 2; do for v in main c1 4 42; do for w in c3 c2; do
L=build/variants/libv_$v.so; [ $v = main ] && L=paimon_b200/libpaimon_gpu.so
PAIMON_GPU_LIB=$L timeout 600 python bench.py --source columns --workload $w --no-e2e --no-extra --no-cpu-baseline --no-parity-sample --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('VARIANT $v $w rep $rep emit', round(r['phase_ms']['emit'],2), 'plan', round(r['phase_ms']['plan+scan'],2), 'frac', round(r['frac'],3), d['clocks'])"
done; done; done
for v in t 4t; do for w in c3 c2; do
echo "=== TIMING $v $w"
PAIMON_GPU_LIB=build/variants/libv_$v.so timeout 600 python bench.py --source columns --workload $w --no-e2e --no-extra --no-cpu-baseline --no-parity-sample --steps 1 --warmup 3 2>&1 | grep "emit timing" | tail -2 | cut -c1-1800
done; done
echo "=== NCU LIST"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file /tmp/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-parity-sample > /tmp/b.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('/tmp/launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    a=agg.setdefault(r[ki].split('(')[0],[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k_,a in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k_:60s} n={a[0]:4d} us/launch={a[1]/a[0]:10.1f} share={a[1]/tot:.3f}")
PY
mkdir -p gpurun_out; cp /tmp/launches.csv gpurun_out/r02_launches_c3_parquet_b.csv
