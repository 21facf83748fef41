# one GPU call of the round-2 loop: parity tests, the bench line, emit A/B (PAIMON_GPU_LIB), launch list / ncu captures
export PYTHONPATH=.
echo "=== TESTS"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== BENCH parquet c3 (main lib)"
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_FLAGS:-} 2>/tmp/bench_err.log | tail -1 > /tmp/bench_c3.json; tail -3 /tmp/bench_err.log; python - <<'PY'
import json
d=json.loads(open('/tmp/bench_c3.json').read())
print(d['value'], d['ms_per_step'], json.dumps(d['roofline']['phase_ms']), 'decode', json.dumps({k:d['roofline_decode'][k] for k in ('frac','stage_ms')}), 'parity', d.get('parity_sample'))
if d.get('e2e'): print('e2e', json.dumps({k:v for k,v in d['e2e'].items() if k not in ('api','sample')}))
print(json.dumps(d.get('extra'))[:6000])
PY
mkdir -p gpurun_out; cp /tmp/bench_c3.json gpurun_out/r02_bench_c3.json
for rep in 1 2; do for v in main ${VARIANTS:-}; do for w in c3 c2; do
L=build/variants/libv_$v.so; [ $v = main ] && L=paimon_b200/libpaimon_gpu.so
PAIMON_GPU_LIB=$L timeout 600 python bench.py --source columns --workload $w --no-e2e --no-extra --no-cpu-baseline --no-parity-sample --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('VARIANT $v $w rep $rep emit', round(r['phase_ms']['emit'],2), 'plan', round(r['phase_ms']['plan+scan'],2), 'frac', round(r['frac'],3))"
done; done; done
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
cp /tmp/launches.csv gpurun_out/r02_launches_c3_parquet.csv
if [ -n "${NCU_FULL:-}" ]; then
echo "=== NCU FULL"
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_emit|k_pq_expand|k_pq_walk_values|k_plan" --launch-skip 10 -c 5 -o /tmp/r02_full python bench.py --steps 1 --warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-parity-sample > /tmp/b2.log 2>&1
tail -2 /tmp/b2.log | cut -c1-200; ls -la /tmp/r02_full.ncu-rep && cp /tmp/r02_full.ncu-rep gpurun_out/r02_full.ncu-rep
fi
