export PYTHONPATH=.
echo "=== TESTS"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== BENCH parquet c3 (main lib)"
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/tmp/bench_err.log | tail -1 > /tmp/bench_c3.json; tail -3 /tmp/bench_err.log; python - <<'PY'
import json
d=json.loads(open('/tmp/bench_c3.json').read())
print(d['value'], d['ms_per_step'], json.dumps(d['roofline']['phase_ms']), 'decode', json.dumps({k:d['roofline_decode'][k] for k in ('frac','stage_ms')}), 'parity', d.get('parity_sample'))
print('e2e', json.dumps({k:v for k,v in d['e2e'].items() if k not in ('api','sample')}))
print(json.dumps(d.get('extra'))[:1500])
PY
for v in main 42; do for w in c3 c2; do
echo "=== VARIANT $v $w"
L=build/variants/libv_$v.so; [ $v = main ] && L=paimon_b200/libpaimon_gpu.so
PAIMON_GPU_LIB=$L timeout 600 python bench.py --source columns --workload $w --no-e2e --no-extra --no-cpu-baseline --no-parity-sample --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'frac', r['frac'], json.dumps(r['phase_ms']))"
done; done
for v in t 4t; do for w in c3 c2; do
echo "=== TIMING $v $w"
PAIMON_GPU_LIB=build/variants/libv_$v.so timeout 600 python bench.py --source columns --workload $w --no-e2e --no-extra --no-cpu-baseline --no-parity-sample --steps 1 --warmup 3 2>&1 | grep "emit timing" | tail -3 | cut -c1-1800
done; done
