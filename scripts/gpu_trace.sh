export PYTHONPATH=.
echo "=== TESTS (parquet, orc, scale)"; timeout 900 python -m pytest tests/test_gpu_parquet.py tests/test_gpu_orc.py tests/test_gpu_scale.py -m gpu -q 2>&1 | tail -3
echo "=== C5 probe"
timeout 600 python scripts/c5_probe.py 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('run_A_none','run_B_zstd1'): print(k, {x:d[k][x] for x in ('rows_per_s','ms_per_step','decode_ms','merge_ms','parity')})"
echo "=== bench (no trace)"
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-extra --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d['roofline']['phase_ms']), d['roofline_decode']['frac'], d.get('parity_sample'))"
