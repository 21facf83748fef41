export PYTHONPATH=.
echo "=== TESTS (parquet)"; timeout 900 python -m pytest tests/test_gpu_parquet.py tests/test_gpu_scale.py -m gpu -q 2>&1 | tail -3
echo "=== decode trace"
PAIMON_GPU_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity-sample 2>&1 | grep -E "decode trace" | tail -3 | cut -c1-260
echo "=== bench (no trace)"
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-extra --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d['roofline']['phase_ms']), d['roofline_decode']['frac'], d.get('parity_sample'))"
