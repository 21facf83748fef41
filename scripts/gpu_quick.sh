export PYTHONPATH=.
echo "=== SMOKE"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== TESTS"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== BENCH short"; timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['parity_sample'], d['roofline']['frac'], d['roofline_decode']['frac'])"
