export PYTHONPATH=.
echo "=== TESTS"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "=== C5 probe"
timeout 600 python scripts/c5_probe.py 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('run_A_none','run_B_zstd1'): print(k, {x:d[k][x] for x in ('rows_per_s','ms_per_step','decode_ms','merge_ms','data_pages','decode_frac_of_hbm_peak','parity')})"
