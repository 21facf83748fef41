export PYTHONPATH=.
timeout 900 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --e2e-steps 2 2>/tmp/e.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['parity_sample'], json.dumps({k:v for k,v in d['e2e'].items() if k not in ('api','sample')})[:700])"; tail -3 /tmp/e.log
