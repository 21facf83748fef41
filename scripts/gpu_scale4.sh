export PYTHONPATH=.
free -g | head -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 4 2>/tmp/err4.log | tail -1 > /tmp/bench4.json; tail -4 /tmp/err4.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('/tmp/bench4.json').read())
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', json.dumps({k:v for k,v in (d.get('e2e') or {}).items() if k not in ('api','sample')})[:900])
PY
mkdir -p gpurun_out; cp /tmp/bench4.json gpurun_out/r02_bench_c3_4gpu.json
