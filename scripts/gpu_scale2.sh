export PYTHONPATH=.
echo "=== C5 probe"
timeout 600 python scripts/c5_probe.py 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('run_A_none','run_B_zstd1'): print(k, {x:d[k][x] for x in ('rows_per_s','ms_per_step','decode_ms','merge_ms','data_pages','decode_frac_of_hbm_peak','parity')})"
echo "=== 2-GPU bench"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/tmp/err2.log | tail -1 > /tmp/bench2.json; tail -5 /tmp/err2.log
python - <<'PY'
import json
d=json.loads(open('/tmp/bench2.json').read())
print(d['n_gpus'], d['value'], d['ms_per_step'], json.dumps(d['roofline']['phase_ms']), 'e2e', json.dumps({k:v for k,v in (d.get('e2e') or {}).items() if k not in ('api','sample')})[:900])
print(json.dumps(d.get('extra'))[:600])
PY
mkdir -p gpurun_out; cp /tmp/bench2.json gpurun_out/r02_bench_c3_2gpu.json
echo "=== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-300
