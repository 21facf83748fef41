# final evidence of a round: parity tests, the full default bench line (+ reference arm), launch list, full ncu capture
export PYTHONPATH=.
mkdir -p gpurun_out
echo "=== TESTS"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== SMOKE"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== BENCH (default)"
timeout 1800 python bench.py 2>/tmp/bench_err.log | tail -1 > gpurun_out/r02_bench_c3.json; tail -3 /tmp/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c3.json').read())
print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], json.dumps(d['roofline']['phase_ms']), 'decode', json.dumps({k:d['roofline_decode'][k] for k in ('frac','stage_ms')}), 'parity', d.get('parity_sample'))
print('e2e', json.dumps({k:v for k,v in d['e2e'].items() if k not in ('api','sample')}))
print('cpu', json.dumps(d.get('cpu_baseline'))[:300], 'launches', d.get('gpu_launches'), 'clocks', d.get('clocks'))
print(json.dumps(d.get('extra'))[:3000])
PY
echo "=== BENCH reference arm"
timeout 900 python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/r02_bench_c3_reference_arm.json; cut -c1-400 gpurun_out/r02_bench_c3_reference_arm.json
echo "=== NCU LIST"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r02_c3_parquet_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-parity-sample > /tmp/b.log 2>&1
tail -1 /tmp/b.log | cut -c1-200
echo "=== NCU FULL"
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_emit|k_pq_expand|k_pq_walk_values|k_plan|k_pq_levels" --launch-skip 12 -c 6 -o /tmp/r02_full python bench.py --steps 1 --warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-parity-sample > /tmp/b2.log 2>&1
tail -2 /tmp/b2.log | cut -c1-200; ls -la /tmp/r02_full.ncu-rep && cp /tmp/r02_full.ncu-rep gpurun_out/r02_full_final.ncu-rep
