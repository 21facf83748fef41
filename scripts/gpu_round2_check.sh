# one GPU call of the round-2 loop: parity tests, emit kernel variants, launch list, full ncu capture
export PYTHONPATH=.
echo "=== TESTS"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
for v in ${VARIANTS:-}; do for w in c3 c2; do
echo "=== VARIANT $v $w"
PAIMON_GPU_LIB=build/variants/libv_$v.so timeout 600 python bench.py --source columns --workload $w --no-e2e --no-extra --no-cpu-baseline --no-parity-sample --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'frac', r['frac'], json.dumps(r['phase_ms']))"
done; done
if [ -n "${NCU_FULL:-}" ]; then
echo "=== NCU FULL"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_emit|k_pq_expand|k_pq_levels|k_pq_walk_bytes|k_plan" --launch-skip 10 -c 10 -o /tmp/r02_full python bench.py --steps 1 --warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-parity-sample > /tmp/b2.log 2>&1
tail -2 /tmp/b2.log | cut -c1-200; ls -la /tmp/r02_full.ncu-rep && mkdir -p gpurun_out && cp /tmp/r02_full.ncu-rep gpurun_out/
fi
