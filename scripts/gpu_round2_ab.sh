# emit kernel A/B on one box: same library, only emit.cu differs (PAIMON_GPU_LIB)
export PYTHONPATH=.
for rep in 1 2; do for v in main ${VARIANTS:-}; do for w in c3 c2; do
L=build/variants/libv_$v.so; [ $v = main ] && L=paimon_b200/libpaimon_gpu.so
PAIMON_GPU_LIB=$L timeout 600 python bench.py --source columns --workload $w --no-e2e --no-extra --no-cpu-baseline --no-parity-sample --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('VARIANT $v $w rep $rep emit', round(r['phase_ms']['emit'],2), 'plan', round(r['phase_ms']['plan+scan'],2), 'frac', round(r['frac'],3))"
done; done; done
