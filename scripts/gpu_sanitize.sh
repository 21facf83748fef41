# compute-sanitizer over the decode / merge parity tests (slow: minutes per tool)
export PYTHONPATH=.
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r02_compute_sanitizer_$tool.log \
    python -m pytest tests/test_gpu_parquet.py tests/test_gpu_orc.py tests/test_gpu_merge.py -m gpu -q -x -k "not more_than_32 and not wide" 2>&1 | tail -3
  tail -3 gpurun_out/r02_compute_sanitizer_$tool.log
done
echo "=== racecheck (decode + merge of one section)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_compute_sanitizer_racecheck.log \
  python -m pytest tests/test_gpu_parquet.py -m gpu -q -x -k "section_runs_are or zstd or asynchronous" 2>&1 | tail -3
tail -3 gpurun_out/r02_compute_sanitizer_racecheck.log
