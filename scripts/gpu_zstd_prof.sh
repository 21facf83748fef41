export PYTHONPATH=.
timeout 1200 ncu --section SourceCounters --section WarpStateStats --section LaunchStats --section SpeedOfLight --clock-control none --import-source on -k regex:k_pq_zstd --launch-skip 1 -c 1 -o /tmp/zstd_prof python scripts/c5_probe.py zstd > /tmp/z.log 2>&1; tail -2 /tmp/z.log | cut -c1-200
ls -la /tmp/zstd_prof.ncu-rep && mkdir -p gpurun_out && cp /tmp/zstd_prof.ncu-rep gpurun_out/zstd_prof2.ncu-rep
