ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --clock-control none --import-source on -k regex:"k_match_small" -c 2 -o gpurun_out/prof_r1d -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1d.log 2>&1
tail -3 gpurun_out/ncu_r1d.log
