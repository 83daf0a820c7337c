ncu --set full --clock-control none --import-source on -k regex:"k_order_tile" -c 1 -o gpurun_out/prof_r1e -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1e.log 2>&1
tail -2 gpurun_out/ncu_r1e.log
