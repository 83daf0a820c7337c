python -m pytest tests/test_gpu_parity.py -x -q -k "config4" 2>&1 | tail -3
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_c4.json 2> gpurun_out/b_c4.err; tail -3 gpurun_out/b_c4.err
