python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err
python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_c4.json 2> gpurun_out/b_c4.err
