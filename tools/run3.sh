python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err
