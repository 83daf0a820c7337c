python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err
python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_c1.json 2> gpurun_out/b_c1.err
