python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err
