python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_p1.json 2> gpurun_out/b_p1.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-kernels 0 > gpurun_out/b_p0.json 2> gpurun_out/b_p0.err
