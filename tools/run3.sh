python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
for v in "" _mt4 _mt5 _mt6; do
RK_LIB_SUFFIX=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b$v.json 2> gpurun_out/b$v.err
done
