set -x
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA node\(s\)"
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_new.json 2> gpurun_out/b_new.err
RK_LIB_SUFFIX=_mt5 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_mt5.json 2> gpurun_out/b_mt5.err
RK_L2_GRAN=32 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_g32.json 2> gpurun_out/b_g32.err
RK_L2_GRAN=128 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_g128.json 2> gpurun_out/b_g128.err
python tools/pcie_bw.py
