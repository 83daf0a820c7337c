python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -6
python tools/cli_e2e.py 2000000 2>&1 | tail -12
