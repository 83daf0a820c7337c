python -m pytest tests/test_gpu_parity.py -x -q -k "cli or fuzz" 2>&1 | tail -3
python tools/cli_e2e.py 10000000 2>&1 | tail -9
