python tools/profile_part.py 2>&1 | cut -c1-60,150-215 | grep -E "ms per call|k_interleave|k_unpack|k_gather|k_onesweep_pass|Self CUDA time"
python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -2
