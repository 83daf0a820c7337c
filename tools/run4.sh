python -m pytest tests/test_gpu_parity.py -x -q -k sort_pairs 2>&1 | tail -2
for v in "" _a _d _f; do
RK_LIB_SUFFIX=$v python tools/bench_sort.py 10000000 24 2>&1 | grep -E "pass:|hist" | tr '\n' ' '; echo " <= variant '$v'"
done
