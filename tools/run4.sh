for v in "" _a _b _c _d _e _f; do
RK_LIB_SUFFIX=$v python tools/bench_sort.py 10000000 24 2>&1 | grep -E "pass:" | tr '\n' ' '; echo " <= variant '$v'"
done
