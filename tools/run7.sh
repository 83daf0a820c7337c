python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r1c_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_match_small|k_onesweep_pass|k_order_tile|k_keys|k_groupsort_warp|k_hkey|k_decode|k_chase" -c 16 -o gpurun_out/prof_r1g -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1g.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err
