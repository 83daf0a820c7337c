python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r1b_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_match_small|k_onesweep_pass|k_order_tile|k_keys|k_groupsort_warp|k_hkey|k_decode" -c 14 -o gpurun_out/prof_r1f -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1f.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1b.json 2> gpurun_out/bench_ref_r1b.err
