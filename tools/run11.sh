TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8b_c2.json 2> gpurun_out/n8b_c2.err
$TR bench.py --gpus 8 --workload c5 --total-n 1e9 --no-e2e --steps 3 --warmup 2 --sections 1 > gpurun_out/n8b_c5.json 2> gpurun_out/n8b_c5.err
$TR bench.py --gpus 8 --n 4000000 --total-n 4000000 --checksum 1 --steps 2 --warmup 1 --no-e2e > gpurun_out/n8b_cs.json 2> gpurun_out/n8b_cs.err
python bench.py --n 4000000 --checksum 1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/n1b_cs.json 2> gpurun_out/n1b_cs.err
tail -2 gpurun_out/n8b_c5.err
