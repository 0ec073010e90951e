timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/bench_r02_n8_strong.json 2> gpurun_out/bench_r02_n8_strong.err
tail -c 600 gpurun_out/bench_r02_n8_strong.json
