set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "two_gpu" 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/b_n2.json 2> gpurun_out/b_n2.err; tail -5 gpurun_out/b_n2.err; tail -c 1500 gpurun_out/b_n2.json
