# 2 GPUs: the multi-GPU parity tests (skipped on one GPU) and the N = 2 bench line (weak + strong)
python -m pytest tests -m gpu -x -q -k "two_gpu or cpp_mirror" 2>&1 | tail -5 > gpurun_out/gputests_s2_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r02_ref_n2.json 2> gpurun_out/bench_r02_ref_n2.err
cat gpurun_out/gputests_s2_n2.txt; tail -c 1500 gpurun_out/bench_r02_n2.json; tail -c 600 gpurun_out/bench_r02_ref_n2.json
