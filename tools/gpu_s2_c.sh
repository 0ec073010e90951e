python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "device_side" 2>&1 | tail -30 > gpurun_out/gputests_s2c.txt
cat gpurun_out/gputests_s2c.txt
