python __graft_entry__.py smoke 2>&1 | tail -1
python -m pytest tests/test_gpu_ec.py tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -2
