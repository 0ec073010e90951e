python -m pytest tests/test_gpu_modp.py -m gpu -x -q -k "full_round or tampering or config2" 2>&1 | tail -2
