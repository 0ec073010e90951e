python -m pytest tests/test_gpu_modp.py -m gpu -x -q -k "tampering" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck_r02.txt 2>&1; tail -4 gpurun_out/sanitizer_memcheck_r02.txt
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/sanitizer_racecheck_r02.txt 2>&1; tail -2 gpurun_out/sanitizer_racecheck_r02.txt
