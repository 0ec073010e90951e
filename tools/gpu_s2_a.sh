# session check: GPU parity suite, default bench line, the small box (chunk heuristic: K must stay 1 at t = 683)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests_s2.txt
python bench.py > gpurun_out/bench_s2_default.json 2> gpurun_out/bench_s2_default.err
python bench.py --n 1024 --t 683 --no-also --no-cpu-baseline > gpurun_out/bench_s2_c1.json 2>&1
tail -3 gpurun_out/gputests_s2.txt; cut -c1-600 gpurun_out/bench_s2_default.json
