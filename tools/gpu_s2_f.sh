# final single-GPU refresh: parity suite, default + ristretto255 bench lines, phase table
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/gputests_final.txt
timeout 600 python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err
timeout 300 python bench.py --group ristretto255 > gpurun_out/bench_r02_ristretto255.json 2> gpurun_out/bench_r02_rist.err
timeout 900 python tools/phase_timing.py > gpurun_out/phases_r02.jsonl 2> gpurun_out/phases.err
python __graft_entry__.py smoke 2>&1 | tail -2
cat gpurun_out/gputests_final.txt; tail -c 300 gpurun_out/bench_r02_default.json
