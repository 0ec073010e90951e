# Round-end refresh on one B200: GPU tests, ncu launch list + full captures, bench lines, phase table.
# Outputs land in gpurun_out/; tools/summarize_profiles.py turns them into profiles/.
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
MPVSS_SKIP_PEAK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_horner_final python tools/profile_verify.py --n 4096 --t 2731 2>&1 | tail -3
MPVSS_SKIP_PEAK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_ec_horner_secp python bench.py --group secp256k1 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-300
MPVSS_SKIP_PEAK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_ec_horner_rist python bench.py --group ristretto255 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-400
timeout 300 python bench.py --group ristretto255 > gpurun_out/bench_rist.json 2> gpurun_out/bench_rist.err; tail -1 gpurun_out/bench_rist.json | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_reference.json | cut -c1-300
timeout 600 python tools/phase_timing.py > gpurun_out/phases.json 2> gpurun_out/phases.err; tail -c 600 gpurun_out/phases.json
