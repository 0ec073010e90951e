# Round-end refresh on one B200: GPU tests, ncu launch list + full captures, bench lines, phase table.
# Outputs land in gpurun_out/; tools/summarize_profiles.py r02 turns them into profiles/.
set -x
R=r02
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
MPVSS_SKIP_PEAK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_horner_$R python tools/profile_verify.py --n 4096 --t 2731 2>&1 | tail -3
MPVSS_SKIP_PEAK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_ec_horner_secp_$R python bench.py --group secp256k1 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-300
MPVSS_SKIP_PEAK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_ec_horner_rist_$R python bench.py --group ristretto255 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-300
MPVSS_SKIP_PEAK=1 timeout 300 ncu --set full --clock-control none -k regex:exp2_comb_kernel -c 1 -f -o gpurun_out/prof_ec_exp2_secp_$R python bench.py --group secp256k1 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-200
tools/imad_peak > gpurun_out/imad_peak_$R.json
timeout 600 python bench.py > gpurun_out/bench_${R}_default.json 2> gpurun_out/bench_${R}_default.err; tail -1 gpurun_out/bench_${R}_default.json | cut -c1-400
timeout 300 python bench.py --group ristretto255 > gpurun_out/bench_${R}_ristretto255.json 2> gpurun_out/bench_${R}_rist.err; tail -1 gpurun_out/bench_${R}_ristretto255.json | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_${R}_ref.err; tail -1 gpurun_out/bench_${R}_reference_arm.json | cut -c1-300
timeout 900 python tools/phase_timing.py > gpurun_out/phases_$R.jsonl 2> gpurun_out/phases.err; tail -c 600 gpurun_out/phases_$R.jsonl
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck_$R.txt 2>&1; tail -3 gpurun_out/sanitizer_memcheck_$R.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/sanitizer_racecheck_$R.txt 2>&1; tail -3 gpurun_out/sanitizer_racecheck_$R.txt
