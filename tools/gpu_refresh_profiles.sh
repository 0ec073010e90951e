set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_horner_final python tools/profile_verify.py --n 4096 --t 2731 2>&1 | tail -3
MPVSS_SKIP_PEAK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_ec_horner_secp python bench.py --group secp256k1 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -3
MPVSS_SKIP_PEAK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 -f -o gpurun_out/prof_ec_horner_rist python bench.py --group ristretto255 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -3
MPVSS_SKIP_PEAK=1 timeout 200 python bench.py --tpi 16 --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/bench_tpi16_final.json 2>gpurun_out/tpi16.err
tail -1 gpurun_out/bench_tpi16_final.json
