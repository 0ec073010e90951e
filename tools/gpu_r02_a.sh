# Round-2 first GPU pass: parity tests, default bench line, bank-conflict counters of the Horner launch.
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -c 3000 gpurun_out/bench_r02_a.json; tail -5 gpurun_out/bench_r02_a.err
MPVSS_SKIP_PEAK=1 timeout 600 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:horner_kernel -c 1 python tools/profile_verify.py --n 4096 --t 2731 2>&1 | tail -25
