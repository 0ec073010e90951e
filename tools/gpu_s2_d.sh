for r in 1 2; do
  python bench.py --no-also --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab3_uni_$r.json 2>&1
  MPVSS_HORNER_ARRAYS=1 python bench.py --no-also --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab3_arr_$r.json 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab3_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j['ms_per_step'], j['kernel_ms_per_step'], j['roofline']['kernel_ms'])
    except Exception as e: print(f,'ERR',open(f).read()[-300:])
PY
