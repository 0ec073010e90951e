# small box (BASELINE config 2, n = 1024, t = 683): where should a2 run?  0 = before the X_i launch, 2 = beside it (regular launch on a side stream), 3 = persistent filler
for o in -1 0 2 3; do
  python bench.py --n 1024 --t 683 --no-also --no-cpu-baseline --steps 5 --warmup 3 --overlap $o > gpurun_out/ov_1024_$o.json 2>&1
  python bench.py --n 2048 --t 1366 --no-also --no-cpu-baseline --steps 5 --warmup 3 --overlap $o > gpurun_out/ov_2048_$o.json 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ov_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(j['ms_per_step'],2), round(j['kernel_ms_per_step'],2), round(j['roofline']['kernel_ms'],2))
    except Exception as e: print(f,'ERR',open(f).read()[-300:])
PY
