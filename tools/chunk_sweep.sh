# X_i launch with K chunks per position at small position counts (what one GPU sees when a box is split over
# 2 / 4 / 8 GPUs), t = 2731; and the small box n = 1024, t = 683
for m in 2048 1024 512; do for k in 1 2 4 8; do
  timeout 300 python bench.py --no-cpu-baseline --no-also --steps 2 --warmup 1 --subset $m --chunks $k > gpurun_out/ck_${m}_$k.json 2>&1
done; done
for k in 1 2 4; do timeout 300 python bench.py --no-cpu-baseline --no-also --steps 2 --warmup 1 --n 1024 --t 683 --subset 1024 --chunks $k > gpurun_out/ck_s1024_$k.json 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ck_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'kernel', round(j['kernel_ms'],1), 'horner', round(j['horner_ms'],1), 'step', round(j['ms'],1))
    except Exception as e: print(f,'ERR',open(f).read()[-300:])
PY
