# Horner launch: a2 filler on/off at exact multiples of 4 warps per SM, and warps per CTA
for n in 4096 4736 9472; do for o in -1 0; do
  timeout 600 python bench.py --no-cpu-baseline --no-also --steps 2 --warmup 1 --n $n --t 2731 --overlap $o > gpurun_out/ov_${n}_$o.json 2>&1
done; done
python - <<'PY'
import json
for n in (4096,4736,9472):
  for o in (-1,0):
    try:
        j=json.loads(open(f'gpurun_out/ov_{n}_{o}.json').read().strip().splitlines()[-1]); r=j['roofline']
        print(n, 'overlap', 'auto' if o<0 else o, 'horner ms', round(r['kernel_ms'],2), 'frac', round(r['frac'],3), 'kernels', round(j['kernel_ms_per_step'],1), 'step', round(j['ms_per_step'],1), 'shares/s', round(j['value']))
    except Exception as e: print(n,o,'ERR',e, open(f'gpurun_out/ov_{n}_{o}.json').read()[-200:])
PY
