# Horner launch time against warps per scheduler (n positions = n/4 one-warp CTAs on 148 x 4 schedulers)
for n in 4096 4736 7104 9472 14208; do
  timeout 600 python bench.py --no-cpu-baseline --no-also --steps 3 --warmup 2 --n $n --t 2731 > gpurun_out/occ_$n.json 2>&1
done
python - <<'PY'
import json
for n in (4096,4736,7104,9472,14208):
    try:
        j=json.loads(open(f'gpurun_out/occ_{n}.json').read().strip().splitlines()[-1]); r=j['roofline']
        print(n, 'warps/SM', round(n/4/148,2), 'horner ms', round(r['kernel_ms'],2), 'us per position', round(1e3*r['kernel_ms']/n,2), 'frac', round(r['frac'],3), 'shares/s', round(j['value']))
    except Exception as e: print(n,'ERR',e)
PY
