for v in default inl; do
  if [ $v = default ]; then L=""; else L="MPVSS_B200_LIB=$PWD/variants/libmpvss_$v.so"; fi
  env $L timeout 600 python bench.py --group secp256k1 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ecv_$v.json 2>&1
done
python - <<'PY'
import json
for v in ('default','inl'):
    try:
        j=json.loads(open(f'gpurun_out/ecv_{v}.json').read().strip().splitlines()[-1]); r=j['roofline']
        print(v, round(j['value']), 'ms', round(j['ms_per_step'],2), 'horner', round(r['kernel_ms'],2), 'frac', round(r['frac'],3))
    except Exception as e: print(v,'ERR',e,open(f'gpurun_out/ecv_{v}.json').read()[-300:])
PY
