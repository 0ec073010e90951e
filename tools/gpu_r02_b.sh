set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="timeout 600 python bench.py --no-cpu-baseline"
$B > gpurun_out/b_default.json 2> gpurun_out/b_default.err; tail -3 gpurun_out/b_default.err
MPVSS_B200_LIB=$PWD/variants/libmpvss_u1.so $B --no-also > gpurun_out/b_u1.json 2>&1
for g in secp256k1 ristretto255; do
  $B --group $g > gpurun_out/b_${g}.json 2>&1
  MPVSS_B200_LIB=$PWD/variants/libmpvss_mb3.so $B --group $g > gpurun_out/b_${g}_mb3.json 2>&1
  $B --group $g --ec-threads 131072 > gpurun_out/b_${g}_t131072.json 2>&1
  $B --group $g --ec-threads 151552 > gpurun_out/b_${g}_t151552.json 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j.get('roofline') or {}
        print(f, round(j['value']), 'ms',round(j['ms_per_step'],2),'kern',round(j['kernel_ms_per_step'],2),'horner',round(r.get('kernel_ms',0),2),'frac',round(r.get('frac') or 0,3),'e2e',round(j['e2e']['value']))
        if 'also' in j and 'secp256k1' in j['also']: 
            s=j['also']['secp256k1']; print('   also secp', round(s['value']), s['roofline']['frac'])
    except Exception as e: print(f,'ERR',e, open(f).read()[-300:])
PY
