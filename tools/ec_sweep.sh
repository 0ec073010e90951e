# tuning sweep: EC Horner occupancy variants (variants/lib_mb*.so) x chunk-thread targets
for lib in "" variants/lib_mb5.so variants/lib_mb6.so variants/lib_mb8.so; do
  for et in 65536 98304 131072; do
    for g in secp256k1 ristretto255; do
      MPVSS_B200_LIB=${lib:+$PWD/$lib} MPVSS_SKIP_PEAK=1 timeout 200 python bench.py --group $g --ec-threads $et --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${lib:-default}', $et, '$g', round(d['value']), d['phase_ms'])"
    done
  done
done
