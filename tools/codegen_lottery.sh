# X_i launch with the product loop unrolled 1 / 2 (default) / 3 / 6 periods per trip, and ptxas with expensive
# optimisations allowed; same box, one short run each (variants built with tools/build_variant.sh-style commands:
# nvcc ... -DMPVSS_MODP_UNROLL=k -c csrc/modp.cu, linked with the other objects)
for v in "" u1 u3 u6 xp ""; do
  MPVSS_SKIP_PEAK=1 MPVSS_B200_LIB=${v:+$PWD/variants/libmpvss_$v.so} python bench.py --no-also --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${v:-default}', round(d['ms_per_step'],2), 'horner_ms', round(d['roofline']['kernel_ms'],2))"
done
