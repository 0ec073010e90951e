# Build first (from the repo root, after `make`):
#   for v in DEDICATED SPLIT; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC \
#     -DMPVSS_MODP_${v}_SQR -c mpvss_rs_b200/csrc/modp.cu -o variants/modp_$v.o && nvcc -shared -o variants/lib_$v.so \
#     mpvss_rs_b200/csrc/{api,modp_api,ec_api,ec,sha256_ni}.o variants/modp_$v.o; done
# Large-batch comparison of the three MODP squarings (variants/lib_{DEDICATED,SPLIT}.so are builds of
# modp.cu with -DMPVSS_MODP_DEDICATED_SQR / -DMPVSS_MODP_SPLIT_SQR), t kept small so a step stays short.
for n in 4096 32768; do
  for lib in "" variants/lib_DEDICATED.so variants/lib_SPLIT.so; do
    MPVSS_B200_LIB=${lib:+$PWD/$lib} MPVSS_SKIP_PEAK=1 timeout 300 python bench.py --n $n --t 342 --steps 3 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n=$n', '${lib:-default}', round(d['value']), 'horner_ms', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],3))"
  done
done
