# a2 placement (modp_overlap 0 / 3) at several box sizes
for cfg in "1024 683" "2048 683" "4096 683" "8192 683"; do set -- $cfg; for ov in 0 3; do  # 3 = adaptive (falls back to 0 when the filler would outlast the Horner launch)
  MPVSS_SKIP_PEAK=1 timeout 300 python bench.py --n $1 --t $2 --overlap $ov --steps 4 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n', $1, 't', $2, 'overlap', $ov, round(d['value']), 'step_ms', round(d['ms_per_step'],1), 'horner_ms', round(d['roofline']['kernel_ms'],1), 'kernels', round(d['kernel_ms_per_step'],1))"
done; done
