python -m pytest tests/test_gpu_modp.py tests/test_gpu_configs.py -m gpu -x -q -k "overlap or config2 or chunked or headline or medium" 2>&1 | tail -3
python bench.py --n 1024 --t 683 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_r02_c2_n1024.json 2>&1
python bench.py --no-also --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_s2i_default.json 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/bench_r02_c2_n1024.json','gpurun_out/bench_s2i_default.json'):
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(j['value']), round(j['ms_per_step'],2), round(j['kernel_ms_per_step'],2), round(j['e2e']['value']))
PY
