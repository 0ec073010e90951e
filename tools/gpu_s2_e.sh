# split Lagrange products (ModpGroup reconstruct): parity tests that reconstruct, then the phase figures
python -m pytest tests -m gpu -x -q -k "full_round or config2 or medium or bucket or cpp_mirror or device_side" 2>&1 | tail -5 > gpurun_out/gputests_s2e.txt
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_s2e.json 2> gpurun_out/bench_s2e.err
cat gpurun_out/gputests_s2e.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_s2e.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'])
p=j['also']['phases']; print({k:(round(v['wall_ms'],2),round(v['kernel_ms'],2)) for k,v in p.items() if isinstance(v,dict)})
PY
