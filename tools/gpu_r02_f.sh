set -x
N=$1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -3 gpurun_out/bench_r02_n$N.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().splitlines()[-1])
print(round(j['value']), 'ms',round(j['ms_per_step'],2),'kern',round(j['kernel_ms_per_step'],2),'tail',round(j['host_tail_ms'],2),'e2e',round(j['e2e']['value']), j['config'].get('digest_check'))
for k,v in j.get('also',{}).items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','kernel_ms_per_step','distribute_wall_s')}, (v.get('roofline') or {}).get('frac'))
PY
