set -x
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_modp.py -m gpu -x -q -k "two_gpu or bucket" 2>&1 | tail -5
timeout 900 python tools/msm_timing.py > gpurun_out/msm_timing.json 2> gpurun_out/msm_timing.err; cat gpurun_out/msm_timing.json; tail -2 gpurun_out/msm_timing.err
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline"
$T > gpurun_out/b2_auto.json 2> gpurun_out/b2_auto.err; tail -3 gpurun_out/b2_auto.err
$T --tpi 8 > gpurun_out/b2_tpi8.json 2> gpurun_out/b2_tpi8.err
$T --group secp256k1 > gpurun_out/b2_secp.json 2> gpurun_out/b2_secp.err; tail -3 gpurun_out/b2_secp.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b2_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['value']), 'ms',round(j['ms_per_step'],2),'kern',round(j['kernel_ms_per_step'],2),'tail',round(j['host_tail_ms'],2),'e2e',round(j['e2e']['value']), 'strong', {k:(round(v,1) if isinstance(v,float) else v) for k,v in j.get('also',{}).get('strong',{}).items() if k in ('value','ms_per_step','kernel_ms_per_step')})
    except Exception as e: print(f,'ERR',e, open(f).read()[-300:])
PY
