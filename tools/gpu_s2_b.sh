# device-side per-share transcripts + the unchunked Horner instantiation: parity suite, A/B against the pre-chunk build
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests_s2b.txt
python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_s2b.json 2> gpurun_out/bench_s2b.err
MPVSS_B200_LIB=variants/libmpvss_prechunk.so python bench.py --no-also --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab2_pre.json 2>&1
python bench.py --group ristretto255 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_s2b_rist.json 2> gpurun_out/bench_s2b_rist.err
tail -4 gpurun_out/gputests_s2b.txt
python - <<'PY'
import json
for f in ('gpurun_out/bench_s2b.json','gpurun_out/ab2_pre.json','gpurun_out/bench_s2b_rist.json'):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j['value'], j['ms_per_step'], j['kernel_ms_per_step'], j['roofline']['kernel_ms'])
        for g,a in (('self',j.get('also',{})),('secp',(j.get('also',{}).get('secp256k1') or {}))):
            p=a.get('phases')
            if p: print('  ',g,{k:(round(v['wall_ms'],2),round(v['kernel_ms'],2)) for k,v in p.items() if isinstance(v,dict)})
    except Exception as e: print(f,'ERR',e,open(f).read()[-400:])
PY
