set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python - <<'PY' > /tmp/standin.json
import json
from oracle import pvss
from oracle.groups import GROUPS
from mpvss_rs_b200 import synth
out={}
for gname in GROUPS:
    g=GROUPS[gname](); n,t=5,3
    bound = g.q if gname=="modp" else g.order()
    sks=synth.private_keys(3,n,gname,g.order(),bound)
    pks=[g.generate_public_key(s) for s in sks]
    secret=pvss.string_to_secret("Hello MPVSS Example.")
    box=pvss.distribute_secret(g,secret,pks,t,synth.coefficients(3,t,g.order()),synth.witnesses(3,n,bound))
    e=lambda x:g.element_to_bytes(x).hex(); s=lambda x:g.scalar_to_bytes(x).hex()
    ws=synth.witnesses(4,n,bound)
    sbs=[pvss.extract_secret_share(g,box,sks[i],ws[i]) for i in range(n)]
    keys=[g.element_to_bytes(pk) for pk in pks]
    hx=lambda v: (v.to_bytes(max(1,(v.bit_length()+7)//8),'big')).hex()
    out[gname]=dict(n=n,t=t,secret=hx(secret),U=hx(box.U),challenge=s(box.challenge),commitments=[e(c) for c in box.commitments],
      publickeys=[e(p) for p in pks],positions=[box.positions[k] for k in keys],shares=[e(box.shares[k]) for k in keys],
      responses=[s(box.responses[k]) for k in keys],private_keys=[s(x) for x in sks],extract_w=[s(x) for x in ws],
      sharebox_share=[e(b.share) for b in sbs],sharebox_challenge=[s(b.challenge) for b in sbs],sharebox_response=[s(b.response) for b in sbs],
      reconstructed=hx(pvss.reconstruct(g,sbs[:t],box)))
print(json.dumps(out))
PY
MPVSS_REF_VECTORS=/tmp/standin.json timeout 600 python -m pytest tests/test_ref_vectors.py -q 2>&1 | tail -3
B="timeout 600 python bench.py --no-cpu-baseline --no-also --steps 3"
for n in 2048 1024; do for tpi in 8 16; do $B --n $n --t 2731 --tpi $tpi > gpurun_out/b_n${n}_tpi${tpi}.json 2>&1; done; done
timeout 900 python tools/msm_timing.py > gpurun_out/msm_timing.json 2> gpurun_out/msm_timing.err; cat gpurun_out/msm_timing.json; tail -3 gpurun_out/msm_timing.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b_n*_tpi*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j['roofline']
        print(f, round(j['value']), 'ms',round(j['ms_per_step'],2),'horner',round(r['kernel_ms'],2),'frac',round(r['frac'],3))
    except Exception as e: print(f,'ERR',e, open(f).read()[-300:])
PY
