"""multi_exp (the reconstruct fold) by one exponentiation per base vs the bucket method, kernel ms per size."""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpvss_rs_b200 as m
from mpvss_rs_b200.participant import RFC3526_2048 as Q

g = m.Group("modp")
rng = random.Random(1)
out = {}
for k in (683, 2731, 10923, 43691):
    bases = g.fixed_base_exp([rng.randrange(Q - 1) for _ in range(k)])
    exps = [rng.randrange((Q - 1) // 2) for _ in range(k)]
    row = {}
    for mode, name in ((0, "direct"), (1, "buckets")):
        g.ctx.set_int("modp_msm", mode)
        res = []
        for rep in range(3):
            r = g.multi_exp(bases, exps)
            res.append(g.ctx.last_kernel_ms)
        row[name] = {"kernel_ms": min(res), "result_low64": r & (2**64 - 1)}
    assert row["direct"]["result_low64"] == row["buckets"]["result_low64"]
    out[k] = row
    print(k, row, file=sys.stderr)
print(json.dumps(out))
