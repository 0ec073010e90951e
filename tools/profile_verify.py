"""Small driver for ncu captures: stage one synthetic box and run verify_distribution a few times."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mpvss_rs_b200 as m
from mpvss_rs_b200.lib import buf, ptr

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--t", type=int, default=128)
ap.add_argument("--tpi", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
g = m.Group("modp")
if a.tpi:
    g.ctx.set_int("modp_tpi", a.tpi)
box = bench.build_box(g, a.n, a.t, 7)
ok = ctypes.c_int(0)
for _ in range(a.reps):
    g.ctx.check(g.ctx.lib.mpvss_verify_distribution(
        g.ctx.h, a.n, a.t, ptr(buf(box["commitments"])), None, ptr(buf(box["publickeys"])), ptr(buf(box["shares"])),
        ptr(buf(box["responses"])), ptr(buf(box["challenge"])), ctypes.byref(ok), None, None, None, None))
    assert ok.value == 1
    print("verify ok: kernel ms", g.ctx.last_kernel_ms, "phase0", g.ctx.last_phase_ms(0))
