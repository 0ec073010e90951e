"""Turns the raw ncu outputs under gpurun_out/ into the text summaries committed under profiles/."""
import csv, io, os, shutil, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
RND = sys.argv[1] if len(sys.argv) > 1 else "r02"

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct',
        'sm__inst_executed_pipe_tc', 'sm__inst_executed_pipe_tensor', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__average_warps_issue_stalled',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_thread_inst_executed_op_integer_pred_on.sum', 'sm__sass_thread_inst_executed_op_imad', 'local_load', 'local_store',
        'smsp__inst_executed_op_local']


def launches():
    src = os.path.join(G, f"launches_{RND}.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split('(')[0].replace('void ', '')
        agg[name][0] += 1
        agg[name][1] += float(r[-1]) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"launches_{RND}_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none --csv: MPVSS_SKIP_PEAK=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also\n")
        f.write("(includes the launches of the synthetic-box build and the correctness gate; cold-cache, serialised: compare SHARES)\n")
        f.write("%-62s %6s %12s %7s\n" % ("kernel", "count", "total ms", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-62s %6d %12.3f %6.1f%%\n" % (k[:62], v[0], v[1], 100 * v[1] / tot))
        f.write("\nlast timed step (mul_kernel = Montgomery conversion of the commitments, horner_kernel = X_i,\nexp2_kernel x2 = a2 = y^r Y^c and a1 = g^r X^c with the fixed-base table, frame_kernel = transcript rows):\n")
        own = [r for r in rows if 'modp::' in r[4] or 'ec::' in r[4]]
        for r in own[-5:]:
            f.write("  %-58s grid %-14s block %-12s %10.3f ms\n" % (r[4][:58], r[8], r[7], float(r[-1]) / 1e6))
    shutil.copy(src, os.path.join(P, f"launches_{RND}.csv"))


def full(rep, out, title):
    src = os.path.join(G, rep)
    if not os.path.exists(src):
        return
    txt = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    with open(os.path.join(P, out), "w") as f:
        f.write(title + "\n\n")
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) or k in h for k in WANT):
                if any(x in h for x in ['.max', '.min', '_elapsed', '.sum.p', 'pcsamp']) and 'dram' not in h:
                    continue
                f.write("%-100s %s %s\n" % (h, v, u))


launches()
full(f"prof_horner_{RND}.ncu-rep", f"horner_{RND}_ncu.txt",
     "ncu --set full --clock-control none --import-source on -k regex:horner_kernel -s 1 -c 1 python tools/profile_verify.py --n 4096 --t 2731\n"
     "modp::horner_kernel<8, true> at the bench configuration (n = 4096, t = 2731); selected raw metrics")
full(f"prof_ec_horner_secp_{RND}.ncu-rep", f"ec_horner_secp256k1_{RND}_ncu.txt",
     "ncu --set full --clock-control none -k regex:horner_kernel -c 1 python bench.py --group secp256k1 --steps 1 --warmup 0\n"
     "ec::horner_kernel<secp::SecpCurve> (n = 4096, t = 2731, 18 chunks: one wave); selected raw metrics")
full(f"prof_ec_horner_rist_{RND}.ncu-rep", f"ec_horner_ristretto255_{RND}_ncu.txt",
     "ncu --set full --clock-control none -k regex:horner_kernel -s 1 -c 1 python bench.py --group ristretto255 --steps 1 --warmup 1\n"
     "ec::horner_kernel<rist::RistCurve> (n = 4096, t = 2731, 18 chunks: one wave); selected raw metrics")
full(f"prof_ec_exp2_secp_{RND}.ncu-rep", f"ec_exp2_comb_secp256k1_{RND}_ncu.txt",
     "ncu --set full --clock-control none -k regex:exp2_comb_kernel -c 1 python bench.py --group secp256k1 --steps 1 --warmup 1\n"
     "ec::exp2_comb_kernel<secp::SecpCurve>: a1 = r*G + c*X with the generator table staged through shared memory; selected raw metrics")
print(open(os.path.join(P, f"launches_{RND}_summary.txt")).read())
