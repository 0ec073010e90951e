#!/usr/bin/env python
"""bench.py -- verified DLEQ shares/sec on B200 (BASELINE.json metric).

A "step" is one pass of Participant::verify_distribution_shares (participant.rs:399-455)
over one synthetic DistributionSharesBox: for every participant recompute
X_i = prod_j C_j^(i^j) from the t commitments, a1 = g^r X^c, a2 = y^r Y^c, then hash the
framed transcript on the host and compare the challenge.  Workload at N=1: the headline
configuration of the metric, ModpGroup n=4096 t=2731.  With N GPUs every rank verifies a
contiguous slice of n participants of one box of N*n participants (weak scaling, t fixed);
the X/a1/a2 rows are combined with one NCCL all-gather and rank 0 hashes them.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n N] [--t T]

`--impl reference` times the reference's CPU schedule (oracle/cpu_baseline.c, an OpenSSL
proxy for num-bigint since the Rust reference cannot be built here) on all host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRICS = {"modp": "verified DLEQ shares/sec (MODP n=4096 t=2731, verify_distribution_shares)",
           "secp256k1": "verified DLEQ shares/sec (secp256k1 n=4096 t=2731, verify_distribution_shares)",
           "ristretto255": "verified DLEQ shares/sec (ristretto255 n=4096 t=2731, verify_distribution_shares)"}
CPU_PROXY = {"modp": "OpenSSL BN_mod_exp_mont (proxy for num-bigint 0.2)",
             "secp256k1": "OpenSSL EC_POINT_mul on secp256k1 (proxy for k256 0.13; OpenSSL has no specialised "
                          "secp256k1 code, so this under-states k256)"}
DTYPE = {"modp": "u32 limbs (2048-bit integers)", "secp256k1": "u32 limbs (256-bit prime fields)",
         "ristretto255": "u32 limbs (255-bit prime field)"}
SQR_MACS, MUL_MACS = 6240, 8256          # 2048-bit Montgomery sqr / mul, 32x32->64 MACs
Q = None


def env_int(k, d):
    return int(os.environ.get(k, d))


def le256(x):
    return int(x).to_bytes(256, "little")


# ------------------------------------------------------------------ clocks ----
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v == "Active":
                    reasons.add(name)
        load = [x for x in sm if x > 200]
        return {"sm_mhz": statistics.median(load) if load else (statistics.median(sm) if sm else None),
                "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------- workload ----
def build_box(group, n_total, t, seed):
    """Synthetic DistributionSharesBox made with the library's own dealer path.
    Returns flat boundary-encoded arrays in publickeys order."""
    from mpvss_rs_b200 import synth
    from mpvss_rs_b200.lib import buf, ptr
    c = group.codec
    eb, sb = c.eb, c.sb
    sks = synth.private_keys(seed, n_total, c.name, c.order, c.key_bound)
    coeffs = synth.coefficients(seed, t, c.order)
    ws = synth.witnesses(seed, n_total, c.key_bound)
    pks = group.fixed_base_exp(sks)
    pk_b = c.enc_elems(pks)
    comm, shares = buf(size=t * eb), buf(size=n_total * eb)
    chal, resp, u, x = buf(size=sb), buf(size=n_total * sb), buf(size=eb), buf(size=n_total * eb)
    secret = b"Hello MPVSS Example."
    group.ctx.check(group.ctx.lib.mpvss_distribute(
        group.ctx.h, n_total, t, ptr(buf(secret)), len(secret), ptr(buf(c.enc_scalars(coeffs))),
        ptr(buf(c.enc_scalars(ws))), ptr(buf(pk_b)), ptr(comm), ptr(shares), ptr(chal), ptr(resp), ptr(u), ptr(x)))
    return {"commitments": bytes(comm), "publickeys": pk_b, "shares": bytes(shares), "responses": bytes(resp),
            "challenge": bytes(chal), "x_dealer": bytes(x), "n": n_total, "t": t, "eb": eb, "sb": sb,
            "group": c.name}


def horner_macs(positions, t, tpi=8):
    """Algorithmic MACs of the X_i kernel for the given 1-based positions, following the schedule the
    kernel executes: per Horner step (2(d-1)+1) squarings and (d-1)+2
    multiplications, d = base-4 digits of the position, minus the window multiplications skipped
    because the digit is zero for every lane group of the warp (positions are dealt to warps in
    increasing order inside each digit class, exactly as modp_api.cu::prep_positions does)."""
    gpw = 32 // tpi
    by = {}
    for p in positions:
        d = 1
        while p >> (2 * d):
            d += 1
        by.setdefault(d, []).append(p)
    total = 0
    for d, ps in by.items():
        for c in range(0, len(ps), gpw):
            chunk = ps[c:c + gpw]
            full = chunk + [chunk[-1]] * (gpw - len(chunk))
            skipped = sum(1 for s_ in range(d - 1) if all(((q >> (2 * s_)) & 3) == 0 for q in full))
            total += len(chunk) * (t - 1) * ((2 * (d - 1) + 1) * SQR_MACS + ((d - 1) + 2 - skipped) * MUL_MACS)
    return total


def dleq_macs(n, rwin=512, cwin=64):
    """Algorithmic MACs of the two DLEQ commitment launches (4-bit fixed windows)."""
    def exp(w):
        return 4 * (w - 1) * SQR_MACS + (15 + (w - 1)) * MUL_MACS  # 14 table mults + to-Montgomery
    return 2 * n * (exp(rwin) + exp(cwin) + 2 * MUL_MACS)


def measure_imad_peak():
    """Measured 32-bit integer multiply-add issue peak of this GPU (tools/imad_peak.cu)."""
    exe = os.path.join(ROOT, "tools", "imad_peak")
    try:
        if os.environ.get("MPVSS_SKIP_PEAK"):
            raise RuntimeError("skipped (profiling run)")
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        j = json.loads(out)
        lo = max(v["tera_per_s"] for k, v in j.items() if k.startswith("imad_lo_"))
        wide = max(v["tera_per_s"] for k, v in j.items() if k.startswith("imad_wide"))
        return lo, wide, "measured live (tools/imad_peak)"
    except Exception:
        try:
            j = json.load(open(os.path.join(ROOT, "profiles", "imad_peak_r01.json")))
            lo = max(v["tera_per_s"] for k, v in j.items() if k.startswith("imad_lo_"))
            wide = max(v["tera_per_s"] for k, v in j.items() if k.startswith("imad_wide"))
            return lo, wide, "profiles/imad_peak_r01.json (measured on this pool)"
        except Exception:
            return 18.4, 7.7, "fallback constant (round-1 measurement)"


# ------------------------------------------------------------ CPU baseline ----
def cpu_reference_step(box, sample_idx, threads, schedule=0):
    """One bounded sample of the workload on the host cores; returns (seconds, X rows as boundary bytes)."""
    from oracle import cpu_baseline
    lib = cpu_baseline.load()
    s = len(sample_idx)
    t, eb, sb = box["t"], box["eb"], box["sb"]
    pos = (ctypes.c_int64 * s)(*[i + 1 for i in sample_idx])
    row = lambda key, w: [box[key][i * w:(i + 1) * w] for i in sample_idx]
    xo, a1o, a2o = (ctypes.create_string_buffer(eb * s) for _ in range(3))
    if box["group"] == "modp":
        from mpvss_rs_b200.participant import RFC3526_2048 as q
        be = lambda b: bytes(reversed(b))
        comm = b"".join(be(box["commitments"][j * 256:(j + 1) * 256]) for j in range(t))
        sel = lambda key: b"".join(be(r) for r in row(key, 256))
        t0 = time.perf_counter()
        lib.cpu_modp_verify(q.to_bytes(256, "big"), comm, t, pos, sel("publickeys"), sel("shares"), sel("responses"),
                            be(box["challenge"]), s, threads, schedule, xo, a1o, a2o)
        dt = time.perf_counter() - t0
        return dt, [bytes(reversed(xo.raw[i * 256:(i + 1) * 256])) for i in range(s)]
    if box["group"] == "secp256k1":
        t0 = time.perf_counter()
        lib.cpu_secp_verify(box["commitments"], t, pos, b"".join(row("publickeys", 33)), b"".join(row("shares", 33)),
                            b"".join(row("responses", 32)), box["challenge"], s, threads, schedule, xo, a1o, a2o)
        dt = time.perf_counter() - t0
        return dt, [xo.raw[i * 33:(i + 1) * 33] for i in range(s)]
    raise ValueError("no CPU baseline for " + box["group"])


def spread_sample(n, s):
    return sorted({min(n - 1, (k * n) // s + (n // (2 * s))) for k in range(s)})


# ------------------------------------------------------------------- main ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # long spellings too: under torchrun, `--n` / `--t` collide with abbreviations of its own options
    ap.add_argument("--n", "--participants-per-gpu", dest="n", type=int, default=4096, help="participants per GPU")
    ap.add_argument("--t", "--threshold", dest="t", type=int, default=0, help="threshold (default ceil(2n/3))")
    ap.add_argument("--tpi", type=int, default=0, help="override lanes per 2048-bit value")
    ap.add_argument("--seed", type=int, default=0x6D70767373)
    ap.add_argument("--group", default="modp", choices=["modp", "secp256k1", "ristretto255"])
    ap.add_argument("--dual", type=int, default=-1, help="override modp_dual (0/1/2)")
    ap.add_argument("--ec-threads", type=int, default=0)
    ap.add_argument("--overlap", type=int, default=-1, help="override modp_overlap (0, 2 or 3; default 3 = a2 as persistent one-warp CTAs in the idle warp slots)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary secp256k1 measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n = args.n
    t = args.t or -(-2 * n // 3)
    n_total = n * world
    cores = os.cpu_count() or 1

    import torch
    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mpvss_rs_b200 has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import mpvss_rs_b200 as m
    from mpvss_rs_b200.lib import buf, ptr
    METRIC = METRICS[args.group]
    group = m.Group(args.group, device=local)
    eb, sb = group.codec.eb, group.codec.sb
    if args.tpi and args.group == "modp":
        group.ctx.set_int("modp_tpi", args.tpi)
    if args.ec_threads and args.group != "modp":
        group.ctx.set_int("ec_threads", args.ec_threads)
    if args.overlap >= 0 and args.group == "modp":
        group.ctx.set_int("modp_overlap", args.overlap)
    if args.dual >= 0 and args.group == "modp":
        group.ctx.set_int("modp_dual", args.dual)
    lib, h = group.ctx.lib, group.ctx.h
    box = build_box(group, n_total, t, args.seed)
    config = {"workload": f"{args.group} verify_distribution_shares n={n} per GPU (box of {n_total}), t={t}",
              "group": args.group, "n_per_gpu": n, "n_total": n_total, "t": t,
              "x_schedule": "Horner in the exponent, fixed 2-bit windows (DESIGN.md)",
              "l2": "flushed between timed steps (256 MiB write)", "sharding": f"participants round-robin over {world} rank(s), one NCCL all-gather per step"}

    # ---------------------------------------------------------- reference arm ----
    if args.impl == "reference":
        sample = spread_sample(n_total, min(cores, n_total))
        for _ in range(args.warmup):
            cpu_reference_step(box, sample[: max(1, len(sample) // 4)], cores)
        times = [cpu_reference_step(box, sample, cores)[0] for _ in range(args.steps)]
        per = statistics.mean(times)
        val = len(sample) / per
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "shares/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[args.group],
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "shares/s", "cores": cores, "kind": "port",
                                 "sample": f"{len(sample)} participants per step at positions spread over "
                                           f"1..{n_total}, full t={t}, reference schedule (t+4 exponentiations per "
                                           "share, participant.rs:423-447) on " + CPU_PROXY[args.group]},
                "e2e": {"value": val, "unit": "shares/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # --------------------------------------------------------------- our arm ----
    # rank r takes positions r+1, r+1+N, ... (round robin): every rank sees the same mix of small and
    # large positions, so per-rank work is equal; the gathered rows are re-interleaved for the transcript
    from mpvss_rs_b200.sharding import interleave, shard_indices
    mine = shard_indices(rank, world, n_total)

    def sl(key):
        w = sb if key == "responses" else eb
        if world == 1:
            return box[key]
        return b"".join(box[key][i * w:(i + 1) * w] for i in mine)
    positions = (ctypes.c_int64 * n)(*[i + 1 for i in mine])
    pin = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).pin_memory()
    host = {k: pin(v) for k, v in (("commitments", box["commitments"]), ("publickeys", sl("publickeys")),
                                   ("shares", sl("shares")), ("responses", sl("responses")),
                                   ("challenge", box["challenge"]))}
    P = lambda tns: ctypes.cast(tns.data_ptr(), ctypes.POINTER(ctypes.c_uint8))
    ok = ctypes.c_int(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out_local = torch.empty((3, n, eb), dtype=torch.uint8, device="cuda")
    out_all = torch.empty((world, 3, n, eb), dtype=torch.uint8, device="cuda") if world > 1 else None
    out_host = torch.empty((world, 3, n, eb), dtype=torch.uint8).pin_memory() if world > 1 else None
    x_chk = buf(size=n * eb)
    shares_all, chal_all = buf(box["shares"]), buf(box["challenge"])
    import numpy as np

    PH = 2 if args.group == "modp" else 0   # phase index of the dominant (Horner) launch

    def stage():
        group.ctx.check(lib.mpvss_verify_distribution_stage(
            h, n, t, P(host["commitments"]), positions, P(host["publickeys"]), P(host["shares"]),
            P(host["responses"]), P(host["challenge"])))

    def run_resident(want_x=False):
        """one step with the box resident in HBM; returns (ok, kernel_ms, Horner-launch ms)"""
        if world == 1:
            group.ctx.check(lib.mpvss_verify_distribution_run(h, ctypes.byref(ok), ptr(x_chk) if want_x else None,
                                                              None, None, None))
            return ok.value, group.ctx.last_kernel_ms, group.ctx.last_phase_ms(PH)
        group.ctx.check(lib.mpvss_verify_distribution_compute(
            h, out_local[0].data_ptr(), out_local[1].data_ptr(), out_local[2].data_ptr()))
        kms, p0 = group.ctx.last_kernel_ms, group.ctx.last_phase_ms(PH)
        dist.all_gather_into_tensor(out_all, out_local)          # one NCCL all-gather per phase
        res = 1
        if rank == 0:
            # [rank, kind, j, eb] -> [kind, j, rank, eb] (participant j*N + rank) on the device, then one
            # contiguous D2H copy: the re-interleave is a 25 MB permute at HBM speed instead of a host pass
            out_host.copy_(out_all.permute(1, 2, 0, 3).contiguous().view(world, 3, n, eb), non_blocking=False)
            arr = out_host.numpy().reshape(3, n, world, eb)
            u8 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
            group.ctx.check(lib.mpvss_transcript_check(h, n_total, u8(arr[0]), ptr(shares_all), u8(arr[1]), u8(arr[2]),
                                                       ptr(chal_all), ctypes.byref(ok), None))
            res = ok.value
        return res, kms, p0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(x):
        if dist is None:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    stage()
    # correctness gate before timing: the box verifies and the verifier's X equals the dealer's g^P(i)
    okv, _, _ = run_resident(want_x=(world == 1))
    if rank == 0 and okv != 1:
        raise SystemExit("verification of the synthetic box failed -- refusing to report a number")
    if world == 1 and bytes(x_chk) != box["x_dealer"]:
        raise SystemExit("verifier X_i differs from dealer X_i -- refusing to report a number")

    imad_lo, imad_wide, peak_src = (measure_imad_peak() if rank == 0 and args.group == "modp" else (0, 0, ""))
    for _ in range(args.warmup):
        run_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_s, kern_ms, p0_ms, launches = [], [], [], 0
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        okv, kms, p0 = run_resident()
        barrier()
        step_s.append(maxr(time.perf_counter() - t0))
        kern_ms.append(maxr(kms))
        p0_ms.append(maxr(p0))
        launches += group.ctx.last_kernel_launches
        assert rank != 0 or okv == 1
    # end to end through the reference-facing call, host buffers, copies inside the timed region
    e2e_s = []
    if world == 1:
        for i in range(args.warmup + args.steps):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            group.ctx.check(lib.mpvss_verify_distribution(
                h, n, t, P(host["commitments"]), positions, P(host["publickeys"]), P(host["shares"]),
                P(host["responses"]), P(host["challenge"]), ctypes.byref(ok), None, None, None, None))
            barrier()
            if i >= args.warmup:
                e2e_s.append(time.perf_counter() - t0)
            assert ok.value == 1
    else:
        for i in range(args.warmup + args.steps):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            stage()
            okv, _, _ = run_resident()
            barrier()
            if i >= args.warmup:
                e2e_s.append(maxr(time.perf_counter() - t0))
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        dist.destroy_process_group()
        return 0

    ms = statistics.mean(step_s) * 1e3
    value = n_total / (ms * 1e-3)
    e2e_ms = statistics.mean(e2e_s) * 1e3
    h2d = t * eb + 2 * n * eb + n * sb + sb + 8 * n
    d2h = 3 * n * eb
    p0 = statistics.mean(p0_ms)
    line = {
        "metric": METRIC, "value": value, "unit": "shares/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE[args.group], "data": "synthetic", "config": config,
        "timing": "per step: cuda-synchronize + barrier bracketed wall clock (kernels + D2H + host SHA-256), "
                  "max over ranks; kernel_ms / roofline from CUDA events on the library's stream",
        "kernel_ms_per_step": statistics.mean(kern_ms),
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "shares/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "mpvss_verify_distribution (pinned host buffers in, verdict out)"},
    }
    if args.group == "modp":
        hm = horner_macs([i + 1 for i in mine], t, args.tpi or (4 if n >= 32768 else 8))
        achieved = 2.0 * hm / (p0 * 1e-3) / 1e12        # TIMAD/s, 1 MAC = 2 IMAD issues
        dual = (args.dual > 0) and t >= 8   # library default: single chain (modp_dual = 0)
        combine = n * (4 * 511 * SQR_MACS + (15 + 511 + 15 + 2) * MUL_MACS) if dual else 0
        total_macs = hm + combine + dleq_macs(n)
        step_timad = 2.0 * total_macs / (statistics.mean(kern_ms) * 1e-3) / 1e12
        line["roofline"] = {
            "bound": "imad", "kernel": "modp::horner_kernel (X_i multi-exponentiation)",
            "achieved": achieved, "peak": imad_lo, "unit": "TIMAD/s", "frac": achieved / imad_lo if imad_lo else None,
            "peak_source": peak_src + "; 32-bit IMAD issue rate, 1 MAC (32x32->64) = 2 IMAD",
            "achieved_tmac_per_s": achieved / 2, "wide_mac_peak_tmac_per_s": imad_wide,
            "frac_of_wide_mac_peak": (achieved / 2) / imad_wide if imad_wide else None,
            "algorithmic_macs_per_launch": hm, "kernel_ms": p0,
            "share_of_step_macs": hm / total_macs, "traffic": 976640,
            "traffic_note": "dram bytes read+written by the Horner launch, ncu --set full, profiles/horner_r01_ncu.txt",
            "note": "kernel_ms = CUDA events around the Horner launch on the library stream; algorithmic MACs "
                    "follow the executed fixed-window schedule, skipped zero-digit multiplications excluded",
            "whole_step": {"achieved": step_timad, "frac": step_timad / imad_lo if imad_lo else None,
                           "algorithmic_macs": total_macs, "kernel_ms": statistics.mean(kern_ms),
                           "what": "all kernels of the step (Horner + chunk combination + both DLEQ launches)"}}
    else:
        line["roofline"] = None
        line["phase_ms"] = {"x_horner": p0, "dleq": statistics.mean(kern_ms) - p0}
    if not args.no_cpu_baseline and world == 1 and args.group in CPU_PROXY:
        sample = spread_sample(n_total, min(cores, n_total))
        dt, xs = cpu_reference_step(box, sample, cores)
        for i, x in zip(sample, xs):       # the CPU restatement and the GPU agree on X_i
            assert x == bytes(x_chk)[i * eb:(i + 1) * eb], "CPU baseline X_i != GPU X_i"
        dt_h, _ = cpu_reference_step(box, sample, cores, schedule=1)
        line["cpu_baseline"] = {
            "value": len(sample) / dt, "unit": "shares/s", "cores": cores, "kind": "port",
            "sample": f"{len(sample)} participants (positions spread over 1..{n_total}), full t={t}, reference "
                      "schedule (t+4 full exponentiations per share) on " + CPU_PROXY[args.group] +
                      ", one participant per thread",
            "same_algorithm_value": len(sample) / dt_h,
            "same_algorithm_note": "CPU running the GPU's Horner schedule (baseline B, BASELINE.md section 3)"}
    if args.group == "modp" and world == 1 and not args.no_also:
        # the metric names MODP + secp256k1: the same step for Secp256k1Group, reported alongside
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", "secp256k1", "--steps",
                                  str(args.steps), "--warmup", str(args.warmup), "--n", str(n), "--t", str(t)] +
                                 (["--no-cpu-baseline"] if args.no_cpu_baseline else []),
                                 capture_output=True, text=True, timeout=900)
            sec = json.loads(out.stdout.strip().splitlines()[-1])
            line["also"] = {"secp256k1": {k: sec.get(k) for k in ("metric", "value", "unit", "ms_per_step",
                                                                 "kernel_ms_per_step", "e2e", "cpu_baseline",
                                                                 "phase_ms", "gpu_launches")}}
        except Exception as ex:  # the primary line stands on its own
            line["also"] = {"secp256k1": {"error": str(ex)[:200]}}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
