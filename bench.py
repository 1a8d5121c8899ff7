#!/usr/bin/env python
"""bench.py -- 128by64 aggregated reciprocal range proofs proved+verified per second (BASELINE.json).

A "step" is one pass of the hot path over one batch: `batch` independent 128by64 proofs
(examples/128by64: 128 x 64-bit values, base-256 shared digits, norm argument; N = 1024,
M = 261, 9 rounds) PROVED and then VERIFIED through the C ABI (bppp_rp_prove_batch +
bppp_rp_verify_batch): host C++ threads run the Fiat-Shamir transcript and the scalar phases,
every group operation runs in the sm_100a kernels.  Proof b uses randomSeed
"default random seed#b" and values 10000 + b, so transcripts differ.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): every rank proves+verifies its own batch
(weak scaling, no data-path collective; NCCL only for the barrier and the max over ranks).
`--impl reference` times the reference's own CPU algorithm (oracle port, oracle/ + oracle/c) on
all host cores instead.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "128by64 range proofs proved+verified/sec"
UNIT = "proofs/s"
R_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


def workload_schema():
    from bulletproofspp_b200.workloads import EXAMPLES
    return EXAMPLES["128by64"][0]


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [c for c in sm if c >= 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm
def _ref_worker(args, pippenger=False):
    idx, n = args
    sys.path.insert(0, ROOT)
    from bulletproofspp_b200.workloads import batched
    from oracle.curve import SecpPip, SecpRef
    from oracle.rangeproof import load_schema, load_witness, prove, verify
    from oracle.transcript import ZKPT
    Grp = SecpPip if pippenger else SecpRef
    schema, wits, seeds = batched("128by64", idx + n)
    ok = True
    for b in range(idx, idx + n):
        s = load_schema(dict(schema, randomSeed=seeds[b]), Grp, points=_ref_worker.points)
        proof = prove(s, ZKPT(Grp, s.random_seed), load_witness(s, wits[b]))
        ok = ok and verify(s, ZKPT(Grp, None), proof)
    return ok


def _ref_init():
    sys.path.insert(0, ROOT)
    from oracle.curve import SecpRef
    from oracle.transcript import get_points
    _ref_worker.points = get_points(SecpRef, "test points", 1300)
    SecpRef.lib()


def run_reference(args, rank, world):
    """The reference's CPU algorithm as written (256-row Straus `innerProduct`, 129-row
    `projectivePairIP`, zero-padded openings) via the oracle port, one proof per task on all
    host cores.  The Haskell reference itself cannot be built here (no GHC)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = cores                      # one 128by64 prove+verify per core per step (~2 s each)
    with mp.Pool(cores, initializer=_ref_init) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, [(i, 1) for i in range(min(cores, 2))])
        t0 = time.time()
        for k in range(args.steps):
            oks = pool.map(_ref_worker, [(k * per_step + i, 1) for i in range(per_step)])
            assert all(oks)
        dt = time.time() - t0
    value = args.steps * per_step / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u256 (Fq/Fr integers)", "data": "synthetic",
            "config": {"workload": "examples/128by64 prove+verify, %d proofs per step (one per host core)" % per_step,
                       "batch": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x %d proofs of 128by64 prove+verify; oracle port of the reference "
                                       "algorithm (Python scalar phases + C group law), one process per core" % (args.steps, per_step)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(n_proofs=4):
    """Rank-0, N=1 only: the oracle port on ONE core (the reference runs its hot path on one core:
    parallel strategies are commented out, NormArgument.hs:71,129), bounded sample."""
    _ref_init()
    t0 = time.time()
    assert _ref_worker((0, n_proofs))
    dt = time.time() - t0
    t1 = time.time()
    assert _ref_worker((0, n_proofs), pippenger=True)
    dt2 = time.time() - t1
    return {"value": n_proofs / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d proofs of 128by64 prove+verify, oracle port of the reference algorithm "
                      "(256-row Straus MSM, 129-row pair folds, zero-padded openings), single thread" % n_proofs,
            "pippenger": {"value": n_proofs / dt2, "unit": UNIT, "cores": 1,
                          "sample": "the same %d proofs with an honest CPU MSM (Pippenger, signed windows, no padding; "
                                    "oracle/c/ref_ec.c pip_msm) in place of the reference's Straus loop" % n_proofs}}


# --------------------------------------------------------------------------- our arm
def make_inputs(batch, offset, n_inputs):
    """host buffers of one step: values (10000 + proof index), types (0), per-proof seeds"""
    import numpy as np
    v = np.zeros((batch, n_inputs, 4), dtype="<u8")                 # 32-byte little-endian scalars
    v[:, :, 0] = (10000 + offset + np.arange(batch, dtype=np.uint64))[:, None]
    vals = v.tobytes()
    tys = bytes(32 * n_inputs * batch)
    seeds = ["default random seed#%d" % (offset + b) for b in range(batch)]
    return vals, tys, seeds


def ncu_summary():
    """per-kernel numbers that only a profiler can see (DRAM bytes per launch, pipe-active percentage), read from the
    summary of the committed `ncu --set full` captures of this same command (profiles/ncu_summary.json; the raw
    CSVs sit next to it).  bench.py cannot run under ncu itself: a number printed under a profiler is not a bench value."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="proofs per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-threads", type=int, default=0)
    ap.add_argument("--sweep-sizes", default="10,14,18,20", help="log2 N of the synthetic norm-argument sweep (N=1 only)")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--sharded-size", type=int, default=20, help="log2 N of the argument sharded over the GPUs (N > 1 only)")
    ap.add_argument("--provers", type=int, default=2,
                    help="prover setups whose bppp_rp_prove_batch calls overlap (consecutive batches in flight: the ramp-up of one "
                         "call's lanes -- host phases first -- fills the ramp-down of the previous call's)")
    ap.add_argument("--lut-gb", type=float, default=48.0,
                    help="memory budget of the generators' full-multiples table (csrc/lut.cuh); 0 = nine-bit bucket kernel")
    ap.add_argument("--verify", default="batch", choices=["batch", "per-proof"],
                    help="batch: one random linear combination per lane sub-batch, per-proof checks only on failure (exact verdicts)")
    ap.add_argument("--transcript", default="device", choices=["device", "host"],
                    help="where the Fiat-Shamir transcript of the batch prover / verifier runs (bit-identical proofs either way)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import bulletproofspp_b200 as bp
    ctx = bp.Context(local)             # before torch creates the primary context: bppp_init asks for blocking syncs
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.host_threads or world > 1:
        # ranks share the host: split the cores between them
        nt = args.host_threads or max(1, (os.cpu_count() or 1) // world)
        ctx.lib.bppp_set_host_threads(nt)
    # prover and verifier are different parties: each gets its own setup (generator tables, lanes,
    # streams), and consecutive batches are pipelined -- batch k is verified while batch k+1 is proved
    setup = bp.RangeProofSetup(ctx, workload_schema())
    vctx = bp.Context(local)
    vsetup = bp.RangeProofSetup(vctx, workload_schema())
    dev_tr = args.transcript == "device"
    setup.set_device_transcript(dev_tr)
    vsetup.set_device_transcript(dev_tr)
    vsetup.set_batch_verify(args.verify == "batch")
    t_lut = time.time()
    lut_c, lut_err = 0, None
    if args.lut_gb > 0:
        try:                                    # a box without the memory keeps the nine-bit bucket kernel (same results)
            lut_c = setup.enable_lut(args.lut_gb)
            if lut_c:
                vsetup.enable_lut(args.lut_gb)
        except bp.BpppError as ex:
            lut_c, lut_err = 0, str(ex)
    t_lut = time.time() - t_lut
    # further prover setups (own lanes, streams and staging buffers; the worker pool and the table are shared)
    psetups = [setup]
    for _ in range(max(1, args.provers) - 1):
        s2 = bp.RangeProofSetup(bp.Context(local), workload_schema())
        s2.set_device_transcript(dev_tr)
        if lut_c:
            s2.enable_lut(args.lut_gb)
        psetups.append(s2)
    lanes = [c for st in psetups for c in st.contexts()] + vsetup.contexts()
    B, n = args.batch, setup.n_inputs
    assert (setup.nrm_len, setup.lin_len, setup.rounds) == (1024, 261, 9)

    def merged_report():
        tot = {"kernels": {}, "h2d_bytes": 0, "d2h_bytes": 0}
        for c in lanes:
            r = c.profile_report()
            tot["h2d_bytes"] += r["h2d_bytes"]
            tot["d2h_bytes"] += r["d2h_bytes"]
            for name, kk in r["kernels"].items():
                d = tot["kernels"].setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
                for f in d:
                    d[f] += kk[f]
        return tot

    def barrier():
        for c in lanes:
            c.sync()
        if world > 1:
            dist.barrier()

    prove_wall = []                     # wall seconds of every bppp_rp_prove_batch call, all legs (diagnostic)

    def run_steps(input_fn, steps):
        """prove `steps` batches; a second host thread verifies each batch as soon as it is proved.
        Returns the number of proofs that verified."""
        import queue
        q, ok_count, errs = queue.Queue(maxsize=2), [0], []

        def verifier():
            while True:
                item = q.get()
                if item is None:
                    return
                try:
                    ok_count[0] += sum(vsetup.verify_batch_raw(B, *item))
                except Exception as ex:          # surfaced after join
                    errs.append(ex)
        th = threading.Thread(target=verifier)
        th.start()
        nxt, lock = [0], threading.Lock()

        def prover(st):
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= steps:
                    return
                try:
                    vals, tys, seeds = input_fn(k)
                    tk = time.time()
                    proof = st.prove_batch_raw(B, vals, tys, None, seeds)
                    prove_wall.append(round(time.time() - tk, 4))
                    q.put(proof)
                except Exception as ex:
                    errs.append(ex)
                    return
        pth = [threading.Thread(target=prover, args=(st,)) for st in psetups[:max(1, min(len(psetups), steps))]]
        for t in pth:
            t.start()
        for t in pth:
            t.join()
        q.put(None)
        th.join()
        if errs:
            raise errs[0]
        return ok_count[0]

    base = rank * B
    inputs = make_inputs(B, base, n)
    # the timed `value` leg proves DIFFERENT proofs every step; all of them are staged before the clock starts
    staged = [make_inputs(B, base + (1000 + k) * world * B, n) for k in range(args.steps)]
    assert run_steps(lambda k: inputs, args.warmup) == B * args.warmup, "a warm-up proof failed to verify"
    # pools, pinned staging buffers and worker threads are sized on demand during the first steps; on a
    # box that just ran something else three steps are sometimes not enough.  Keep warming up (untimed,
    # at most 4 more steps) until a step is within 10 % of the one before it.
    extra_warmup = 0
    while extra_warmup < 4 and len(prove_wall) >= 2 and prove_wall[-1] > 0 and abs(prove_wall[-1] - prove_wall[-2]) > 0.10 * prove_wall[-1]:
        assert run_steps(lambda k: inputs, 1) == B
        extra_warmup += 1
    # ---- timed region 1: `value` -- inputs staged before the clock starts, device-timed
    for c in lanes:
        c.profile_reset()               # zero the H2D / D2H byte counters
    launches0 = sum(c.launch_count() for c in lanes)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ctx.timer_start()
    t0 = time.time()
    cpu0 = os.times()
    n_ok = run_steps(lambda k: staged[k], args.steps)
    for c in lanes:
        c.sync()
    ms = ctx.timer_stop()
    wall = time.time() - t0
    cpu1 = os.times()
    host_cpu_s = (cpu1.user - cpu0.user) + (cpu1.system - cpu0.system)
    barrier()
    clocks = sampler.stop()
    assert n_ok == B * args.steps
    launches = sum(c.launch_count() for c in lanes) - launches0
    h2d0 = merged_report()      # byte counters are always on; kernel events only when profiling is enabled
    # ---- timed region 2: `e2e` -- inputs built on the host every step, verdicts read back.  One untimed step of the same
    # kind first: the first call after the profile counters were collected has been seen to take 2-4x longer on
    # multi-GPU boxes (prove_call_wall_s shows every call), which is a transition artefact, not throughput
    assert run_steps(lambda k: make_inputs(B, base + 7 * world * B, n), 1) == B
    barrier()
    t0 = time.time()
    n_ok = run_steps(lambda k: make_inputs(B, base + (k + 1) * world * B, n), args.steps)
    assert n_ok == B * args.steps
    barrier()
    e2e_s = time.time() - t0
    t_dev, t_e2e = ms / 1e3, e2e_s
    # ---- per-kernel times for the rooflines: ONE lane's share of a step (B / lanes proofs, prove then
    # verify) on a single stream with a CUDA-event pair around every launch.  In the timed regions the
    # lanes' streams overlap, so an event pair there also times other lanes' kernels; serialised, the
    # shares line up with the ncu launch list in profiles/.
    lanes_per_setup = len(setup.contexts())
    Bp = max(1, B // lanes_per_setup)
    os.environ["BPPP_LANES"] = "1"
    pctx = bp.Context(local)
    psetup = bp.RangeProofSetup(pctx, workload_schema())
    psetup.set_device_transcript(dev_tr)
    psetup.set_batch_verify(args.verify == "batch")
    if lut_c:
        psetup.enable_lut(args.lut_gb)
    pin = make_inputs(Bp, base, n)
    proof = psetup.prove_batch_raw(Bp, pin[0], pin[1], None, pin[2])            # warm-up (tables, pools)
    pctx.profile_enable(True)
    pctx.profile_reset()
    tp0 = time.time()
    proof = psetup.prove_batch_raw(Bp, pin[0], pin[1], None, pin[2])
    assert sum(psetup.verify_batch_raw(Bp, *proof)) == Bp
    pctx.sync()
    prof_wall = time.time() - tp0
    rep = pctx.profile_report()
    pctx.profile_enable(False)
    # one proof alone (the drop-in seams' batch size): prove + verify wall time through the same entry points
    lat = []
    one = make_inputs(1, base, n)
    for _ in range(6):
        tl0 = time.time()
        pr1 = psetup.prove_batch_raw(1, one[0], one[1], None, one[2])
        assert sum(psetup.verify_batch_raw(1, *pr1)) == 1
        lat.append(1e3 * (time.time() - tl0))
    single_ms = statistics.median(lat[1:])
    rep["h2d_bytes"], rep["d2h_bytes"] = h2d0["h2d_bytes"], h2d0["d2h_bytes"]
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = t.tolist()
    sharded = None
    if world > 1 and not args.no_sweep:
        # BASELINE.json config 5 on several GPUs: ONE 2^20-element norm argument sharded over the ranks
        # (strong scaling; NCCL all-gather inside the library, bppp_nl_prove_sharded), after the timed regions
        from bulletproofspp_b200 import sweep
        vsetup.close(); psetup.close()
        for st in psetups:
            st.close()
        try:
            sharded = sweep.run_sharded(ctx, args.sharded_size, rank, world, dist)
        except Exception as ex:                    # reported, never silently dropped
            sharded = {"error": repr(ex)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total = world * B * args.steps
    # ---- rooflines from the per-kernel CUDA-event times of timed region 1
    kern = rep["kernels"]
    tot_ms = sum(k["ms"] for k in kern.values()) or 1.0
    shares = {name: round(k["ms"] / tot_ms, 4) for name, k in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
    imad_wide, imad_lo = ctx.measure_imad_peak()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    top = max((name for name in kern if kern[name]["work"] > 0 and name != "k_fold_dots"), key=lambda nme: kern[nme]["ms"])
    kt = kern[top]
    ach = kt["work"] / (kt["ms"] * 1e-3) / 1e12
    ncu = ncu_summary().get(top, {})
    peak_t = imad_wide / 1e12
    # IMAD.WIDE instructions the kernel really issues for its group arithmetic, counted live: every table lookup of
    # k_msm_lut is one XYZZ mixed addition = 8 fq::mul + 2 fq::sqr = 8 x 72 + 2 x 43 IMAD.WIDE (SASS of this build)
    issued = None
    if top == "k_msm_lut" and rep.get("lut_lookups"):
        issued = rep["lut_lookups"] * (8 * 72 + 2 * 43) / (kt["ms"] * 1e-3) / 1e12
    frac_pipe = issued / peak_t if issued else None
    roofline = {"kernel": top, "bound": "imad", "achieved": issued if issued else ach, "peak": peak_t, "unit": "TIMAD/s",
                "frac": frac_pipe if frac_pipe else ach / peak_t,
                "frac_pipe": frac_pipe, "frac_alg": ach / peak_t, "achieved_alg": ach,
                "pipe_active_ncu": (ncu.get("pipe_fmaheavy_active_pct", 0) / 100.0) or None,
                "traffic": ncu.get("dram_bytes_per_launch"),
                "lookups_per_launch": (rep.get("lut_lookups", 0) / kt["launches"]) if issued else None,
                "avg_launch_ms": kt["ms"] / kt["launches"], "share_of_gpu_time": shares[top],
                "note": "integer-pipe bound (no hbm/tensor roofline applies).  frac = frac_pipe = IMAD.WIDE instructions issued for the group "
                        "arithmetic (counted live: lookups x 662 per mixed addition) / CUDA-event time / the IMAD.WIDE issue rate measured in "
                        "this run.  pipe_active_ncu = sm__pipe_fmaheavy_cycles_active of the committed ncu --set full capture (%s): the "
                        "carry handling and register moves of a field multiplication (IMAD.X, IMAD.MOV: a quarter of its IMAD-class "
                        "instructions) occupy the same pipe, which is why %s of useful multiplies is %s of the pipe.  frac_alg = the reference schedule's IMADs "
                        "(SURVEY 8(d): Pippenger MSMs over each round's CURRENT lengths + the generator folds this kernel absorbs, 136 IMAD "
                        "per multiplication) / time / peak -- it exceeds 1 because a full-multiples table needs neither buckets nor "
                        "reductions.  traffic = dram read + write bytes of one launch from the same capture (a profiler cannot run inside "
                        "the timed bench)" % (ncu.get("source", "profiles/ncu_summary.json missing"),
                                              "%.2f" % frac_pipe if frac_pipe else "this share",
                                              "%.2f" % (ncu.get("pipe_fmaheavy_active_pct", 0) / 100.0) if ncu else "more")}
    rooflines = {}
    for name in ("k_msm_bucket", "k_pair_fold"):
        if name in kern and kern[name]["ms"] > 0 and kern[name]["work"] > 0:
            a = kern[name]["work"] / (kern[name]["ms"] * 1e-3) / 1e12
            rooflines[name] = {"bound": "imad", "achieved": a, "peak": imad_wide / 1e12, "unit": "TIMAD/s",
                               "frac": a / (imad_wide / 1e12), "share_of_gpu_time": shares[name]}
    if "k_fold_dots" in kern:
        kf = kern["k_fold_dots"]
        a = kf["work"] / (kf["ms"] * 1e-3) / 1e9
        rooflines["k_fold_dots"] = {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                                    "traffic": None, "share_of_gpu_time": shares["k_fold_dots"],
                                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}
    line = {"metric": METRIC, "value": total / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u256 (Fq/Fr integers, 8x32-bit limbs)", "data": "synthetic",
            "config": {"workload": "examples/128by64 prove+verify (N=1024, M=261, 9 rounds), %d proofs per GPU per step" % B,
                       "batch_per_gpu": B, "parallelism": "batch sharded over %d GPU(s), no data-path collective" % world,
                       "l2": "working set per step (%.0f MB of generators+witness vectors) exceeds the 126 MB L2" % (B * 0.33),
                       "host_threads": args.host_threads or max(1, (os.cpu_count() or 1) // world), "lanes": len(lanes),
                       "batches_in_flight": "%d prover setups: consecutive bppp_rp_prove_batch calls overlap, a third thread verifies "
                                            "finished batches" % len(psetups),
                       "msm_table": ("full-multiples table in HBM, window %d bits: %d lookups + mixed additions per term, %.1f GB, built once "
                                     "per process in %.2f s before the timed regions (csrc/lut.cuh)" % (
                                         lut_c, (256 + lut_c - 1) // lut_c, 1286 * ((256 + lut_c - 1) // lut_c) * 2 ** (lut_c - 1) * 64 / 1e9, t_lut))
                                    if lut_c else "nine-bit window table + bucket kernel (k_msm_gens)" + (" -- table not built: %s" % lut_err if lut_err else ""),
                       "verify": "batch verification across proofs: one random linear combination per lane sub-batch of %d proofs (128-bit "
                                 "weights from getrandom), per-proof checks locate failures (SURVEY 8 f2)" % max(1, B // (len(lanes) // 2))
                                 if args.verify == "batch" else "per proof (the reference's verifyM)",
                       "transcript": "device (SHA-256 + decimal show of the commitments in k_tr_prepend / k_tr_squeeze, challenges "
                                     "bit-identical to the host transcript; SURVEY 8 f4)" if dev_tr else "host (the reference's arrangement)",
                       "extra_warmup_steps": extra_warmup},
            "e2e": {"value": total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": rep["h2d_bytes"] // args.steps,
                    "d2h_bytes_per_step": rep["d2h_bytes"] // args.steps,
                    "note": "host buffers in, proofs + verdicts out through bppp_rp_prove_batch/bppp_rp_verify_batch; "
                            "inputs rebuilt on the host every step, wall clock"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "rooflines": rooflines,
            "kernel_time_shares": shares,
            "gpu_busy_estimate": tot_ms * lanes_per_setup / (1e3 * t_dev / args.steps),
            "profile_pass": {"proofs": Bp, "kernel_ms": tot_ms, "wall_s": prof_wall,
                             "note": "one lane's share of a step (prove then verify) on a single stream with a CUDA-event pair "
                                     "around every launch, run after the timed regions; gpu_busy_estimate = kernel_ms x lanes / ms_per_step"},
            "single_proof_latency_ms": {"prove_plus_verify": round(single_ms, 3),
                                        "note": "batch of ONE 128by64 proof through bppp_rp_prove_batch + bppp_rp_verify_batch (one lane, "
                                                "host buffers, wall clock, median of 5): the latency the per-proof Haskell seams see"},
            "imad_peak": {"wide_per_s": imad_wide, "lo32_per_s": imad_lo}, "wall_s_value_leg": wall,
            "prove_call_wall_s": prove_wall,
            "host": {"cpu_ms_per_proof": 1e3 * host_cpu_s / (B * args.steps), "cores_busy": host_cpu_s / wall,
                     "cores_available": (os.cpu_count() or 1) / world,
                     "note": "rank 0's process CPU time over the value leg (blinders, linear slots, round sequencing"
                             + ("" if dev_tr else ", transcript hashing") + ")"}}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_sample()
    if not args.no_sweep and world == 1:
        # second half of BASELINE.json's metric ("norm-arg fold GB/s") and its config 5: the synthetic norm-argument
        # sweep, one large argument per size, after the timed regions of the headline metric
        from bulletproofspp_b200 import sweep
        vsetup.close(); psetup.close()
        for st in psetups:
            st.close()
        line["norm_arg_sweep"] = sweep.run(ctx, [int(x) for x in args.sweep_sizes.split(",") if x], imad_wide, hbm_peak)
    if sharded is not None:
        line["norm_arg_sharded"] = sharded
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
