#!/usr/bin/env python
"""bench.py - headline benchmark of the Coprocessor GPU hot path (BASELINE.json: subsumption checks/s, simplify wall time,
% of the HBM roofline).

A "check" is the reference's own unit: one Stepper::increaseSteps(1) - one (subsumer | strengthener, candidate) pair taken from
an occurrence list (Subsumption.cc:268,1330,1431) or one (pos, neg) clause pair of a BVE candidate (BVE.cc:756,910).  Parity
makes these counters identical to the reference's, so checks/s of both arms divide the same number by their own time.

A "step" is one whole simplification fixpoint over one formula:

  value  checks / time of CPpreprocess alone: the clause arena is already resident in HBM (CPaddLits put it there), the timed
         region is the fixpoint itself - every kernel launch, every ordered commit, results on the host.  Like for like with
         the reference arm, which times Preprocessor::preprocess() only.
  e2e    checks / time of CPaddLits (DIMACS stream in pinned HOST memory -> device bulk ingest) + CPpreprocess + read-back of
         the simplified CNF into a host buffer: what a client of libcoprocessorc.h pays.
  roofline  the full-queue kernel k_full (K3 subsumption + K4 strengthening launches inside the timed steps): ALGORITHMIC bytes
         per launch (4 + 8 + 4*#D per check, counted by the kernel) / launch duration (CUDA events on the
         library's stream), against MEASURED_PEAKS.json; `per_kernel` lists every kernel family of the step the same way.

--config c2 (default)  BASELINE configs[1]: planted random 3-SAT, 1M vars / 4.26M clauses, subsumption + strengthening fixpoint.
                       At N=1 the default run appends one fixpoint at BASELINE configs[2] scale under the key "c3"
                       (community k-SAT 10M vars / 54M clauses, -up -subsimp -bve): the target configuration takes ~20 s per
                       step, so it is measured once per run, not K times (DESIGN.md "Measurement").
--config c3            configs[2] as the K-step workload.
--config c4            configs[3]: the Riss427-style pipeline (-up -subsimp -bve -bce -dense) on one 25M-clause shard per rank of the
                       200M-clause community k-SAT formula (--gpus 8: weak, one shard each); --vars scales the shard.
--config zipf          configs[4]: literal-frequency skew sweep, one JSON line per exponent s.
--impl reference       the reference's own CPU Coprocessor (oracle/_ref, unmodified sources, all host threads) on the same
                       workload and metric.

One JSON line on stdout (rank 0) - except --config zipf.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simplification checks/s over the whole fixpoint (reference Stepper steps: subsumption + strengthening + BVE)"
UNIT = "checks/s"
REFSHIM = os.path.join(ROOT, "oracle", "_ref", "refshim")
SUSI = ["-enabled_cp3", "-subsimp", "-no-cp3_limited"]
FULL = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
R427 = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-bce", "-dense", "-no-cp3_limited"]
CHECK_STATS = ("sub_subsSteps", "sub_strengthSteps", "bve_steps", "bce_steps")


class Config:
    def __init__(self, name, n_vars):
        self.name, self.n = name, n_vars
        self.opts = SUSI if name in ("c2", "zipf") else (R427 if name == "c4" else FULL)

    def generate(self, seed_offset=0, zipf_s=0.0):
        import riss_solver_b200 as rs
        if self.name in ("c3", "c4"):
            lits, sizes = rs.gen.community_ksat(self.n, 5 * self.n, seed=20260001 + seed_offset)
        else:
            lits, sizes = rs.gen.random_ksat(self.n, int(4.26 * self.n), 3, seed=(12345 if zipf_s == 0 else 777) + seed_offset, zipf_s=zipf_s)
        return lits, sizes, rs.gen.csr_to_stream(lits, sizes)

    def workload(self):
        """the string both arms print: what is simplified, with which options"""
        if self.name == "c4":
            return (f"BASELINE configs[3]: Riss427-style pipeline (unit propagation, subsumption + strengthening, BVE, blocked clause elimination, Dense) on one "
                    f"shard of the 200M-clause community-attachment k-SAT formula: {self.n} vars / {5 * self.n} base clauses per rank (8 ranks x 25M = 200M), "
                    f"seed 20260001 + rank ({' '.join(R427)})")
        if self.name == "c3":
            return (f"BASELINE configs[2]: community-attachment k-SAT {self.n} vars / {5 * self.n} base clauses, k in 3/4/5 (50/30/20 %), 5 % rare variables, "
                    f"+2 % duplicates, 3 % supersets, 3 % flipped copies, seed 20260001; unit propagation + subsumption + strengthening + BVE to the fixpoint ({' '.join(FULL)})")
        return (f"BASELINE configs[1]: planted random 3-SAT {self.n} vars / {int(4.26 * self.n)} base clauses +2 % duplicates, 3 % supersets, 3 % flipped copies, "
                f"seed 12345; subsumption + strengthening to the fixpoint ({' '.join(SUSI)})")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference on the host cores
def run_refshim(stream, opts, tag):
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, f"{tag}.bin"); out = os.path.join(td, "out.txt")
        np.ascontiguousarray(stream, dtype=np.int32).tofile(inp)
        subprocess.check_call([REFSHIM, "time", inp, out] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        st = {}
        for line in open(out):
            t = line.split()
            if t and t[0] == "stat":
                st[t[1]] = float(t[2])
            elif t and t[0] == "seconds":
                st["seconds"] = float(t[1])
        return st


def checks_of(st):
    return sum(st.get(k, 0) for k in CHECK_STATS)


def cpu_baseline(cfg: Config, sample: Config):
    """the sequential reference (the parity oracle's source of truth) on a bounded sample of the workload, one host core"""
    _, sizes, stream = sample.generate()
    what = f"{sample.name} shape, {sample.n} vars / {sizes.size} clauses ({'the full workload' if sample.n == cfg.n else 'bounded sample of the workload'}), options {' '.join(sample.opts)}"
    if not os.path.exists(REFSHIM):
        # the compiled reference did not travel: time the plain-C port instead
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from harness import Oracle
        o = Oracle(sample.opts); o.add(stream)
        t = time.time(); o.preprocess(); dt = time.time() - t
        r = o.result(); o.close()
        return {"value": checks_of(r.stats) / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt, "sample": what}
    st = run_refshim(stream, sample.opts, "cpu")
    return {"value": checks_of(st) / st["seconds"], "unit": UNIT, "cores": 1, "kind": "reference", "seconds": st["seconds"],
            "sample": what + f"; sequential Preprocessor::preprocess() {st['seconds']:.2f} s, {int(checks_of(st))} checks"}


def reference_arm(args, cfg: Config):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REFSHIM):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/refshim missing (reference not built)"}))
        return
    # c2 runs the full workload (2.5 s per step with the thread pool); c3 a tenth of it (the full one takes minutes per step)
    sample = cfg if cfg.name not in ("c3", "c4") else Config(cfg.name, max(1, cfg.n // 10))
    _, sizes, stream = sample.generate()
    cores = os.cpu_count() or 1
    seq = run_refshim(stream, sample.opts, "seq")                 # untimed: the step counters of this sample (thread-independent unit)
    checks = checks_of(seq)
    par = [f"-cp3_threads={cores - 1}", "-cp3_par_subs=2", "-cp3_par_strength=2"] if cores > 1 else []
    times = []
    for i in range(args.warmup + args.steps):
        st = run_refshim(stream, sample.opts + par, "par")
        if i >= args.warmup:
            times.append(st["seconds"])
    T = float(np.mean(times))
    v = checks / T
    desc = (f"{sample.n} vars / {sizes.size} clauses ({'the full workload' if sample.n == cfg.n else 'a tenth of the workload: bounded sample'}); unmodified reference "
            f"Preprocessor::preprocess() with {' '.join(par) or 'one thread'}; checks = sequential step counters of the same input ({int(checks)}); "
            f"sequential run of the same input {seq['seconds']:.2f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": T * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": cfg.workload(), "parallelism": f"cpu pthreads x{max(cores - 1, 1)}", "sample": desc},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc, "sequential_seconds": seq["seconds"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def to_codes_sorted(lits, sizes):
    """DIMACS literals -> Riss literal codes 2*var+sign, every clause sorted (what cp3g_upload expects; used by scripts/)"""
    v = np.abs(lits).astype(np.int64) - 1
    codes = (2 * v + (lits < 0)).astype(np.int32)
    k = int(sizes.max())
    if (sizes == k).all():
        return np.sort(codes.reshape(-1, k), axis=1).reshape(-1)
    out = codes.copy(); ends = np.cumsum(sizes); starts = ends - sizes
    for s in np.unique(sizes):
        idx = np.flatnonzero(sizes == s)
        pos = (starts[idx][:, None] + np.arange(s)[None, :]).reshape(-1)
        out[pos] = np.sort(codes[pos].reshape(-1, s), axis=1).reshape(-1)
    return out


KERNEL_FAMILIES = [   # (label, ns counter, algorithmic-bytes counter)
    ("K0 k_ing_* (bulk ingest)", "gpu_ingest_ns", "gpu_ingest_bytes"),
    ("K1+K2 k_occ_* (signatures + occurrence lists)", "gpu_build_ns", "gpu_build_bytes"),
    ("K3 k_full<subsume>", "gpu_full0_ns", "gpu_full0_bytes"),
    ("K4 k_full<strengthen>", "gpu_full1_ns", "gpu_full1_bytes"),
    ("commit k_sp_* (device pass commit)", "gpu_commit_ns", "gpu_commit_bytes"),
    ("K5'+K6' k_bve_eval/k_bve_stream", "gpu_bve_ns", "gpu_bve_bytes"),
    ("K9 k_bce_masks", "gpu_bce_ns", "gpu_bce_bytes"),
]
STEP_STATS = ["sub_subsSteps", "sub_strengthSteps", "bve_steps", "bce_steps", "bce_remBCE", "gpu_launches", "gpu_h2d_bytes", "gpu_d2h_bytes", "sub_subsumedClauses", "sub_removedLiterals",
              "bve_eliminatedVars", "scan_batches", "scan_ondemand", "device_sub_passes", "eval_batches", "gpu_kernel_ns", "gpu_scan_ns",
              "gpu_full0_checks", "gpu_full1_checks", "gpu_full0_launches", "gpu_full1_launches", "gpu_build_count"] + [k for f in KERNEL_FAMILIES for k in f[1:]]


def fixpoint_step(rs, opts, stream, barrier, distributed=None):
    """one step: CPinit(options) -> CPaddLits -> CPpreprocess -> CPgetOutput; returns (t_add, t_preprocess, t_output, stats, output size)"""
    cp = rs.Coprocessor(opts)
    if distributed is not None:
        rank, world, uid = distributed
        if cp.set_distributed(rank, world, uid) != 0:
            raise SystemExit("CPsetDistributed failed (NCCL communicator)")
    barrier()
    t0 = time.perf_counter()
    cp.add_stream(stream)                         # CPaddLits: the DIMACS stream from pinned host memory (device bulk ingest K0)
    t1 = time.perf_counter()
    cp.preprocess()                               # the fixpoint
    t2 = time.perf_counter()
    out = cp.output_stream()                      # the simplified CNF in a host buffer
    barrier()
    t3 = time.perf_counter()
    st = {k: cp.stat(k) for k in STEP_STATS}
    cp.close()
    return t1 - t0, t2 - t1, t3 - t2, st, int(out.size)


def roofline_of(steps_stats, peak, peak_src, traffic):
    tot = {k: float(np.mean([s[k] for s in steps_stats])) for k in STEP_STATS}
    per = {}
    for label, kns, kby in KERNEL_FAMILIES:
        if tot[kns] <= 0:
            continue
        per[label] = {"ms_per_step": tot[kns] * 1e-6, "algorithmic_GB_per_step": tot[kby] * 1e-9, "GB/s": tot[kby] / tot[kns], "frac": tot[kby] / tot[kns] / peak}
    nl = max(tot["gpu_full0_launches"] + tot["gpu_full1_launches"], 1.0)
    fb = tot["gpu_full0_bytes"] + tot["gpu_full1_bytes"]; fn = tot["gpu_full0_ns"] + tot["gpu_full1_ns"]
    achieved = fb / fn if fn > 0 else 0.0                          # bytes / ns = GB/s
    return {"bound": "hbm", "kernel": "k_full (full-queue K3 + K4 launches of the timed steps)", "achieved": achieved, "peak": peak, "peak_source": peak_src,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": fb / nl, "launch_ms": fn * 1e-6 / nl,
            "launches_per_step": nl, "checks_per_launch": (tot["gpu_full0_checks"] + tot["gpu_full1_checks"]) / nl, "per_kernel": per}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "zipf"])
    ap.add_argument("--vars", type=int, default=0, help="variables of the workload (default: 1M for c2/zipf, 10M for c3)")
    ap.add_argument("--mode", default="weak", choices=["weak", "sharded"],
                    help="N>1: weak = one independent formula per rank (each rank its own GPU and share of the host cores); sharded = one formula, "
                         "scan requests partitioned over the ranks and hits all-gathered over NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the configs[2]-scale fixpoint that the default c2 run appends at N=1")
    ap.add_argument("--c3-vars", type=int, default=10000000)
    args = ap.parse_args()
    cfg = Config(args.config, args.vars or (10000000 if args.config == "c3" else (5000000 if args.config == "c4" else 1000000)))
    if args.impl == "reference":
        return reference_arm(args, cfg)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # one engine per rank: split the host cores between the ranks of this node (the ordered commit uses a thread pool)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("CP3_THREADS", str(max(1, min(24, (os.cpu_count() or 1) // max(1, local_world)))))
    # ... and give every rank its own contiguous set of them, like numactl would: the ranks' thread pools then do not migrate over
    # each other's cores and caches (CP3_BENCH_PIN=0 leaves the affinity alone)
    affinity = "inherited"
    if local_world > 1 and os.environ.get("CP3_BENCH_PIN", "1") != "0" and hasattr(os, "sched_setaffinity"):
        cpus = sorted(os.sched_getaffinity(0)); per = len(cpus) // local_world
        if per >= 1:
            os.sched_setaffinity(0, cpus[local * per:(local + 1) * per])
            affinity = f"rank-contiguous, {per} cpus"
    # a process that calls CPpreprocess repeatedly keeps the engine's work arrays in its heap (opt-in, capi.cpp)
    os.environ.setdefault("CP3_KEEP_HEAP", "1")
    import torch
    import torch.distributed as dist
    import riss_solver_b200 as rs
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the Coprocessor GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fresh_uid():
        # one ncclUniqueId per communicator: every Coprocessor instance of the sharded mode gets its own
        if not (world > 1 and args.mode == "sharded"):
            return None
        box = [rs.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return (rank, world, box[0])

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def pinned(stream):
        return torch.from_numpy(np.ascontiguousarray(stream, dtype=np.int32)).pin_memory().numpy()

    def run_steps(c: Config, warmup, steps, zipf_s=0.0):
        lits, sizes, stream = c.generate(seed_offset=(rank if args.mode == "weak" and world > 1 else 0), zipf_s=zipf_s)
        del lits
        stream = pinned(stream)
        opts = c.opts + [f"-cp3_gpu_device={local}"]
        for _ in range(warmup):
            flush.fill_(1); fixpoint_step(rs, opts, stream, barrier, fresh_uid())
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        res = []
        for _ in range(steps):
            flush.fill_(1); torch.cuda.synchronize()
            res.append(fixpoint_step(rs, opts, stream, barrier, fresh_uid()))
        barrier()
        clocks = sampler.stop() if sampler else None
        return sizes.size, int(stream.size), res, clocks

    def summarize(res):
        t_add = float(np.mean([r[0] for r in res])); t_pp = float(np.mean([r[1] for r in res])); t_out = float(np.mean([r[2] for r in res]))
        st = res[-1][3]
        return t_add, t_pp, t_out, st, float(checks_of(st))

    if cfg.name == "zipf":
        for s in (0.0, 0.5, 0.8, 1.0, 1.2):
            # the fixpoint over degenerate lists (s >= 1: lists of 10^5-10^6 entries, 10^10-10^11 checks) takes minutes per step - one
            # untimed-free step there; the kernel figures (roofline) are what this sweep is about
            ncl, nints, res, clocks = run_steps(cfg, 1 if s < 1.0 else 0, max(1, min(args.steps, 3)) if s < 1.0 else 1, zipf_s=s)
            t_add, t_pp, t_out, st, checks = summarize(res)
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": checks / t_pp, "unit": UNIT, "n_gpus": world, "ms_per_step": t_pp * 1e3, "higher_is_better": True,
                                  "dtype": "int32", "data": "synthetic",
                                  "config": {"workload": f"BASELINE configs[4]: Zipf literal-frequency skew s={s}, C2 shape ({cfg.n} vars / {ncl} clauses), subsumption + strengthening fixpoint"},
                                  "roofline": roofline_of([r[3] for r in res], peak, peak_src, None),
                                  "e2e": {"value": checks / (t_add + t_pp + t_out), "unit": UNIT, "ms_per_step": (t_add + t_pp + t_out) * 1e3},
                                  "checks_per_step": checks, "clocks": clocks}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    ncl, nints, res, clocks = run_steps(cfg, args.warmup, args.steps)
    t_add, t_pp, t_out, st, checks = summarize(res)
    T_e2e = t_add + t_pp + t_out
    # max over ranks (time), sum over ranks (work)
    if world > 1:
        t = torch.tensor([T_e2e, t_pp], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        T_e2e, t_pp = float(t[0]), float(t[1])
        w = torch.tensor([checks], device="cuda", dtype=torch.float64)
        if args.mode == "weak":
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
        checks = float(w[0])
    if rank == 0:
        line = {
            "metric": METRIC, "value": checks / t_pp, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_pp * 1e3, "higher_is_better": True, "scaling": "weak" if args.mode == "weak" else "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": cfg.workload(),
                       "clauses": ncl, "value_leg": "CPpreprocess: the whole fixpoint, clause arena resident in HBM at the start (wall clock, barrier + device sync on both sides)",
                       "e2e_leg": "CPaddLits (DIMACS stream in pinned host memory) + CPpreprocess + read-back of the simplified CNF",
                       "l2": "working set (occurrence entries + headers + literals) larger than the 126 MB L2; a 256 MB buffer is also written between steps",
                       "host_threads_per_rank": int(os.environ["CP3_THREADS"]), "cpu_affinity": affinity,
                       "parallelism": ("1 gpu" if world == 1 else (f"{world} ranks, one independent formula per rank (no data-path collective)" if args.mode == "weak"
                                       else f"{world} ranks, one formula: scan requests partitioned, hits all-gathered over NCCL"))},
            "roofline": roofline_of([r[3] for r in res], peak, peak_src, traffic.get("k_full_bytes_per_launch")),
            "e2e": {"value": checks / T_e2e, "unit": UNIT, "ms_per_step": T_e2e * 1e3, "add_ms": t_add * 1e3, "preprocess_ms": t_pp * 1e3, "output_ms": t_out * 1e3,
                    "h2d_bytes_per_step": st["gpu_h2d_bytes"], "d2h_bytes_per_step": st["gpu_d2h_bytes"], "checks_per_step": checks / (world if args.mode == "weak" else 1),
                    "subsumed_clauses": st["sub_subsumedClauses"], "strengthened_literals": st["sub_removedLiterals"], "eliminated_variables": st["bve_eliminatedVars"],
                    "scan_batches": st["scan_batches"], "scan_ondemand": st["scan_ondemand"], "device_sub_passes": st["device_sub_passes"],
                    "gpu_kernel_ms_per_step": st["gpu_kernel_ns"] * 1e-6},
            "gpu_launches": int(sum(r[3]["gpu_launches"] for r in res)),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, cfg if cfg.name not in ("c3", "c4") else Config(cfg.name, max(1, cfg.n // 10)))
        if world == 1 and cfg.name == "c2" and not args.no_c3:
            # the target configuration, once per run (one untimed + one timed fixpoint)
            c3 = Config("c3", args.c3_vars)
            t0 = time.time()
            ncl3, _, res3, _ = run_steps(c3, 1, 1)
            a3, p3, o3, st3, chk3 = summarize(res3)
            sec = {"workload": c3.workload(), "clauses": ncl3, "steps": 1, "warmup": 1, "value": chk3 / p3, "unit": UNIT, "preprocess_s": p3,
                   "e2e": {"value": chk3 / (a3 + p3 + o3), "unit": UNIT, "seconds": a3 + p3 + o3, "add_s": a3, "output_s": o3,
                           "h2d_bytes": st3["gpu_h2d_bytes"], "d2h_bytes": st3["gpu_d2h_bytes"]},
                   "checks": chk3, "sub_subsSteps": st3["sub_subsSteps"], "sub_strengthSteps": st3["sub_strengthSteps"], "bve_steps": st3["bve_steps"],
                   "subsumed_clauses": st3["sub_subsumedClauses"], "strengthened_literals": st3["sub_removedLiterals"], "eliminated_variables": st3["bve_eliminatedVars"],
                   "gpu_launches": st3["gpu_launches"], "roofline": roofline_of([res3[-1][3]], peak, peak_src, traffic.get("k_full_c3_bytes_per_launch")),
                   "section_seconds": time.time() - t0}
            if not args.no_cpu_baseline:
                sec["cpu_baseline"] = cpu_baseline(c3, Config("c3", max(1, c3.n // 10)))
            line["c3"] = sec
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
