#!/usr/bin/env python
"""bench.py - headline benchmark of the Coprocessor GPU hot path (BASELINE.json: subsumption checks/s).

Workload (config.workload): BASELINE configs[1] - synthetic random 3-SAT, 1M vars / 4.26M clauses with planted
duplicates / supersets / one-literal-flipped copies (seed 12345): subsumption + strengthening
fixpoint.  A "check" is the reference's own unit: one Stepper::increaseSteps(1), i.e. one (subsumer|strengthener,
candidate) pair taken from an occurrence list (Subsumption.cc:268,1330,1431).

  value  checks/s of the device-resident bulk pass: clause arena already in HBM, K1 signatures + K2 occurrence
         CSR + K3 full-queue subsumption scan + K4 full-queue strengthening scan, timed with CUDA events on the
         library's stream (cp3g_* scan service, hits stay on the device).
  e2e    checks/s through the reference-facing C ABI on HOST buffers: CPaddLits of the DIMACS stream (host -> device
         inside the timed region), CPpreprocess (all scans, ordered commit on the host), read-back of the output CNF -
         the whole fixpoint; checks = the reference-defined step counters, which parity makes identical to the
         reference's.  (The reference arm times Preprocessor::preprocess() only, its own ingest is not charged.)
  --impl reference   the reference's own CPU Coprocessor (oracle/_ref, unmodified sources) on a bounded sample of
         the same workload with all host threads (-cp3_threads), same metric.

One JSON line on stdout (rank 0).
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

METRIC = "subsumption+strengthening checks/s (reference Stepper steps)"
UNIT = "checks/s"
REFSHIM = os.path.join(ROOT, "oracle", "_ref", "refshim")
SUSI = ["-enabled_cp3", "-subsimp", "-no-cp3_limited"]


def workload(n_vars: int, seed: int):
    import riss_solver_b200 as rs
    lits, sizes = rs.gen.random_ksat(n_vars, int(4.26 * n_vars), 3, seed=seed)
    return lits, sizes, rs.gen.csr_to_stream(lits, sizes)


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


# ------------------------------------------------------------------------------------------------ reference arm
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


def cpu_baseline(sample_vars: int, seed: int):
    """sequential reference (the parity oracle's source of truth) on a bounded sample, one host core"""
    if not os.path.exists(REFSHIM):
        # the compiled reference did not travel: time the plain-C port instead
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from harness import Oracle
        _, _, stream = workload(sample_vars, seed)
        o = Oracle(SUSI); o.add(stream)
        t = time.time(); o.preprocess(); dt = time.time() - t
        r = o.result(); o.close()
        checks = r.stats["sub_subsSteps"] + r.stats["sub_strengthSteps"]
        return {"value": checks / dt, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"planted 3-SAT {sample_vars} vars / {int(4.26 * sample_vars)} clauses, seed {seed}, {dt:.2f} s"}
    _, _, stream = workload(sample_vars, seed)
    st = run_refshim(stream, SUSI, "cpu")
    checks = st["sub_subsSteps"] + st["sub_strengthSteps"]
    return {"value": checks / st["seconds"], "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"planted 3-SAT {sample_vars} vars / {int(4.26 * sample_vars)} clauses, seed {seed}, sequential "
                      f"Preprocessor::preprocess() {st['seconds']:.2f} s, {int(checks)} checks"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REFSHIM):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/refshim missing (reference not built)"}))
        return
    nv = args.ref_vars
    _, _, stream = workload(nv, args.seed)
    cores = os.cpu_count() or 1
    seq = run_refshim(stream, SUSI, "seq")                       # untimed: the step counters of this sample
    checks = seq["sub_subsSteps"] + seq["sub_strengthSteps"]
    opts = SUSI + ([f"-cp3_threads={cores - 1}", "-cp3_par_subs=2", "-cp3_par_strength=2"] if cores > 1 else [])
    times = []
    for i in range(args.warmup + args.steps):
        st = run_refshim(stream, opts, "par")
        if i >= args.warmup:
            times.append(st["seconds"])
    T = float(np.mean(times))
    v = checks / T
    sample = (f"planted 3-SAT {nv} vars / {int(4.26 * nv)} clauses, seed {args.seed}; unmodified reference "
              f"Preprocessor::preprocess() with -cp3_threads={cores - 1} -cp3_par_subs=2 -cp3_par_strength=2; "
              f"checks = sequential step counters of the same sample ({int(checks)})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": T * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"C2-shape planted random 3-SAT sample, {nv} vars / {int(4.26 * nv)} clauses, subsumption+strengthening fixpoint",
                   "parallelism": f"cpu pthreads x{cores - 1}"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def to_codes_sorted(lits, sizes):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vars", type=int, default=1000000, help="variables of the C2 workload (clauses = 4.26x)")
    ap.add_argument("--ref-vars", type=int, default=1000000, help="size of the sample the CPU reference legs run (default: the full workload)")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--mode", default="weak", choices=["weak", "sharded"],
                    help="N>1: weak = one independent formula per rank; sharded = one formula, scan requests "
                         "partitioned over the ranks and hits all-gathered over NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # one engine per rank: split the host cores between the ranks of this node (the ordered commit uses a thread pool)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("CP3_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // max(1, local_world)))))
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

    seed = args.seed + (rank if args.mode == "weak" else 0)
    lits, sizes, stream = workload(args.vars, seed)
    nvars = int(np.abs(lits).max())
    codes = to_codes_sorted(lits, sizes)
    opts = SUSI + [f"-cp3_gpu_device={local}"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    stream = torch.from_numpy(np.ascontiguousarray(stream, dtype=np.int32)).pin_memory().numpy()   # e2e input: pinned host memory

    def fresh_uid():
        # one ncclUniqueId per communicator: every Coprocessor instance of the sharded mode gets its own
        if not (world > 1 and args.mode == "sharded"):
            return None
        box = [rs.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    # ---------------- leg 1: device-resident bulk pass (value, roofline)
    svc = rs.ScanService(local)
    svc.upload(nvars, codes, sizes)
    ids = np.arange(sizes.size, dtype=np.uint32)
    # algorithmic bytes of one check: 4 (occ entry) + 8 (header) + 4*#D (candidate literals)
    def resident_pass():
        c0 = {k: svc.counter(k) for k in ("kernel_ns", "launches", "scan_ns", "scan_checks", "scan_alg_bytes")}
        svc.build()
        b = svc.counter("kernel_ns") - c0["kernel_ns"]
        s0 = svc.counter("scan_ns"); a0 = svc.counter("scan_alg_bytes"); k0 = svc.counter("scan_checks")
        svc.scan_all(rs.SUBSUME, count_only=True)
        s1 = svc.counter("scan_ns"); a1 = svc.counter("scan_alg_bytes"); k1 = svc.counter("scan_checks")
        svc.scan_all(rs.STRENGTHEN, count_only=True)
        r = {k: svc.counter(k) - c0[k] for k in c0}
        r.update(build_ns=b, k3_ns=s1 - s0, k3_bytes=a1 - a0, k3_checks=k1 - k0, k4_ns=svc.counter("scan_ns") - s1,
                 k4_bytes=svc.counter("scan_alg_bytes") - a1, k4_checks=svc.counter("scan_checks") - k1)
        return r

    for _ in range(args.warmup):
        flush.fill_(1); resident_pass()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    res = []
    for _ in range(args.steps):
        flush.fill_(1); torch.cuda.synchronize()
        res.append(resident_pass())
    barrier()
    dev_ns = float(np.mean([r["kernel_ns"] for r in res]))
    scan_ns = float(np.mean([r["scan_ns"] for r in res]))
    dev_checks = float(np.mean([r["scan_checks"] for r in res]))
    alg_bytes = float(np.mean([r["scan_alg_bytes"] for r in res]))
    launches_res = int(np.mean([r["launches"] for r in res]))
    svc.close()

    # ---------------- leg 2: end to end through the C ABI (host buffers in, host buffers out)
    def e2e_step():
        cp = rs.Coprocessor(opts)
        uid = fresh_uid()
        if uid is not None and cp.set_distributed(rank, world, uid) != 0:
            raise SystemExit("CPsetDistributed failed (NCCL communicator)")
        barrier()
        t0 = time.perf_counter()
        cp.add_stream(stream)                         # CPaddLits: the DIMACS stream from host memory (device bulk ingest K0)
        cp.preprocess()
        out = cp.output_stream()                      # device->host results are part of the call; read the CNF back
        barrier()
        dt = time.perf_counter() - t0
        st = {k: cp.stat(k) for k in ("sub_subsSteps", "sub_strengthSteps", "gpu_launches", "gpu_h2d_bytes", "gpu_d2h_bytes",
                                      "sub_subsumedClauses", "sub_removedLiterals", "scan_batches", "scan_ondemand")}
        cp.close()
        return dt, st, out.size

    for _ in range(args.warmup):
        flush.fill_(1); e2e_step()
    e2e = []
    for _ in range(args.steps):
        flush.fill_(1); e2e.append(e2e_step())
    clocks = sampler.stop() if sampler else None
    T = float(np.mean([x[0] for x in e2e]))
    st = e2e[-1][1]
    checks = st["sub_subsSteps"] + st["sub_strengthSteps"]

    # max over ranks (time), sum over ranks (work)
    if world > 1:
        t = torch.tensor([T, dev_ns], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        T, dev_ns = float(t[0]), float(t[1])
        w = torch.tensor([float(checks), dev_checks], device="cuda", dtype=torch.float64)
        if args.mode == "weak":
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
        checks, dev_checks = float(w[0]), float(w[1])

    def mean(k):
        return float(np.mean([r[k] for r in res]))

    # DRAM bytes per launch of the dominant kernel from the last `ncu --set full` capture (profiles/), or None
    TRAFFIC = None
    try:
        TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_full_bytes_per_launch")
    except Exception:
        pass
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / scan_ns if scan_ns > 0 else 0.0                 # bytes/ns = GB/s
        line = {
            "metric": METRIC, "value": dev_checks / (dev_ns * 1e-9), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ns * 1e-6, "higher_is_better": True,
            "scaling": "weak" if args.mode == "weak" else "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: planted random 3-SAT {args.vars} vars / {sizes.size} clauses "
                                   f"(2% duplicates, 3% supersets, 3% flipped copies), seed {args.seed}, subsumption+strengthening fixpoint",
                       "value_leg": "device-resident bulk pass: K1 signatures + K2 occurrence CSR build + K3 + K4 over the full queue, CUDA events on the library stream",
                       "e2e_leg": "CPaddLits (DIMACS stream from host memory) + CPpreprocess + output read-back, wall clock per step",
                       "l2": "working set (occ entries + headers + literals, ~350 MB) larger than the 126 MB L2; a 256 MB buffer is also written between steps",
                       "host_threads_per_rank": int(os.environ["CP3_THREADS"]),
                       "parallelism": ("1 gpu" if world == 1 else (f"{world} ranks, one independent formula per rank" if args.mode == "weak"
                                       else f"{world} ranks, one formula: scan requests partitioned, hits all-gathered over NCCL"))},
            "roofline": {"bound": "hbm", "kernel": "k_full (K3+K4 full-queue form, both launches)", "achieved": achieved, "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": TRAFFIC,
                         "algorithmic_bytes_per_launch": alg_bytes / 2, "launch_ms": scan_ns * 1e-6 / 2,
                         "per_kernel": {
                             "k_full<subsume>": {"ms": mean("k3_ns") * 1e-6, "checks": mean("k3_checks"), "GB/s": mean("k3_bytes") / max(mean("k3_ns"), 1.0)},
                             "k_full<strengthen>": {"ms": mean("k4_ns") * 1e-6, "checks": mean("k4_checks"), "GB/s": mean("k4_bytes") / max(mean("k4_ns"), 1.0)},
                             "K1+K2 build": {"ms": mean("build_ns") * 1e-6}}},
            "e2e": {"value": checks / T, "unit": UNIT, "ms_per_step": T * 1e3, "h2d_bytes_per_step": st["gpu_h2d_bytes"],
                    "d2h_bytes_per_step": st["gpu_d2h_bytes"], "checks_per_step": checks,
                    "subsumed_clauses": st["sub_subsumedClauses"], "strengthened_literals": st["sub_removedLiterals"],
                    "scan_batches": st["scan_batches"], "scan_ondemand": st["scan_ondemand"]},
            "gpu_launches": launches_res * args.steps + sum(x[1]["gpu_launches"] for x in e2e),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.ref_vars, args.seed)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
