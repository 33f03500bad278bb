"""BASELINE configs[2] data point: community-attachment k-SAT, 10M vars / 50M clauses (+ planted redundancy),
subsumption + strengthening + BVE on 1 B200.  (1) device-resident K2 build + K3 + K4 with the roofline figures,
(2) end to end through the C ABI with the full option set, (3) the sequential reference on the same formula.
usage: python scripts/c3_point.py [n_vars] [--no-ref]"""
import os, sys, time, subprocess, tempfile, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
from bench import to_codes_sorted
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 10000000
FULL = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
t = time.time(); lits, sizes = rs.gen.community_ksat(n, 5 * n); stream = rs.gen.csr_to_stream(lits, sizes)
print(f"community k-SAT {n} vars / {sizes.size} clauses / {lits.size} literals generated in {time.time()-t:.1f}s", flush=True)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6553.9)
codes = to_codes_sorted(lits, sizes)
svc = rs.ScanService(0); svc.upload(int(np.abs(lits).max()), codes, sizes)
for rep in range(3):
    c0 = {k: svc.counter(k) for k in ("kernel_ns", "scan_ns", "scan_checks", "scan_alg_bytes")}
    svc.build(); b = svc.counter("kernel_ns") - c0["kernel_ns"]
    s0 = svc.counter("scan_ns"); n3, c3 = svc.scan_all(rs.SUBSUME, count_only=True); s1 = svc.counter("scan_ns")
    n4, c4 = svc.scan_all(rs.STRENGTHEN, count_only=True); s2 = svc.counter("scan_ns")
    ab = svc.counter("scan_alg_bytes") - c0["scan_alg_bytes"]
    print(f"resident rep {rep}: build {b/1e6:.2f} ms | K3 {(s1-s0)/1e6:.3f} ms {c3} checks {n3} hits | K4 {(s2-s1)/1e6:.3f} ms {c4} checks {n4} hits | "
          f"K3+K4 {ab/(s2-s0):.0f} GB/s algorithmic = {ab/(s2-s0)/peak:.3f} of {peak} | whole pass {(c3+c4)/((b+s2-s0)*1e-9)/1e9:.1f} G checks/s", flush=True)
svc.close(); del codes
for rep in range(2):
    cp = rs.Coprocessor(FULL); t = time.time(); cp.add_stream(stream); ta = time.time() - t
    t = time.time(); cp.preprocess(); tp = time.time() - t
    names = ["sub_subsSteps", "sub_strengthSteps", "sub_subsumedClauses", "sub_removedLiterals", "bve_steps", "bve_eliminatedVars", "bve_createdClauses", "scan_batches",
             "scan_ondemand", "pair_batches", "resolve_batches", "gpu_launches", "host_gpuwait_ns", "host_commit_ns", "ph_init_ns", "ph_sub_ns", "ph_str_ns", "gpu_h2d_bytes", "gpu_d2h_bytes"]
    st = {k: cp.stat(k) for k in names}
    print(f"e2e rep {rep}: add {ta:.2f}s preprocess {tp:.2f}s ok={cp.ok()} clauses_out={cp.n_clauses()} "
          f"{(st['sub_subsSteps']+st['sub_strengthSteps']+st['bve_steps'])/tp/1e6:.1f} M checks/s", {k: (round(v/1e6,1) if k.endswith('_ns') else v) for k, v in st.items()}, flush=True)
    cp.close()
ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "refshim")
if "--no-ref" not in sys.argv and os.path.exists(ref):
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.bin"); out = os.path.join(td, "o.txt"); stream.tofile(inp)
        t = time.time()
        subprocess.check_call([ref, "time", inp, out] + FULL, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        print(f"reference sequential ({time.time()-t:.1f}s wall incl. ingest):", " ".join(l.strip() for l in open(out) if any(k in l for k in ("seconds", "Steps", "bve_steps", "eliminatedVars"))), flush=True)
