"""Data points beyond the bench workload: Zipf skew sweep (BASELINE configs[4]) at the kernel level and the
community-attachment k-SAT of configs[2] end to end."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
from bench import to_codes_sorted
what = sys.argv[1]
if what == "zipf":
    for i, s in enumerate([0.0, 0.5, 0.8, 1.0, 1.2]):
        lits, sizes = rs.gen.random_ksat(1000000, 4260000, 3, seed=777 + i, zipf_s=s)
        codes = to_codes_sorted(lits, sizes)
        svc = rs.ScanService(0); svc.upload(int(np.abs(lits).max()), codes, sizes)
        best = None
        for rep in range(3):
            c0 = {k: svc.counter(k) for k in ("kernel_ns", "scan_ns", "scan_checks", "scan_alg_bytes")}
            svc.build(); b = svc.counter("kernel_ns") - c0["kernel_ns"]
            svc.scan_all(rs.SUBSUME, count_only=True); svc.scan_all(rs.STRENGTHEN, count_only=True)
            d = {k: svc.counter(k) - c0[k] for k in c0}
            r = (d["scan_ns"], b, d["scan_checks"], d["scan_alg_bytes"])
            if best is None or r[0] < best[0]: best = r
        off, refs = svc.download_occ()
        print(f"zipf s={s}: max list {int(np.diff(off).max())}  build {best[1]/1e6:.2f} ms  K3+K4 {best[0]/1e6:.3f} ms  {best[2]/1e6:.1f} M checks  "
              f"{best[2]/best[0]:.2f} G checks/s  {best[3]/best[0]:.0f} GB/s algorithmic", flush=True)
        svc.close()
else:
    n = int(sys.argv[2]); opts = sys.argv[3:]
    t = time.time(); lits, sizes = rs.gen.community_ksat(n, 5 * n); stream = rs.gen.csr_to_stream(lits, sizes)
    print(f"community k-SAT {n} vars / {sizes.size} clauses generated in {time.time()-t:.1f}s", flush=True)
    cp = rs.Coprocessor(opts); t = time.time(); cp.add_stream(stream); ta = time.time() - t
    t = time.time(); cp.preprocess(); tp = time.time() - t
    names = ["sub_subsSteps", "sub_strengthSteps", "sub_subsumedClauses", "sub_removedLiterals", "bve_steps", "bve_eliminatedVars", "scan_batches", "scan_ondemand", "pair_batches",
             "gpu_launches", "host_gpuwait_ns", "host_commit_ns", "ph_init_ns", "gpu_h2d_bytes"]
    print(f"add {ta:.2f}s preprocess {tp:.2f}s ok={cp.ok()} clauses_out={cp.n_clauses()}", {k: (round(cp.stat(k)/1e6,1) if k.endswith('_ns') else cp.stat(k)) for k in names}, flush=True)
