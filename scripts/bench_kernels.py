"""Kernel-level timing of the device-resident pass (CUDA events inside the library): K1+K2 build, K3, K4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Config, to_codes_sorted
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
# a third argument selects the Zipf sweep formula (BASELINE configs[4]) with that skew exponent, 0 included
lits, sizes, stream = (Config("zipf", n).generate(zipf_s=float(sys.argv[3])) if len(sys.argv) > 3 else Config("c2", n).generate())
codes = to_codes_sorted(lits, sizes)
svc = rs.ScanService(0)
svc.upload(int(np.abs(lits).max()), codes, sizes)
def delta(f):
    a = {k: svc.counter(k) for k in ("kernel_ns", "scan_ns", "scan_checks", "scan_alg_bytes")}
    r = f()
    return r, {k: svc.counter(k) - a[k] for k in a}
for rep in range(reps):
    _, b = delta(svc.build)
    (nh0, c0), s0 = delta(lambda: svc.scan_all(rs.SUBSUME, count_only=True))
    (nh1, c1), s1 = delta(lambda: svc.scan_all(rs.STRENGTHEN, count_only=True))
    print(f"rep {rep}: build {b['kernel_ns']/1e6:.3f} ms | K3 {s0['scan_ns']/1e6:.3f} ms {c0} checks {nh0} hits {s0['scan_alg_bytes']/max(s0['scan_ns'],1):.0f} GB/s | "
          f"K4 {s1['scan_ns']/1e6:.3f} ms {c1} checks {nh1} hits {s1['scan_alg_bytes']/max(s1['scan_ns'],1):.0f} GB/s")
