"""Per-kernel view of the device-resident pass at BASELINE configs[2] size (run under ncu for the launch list)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
from bench import to_codes_sorted
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
lits, sizes = rs.gen.community_ksat(n, 5 * n)
codes = to_codes_sorted(lits, sizes)
svc = rs.ScanService(0); svc.upload(int(np.abs(lits).max()), codes, sizes)
for rep in range(2):
    k0 = svc.counter("kernel_ns"); svc.build(); b = svc.counter("kernel_ns") - k0
    s0 = svc.counter("scan_ns"); svc.scan_all(rs.SUBSUME, count_only=True); svc.scan_all(rs.STRENGTHEN, count_only=True)
    print(f"rep {rep}: build {b/1e6:.2f} ms, K3+K4 {(svc.counter('scan_ns')-s0)/1e6:.2f} ms", flush=True)
