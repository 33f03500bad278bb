import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
opts = sys.argv[2:] or ["-enabled_cp3", "-subsimp", "-no-cp3_limited"]
t = time.time(); lits, sizes = rs.gen.random_ksat(n, int(4.26 * n), 3, seed=12345); stream = rs.gen.csr_to_stream(lits, sizes); print(f"gen {time.time()-t:.2f}s clauses {sizes.size}")
for rep in range(3):
    cp = rs.Coprocessor(opts)
    print(f"=== rep {rep}", file=sys.stderr, flush=True)
    t = time.time(); cp.add_stream(stream); t_add = time.time() - t
    t = time.time(); cp.preprocess(); t_pp = time.time() - t
    names = ["sub_subsSteps", "sub_strengthSteps", "sub_subsumedClauses", "sub_removedLiterals", "bve_eliminatedVars", "bve_steps", "scan_batches", "scan_ondemand",
             "queue_fast", "queue_slow", "spec_requests", "spec_rounds", "spec_used", "ph_total_ns", "ph_simplify_ns", "ph_init_ns", "ph_upload_ns", "ph_download_ns", "queue_static", "ph_sub_ns", "ph_str_ns", "ph_prefetch_ns", "pair_batches", "resolve_batches", "host_commit_ns", "host_gpuwait_ns", "gpu_launches", "gpu_kernel_ns", "gpu_h2d_bytes", "gpu_d2h_bytes", "gpu_scan_checks"]
    st = {k: cp.stat(k) for k in names}
    print(f"rep {rep}: add {t_add:.3f}s preprocess {t_pp:.3f}s")
    print("   ", {k: (round(v / 1e6, 1) if k.endswith("_ns") else v) for k, v in st.items()})
    cp.close()
