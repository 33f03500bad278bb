"""A/B of two library builds on the configs[2] fixpoint: CP3B200_LIB=<other .so> python scripts/ab_c3.py [n_vars] (3 reps, phase timers and batch counters)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
lits, sizes = rs.gen.community_ksat(n, 5 * n); stream = rs.gen.csr_to_stream(lits, sizes)
for rep in range(3):
    cp = rs.Coprocessor(["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"])
    t = time.time(); cp.add_stream(stream); ta = time.time() - t
    t = time.time(); cp.preprocess(); tp = time.time() - t
    print(os.environ.get("CP3B200_LIB", "new"), f"rep {rep} add {ta:.2f}s preprocess {tp:.2f}s", {k: round(cp.stat(k) / 1e6, 1) for k in ["ph_init_ns", "ph_sub_ns", "ph_str_ns", "host_gpuwait_ns", "host_commit_ns", "ph_total_ns"]},
          {k: cp.stat(k) for k in ["gpu_launches", "scan_batches", "scan_ondemand", "eval_batches", "eval_used", "eval_stale", "eval_host", "pair_batches", "resolve_batches", "gpu_build_count", "arena_compactions", "gpu_kernel_ns"]}, flush=True)
    cp.close()
