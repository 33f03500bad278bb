"""A/B of two library builds on the configs[1] fixpoint: CP3B200_LIB=<other .so> python scripts/ab_c2.py [reps] (phase timers from CPstat)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
lits, sizes = rs.gen.random_ksat(1000000, 4260000, 3, seed=12345); stream = rs.gen.csr_to_stream(lits, sizes)
ts = []
for rep in range(reps):
    cp = rs.Coprocessor(["-enabled_cp3", "-subsimp", "-no-cp3_limited"])
    t = time.time(); cp.add_stream(stream); ta = time.time() - t
    t = time.time(); cp.preprocess(); tp = time.time() - t
    ts.append(tp)
    print(os.environ.get("CP3B200_LIB", "new"), f"rep {rep} add {ta*1e3:.1f} ms preprocess {tp*1e3:.1f} ms", {k: round(cp.stat(k) / 1e6, 1) for k in ["ph_init_ns", "ph_sub_ns", "ph_str_ns", "host_gpuwait_ns", "host_commit_ns"]}, flush=True)
    cp.close()
print(os.environ.get("CP3B200_LIB", "new"), "min", round(min(ts) * 1e3, 1), "median", round(sorted(ts)[len(ts) // 2] * 1e3, 1))
