import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
lits, sizes = rs.gen.random_ksat(1000000, 4260000, 3, seed=12345); stream = rs.gen.csr_to_stream(lits, sizes)
opts = ["-enabled_cp3", "-subsimp", "-no-cp3_limited"]
ts = []
for rep in range(4):
    cp = rs.Coprocessor(opts); cp.add_stream(stream)
    t = time.perf_counter(); cp.preprocess(); out = cp.output_stream(); dt = time.perf_counter() - t
    ts.append(dt)
    st = {k: round(cp.stat(k) / 1e6, 1) for k in ("ph_total_ns", "ph_init_ns", "ph_sub_ns", "ph_str_ns", "ph_prefetch_ns", "host_commit_ns", "host_gpuwait_ns")}
    cp.close()
print(os.environ.get("TAG", ""), "best %.3f s median %.3f s" % (min(ts), sorted(ts)[len(ts) // 2]), st)
