"""Phase breakdown of the configs[2] fixpoint (CP3_DEBUG_PAR=1 for the laps): two runs in one process, the second one warm.
usage: CP3_KEEP_HEAP=1 CP3_DEBUG_PAR=1 python scripts/c3_bve_breakdown.py [n_vars]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
lits, sizes = rs.gen.community_ksat(n, 5 * n); stream = rs.gen.csr_to_stream(lits, sizes)
for rep in range(2):
    print(f"===== rep {rep}", file=sys.stderr, flush=True)
    cp = rs.Coprocessor(["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"])
    t = time.time(); cp.add_stream(stream); ta = time.time() - t
    t = time.time(); cp.preprocess(); tp = time.time() - t
    print(f"rep {rep} add {ta:.2f}s preprocess {tp:.2f}s", {k: round(cp.stat(k) / 1e6, 1) for k in ["ph_init_ns", "ph_sub_ns", "ph_str_ns", "host_gpuwait_ns", "host_commit_ns", "ph_total_ns"]},
          {k: cp.stat(k) for k in ["bve_testedVars", "bve_eliminatedVars", "bve_anticipations", "bve_skippedVars", "bulk_ingests", "arena_compactions", "component_passes", "component_fallbacks", "device_static_passes"]}, file=sys.stderr, flush=True)
    cp.close()
