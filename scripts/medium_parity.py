"""Medium-size parity sweep of the CUDA path against the CPU oracle (ordered stream, counters, undo, model)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from harness import run_oracle, compare, load_gen
from fuzz_engine import run_engine
gen = load_gen()
FULL = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
cases = [("planted 3-SAT 100k", gen.random_ksat(100000, 426000, 3, seed=71), FULL),
         ("planted 3-SAT 100k iters=3 dense", gen.random_ksat(100000, 426000, 3, seed=72), FULL + ["-cp3_iters=3", "-dense"]),
         ("community 100k", gen.community_ksat(100000, 500000, seed=73), FULL),
         ("community 100k limited", gen.community_ksat(100000, 500000, seed=74), ["-enabled_cp3", "-subsimp", "-bve"]),
         ("zipf 0.9 60k susi", gen.random_ksat(60000, 255600, 3, seed=75, zipf_s=0.9), ["-enabled_cp3", "-subsimp", "-no-cp3_limited"]),
         ("zipf 0.7 60k full", gen.random_ksat(60000, 255600, 3, seed=76, zipf_s=0.7), FULL),
         ("4-SAT 80k", gen.random_ksat(80000, 700000, 4, seed=77), FULL + ["-bve_BCElim"]),
         ("community 100k riss427-style + bce", gen.community_ksat(100000, 500000, seed=78), FULL + ["-bce", "-dense"]),
         ("planted 3-SAT 100k bce only", gen.random_ksat(100000, 426000, 3, seed=79), ["-enabled_cp3", "-bce", "-no-cp3_limited"]),
         ("planted 3-SAT 100k naive strengthening", gen.random_ksat(100000, 426000, 3, seed=80), ["-enabled_cp3", "-subsimp", "-naive_strength", "-no-cp3_limited"]),
         ("community 100k heap order 5", gen.community_ksat(100000, 500000, seed=81), FULL + ["-cp3_bve_heap=5"])]
bad = 0
for name, (lits, sizes), opts in cases:
    stream = gen.csr_to_stream(lits, sizes)
    t = time.time(); a = run_engine(stream, opts); ta = time.time() - t
    t = time.time(); b = run_oracle(stream, opts); tb = time.time() - t
    d = compare(a, b); bad += bool(d)
    print(f"{name}: {'OK' if not d else d[:3]}  gpu path {ta:.2f}s oracle {tb:.2f}s  subsumed {a.stats['sub_subsumedClauses']} strengthened {a.stats['sub_removedLiterals']} eliminated {a.stats['bve_eliminatedVars']}", flush=True)
print("MEDIUM_PARITY", "OK" if not bad else "FAILED")
sys.exit(1 if bad else 0)
