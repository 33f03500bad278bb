import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import riss_solver_b200 as rs
from harness import run_oracle, compare, load_gen
from fuzz_engine import run_engine
gen = load_gen()
FULL = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
rng = np.random.default_rng(3)
cases = {}
cases["empty"] = np.zeros(0, np.int32)
cases["only units"] = np.array([1, 0, -2, 0, 3, 0], np.int32)
cases["one huge clause + 3sat"] = np.concatenate([np.arange(1, 6001, dtype=np.int32), [0], gen.csr_to_stream(*gen.random_ksat(6000, 20000, 3, seed=5))]).astype(np.int32)
l, s = gen.random_ksat(2000, 9000, 3, seed=6); cases["sparse vars up to 4M"] = gen.csr_to_stream((np.sign(l) * (np.abs(l) * 2000)).astype(np.int32), s)
cases["tautologies + dup lits"] = np.concatenate([np.array([1, -1, 2, 0, 3, 3, 4, 0], np.int32), gen.csr_to_stream(*gen.random_ksat(300, 1300, 3, seed=7))]).astype(np.int32)
cases["all binary"] = gen.csr_to_stream(*gen.random_ksat(500, 800, 2, seed=8))
cases["wide clauses k=9"] = gen.csr_to_stream(*gen.random_ksat(3000, 9000, 9, seed=9))
bad = 0
for name, st in cases.items():
    for opts in (FULL, FULL + ["-dense"], FULL + ["-bce"]):
        a = run_engine(st, opts); b = run_oracle(st, opts)
        hu = bool(((np.diff(np.concatenate([[-1], np.flatnonzero(st == 0)])) - 1) == 1).any()) if st.size else False
        d = compare(a, b)
        bad += bool(d)
        print(name, opts[-1], "OK" if not d else d[:3], flush=True)
print("ODD", "OK" if not bad else "FAILED")
