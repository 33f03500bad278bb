"""Differential fuzz of the ordered-commit host engine (linked against the CPU mock of the scan service,
tests/mock) vs the CPU oracle.  python tests/fuzz_engine.py [rounds] [seed]"""
import os, subprocess, sys
import numpy as np
from harness import run_oracle, compare, load_gen, Result, STAT_NAMES, ROOT, engine_xmodel
import riss_solver_b200 as rs
from fuzz_oracle import OPTSETS, random_instance

MOCK = os.path.join(ROOT, "tests", "_build", "libcp3b200_mock.so")


def build_mock():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    return MOCK


def run_engine(stream, opts, lib=None) -> Result:
    cp = rs.Coprocessor(opts, lib_path=lib)
    cp.add_stream(stream)
    cp.preprocess()
    out = cp.output_stream(); nt = cp.output_units()
    res = Result(int(cp.ok()), cp.n_vars(), out[0:2 * nt:2].copy(), out[2 * nt:].copy(), cp.undo(),
                 {k: cp.stat(k) for k in STAT_NAMES})
    res.xmodel = engine_xmodel(cp, res.nvars, res.trail)
    res.extra = {k: cp.stat(k) for k in ("scan_batches", "scan_ondemand", "queue_fast", "queue_slow", "pair_batches", "resolve_batches", "gpu_launches")}
    cp.close()
    return res


def main():
    gen = load_gen()
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = build_mock()
    rng = np.random.default_rng(seed)
    bad = 0
    for r in range(rounds):
        lits, sizes = random_instance(rng)
        stream = gen.csr_to_stream(lits, sizes)
        opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]
        a = run_engine(stream, opts, lib)
        b = run_oracle(stream, opts)
        d = compare(a, b)
        if d:
            bad += 1
            fn = f"/tmp/engine_fail_{seed}_{r}.bin"
            stream.tofile(fn)
            print(f"[{r}] MISMATCH opts={' '.join(opts)} clauses={sizes.size} -> {fn}")
            for x in d[:6]:
                print("     ", x)
            if bad >= 5:
                break
    print(f"done: {rounds} rounds, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
