"""Differential fuzz: CPU oracle (oracle/cp3_oracle.c) vs the compiled reference (oracle/_ref/refshim).
Run here (container with /root/reference built): python tests/fuzz_oracle.py [rounds] [seed]"""
import sys
import numpy as np
from harness import run_oracle, run_ref, compare, load_gen

gen = load_gen()

OPTSETS = [
    ["-enabled_cp3", "-subsimp"],
    ["-enabled_cp3", "-subsimp", "-no-cp3_strength"],
    ["-enabled_cp3", "-up", "-subsimp", "-no-cp3_limited"],
    ["-enabled_cp3", "-subsimp", "-bve"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"],
    ["-enabled_cp3", "-subsimp", "-bve", "-no-bve_gates"],
    ["-enabled_cp3", "-subsimp", "-bve", "-no-bve_gates", "-bve_cgrow=0", "-bve_cgrow_t=0", "-no-cp3_limited"],
    ["-enabled_cp3", "-bve", "-bve_BCElim"],
    ["-enabled_cp3", "-subsimp", "-bve", "-bve_unlimited", "-bve_cgrow=2"],
    ["-enabled_cp3", "-subsimp", "-bve", "-no-bve_strength"],
    ["-enabled_cp3", "-subsimp", "-bve", "-cp3_bve_heap=1"],
    ["-enabled_cp3", "-subsimp", "-bve", "-bve_heap_updates=0"],
    ["-enabled_cp3", "-subsimp", "-bve", "-cp3_bve_heap=3"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-cp3_bve_heap=6", "-no-cp3_limited"],
    ["-enabled_cp3", "-subsimp", "-bve", "-cp3_bve_heap=9", "-bve_heap_updates=2"],
    ["-enabled_cp3", "-subsimp", "-bve", "-bve_early"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-cp3_iters=2"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-dense", "-no-cp3_limited"],
    ["-enabled_cp3", "-subsimp", "-dense"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-dense", "-cp3_dense_frag=20"],
    # small step budgets (Stepper limits, Technique.h:205-221): the passes stop in the middle of their queues
    ["-enabled_cp3", "-subsimp", "-cp3_sub_limit=50", "-cp3_str_limit=80", "-cp3_call_inc=10"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-cp3_bve_limit=40", "-cp3_sub_limit=300", "-cp3_str_limit=200", "-cp3_call_inc=20"],
    # size gates (Coprocessor.cc:1002-1006): sized to fire on part of the random instances
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-cp3_lits=400"],
    ["-enabled_cp3", "-subsimp", "-cp3_cls=150", "-cp3_vars=40"],
    # naive strengthening (fullStrengthening, Subsumption.cc:1056-1197)
    ["-enabled_cp3", "-subsimp", "-naive_strength"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-naive_strength", "-no-cp3_limited"],
    # stand-alone blocked clause elimination (BCE.cc), alone and behind the other techniques as in Riss427
    ["-enabled_cp3", "-bce"],
    ["-enabled_cp3", "-up", "-subsimp", "-bve", "-bce", "-no-cp3_limited"],
    ["-enabled_cp3", "-bce", "-bce-bin", "-no-bce-compl"],
    ["-enabled_cp3", "-subsimp", "-bce", "-bce-limit=30", "-cp3_iters=2"],
    ["-enabled_cp3", "-bve", "-bce", "-dense", "-cp3_bce_vars=3", "-cp3_bce_cls=3", "-cp3_bce_lits=3"],
]


def random_instance(rng):
    kind = rng.integers(0, 4)
    n = int(rng.integers(4, 60))
    if kind == 0:      # tiny dense mixed-size
        m = int(rng.integers(3, 6 * n))
        sizes = rng.integers(1 if rng.random() < 0.3 else 2, 5, size=m)
        lits = []
        for s in sizes:
            vs = rng.choice(n, size=min(int(s), n), replace=False) + 1
            lits.append((vs * (rng.integers(0, 2, size=vs.size) * 2 - 1)).astype(np.int32))
        sizes = np.array([l.size for l in lits], dtype=np.int32)
        lits = np.concatenate(lits)
        return lits, sizes
    if kind == 1:      # planted 3-SAT
        n = int(rng.integers(20, 400)); m = int(n * rng.uniform(2.0, 5.0))
        return gen.random_ksat(n, m, 3, seed=int(rng.integers(1 << 30)), dup=0.05, sup=0.08, flip=0.08)
    if kind == 2:      # binary-heavy (unit cascades, gates)
        n = int(rng.integers(6, 80)); m = int(n * rng.uniform(1.0, 3.5))
        l2, s2 = gen.random_ksat(n, m, 2, seed=int(rng.integers(1 << 30)), dup=0.05, sup=0.1, flip=0.1)
        l3, s3 = gen.random_ksat(n, max(1, m // 2), 3, seed=int(rng.integers(1 << 30)), planted=False)
        L = np.concatenate([l2, l3]); S = np.concatenate([s2, s3])
        return gen.permute_clauses(rng, L, S)
    n = int(rng.integers(30, 300)); m = int(n * rng.uniform(2.5, 5))
    return gen.community_ksat(n, m, seed=int(rng.integers(1 << 30)))


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    bad = 0
    for r in range(rounds):
        lits, sizes = random_instance(rng)
        stream = gen.csr_to_stream(lits, sizes)
        has_units = bool((sizes == 1).any())
        opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]
        a = run_oracle(stream, opts)
        b = run_ref(stream, opts)
        d = compare(a, b)          # inputs with unit clauses included: the trail order of Solver::propagate is reproduced
        if d:
            bad += 1
            fn = f"/tmp/fuzz_fail_{seed}_{r}.bin"
            stream.tofile(fn)
            print(f"[{r}] MISMATCH opts={' '.join(opts)} clauses={sizes.size} units={has_units} -> {fn}")
            for x in d[:6]:
                print("     ", x)
            if bad >= 5:
                break
    print(f"done: {rounds} rounds, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
