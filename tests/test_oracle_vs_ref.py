"""Differential test of the CPU oracle against the compiled reference itself (oracle/_ref, present only where
/root/reference was built - the container; skipped elsewhere, where the golden fixtures take over)."""
import numpy as np
import pytest

from harness import have_ref, run_oracle, run_ref, compare
from fuzz_oracle import OPTSETS, random_instance

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def test_fuzz_against_reference(gen):
    rng = np.random.default_rng(4242)
    for r in range(120):
        lits, sizes = random_instance(rng)
        stream = gen.csr_to_stream(lits, sizes)
        opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]
        a = run_oracle(stream, opts); b = run_ref(stream, opts)
        # ordered, with counters and undo stack - also for inputs with unit clauses (two-watched-literal trail order)
        assert compare(a, b) == [], (r, opts)


@pytest.mark.parametrize("opts", [["-enabled_cp3", "-subsimp"], ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]])
def test_planted_20k_against_reference(gen, opts):
    lits, sizes = gen.random_ksat(20000, 85200, 3, seed=99)
    stream = gen.csr_to_stream(lits, sizes)
    assert compare(run_oracle(stream, opts), run_ref(stream, opts)) == []


@pytest.mark.parametrize("opts", [["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"],
                                  ["-enabled_cp3", "-up", "-subsimp", "-bve", "-dense", "-no-cp3_limited"]])
def test_oracle_output_is_equisatisfiable_riss_cross_check(gen, opts):
    """satisfiability cross-checked by riss (oracle/_ref/riss-core): same verdict on input and simplified formula;
    a model of the simplified formula, extended (Dense map + undo stack), satisfies the input"""
    import satcheck
    from harness import Oracle
    if not satcheck.have_riss():
        pytest.skip("oracle/_ref/riss-core not built")
    verdicts = set()
    for seed in range(10):
        n = 120
        lits, sizes = gen.random_ksat(n, int(n * (3.9 + 0.1 * seed)), 3, seed=500 + seed, dup=0.03, sup=0.05, flip=0.05)
        stream = gen.csr_to_stream(lits, sizes)
        o = Oracle(opts); o.add(stream); o.preprocess(); r = o.result()
        rc0, _ = satcheck.solve(stream, n)
        if not r.ok:
            assert rc0 == 20
            o.close(); verdicts.add(20); continue
        simp = satcheck.simplified_stream(r)
        rc1, model = satcheck.solve(simp, max(r.nvars, 1))
        assert rc0 == rc1, (seed, rc0, rc1)
        verdicts.add(rc0)
        if rc1 == 10:
            full = int(o.L.cp3o_model_vars(o.h))
            buf = np.full(max(full, r.nvars, 1), 2, dtype=np.uint8)
            for l in model:
                if abs(l) <= r.nvars:
                    buf[abs(l) - 1] = 0 if l > 0 else 1            # oracle encoding: 0 = true
            buf[:r.nvars][buf[:r.nvars] == 2] = 0
            o.L.cp3o_extend_model(o.h, buf.ctypes.data, r.nvars)
            assert satcheck.satisfies(stream, (1 - buf[:max(full, n)]).astype(np.uint8)), seed
        o.close()
    assert verdicts == {10, 20}          # the sweep crosses the threshold: both verdicts were exercised
