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
        has_units = bool((sizes == 1).any())
        opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]
        a = run_oracle(stream, opts); b = run_ref(stream, opts)
        assert compare(a, b, ordered=not has_units, stats=not has_units, undo=not has_units) == [], (r, opts)


@pytest.mark.parametrize("opts", [["-enabled_cp3", "-subsimp"], ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]])
def test_planted_20k_against_reference(gen, opts):
    lits, sizes = gen.random_ksat(20000, 85200, 3, seed=99)
    stream = gen.csr_to_stream(lits, sizes)
    assert compare(run_oracle(stream, opts), run_ref(stream, opts)) == []
