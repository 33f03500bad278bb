"""Host logic of the product (ordered-commit engine + C ABI, riss-solver_b200/csrc/engine*.cpp, capi.cpp) linked
against the CPU mock of the scan service (tests/mock) - no GPU needed.  The mock answers what the kernels answer,
so these tests pin the commit order, the bookkeeping quirks and the speculative-closure cache against the oracle."""
import json
import os

import numpy as np
import pytest

import riss_solver_b200 as rs
from harness import run_oracle, compare, Result
from fuzz_engine import build_mock, run_engine
from fuzz_oracle import OPTSETS, random_instance
from test_oracle_golden import GOLD, golden_result


@pytest.fixture(scope="module")
def mock():
    return build_mock()


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_engine_matches_reference_golden(mock, case):
    got = run_engine(np.array(case["input"], np.int32), case["options"], mock)
    assert compare(got, golden_result(case)) == []      # ordered - also for inputs with unit clauses


def _clauses(stream):
    out, cur = [], []
    for x in stream:
        if x == 0:
            out.append(cur); cur = []
        else:
            cur.append(x)
    return out


@pytest.mark.parametrize("bulk_min", ["0", "4096", "forced"])
def test_engine_fuzz_against_oracle(mock, gen, bulk_min, monkeypatch):
    # CP3_BULK_MIN=0 forces the full-queue (bulk) scan path on tiny formulas as well; "forced" switches every large-queue form on
    # (time-indexed subsumption commit, static plan + walk plans, parallel id sort / cache build / occurrence updates, 4 threads)
    if bulk_min == "forced":
        for k in ("CP3_BULK_MIN", "CP3_PARSUB_MIN", "CP3_STATIC_MIN", "CP3_PAR_MIN", "CP3_TOUCHED_MIN"):
            monkeypatch.setenv(k, "0")
        monkeypatch.setenv("CP3_THREADS", "4"); bulk_min = "7"
    else:
        monkeypatch.setenv("CP3_BULK_MIN", bulk_min)
    rng = np.random.default_rng(31337 + int(bulk_min))
    for r in range(150):
        lits, sizes = random_instance(rng)
        stream = gen.csr_to_stream(lits, sizes)
        opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]
        a = run_engine(stream, opts, mock); b = run_oracle(stream, opts)
        assert compare(a, b) == [], (r, opts)


def test_engine_planted_50k(mock, gen):
    lits, sizes = gen.random_ksat(50000, 213000, 3, seed=5)
    stream = gen.csr_to_stream(lits, sizes)
    for opts in (["-enabled_cp3", "-subsimp"], ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]):
        a = run_engine(stream, opts, mock); b = run_oracle(stream, opts)
        assert compare(a, b) == []
    # the speculative closure must have answered almost every re-scan of a strengthened clause
    assert a.extra["scan_ondemand"] < 50


def test_cp_api_surface(mock):
    """literal-by-literal API, output iteration, counters, model extension (libcoprocessorc.h call sequence)"""
    cp = rs.Coprocessor(lib_path=mock)
    cp.parse_options("-enabled_cp3 -subsimp -bve")
    for x in [-1, 2, 0, -2, 3, 0, -3, 1, 0, 2, 7, 6, 0, 2, -7, -6, 0, -6, 7, 0, -7, 6, 0]:   # regression/cnfs/sat.cnf
        cp.add_lit(x)
    assert cp.n_vars() == 7 and cp.n_clauses() == 7
    cp.preprocess()
    assert cp.ok() and cp.n_clauses() == 0                # everything eliminated
    assert list(cp.iter_output()) == []
    model = cp.extend_model([])
    assert len(model) == 7
    sat = lambda cl: any((l > 0) == (model[abs(l) - 1] > 0) for l in cl)
    for cl in [[-1, 2], [-2, 3], [-3, 1], [2, 7, 6], [2, -7, -6], [-6, 7], [-7, 6]]:
        assert sat(cl), (cl, model)
    cp.close()


def test_unknown_and_unsupported_options(mock):
    cp = rs.Coprocessor(lib_path=mock)
    cp.parse_options("-enabled_cp3 -subsimp -this_flag_does_not_exist -cp3_threads=4")   # ignored like the reference
    cp.add_stream(np.array([1, 2, 0, 1, 2, 3, 0], np.int32))
    cp.preprocess()
    assert cp.ok() and cp.n_clauses() == 1
    cp.close()


def test_freeze_blocks_elimination(mock):
    stream = np.array([-1, 2, 0, -2, 3, 0, -3, 1, 0, 2, 7, 6, 0, 2, -7, -6, 0, -6, 7, 0, -7, 6, 0], np.int32)
    from harness import Oracle
    o = Oracle(["-enabled_cp3", "-bve"]); o.add(stream)
    for v in (1, 2, 3):
        o.freeze(v)
    o.preprocess(); want = o.result(); o.close()
    cp = rs.Coprocessor(["-enabled_cp3", "-bve"], lib_path=mock)
    cp.add_stream(stream)
    for v in (1, 2, 3):
        cp.freeze(v)
    cp.preprocess()
    out = cp.output_stream()
    assert np.array_equal(out[2 * cp.output_units():], want.stream)
    cp.close()


FORCED = {"CP3_PARSUB_MIN": "0", "CP3_STATIC_MIN": "0", "CP3_BULK_MIN": "0", "CP3_THREADS": "3", "CP3_INGEST_MIN": "0", "CP3_TOUCHED_MIN": "0"}


def test_forced_parallel_paths_fuzz(mock):
    """The data-parallel commit paths (time-indexed subsumption commit, static strengthening plan with up-front list
    cleaning + rollback, full-queue bulk scans, scan/commit overlap) only switch on for large queues; force them on
    tiny formulas (the thresholds are read once per process, hence the subprocess) and fuzz against the oracle."""
    import subprocess, sys
    env = dict(os.environ, **FORCED)
    p = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz_engine.py"), "250", "9001"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "0 mismatches" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_literal_by_literal_clients_get_the_bulk_ingest(mock, gen):
    """CPaddLit into an empty solver is buffered and delivered as one stream (K0) by the first observing call; the
    result equals the literal-by-literal ingest, tautologies / duplicate literals / a trailing partial clause included"""
    import subprocess, sys
    code = (
        "import sys, os; sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "import numpy as np\n"
        "import riss_solver_b200 as rs\n"
        "from harness import load_gen\n"
        "gen = load_gen(); lib = %r\n"
        "lits, sizes = gen.random_ksat(400, 1700, 3, seed=8, dup=0.05, sup=0.05, flip=0.05)\n"
        "stream = np.concatenate([gen.csr_to_stream(lits, sizes), np.array([5, -5, 6, 0, 7, 7, 8, 0, 9, 10], np.int32)])\n"
        "outs = []\n"
        "for mode in ('lit', 'stream'):\n"
        "    cp = rs.Coprocessor(['-enabled_cp3', '-up', '-subsimp', '-bve'], lib_path=lib)\n"
        "    if mode == 'lit':\n"
        "        for x in stream.tolist(): cp.add_lit(x)\n"
        "    else: cp.add_stream(stream)\n"
        "    cp.add_lit(11); cp.add_lit(0)                      # completes the trailing clause (9 10 11)\n"
        "    n0 = cp.n_vars(); cp.preprocess()\n"
        "    outs.append((n0, cp.stat('bulk_ingests'), cp.output_stream().tolist(), cp.undo().tolist())); cp.close()\n"
        "assert outs[0][1] == 1 and outs[1][1] == 1, (outs[0][1], outs[1][1])\n"
        "assert outs[0][0] == outs[1][0] and outs[0][2] == outs[1][2] and outs[0][3] == outs[1][3]\n"
        "print('INGEST_OK')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.dirname(os.path.abspath(__file__))), mock)
    env = dict(os.environ, CP3_INGEST_MIN="0", CP3_TOUCHED_MIN="0")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "INGEST_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
