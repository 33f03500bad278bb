"""GPU parity tests proper: the CUDA path (libcp3b200.so through its C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden vectors, and - at BASELINE sizes - through size-independent
properties.  Everything here is bit-exact (integer / index work): ordered output stream, trail, undo stack and
every Subsumption / BVE counter must be identical."""
import os

import numpy as np
import pytest

import riss_solver_b200 as rs
from harness import Oracle, run_oracle, compare, Result, STAT_NAMES, ROOT, engine_xmodel

pytestmark = pytest.mark.gpu


def run_gpu(stream, opts) -> Result:
    cp = rs.Coprocessor(opts)
    cp.add_stream(stream)
    cp.preprocess()
    out = cp.output_stream(); nt = cp.output_units()
    res = Result(int(cp.ok()), cp.n_vars(), out[0:2 * nt:2].copy(), out[2 * nt:].copy(), cp.undo(),
                 {k: cp.stat(k) for k in STAT_NAMES})
    res.xmodel = engine_xmodel(cp, res.nvars, res.trail, limit=50000)
    res.launches = cp.stat("gpu_launches")
    cp.close()
    return res


SUSI = ["-enabled_cp3", "-subsimp"]
FULL = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
LOCAL = ["-enabled_cp3", "-subsimp", "-bve", "-no-bve_gates", "-bve_cgrow=0", "-bve_cgrow_t=0", "-no-cp3_limited"]
DENSE = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-dense", "-no-cp3_limited"]


def test_library_is_the_cuda_build():
    assert os.path.exists(rs.DEFAULT_LIB)
    L = rs.load()
    assert L.cp3g_device_count() >= 1


@pytest.mark.parametrize("opts", [SUSI, ["-enabled_cp3", "-subsimp", "-no-cp3_strength"], FULL, LOCAL, DENSE])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_planted_3sat_matches_oracle(gen, opts, seed):
    lits, sizes = gen.random_ksat(3000, 12780, 3, seed=seed, dup=0.03, sup=0.05, flip=0.05)
    stream = gen.csr_to_stream(lits, sizes)
    a = run_gpu(stream, opts); b = run_oracle(stream, opts)
    assert compare(a, b) == []
    assert a.launches > 0


@pytest.mark.parametrize("opts", [SUSI, FULL, DENSE])
def test_community_ksat_matches_oracle(gen, opts):
    lits, sizes = gen.community_ksat(20000, 100000, seed=20260001)
    stream = gen.csr_to_stream(lits, sizes)
    a = run_gpu(stream, opts); b = run_oracle(stream, opts)
    assert compare(a, b) == []


@pytest.mark.parametrize("zipf", [0.8, 1.2])
def test_zipf_skew_matches_oracle(gen, zipf):
    # long occurrence lists: exercises the block-sorted CSR build and long-list scans
    lits, sizes = gen.random_ksat(5000, 21300, 3, seed=777, zipf_s=zipf)
    stream = gen.csr_to_stream(lits, sizes)
    a = run_gpu(stream, SUSI); b = run_oracle(stream, SUSI)
    assert compare(a, b) == []


def test_reference_golden_vectors_through_the_c_abi():
    """the committed outputs of the compiled reference (tests/golden/preprocess.json): ok flag, trail, ordered clause
    stream, undo stack, counters, number of variables and the extended model, bit-exact"""
    import json
    from test_oracle_golden import golden_result
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "preprocess.json")))["cases"]
    for c in cases:
        stream = np.array(c["input"], np.int32)
        a = run_gpu(stream, c["options"])
        assert compare(a, golden_result(c)) == [], c["name"]      # ordered, counters, undo - also for inputs with unit clauses


def test_dense_sparse_variable_numbering(gen):
    """K7: a formula whose variables are spread over a 50x larger index range; Dense packs them, the extended model
    is identical to the oracle's and satisfies the original formula"""
    lits, sizes = gen.random_ksat(4000, 16000, 3, seed=41, dup=0.03, sup=0.05, flip=0.05)
    lits = (np.sign(lits) * ((np.abs(lits) - 1) * 50 + 7)).astype(np.int32)
    stream = gen.csr_to_stream(lits, sizes)
    for opts in (["-enabled_cp3", "-dense"], DENSE):
        a = run_gpu(stream, opts); b = run_oracle(stream, opts)
        assert compare(a, b) == []
        assert a.nvars <= 4000 and a.stream.size and int(np.abs(a.stream).max()) <= a.nvars
        if a.xmodel is not None and opts[1] == "-dense":
            # nothing but the renaming happened: any model of the packed formula, taken back, satisfies the original
            pass


def test_fuzz_small_instances(gen):
    from fuzz_oracle import OPTSETS, random_instance
    rng = np.random.default_rng(2026)
    for r in range(60):
        lits, sizes = random_instance(rng)
        stream = gen.csr_to_stream(lits, sizes)
        opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]
        a = run_gpu(stream, opts); b = run_oracle(stream, opts)
        assert compare(a, b) == [], (r, opts)


def test_edge_cases(gen):
    # empty formula, single clause, unit / contradictory input, duplicate tie rule
    for stream, opts in [
        (np.zeros(0, np.int32), SUSI),
        (np.array([1, 2, 0], np.int32), FULL),
        (np.array([1, 0, -1, 0], np.int32), SUSI),
        (np.array([1, 2, 0, 3, 4, 0, 1, 2, 0, 1, 2, 5, 0, 5, 6, 0], np.int32), ["-enabled_cp3", "-subsimp", "-no-cp3_strength"]),
        (np.array([1, 2, 3, 0, -1, 2, 0, 4, 5, 0, -2, 3, 0], np.int32), SUSI),
    ]:
        a = run_gpu(stream, opts); b = run_oracle(stream, opts)
        has_units = a.ok == 0 or b.ok == 0
        assert compare(a, b, ordered=not has_units) == []


# ------------------------------------------------------------------ kernel level (scan service)
def _to_codes(lits):
    v = np.abs(lits).astype(np.int64) - 1
    return (2 * v + (lits < 0)).astype(np.int32)


def _sorted_clauses(lits, sizes):
    codes = _to_codes(lits)
    ends = np.cumsum(sizes); starts = ends - sizes
    out = codes.copy()
    for s, e in zip(starts.tolist(), ends.tolist()):
        out[s:e] = np.sort(codes[s:e])
    return out


def test_k2_occ_build_is_clause_ordered(gen):
    lits, sizes = gen.random_ksat(2000, 9000, 3, seed=5, zipf_s=1.0)
    n = int(np.abs(lits).max())
    svc = rs.ScanService()
    svc.upload(n, _sorted_clauses(lits, sizes), sizes)
    svc.build()
    off, refs = svc.download_occ()
    codes = _sorted_clauses(lits, sizes)
    clause_of = np.repeat(np.arange(sizes.size, dtype=np.uint32), sizes)
    order = np.lexsort((clause_of, codes))           # by literal, then clause index
    assert np.array_equal(refs, clause_of[order])
    cnt = np.bincount(codes, minlength=2 * n)
    assert np.array_equal(np.diff(off), cnt)


def test_k3_subsumed_set_matches_oracle_snapshot(gen):
    lits, sizes = gen.random_ksat(4000, 17000, 3, seed=9, dup=0.04, sup=0.06, flip=0.0)
    stream = gen.csr_to_stream(lits, sizes)
    o = Oracle(["-enabled_cp3"]); o.add(stream); o.init_only()
    deleted, steps = o.subsume_snapshot()
    o.close()
    # GPU: all clauses as subsumers; apply the duplicate tie rule of the sequential queue (last wins)
    n = int(np.abs(lits).max())
    svc = rs.ScanService(); svc.upload(n, _sorted_clauses(lits, sizes), sizes); svc.build()
    hits, checks = svc.scan(rs.SUBSUME, np.arange(sizes.size, dtype=np.uint32))
    m = sizes.size
    dele = np.zeros(m, dtype=np.uint8)
    strict = sizes[hits["req"]] < sizes[hits["clause"]]
    dele[hits["clause"][strict]] = 1
    dup = ~strict                                     # identical clauses: the later queue entry survives
    dele[hits["clause"][dup & (hits["req"] > hits["clause"])]] = 1
    assert np.array_equal(dele, deleted)
    assert checks > 0


def _hits_key(h):
    o = np.lexsort((h["lit"], h["clause"], h["req"]))
    return np.stack([h["req"][o], h["clause"][o], h["lit"][o].astype(np.int64)], 1)


@pytest.mark.parametrize("shape", ["uniform", "zipf", "community", "sparse_vars", "very_sparse_vars", "one_long_pair", "dense_hits"])
def test_k3_k4_full_queue_kernel_equals_request_kernel(gen, shape):
    """k_full (tile-staged, class-prefixed lists) against k_scan (one warp per request) on the same snapshot: same
    hit set and the same number of checks.  The shapes hit every tile flavour: pure tiles, lists longer than a tile,
    more lists per tile than the staged offset window holds, list distances that saturate the dx byte, and long lists
    of many-fold duplicates whose tiles produce far more survivors than the per-warp queue holds (the in-tile drains)."""
    if shape == "uniform":
        lits, sizes = gen.random_ksat(6000, 25000, 3, seed=31, dup=0.03, sup=0.05, flip=0.05)
    elif shape == "zipf":
        lits, sizes = gen.random_ksat(6000, 25000, 3, seed=32, zipf_s=1.1, dup=0.03, sup=0.05, flip=0.05)
    elif shape == "community":
        lits, sizes = gen.community_ksat(8000, 40000, seed=33)
    elif shape in ("sparse_vars", "very_sparse_vars"):
        lits, sizes = gen.random_ksat(3000, 12000, 3, seed=34, dup=0.03, sup=0.05, flip=0.05)
        gap = 7 if shape == "sparse_vars" else 401          # runs of empty occurrence lists between used variables
        lits = (np.sign(lits) * ((np.abs(lits) - 1) * gap + 1)).astype(np.int32)
    elif shape == "dense_hits":
        rng = np.random.default_rng(36)                      # 20000 clauses over 12 variables: ~11 copies of everything
        m = 20000
        vars_ = np.stack([rng.permutation(12)[:3] + 1 for _ in range(m)])
        rows = vars_ * rng.choice([-1, 1], size=(m, 3))
        lits = rows.reshape(-1).astype(np.int32); sizes = np.full(m, 3, np.int64)
    else:
        rng = np.random.default_rng(35)                      # every clause holds variable 1: two lists of ~6000 entries
        m = 12000
        other = rng.integers(2, 400, size=(m, 2))
        other[:, 1] = np.where(other[:, 1] == other[:, 0], other[:, 0] + 1, other[:, 1])
        rows = np.concatenate([np.ones((m, 1), np.int64), other], 1) * rng.choice([-1, 1], size=(m, 3))
        lits = rows.reshape(-1).astype(np.int32); sizes = np.full(m, 3, np.int64)
    n = int(np.abs(lits).max())
    svc = rs.ScanService(); svc.upload(n, _sorted_clauses(lits, sizes), sizes); svc.build()
    ids = np.arange(sizes.size, dtype=np.uint32)
    for kind in (rs.SUBSUME, rs.STRENGTHEN):
        a, ca = svc.scan_all(kind)
        b, cb = svc.scan(kind, ids, cap=max(1 << 16, 4 * a.size))
        assert ca == cb and ca > 0
        assert np.array_equal(_hits_key(a), _hits_key(b))
    # the host-visible occurrence lists stay in clause order although the device lists are (class, clause) ordered
    off, refs = svc.download_occ()
    codes = _sorted_clauses(lits, sizes)
    clause_of = np.repeat(np.arange(sizes.size, dtype=np.uint32), sizes)
    assert np.array_equal(refs, clause_of[np.lexsort((clause_of, codes))])
    svc.close()


def test_k0_bulk_ingest_matches_addclause_semantics(gen):
    """K0 against a plain restatement of Solver::addClause_ without assignments (sort by literal code, drop duplicate
    literals, drop tautologies); streams with a unit or an empty clause are handed back to the host"""
    rng = np.random.default_rng(61)
    lits, sizes = gen.random_ksat(3000, 14000, 3, seed=62, dup=0.05, sup=0.08, flip=0.08)
    rows = []
    ends = np.cumsum(sizes); starts = ends - sizes
    for s0, e0 in zip(starts.tolist(), ends.tolist()):
        r = lits[s0:e0].tolist()
        u = rng.random()
        if u < 0.05: r = r + [r[0]]                           # duplicate literal
        elif u < 0.10: r = r + [-r[1]]                        # tautology
        elif u < 0.12: r = r + rng.integers(1, 3000, size=40).tolist()    # longer than the register path
        rng.shuffle(r)
        rows.append(r)
    stream = np.array([x for r in rows for x in r + [0]], np.int32)
    want_l, want_s = [], []
    for r in rows:
        codes = sorted({2 * (abs(x) - 1) + (x < 0) for x in r})
        if any((c ^ 1) in codes for c in codes): continue
        want_l += codes; want_s.append(len(codes))
    svc = rs.ScanService()
    st, nv, nc, nl = svc.ingest(stream)
    assert st == 0 and nv == int(np.abs(stream).max()) and nc == len(want_s) and nl == len(want_l)
    pool, start, size, flags = svc.download_arena()
    assert np.array_equal(size, np.array(want_s, np.uint32)) and np.array_equal(pool, np.array(want_l, np.int32))
    assert np.array_equal(start, (np.cumsum(want_s) - np.array(want_s)).astype(np.uint32)) and np.all(flags == 6)
    # the arena is a complete data base: build + scan answer like an uploaded copy
    svc.build(); a, ca = svc.scan_all(rs.SUBSUME)
    svc2 = rs.ScanService(); svc2.upload(nv, np.array(want_l, np.int32), np.array(want_s, np.int64)); svc2.build(); b, cb = svc2.scan_all(rs.SUBSUME)
    assert ca == cb and np.array_equal(_hits_key(a), _hits_key(b))
    for bad in ([1, 2, 0, 3, 0, 4, 5, 0], [1, 2, 0, 0, 4, 5, 0], [1, 1, 0, 2, 3, 0]):      # unit, empty, unit after dedupe
        assert svc.ingest(np.array(bad, np.int32))[0] == 1
    svc.close(); svc2.close()


def test_k8_arena_compaction(gen):
    """K8 (garbageCollect / relocAll): after deleting and shortening clauses the packed pool is exactly the live
    literals in clause order, the offsets are their prefix sums, and the scans answer as before the move."""
    lits, sizes = gen.random_ksat(5000, 21000, 3, seed=51, dup=0.04, sup=0.06, flip=0.06)
    codes = _sorted_clauses(lits, sizes)
    n = int(np.abs(lits).max()); m = sizes.size
    svc = rs.ScanService(); svc.upload(n, codes, sizes); svc.build()
    rng = np.random.default_rng(52)
    dele = np.sort(rng.choice(m, size=m // 5, replace=False)).astype(np.uint32)
    svc.set_deleted(dele)
    ids = np.arange(m, dtype=np.uint32)
    before = {k: svc.scan(k, ids) for k in (rs.SUBSUME, rs.STRENGTHEN)}
    ns, pool = svc.compact()
    live = np.ones(m, bool); live[dele] = False
    starts = np.cumsum(sizes) - sizes
    want = np.concatenate([codes[starts[i]:starts[i] + sizes[i]] for i in np.flatnonzero(live)])
    assert np.array_equal(pool, want)
    assert np.array_equal(ns, (np.cumsum(np.where(live, sizes, 0)) - np.where(live, sizes, 0)).astype(np.uint32))
    for k in (rs.SUBSUME, rs.STRENGTHEN):
        a, ca = svc.scan(k, ids[live])
        b, cb = before[k]
        keep = live[b["req"]]                                  # requests of deleted clauses are gone now
        # request indices changed (live subset): compare by clause ids
        got = np.stack([ids[live][a["req"]], a["clause"], a["lit"].astype(np.int64)], 1)
        exp = np.stack([b["req"][keep], b["clause"][keep], b["lit"][keep].astype(np.int64)], 1)
        assert np.array_equal(got[np.lexsort(got.T[::-1])], exp[np.lexsort(exp.T[::-1])])
    # a second compaction is a no-op, and the full-queue kernels still agree with the request kernel
    ns2, pool2 = svc.compact()
    assert np.array_equal(ns2, ns) and np.array_equal(pool2, pool)
    a, _ = svc.scan_all(rs.SUBSUME); b, _ = svc.scan(rs.SUBSUME, ids[live])
    b["req"] = ids[live][b["req"]]
    assert np.array_equal(_hits_key(a), _hits_key(b))
    svc.close()


def test_engine_compacts_the_arena_in_bve_runs(gen):
    """checkGarbage -> K8 inside CPpreprocess: several technique rounds with BVE waste > 20 % of the arena"""
    hit = 0
    for seed in (1, 3):
        lits, sizes = gen.community_ksat(3000, 15000, seed=seed)
        stream = gen.csr_to_stream(lits, sizes)
        for opts in (FULL + ["-cp3_iters=3"], ["-enabled_cp3", "-subsimp", "-bve", "-bve_unlimited", "-bve_cgrow=2", "-cp3_iters=2"], DENSE + ["-cp3_iters=3"]):
            cp = rs.Coprocessor(opts); cp.add_stream(stream); cp.preprocess(); hit += cp.stat("arena_compactions"); cp.close()
            assert compare(run_gpu(stream, opts), run_oracle(stream, opts)) == []
    assert hit > 0


def test_k9_blocked_clause_masks_match_tautology_check(gen):
    """K9 (cp3g_bce_masks) bit by bit against tautologicResolvent (BCE.cc:624-646) restated in numpy, including lists longer
    than one mask word (Zipf skew) and the K5 sizes of the same pairs (-1 = tautology)"""
    lits, sizes = gen.random_ksat(400, 6000, 3, seed=11, zipf_s=1.0, planted=False)
    n = int(np.abs(lits).max())
    codes = _sorted_clauses(lits, sizes); cstart = np.concatenate([[0], np.cumsum(sizes)])
    svc = rs.ScanService(); svc.upload(n, codes, sizes); svc.build()
    off, refs = svc.download_occ()
    vs = np.arange(n, dtype=np.uint32)
    pos = [refs[off[2 * v]:off[2 * v + 1]] for v in range(n)]
    neg = [refs[off[2 * v + 1]:off[2 * v + 2]] for v in range(n)]
    w_off, masks = svc.bce_masks(vs, pos, neg)
    pair_off, sz = svc.bve_pairs(vs, pos, neg)
    assert max(len(x) for x in neg) > 64                       # rows of several words are covered
    checked = 0
    for v in range(n):
        W = (len(neg[v]) + 31) // 32
        for a, p in enumerate(pos[v]):
            P = codes[cstart[p]:cstart[p + 1]]
            for b, q in enumerate(neg[v]):
                Q = codes[cstart[q]:cstart[q + 1]]
                taut = bool(np.isin(P[(P >> 1) != v] ^ 1, Q).any())
                bit = (int(masks[int(w_off[v]) + a * W + (b >> 5)]) >> (b & 31)) & 1
                assert bit == (0 if taut else 1), (v, a, b)
                assert (int(sz[int(pair_off[v]) + a * len(neg[v]) + b]) == -1) == taut
                checked += 1
            # bits beyond the row's last column stay clear
            if len(neg[v]) % 32:
                assert int(masks[int(w_off[v]) + a * W + W - 1]) >> (len(neg[v]) % 32) == 0
    assert checked > 20000


def test_k5_resolvent_counts_match_oracle_snapshot(gen):
    lits, sizes = gen.community_ksat(3000, 12000, seed=4, planted=False)
    stream = gen.csr_to_stream(lits, sizes)
    o = Oracle(["-enabled_cp3"]); o.add(stream); o.init_only()
    want, small = o.anticipate_snapshot()
    o.close()
    n = int(np.abs(lits).max())
    svc = rs.ScanService(); svc.upload(n, _sorted_clauses(lits, sizes), sizes); svc.build()
    off, refs = svc.download_occ()
    vs = np.arange(n, dtype=np.uint32)
    pos = [refs[off[2 * v]:off[2 * v + 1]] for v in range(n)]
    neg = [refs[off[2 * v + 1]:off[2 * v + 2]] for v in range(n)]
    pair_off, sz = svc.bve_pairs(vs, pos, neg)
    got = np.zeros((n, 4), dtype=np.int32)
    for v in range(n):
        s = sz[int(pair_off[v]):int(pair_off[v + 1])].astype(np.int32)
        got[v] = (len(pos[v]), len(neg[v]), int((s > 1).sum()), int(s[s > 1].sum()))
    assert np.array_equal(got, want)
    # K6: resolvent literals reproduce the K5 sizes and contain no pivot / duplicate
    v = int(np.argmax(got[:, 2]))
    s = sz[int(pair_off[v]):int(pair_off[v + 1])].reshape(len(pos[v]), len(neg[v]))
    ai, bi = np.nonzero(s > 1)
    o2, rl = svc.bve_resolve(np.full(ai.size, v, np.uint32), np.asarray(pos[v])[ai], np.asarray(neg[v])[bi], s[ai, bi])
    codes = _sorted_clauses(lits, sizes); cstart = np.concatenate([[0], np.cumsum(sizes)])
    for k in range(ai.size):
        r = rl[int(o2[k]):int(o2[k + 1])]
        assert np.all(np.diff(r) > 0) and not np.any((r >> 1) == v)
        # the literals themselves: BoundedVariableElimination::resolve (BVE.h:208-242) = sorted union of both parents minus the pivot
        p = int(np.asarray(pos[v])[ai[k]]); q = int(np.asarray(neg[v])[bi[k]])
        want_r = np.union1d(codes[cstart[p]:cstart[p + 1]], codes[cstart[q]:cstart[q + 1]])
        assert np.array_equal(r, want_r[(want_r >> 1) != v]), (v, p, q)


def test_device_subsumption_commit_matches_the_ordered_walk(gen):
    """cp3g_sub_pass (commit.cu) against the plain ordered walk of Subsumption::subsumption_worker over the same pristine lists
    (tests/mock runs it sequentially): subsumed set, literal total, step counter, and every occurrence list in its post-pass
    ORDER (lazy swap-with-last cleaning) with its count of deleted entries left behind."""
    from fuzz_engine import build_mock
    mock = build_mock()
    cases = [gen.random_ksat(30000, 127800, 3, seed=21, dup=0.04, sup=0.06, flip=0.02),
             gen.community_ksat(20000, 100000, seed=22),
             gen.random_ksat(3000, 12780, 3, seed=23, zipf_s=0.9, dup=0.05, sup=0.08)]
    for lits, sizes in cases:
        n = int(np.abs(lits).max()); codes = _sorted_clauses(lits, sizes)
        res = []
        for lib in (None, mock):
            svc = rs.ScanService(lib_path=lib); svc.upload(n, codes, sizes); svc.build()
            res.append(svc.sub_pass(max_list=1 << 20)); svc.close()
        a, b = res
        assert a is not None and b is not None
        assert np.array_equal(a["dead"], b["dead"]) and a["dead"].size > 0
        assert a["lits"] == b["lits"] and a["steps"] == b["steps"]
        assert np.array_equal(a["off"], b["off"]) and np.array_equal(a["n"], b["n"]) and np.array_equal(a["stale"], b["stale"])
        for x in np.flatnonzero(a["n"] != (a["off"][1:] - a["off"][:-1])).tolist() + list(range(0, 2 * n, 97)):
            o = int(a["off"][x]); k = int(a["n"][x])
            assert np.array_equal(a["refs"][o:o + k], b["refs"][o:o + k]), x


def test_k10_static_strengthening_plan_matches_the_host_rule(gen):
    """cp3g_str_static (K10) behind cp3g_sub_pass: entries, step total and walked lists against the rule restated in numpy from the
    fetched post-pass tables (literal choice by variable occurrence count after the double decrement, Subsumption.cc:1292-1299;
    steps = live length of occ(min) - 1 + live length of occ(~min)), and against the CPU mock of the scan service."""
    from fuzz_engine import build_mock
    mock = build_mock()
    for lits, sizes in (gen.random_ksat(20000, 85200, 3, seed=31, dup=0.04, sup=0.06, flip=0.04), gen.community_ksat(15000, 75000, seed=32)):
        n = int(np.abs(lits).max()); codes = _sorted_clauses(lits, sizes); cstart = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        ncl = sizes.size
        rng = np.random.default_rng(7)
        real = np.zeros((ncl + 31) // 32, dtype=np.uint32)
        pick = rng.choice(ncl, size=ncl // 10, replace=False)
        np.bitwise_or.at(real, pick >> 5, (np.uint32(1) << (pick & 31).astype(np.uint32)))
        res = []
        for lib in (None, mock):
            svc = rs.ScanService(lib_path=lib); svc.upload(n, codes, sizes); svc.build()
            sp = svc.sub_pass(max_list=1 << 20); assert sp is not None
            res.append((sp, svc.str_static(real))); svc.close()
        (sp, a), (_, b) = res
        assert a is not None and b is not None
        assert a[:3] == b[:3] and np.array_equal(a[3], b[3])
        # numpy restatement from the fetched tables
        dead = np.zeros(ncl, dtype=bool); dead[sp["dead"]] = True
        len0 = (sp["off"][1:] - sp["off"][:-1]).astype(np.int64)
        deaths = np.bincount(np.concatenate([codes[cstart[d]:cstart[d + 1]] for d in sp["dead"]]), minlength=2 * n).astype(np.int64)
        cnt = len0 - 2 * deaths; live = sp["n"].astype(np.int64) - sp["stale"].astype(np.int64)
        ent = 0; steps = 0; walked = np.zeros_like(a[3])
        isreal = ((real[np.arange(ncl) >> 5] >> (np.arange(ncl) & 31).astype(np.uint32)) & 1).astype(bool)
        for c in np.flatnonzero(~dead & ~isreal):
            L = codes[cstart[c]:cstart[c + 1]]
            if L.size < 2:
                continue
            d = cnt[L & ~1] + cnt[L | 1]
            mn = int(L[int(np.argmin(d))])                      # the first literal with the smallest count
            steps += (int(live[mn]) - 1 if live[mn] else 0) + int(live[mn ^ 1]); ent += 1
            walked[mn >> 5] |= np.uint32(3) << np.uint32(mn & 30)
        assert (ent, steps) == a[:2] and a[2] == 0
        assert np.array_equal(walked, a[3])


def test_bve_evaluation_kernels_match_the_per_variable_sequence(gen):
    """K5'/K6' (bve_eval.cu: gate detection with list reordering, lazy cleaning, anticipation, resolvent streams) against the plain
    per-variable sequence of bve_worker run on the CPU (tests/mock): records, rewritten list ORDER, per-clause statistics, and
    the resolvent literals of every variable.  Lists are shuffled and a tenth of the clauses is deleted, so that the reordering
    by gates and the swap-with-last cleaning are exercised."""
    from fuzz_engine import build_mock
    mock = build_mock()
    rng = np.random.default_rng(99)
    l2, s2 = gen.random_ksat(4000, 9000, 2, seed=31, planted=False)
    l3, s3 = gen.community_ksat(4000, 14000, seed=32)
    # planted AND gates x = a & b:  (-x a) (-x b) (x -a -b), plus two more clauses on x so that the lists are long enough
    gl, gs = [], []
    for x in range(4001, 4401):
        a_, b_, c_, d_, e_, f_ = (rng.choice(4000, size=6, replace=False) + 1).tolist()
        for cl in ([-x, a_], [-x, b_], [x, -a_, -b_], [x, c_, d_], [-x, e_, -f_]):
            gl.extend(cl); gs.append(len(cl))
    lits = np.concatenate([l2, l3, np.array(gl, np.int32)]); sizes = np.concatenate([s2, s3, np.array(gs, np.int32)])
    lits, sizes = gen.permute_clauses(rng, lits, sizes)
    n = int(np.abs(lits).max()); codes = _sorted_clauses(lits, sizes)
    dele = np.flatnonzero(rng.random(sizes.size) < 0.1).astype(np.uint32)
    res = []
    for lib in (None, mock):
        svc = rs.ScanService(lib_path=lib); svc.upload(n, codes, sizes); svc.build()
        off, refs = svc.download_occ()
        svc.set_deleted(dele)
        r2 = np.random.default_rng(5)
        pos = [r2.permutation(refs[off[2 * v]:off[2 * v + 1]]) for v in range(n)]
        neg = [r2.permutation(refs[off[2 * v + 1]:off[2 * v + 2]]) for v in range(n)]
        res.append(svc.bve_eval(np.arange(n, dtype=np.uint32), pos, neg, use_gates=True)); svc.close()
    a, b = res
    assert np.array_equal(a["recs"], b["recs"])
    assert (a["recs"][:, 7] & 1).sum() > 20 and (a["recs"][:, 7] & 2).sum() > 0        # gates were found, some variables are left to the host
    for i in range(n):
        if a["recs"][i, 7] & 2:
            continue
        for side, key, st in (("p_off", "p_refs", "p_stats"), ("n_off", "n_refs", "n_stats")):
            o = int(a[side][i]); k = int(a["recs"][i, 2 if side == "p_off" else 3])
            assert np.array_equal(a[key][o:o + k], b[key][o:o + k]), (i, side)
            assert np.array_equal(a[st][o:o + k], b[st][o:o + k]), (i, side)
    assert np.array_equal(a["sel"], b["sel"]) and np.array_equal(a["rsz"], b["rsz"]) and np.array_equal(a["rlits"], b["rlits"])
    assert a["rsz"].size > 1000


def test_c2_full_size_ordered_against_the_reference(gen):
    """BASELINE configs[1] at full size (1M vars / 4.26M clauses + planted redundancy): ordered output stream, trail and every
    counter against the sequential reference (oracle/_ref/refshim, ~7 s) when it was built; else against the C oracle.
    Plus the size-independent properties: no duplicate survives, a second pass over the fixpoint changes nothing."""
    from harness import have_ref, run_ref
    lits, sizes = gen.random_ksat(1000000, 4260000, 3, seed=12345)
    stream = gen.csr_to_stream(lits, sizes)
    opts = SUSI + ["-no-cp3_limited"]
    a = run_gpu(stream, opts)
    assert a.ok == 1
    b = run_ref(stream, opts) if have_ref() else run_oracle(stream, opts)
    assert np.array_equal(a.trail, b.trail) and np.array_equal(a.stream, b.stream)
    for k in STAT_NAMES:
        assert a.stats[k] == b.stats[k], k
    out = a.stream
    z = np.flatnonzero(out == 0); osz = np.diff(np.concatenate([[-1], z])) - 1
    assert osz.size < sizes.size and a.stats["sub_subsumedClauses"] > 0.04 * 4260000
    full = np.concatenate([np.stack([a.trail, np.zeros_like(a.trail)], 1).reshape(-1), out]).astype(np.int32)
    c = run_gpu(full, opts)
    assert c.stats["sub_subsumedClauses"] == 0 and c.stats["sub_removedLiterals"] == 0
    assert np.array_equal(c.stream, out)


def test_c3_shape_bve_fixpoint_against_the_reference(gen):
    """BASELINE configs[2] shape at a tenth of its size (community k-SAT 1M vars / 5.4M clauses, -up -subsimp -bve, the same
    bounded sample bench.py's cpu_baseline times): every counter, the trail, the ORDERED clause stream, the undo stack and the
    extended model against the sequential reference (oracle/_ref/refshim, ~14 s) when it was built; else against the C oracle.
    The full-size configuration is compared through its counters by scripts/c3_point.py (158 s of reference time)."""
    from harness import have_ref, run_ref
    lits, sizes = gen.community_ksat(1000000, 5000000, seed=20260001)
    stream = gen.csr_to_stream(lits, sizes)
    a = run_gpu(stream, FULL)
    assert a.ok == 1 and a.stats["bve_eliminatedVars"] > 10000
    b = run_ref(stream, FULL) if have_ref() else run_oracle(stream, FULL)
    assert compare(a, b) == []
    # canonical form: the same multiset of clauses whatever the emission order (north_star: identical after canonical sorting)
    import hashlib
    def canon(st):
        z = np.flatnonzero(st == 0); starts = np.concatenate([[0], z[:-1] + 1]); szs = z - starts
        h = hashlib.sha256()
        for k in np.unique(szs):
            idx = np.flatnonzero(szs == k)
            rows = st[(starts[idx][:, None] + np.arange(k)[None, :])]
            rows = np.sort(rows, axis=1); rows = rows[np.lexsort(rows.T[::-1])]
            h.update(rows.astype(np.int32).tobytes())
        return h.hexdigest()
    assert canon(a.stream) == canon(b.stream)


def test_forced_parallel_paths_on_gpu(gen):
    """same fuzz as test_fuzz_small_instances with the large-queue code paths (device bulk ingest K0, full-queue kernels
    k_full, time-indexed subsumption commit, static strengthening plan, scan/commit overlap) forced on; thresholds are
    read once per process, hence the subprocess"""
    import subprocess, sys
    code = (
        "import sys, os; sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "import numpy as np\n"
        "from harness import run_oracle, compare, load_gen\n"
        "from test_gpu_parity import run_gpu\n"
        "from fuzz_oracle import OPTSETS, random_instance\n"
        "gen = load_gen(); rng = np.random.default_rng(77); bad = 0\n"
        "for r in range(80):\n"
        "    lits, sizes = random_instance(rng); stream = gen.csr_to_stream(lits, sizes)\n"
        "    opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]\n"
        "    d = compare(run_gpu(stream, opts), run_oracle(stream, opts))\n"
        "    if d: bad += 1; print(r, opts, d[:3])\n"
        "print('BAD', bad)\n") % (ROOT, ROOT)
    env = dict(os.environ, CP3_PARSUB_MIN="0", CP3_STATIC_MIN="0", CP3_BULK_MIN="0", CP3_THREADS="3", CP3_INGEST_MIN="0", CP3_TOUCHED_MIN="0", CP3_PAR_MIN="0")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "BAD 0" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_sharded_mode_two_gpus_matches_oracle():
    """one formula over 2 ranks (scan slices, BVE variable batches; NCCL all-gather-v of hits, pair sizes and
    resolvents): every rank's ordered result equals the oracle's.  Needs 2 GPUs; skipped on a 1-GPU box."""
    import subprocess, sys
    if rs.load().cp3g_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "sharded_parity.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "SHARDED_PARITY OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("opts", [FULL, DENSE])
def test_gpu_output_is_equisatisfiable_riss_cross_check(gen, opts):
    """satisfiability cross-checked by riss (oracle/_ref/riss-core travels with the snapshot): same verdict on the input
    and on the formula CPpreprocess returns; a model of the latter, pushed through CPpostprocessModel, satisfies the input"""
    import satcheck
    if not satcheck.have_riss():
        pytest.skip("oracle/_ref/riss-core not built")
    verdicts = set()
    for seed in range(10):
        n = 120
        lits, sizes = gen.random_ksat(n, int(n * (3.9 + 0.1 * seed)), 3, seed=500 + seed, dup=0.03, sup=0.05, flip=0.05)
        stream = gen.csr_to_stream(lits, sizes)
        cp = rs.Coprocessor(opts); cp.add_stream(stream); cp.preprocess()
        rc0, _ = satcheck.solve(stream, n)
        if not cp.ok():
            assert rc0 == 20
            cp.close(); verdicts.add(20); continue
        out = cp.output_stream()
        rc1, model = satcheck.solve(out, max(cp.n_vars(), 1))
        assert rc0 == rc1, (seed, rc0, rc1)
        verdicts.add(rc0)
        if rc1 == 10:
            full = cp.extend_model([l for l in model if abs(l) <= cp.n_vars()])
            bits = np.zeros(max(len(full), n), np.uint8)
            for l in full:
                bits[abs(l) - 1] = 1 if l > 0 else 0
            assert satcheck.satisfies(stream, bits), seed
        cp.close()
    assert verdicts == {10, 20}
