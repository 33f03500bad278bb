"""Generates tests/golden/*.json from the compiled, UNMODIFIED reference (oracle/_ref/refshim, built from
/root/reference by oracle/Makefile.ref).  Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin the CPU oracle (and, through it, the CUDA path) on machines where the reference is absent."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from harness import run_ref, run_ref_anticipate, load_gen, STAT_NAMES  # noqa: E402

gen = load_gen()


def dimacs(path):
    out = []
    for line in open(path):
        if line[0] in "pc":
            continue
        out += [int(x) for x in line.split()]
    return np.array(out, dtype=np.int32)


SUSI = ["-enabled_cp3", "-subsimp"]
NOSTR = ["-enabled_cp3", "-subsimp", "-no-cp3_strength"]
FULL = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
BVE = ["-enabled_cp3", "-subsimp", "-bve"]
LOCAL = ["-enabled_cp3", "-subsimp", "-bve", "-no-bve_gates", "-bve_cgrow=0", "-bve_cgrow_t=0", "-no-cp3_limited"]
DENSE = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-dense", "-no-cp3_limited"]      # Riss427-style: ... then Dense

cases = []


def add(name, stream, opts):
    r = run_ref(stream, opts)
    cases.append({"name": name, "options": opts, "input": np.asarray(stream).tolist(), "ok": r.ok, "nvars": r.nvars,
                  "trail": r.trail.tolist(), "stream": r.stream.tolist(), "undo": r.undo.tolist(),
                  "xmodel": None if r.xmodel is None else r.xmodel.tolist(),
                  "stats": {k: int(r.stats[k]) for k in STAT_NAMES if k in r.stats}})


ref_cnfs = "/root/reference/regression/cnfs"
for f in ("sat", "unsat"):
    for tag, o in (("susi", SUSI), ("bve", BVE), ("full", FULL), ("dense", DENSE)):
        add(f"regression_{f}_{tag}", dimacs(os.path.join(ref_cnfs, f + ".cnf")), o)
# the hand vectors of SURVEY.md 8(c): duplicate tie rule, order dependence of strengthening
add("dup_tie_last_survives", np.array([1, 2, 0, 3, 4, 0, 1, 2, 0, 1, 2, 5, 0, 5, 6, 0], np.int32), NOSTR)
add("strengthening_order_dependent", np.array([1, 2, 3, 0, -1, 2, 0, 4, 5, 0, -2, 3, 0], np.int32), SUSI)
add("double_decrement_counts", np.array([1, 2, 0, 1, 2, 3, 0, 3, 4, 0], np.int32), NOSTR)
for seed in (1, 2, 3):
    l, s = gen.random_ksat(300, 1278, 3, seed=seed, dup=0.04, sup=0.06, flip=0.06)
    st = gen.csr_to_stream(l, s)
    for tag, o in (("susi", SUSI), ("nostr", NOSTR), ("full", FULL), ("local", LOCAL), ("dense", DENSE)):
        add(f"planted3sat_s{seed}_{tag}", st, o)
l, s = gen.community_ksat(400, 1800, seed=20260001)
st = gen.csr_to_stream(l, s)
for tag, o in (("susi", SUSI), ("full", FULL), ("bve", BVE), ("dense", DENSE), ("dense_frag", DENSE + ["-cp3_dense_frag=30"])):
    add(f"community_{tag}", st, o)
# stand-alone blocked clause elimination (BCE.cc) and the Riss427-style pipeline around it
BCE = ["-enabled_cp3", "-bce"]
R427 = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-bce", "-dense", "-no-cp3_limited"]
add("regression_sat_bce", dimacs(os.path.join(ref_cnfs, "sat.cnf")), BCE)
add("community_bce", st, BCE)
add("community_riss427", st, R427)
l, s = gen.random_ksat(300, 1278, 3, seed=2, dup=0.04, sup=0.06, flip=0.06)
add("planted3sat_s2_bce", gen.csr_to_stream(l, s), BCE)
add("planted3sat_s2_bce_bin", gen.csr_to_stream(l, s), BCE + ["-bce-bin", "-no-bce-compl"])
# variable gaps in the input: Dense without any simplification
add("gaps_dense_only", np.array([3, -7, 0, 7, 12, -20, 0, -3, 20, 0], np.int32), ["-enabled_cp3", "-dense"])
l, s = gen.random_ksat(200, 900, 3, seed=777, zipf_s=1.1)
add("zipf_susi", gen.csr_to_stream(l, s), SUSI)
json.dump({"generator": "tests/golden/make_golden.py", "reference": "nmanthey/riss-solver via oracle/_ref/refshim", "cases": cases},
          open(os.path.join(HERE, "preprocess.json"), "w"))

# per-variable resolvent counts on a pristine snapshot (anticipateElimination, BVE.cc:695)
ant = []
for seed in (4, 5):
    l, s = gen.community_ksat(300, 1300, seed=seed, planted=False)
    st = gen.csr_to_stream(l, s)
    rows = run_ref_anticipate(st)
    ant.append({"name": f"community_s{seed}", "input": st.tolist(), "rows": rows.tolist()})
json.dump({"generator": "tests/golden/make_golden.py", "columns": ["pos", "neg", "resolvents", "lit_clauses"], "cases": ant},
          open(os.path.join(HERE, "anticipate.json"), "w"))
print(len(cases), "preprocess cases,", len(ant), "anticipate cases")
