"""SAT/UNSAT cross-check with the reference solver (oracle/_ref/riss-core, built from the unmodified sources):
the simplified formula is equisatisfiable with the input, and a model of it - extended through the undo stack
(and the Dense map) - satisfies the input.  Test infrastructure only."""
import os
import subprocess
import tempfile

import numpy as np

from harness import ROOT

RISS = os.path.join(ROOT, "oracle", "_ref", "riss-core")


def have_riss():
    return os.path.exists(RISS)


def _write(path, stream, nvars):
    s = np.asarray(stream, dtype=np.int64)
    ncl = int((s == 0).sum())
    with open(path, "w") as f:
        f.write(f"p cnf {nvars} {ncl}\n")
        line = []
        for x in s.tolist():
            if x == 0:
                f.write(" ".join(line) + " 0\n"); line = []
            else:
                line.append(str(x))


def solve(stream, nvars):
    """-> (10 | 20, model as DIMACS literal list or None)"""
    with tempfile.TemporaryDirectory() as td:
        cnf = os.path.join(td, "f.cnf"); out = os.path.join(td, "f.out")
        _write(cnf, stream, nvars)
        p = subprocess.run([RISS, cnf, out, "-verb=0"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        model = None
        if p.returncode == 10:
            model = []
            for line in open(out):
                if line.startswith("v "):
                    model += [int(x) for x in line[2:].split() if x != "0"]
        return p.returncode, model


def satisfies(stream, model_bits):
    """model_bits[v] = 1 true / 0 false for variable v+1"""
    ok, cur = True, False
    for x in np.asarray(stream).tolist():
        if x == 0:
            ok = ok and cur; cur = False
        else:
            cur = cur or (model_bits[abs(x) - 1] == (1 if x > 0 else 0))
    return ok


def simplified_stream(res):
    """trail units + clauses, as the C API emits them"""
    units = np.stack([res.trail, np.zeros_like(res.trail)], 1).reshape(-1) if res.trail.size else np.zeros(0, np.int32)
    return np.concatenate([units, res.stream]).astype(np.int32)
