"""Test harness: ctypes bindings for the CPU oracle, a runner for the compiled reference (oracle/_ref),
and canonicalisation helpers.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REFSHIM = os.path.join(REF_DIR, "refshim")

sys.path.insert(0, ROOT)


def have_ref() -> bool:
    return os.path.exists(REFSHIM)


def build_oracle() -> str:
    so = os.path.join(ORACLE_DIR, "libcp3oracle.so")
    src = os.path.join(ORACLE_DIR, "cp3_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libcp3oracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.cp3o_new.restype = C.c_void_p
        L.cp3o_free.argtypes = [C.c_void_p]
        L.cp3o_option.argtypes = [C.c_void_p, C.c_char_p]
        L.cp3o_option.restype = C.c_int
        L.cp3o_add.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.cp3o_freeze.argtypes = [C.c_void_p, C.c_int]
        L.cp3o_preprocess.argtypes = [C.c_void_p]
        L.cp3o_init_only.argtypes = [C.c_void_p]
        L.cp3o_ok.argtypes = [C.c_void_p]
        L.cp3o_nvars.argtypes = [C.c_void_p]
        L.cp3o_output.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.cp3o_output.restype = C.c_size_t
        L.cp3o_ntrail.argtypes = [C.c_void_p]
        L.cp3o_ntrail.restype = C.c_size_t
        L.cp3o_nclauses.argtypes = [C.c_void_p]
        L.cp3o_nclauses.restype = C.c_size_t
        L.cp3o_undo.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.cp3o_undo.restype = C.c_size_t
        L.cp3o_model_vars.argtypes = [C.c_void_p]
        L.cp3o_extend_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.cp3o_extend_model.restype = None
        L.cp3o_stat.argtypes = [C.c_void_p, C.c_char_p]
        L.cp3o_stat.restype = C.c_int64
        L.cp3o_subsume_snapshot.argtypes = [C.c_void_p, C.c_void_p]
        L.cp3o_subsume_snapshot.restype = C.c_int64
        L.cp3o_anticipate_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


STAT_NAMES = ["sub_subsumedClauses", "sub_subsumedLiterals", "sub_removedLiterals", "sub_subsSteps",
              "sub_strengthSteps", "bve_eliminatedVars", "bve_createdClauses", "bve_removedClauses",
              "bve_testedVars", "bve_anticipations", "bve_skippedVars", "bve_unitsEnqueued",
              "bve_removedBC", "bve_totallyAddedClauses", "bve_steps", "up_removedClauses",
              "up_removedLiterals", "bce_steps", "bce_testedLits", "bce_remBCE"]


class Result:
    """ok flag, trail units, clause stream (both in emission order), undo stack, counters."""

    def __init__(self, ok, nvars, trail, stream, undo, stats, xmodel=None):
        self.ok, self.nvars, self.trail, self.stream, self.undo, self.stats = ok, nvars, trail, stream, undo, stats
        # Preprocessor::extendModel applied to pseudo_model(): 0/1 per ORIGINAL variable, or None when not computed
        self.xmodel = xmodel

    def clauses(self):
        out, cur = [], []
        for x in self.stream.tolist():
            if x == 0:
                out.append(tuple(cur)); cur = []
            else:
                cur.append(x)
        return out

    def canonical(self):
        """units first (sorted), then clauses with sorted literals, sorted"""
        return (sorted(self.trail.tolist()), sorted(tuple(sorted(c)) for c in self.clauses()))


def pseudo_model(nvars, trail):
    """fixed pseudo-random assignment (1 = true) of the simplified formula's variables that agrees with the trail;
    the same rule is hard-wired in oracle/refbuild/refshim.cc"""
    v = np.arange(1, nvars + 1, dtype=np.uint64)
    m = (((v * np.uint64(2654435761)) & np.uint64(0xffffffff)) >> np.uint64(13)) & np.uint64(1)
    m = m.astype(np.uint8)
    for x in np.asarray(trail).tolist():
        m[abs(x) - 1] = 1 if x > 0 else 0
    return m


def engine_xmodel(cp, nvars, trail, limit=200000):
    """CPpostprocessModel on pseudo_model() through the C ABI (skipped for big formulas: one ctypes call per literal)"""
    if nvars > limit or not cp.ok():
        return None
    m = pseudo_model(nvars, trail)
    full = cp.extend_model([(i + 1) if b else -(i + 1) for i, b in enumerate(m.tolist())])
    return np.array([1 if l > 0 else 0 for l in full], dtype=np.uint8)


class Oracle:
    def __init__(self, options=()):
        self.L = oracle_lib()
        self.h = C.c_void_p(self.L.cp3o_new())
        for o in options:
            if not self.L.cp3o_option(self.h, o.encode()):
                raise ValueError(f"oracle does not know option {o}")

    def add(self, stream):
        s = np.ascontiguousarray(stream, dtype=np.int32)
        self.L.cp3o_add(self.h, s.ctypes.data, s.size)

    def freeze(self, v):
        self.L.cp3o_freeze(self.h, v)

    def preprocess(self):
        self.L.cp3o_preprocess(self.h)

    def result(self) -> Result:
        L, h = self.L, self.h
        n = L.cp3o_output(h, None, 0)
        buf = np.zeros(max(n, 1), dtype=np.int32)
        L.cp3o_output(h, buf.ctypes.data, n)
        nt = L.cp3o_ntrail(h)
        trail = buf[0:2 * nt:2].copy()
        stream = buf[2 * nt:n].copy()
        nu = L.cp3o_undo(h, None, 0)
        undo = np.zeros(max(nu, 1), dtype=np.int32)
        L.cp3o_undo(h, undo.ctypes.data, nu)
        stats = {k: int(L.cp3o_stat(h, k.encode())) for k in STAT_NAMES}
        xm = None
        if L.cp3o_ok(h):
            nv = int(L.cp3o_nvars(h)); full = int(L.cp3o_model_vars(h))
            buf2 = np.full(max(full, nv, 1), 2, dtype=np.uint8)
            buf2[:nv] = 1 - pseudo_model(nv, trail)                 # oracle encoding: 0 = true, 1 = false
            L.cp3o_extend_model(h, buf2.ctypes.data, nv)
            xm = (1 - buf2[:max(full, nv)]).astype(np.uint8)
        return Result(int(L.cp3o_ok(h)), int(L.cp3o_nvars(h)), trail, stream, undo[:nu].copy(), stats, xm)

    def init_only(self):
        self.L.cp3o_init_only(self.h)

    def subsume_snapshot(self):
        n = self.L.cp3o_nclauses(self.h)
        d = np.zeros(max(n, 1), dtype=np.uint8)
        steps = self.L.cp3o_subsume_snapshot(self.h, d.ctypes.data)
        return d[:n], int(steps)

    def anticipate_snapshot(self):
        n = self.L.cp3o_nvars(self.h)
        out = np.zeros((max(n, 1), 4), dtype=np.int32)
        small = np.zeros(max(n, 1), dtype=np.int32)
        self.L.cp3o_anticipate_snapshot(self.h, out.ctypes.data, small.ctypes.data)
        return out[:n], small[:n]

    def close(self):
        if self.h:
            self.L.cp3o_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_oracle(stream, options) -> Result:
    o = Oracle(options)
    o.add(stream)
    o.preprocess()
    r = o.result()
    o.close()
    return r


def parse_shim(path) -> Result:
    ok, nvars, trail, undo, stats = 1, 0, np.zeros(0, np.int32), np.zeros(0, np.int32), {}
    lits = []; xmodel = None
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        t = lines[i].split()
        i += 1
        if not t:
            continue
        if t[0] == "ok":
            ok = int(t[1])
        elif t[0] == "vars":
            nvars = int(t[1])
        elif t[0] == "stat":
            try:
                stats[t[1]] = int(t[2])
            except ValueError:
                stats[t[1]] = float(t[2])
        elif t[0] == "seconds":
            stats["seconds"] = float(t[1])
        elif t[0] == "trail":
            trail = np.array([int(x) for x in t[2:]], dtype=np.int32)
        elif t[0] == "clauses":
            m = int(t[1])
            body = " ".join(lines[i:i + m])
            lits = np.array(body.split(), dtype=np.int32) if m else np.zeros(0, np.int32)
            i += m
        elif t[0] == "undo":
            undo = np.array([int(x) for x in t[2:]], dtype=np.int32)
        elif t[0] == "xmodel":
            xmodel = np.array([int(x) for x in t[2:]], dtype=np.uint8)
    return Result(ok, nvars, trail, np.asarray(lits, dtype=np.int32), undo, stats, xmodel)


def run_ref(stream, options, mode="run") -> Result:
    """Run the compiled, unmodified reference through oracle/_ref/refshim."""
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.bin")
        out = os.path.join(td, "out.txt")
        np.ascontiguousarray(stream, dtype=np.int32).tofile(inp)
        subprocess.check_call([REFSHIM, mode, inp, out] + list(options), stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        return parse_shim(out)


def run_ref_anticipate(stream, options=("-enabled_cp3",)):
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.bin")
        out = os.path.join(td, "out.txt")
        np.ascontiguousarray(stream, dtype=np.int32).tofile(inp)
        subprocess.check_call([REFSHIM, "anticipate", inp, out] + list(options), stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        rows = []
        with open(out) as f:
            for line in f:
                t = line.split("|")[0].split()
                if t and t[0] != "vars":
                    rows.append([int(x) for x in t[1:5]])
        return np.array(rows, dtype=np.int32)


def load_gen():
    import importlib.util
    spec = importlib.util.spec_from_file_location("cp3_gen", os.path.join(ROOT, "riss-solver_b200", "gen.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def compare(a: Result, b: Result, ordered=True, stats=True, undo=True):
    """Returns a list of human-readable differences (empty = identical)."""
    diffs = []
    if a.ok != b.ok:
        diffs.append(f"ok {a.ok} vs {b.ok}")
        return diffs
    if not a.ok:
        return diffs
    if a.nvars != b.nvars:
        diffs.append(f"number of variables {a.nvars} vs {b.nvars}")
    if ordered:
        if not np.array_equal(a.trail, b.trail):
            diffs.append(f"trail differs: {a.trail[:10]} vs {b.trail[:10]}")
        if not np.array_equal(a.stream, b.stream):
            diffs.append(f"ordered clause stream differs ({a.stream.size} vs {b.stream.size} ints)")
    if a.canonical() != b.canonical():
        diffs.append("canonical CNF differs")
    if undo and not np.array_equal(a.undo, b.undo):
        diffs.append(f"undo stack differs ({a.undo.size} vs {b.undo.size})")
    if undo and ordered and a.xmodel is not None and b.xmodel is not None and not np.array_equal(a.xmodel, b.xmodel):
        diffs.append(f"extended model differs ({a.xmodel.size} vs {b.xmodel.size} variables)")
    if stats:
        for k in STAT_NAMES:
            if k in a.stats and k in b.stats and a.stats[k] != b.stats[k]:
                diffs.append(f"stat {k}: {a.stats[k]} vs {b.stats[k]}")
    return diffs
