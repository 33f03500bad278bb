"""The drop-in claim, checked literally: the SAME ctypes script of libcoprocessorc.h calls (coprocessor/libcoprocessorc.h:46-133)
is run against the unmodified reference library (oracle/_ref/libriss-coprocessor.so, built from /root/reference by
oracle/Makefile.ref) and against the product's C ABI; every observable value must be equal.

Covered here and nowhere else against the REFERENCE itself (not the C port): literal-by-literal CPaddLit, CPfreezeVariable,
the CPhasNextOutputLit / CPnextOutputLit loop, CPnVars / CPnClss / CPnLits / CPok / CPlitFalsified, CPgetReplaceLiteral
(with -dense), the CPpushModelBool -> CPpostprocessModel -> CPgetFinalModelLit round trip, CPgiveModelLit for positive
literals, step budgets and size gates through CPparseOptionString.

`-m "not gpu"`: product = host engine + C ABI linked against the CPU mock of the scan service (tests/mock).
`-m gpu`: product = libcp3b200.so (the CUDA path)."""
import ctypes as C
import os

import numpy as np
import pytest

import riss_solver_b200 as rs
from harness import REF_DIR, load_gen, pseudo_model
from fuzz_engine import build_mock
from fuzz_oracle import random_instance

REFLIB = os.path.join(REF_DIR, "libriss-coprocessor.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref/libriss-coprocessor.so not built")

OPTS = [
    "-enabled_cp3 -subsimp",
    "-enabled_cp3 -up -subsimp -bve",
    "-enabled_cp3 -up -subsimp -bve -no-cp3_limited",
    "-enabled_cp3 -up -subsimp -bve -dense",
    "-enabled_cp3 -subsimp -bve -no-bve_gates -bve_BCElim",
    "-enabled_cp3 -subsimp -cp3_sub_limit=50 -cp3_str_limit=80 -cp3_call_inc=10",
    "-enabled_cp3 -up -subsimp -bve -cp3_bve_limit=40",
    "-enabled_cp3 -up -subsimp -bve -cp3_lits=400",
    "-enabled_cp3 -subsimp -bve -cp3_cls=150",
]


def _bind(path):
    """the prototypes of libcoprocessorc.h - valid for both libraries"""
    L = C.CDLL(path)
    vp = C.c_void_p
    L.CPinit.restype = vp; L.CPinit.argtypes = [C.c_char_p]
    L.CPdestroy.argtypes = [C.POINTER(vp)]; L.CPdestroy.restype = None
    L.CPparseOptionString.argtypes = [vp, C.c_char_p]; L.CPparseOptionString.restype = None
    for f in ("CPpreprocess", "CPresetOutput", "CPresetModel", "CPpostprocessModel"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = None
    for f in ("CPhasNextOutputLit", "CPnextOutputLit", "CPgetFinalModelLit", "CPmodelVariables", "CPnVars", "CPnClss", "CPnLits", "CPok"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = C.c_int
    for f in ("CPaddLit", "CPfreezeVariable", "CPpushModelBool", "CPgiveModelLit"):
        getattr(L, f).argtypes = [vp, C.c_int]; getattr(L, f).restype = None
    for f in ("CPgetReplaceLiteral", "CPlitFalsified"):
        getattr(L, f).argtypes = [vp, C.c_int]; getattr(L, f).restype = C.c_int
    return L


def script(L, stream, opts, freeze, give_lits=False, via_init=False):
    """one client session; returns everything a client can observe.  via_init: the option string goes to CPinit (an unknown
    preset name is parsed as options, riss/utils/Config.h:736-739) - the only way cp3_limited and the step limits take
    effect through this interface, they are read by the constructors inside CPinit."""
    obs = {}
    if via_init:
        h = C.c_void_p(L.CPinit(opts.encode()))
    else:
        h = C.c_void_p(L.CPinit(None))
        L.CPparseOptionString(h, opts.encode())
    for x in stream.tolist():
        L.CPaddLit(h, int(x))
    obs["before"] = (L.CPnVars(h), L.CPnClss(h), L.CPnLits(h), L.CPok(h))
    for v in freeze:
        L.CPfreezeVariable(h, int(v))
    L.CPfreezeVariable(h, 0)                                   # freezeExtern(0) is a no-op
    L.CPpreprocess(h)
    obs["after"] = (L.CPnVars(h), L.CPnClss(h), L.CPnLits(h), L.CPok(h))
    out = []
    if L.CPok(h):
        L.CPresetOutput(h)
        while L.CPhasNextOutputLit(h):
            out.append(L.CPnextOutputLit(h))
    obs["output"] = out
    nv0 = int(np.abs(stream).max()) if stream.size else 0
    obs["falsified"] = [L.CPlitFalsified(h, s * v) for v in range(1, min(nv0, L.CPnVars(h)) + 1) for s in (1, -1)] if L.CPok(h) and "-dense" not in opts else []
    obs["replace"] = [L.CPgetReplaceLiteral(h, s * v) for v in range(1, nv0 + 1) for s in (1, -1)]
    if L.CPok(h):
        # a model of the simplified formula's variables: fixed pseudo-random bits that agree with the unit clauses
        units = []
        i = 0
        while i + 1 < len(out) and out[i + 1] == 0 and out[i] != 0:
            units.append(out[i]); i += 2
        nv = L.CPnVars(h)
        m = pseudo_model(nv, [u for u in units if abs(u) <= nv])
        L.CPresetModel(h)
        if give_lits:                                          # CPgiveModelLit: identical for positive literals
            for v in range(nv, 0, -1):
                if m[v - 1]:
                    L.CPgiveModelLit(h, v)
            obs["model_in"] = L.CPmodelVariables(h)
        else:
            for b in m.tolist():
                L.CPpushModelBool(h, int(b))
            L.CPpostprocessModel(h)
            n = L.CPmodelVariables(h)
            obs["model"] = [L.CPgetFinalModelLit(h) for _ in range(n)]
            obs["model_end"] = L.CPgetFinalModelLit(h)
    L.CPdestroy(C.byref(h))
    obs["destroyed"] = h.value is None
    return obs


def _instances(n, seed):
    gen = load_gen()
    rng = np.random.default_rng(seed)
    for r in range(n):
        lits, sizes = random_instance(rng)
        stream = gen.csr_to_stream(lits, sizes)
        nv = int(np.abs(lits).max())
        freeze = sorted(set(int(v) for v in rng.integers(1, nv + 1, size=int(rng.integers(0, 4)))))
        yield r, stream, OPTS[int(rng.integers(0, len(OPTS)))], freeze


def _run(product_path, n, seed):
    ref = _bind(REFLIB)
    prod = _bind(product_path)
    for r, stream, opts, freeze in _instances(n, seed):
        for via_init in (False, True):
            a = script(prod, stream, opts, freeze, via_init=via_init); b = script(ref, stream, opts, freeze, via_init=via_init)
            for k in b:
                assert a[k] == b[k], (r, opts, freeze, k, via_init)
        if r % 5 == 0:
            a = script(prod, stream, opts, freeze, give_lits=True); b = script(ref, stream, opts, freeze, give_lits=True)
            assert a.get("model_in") == b.get("model_in"), (r, opts, "CPgiveModelLit")


def test_cp_api_script_matches_reference_host_engine():
    _run(build_mock(), 60, 20261)


@pytest.mark.gpu
def test_cp_api_script_matches_reference_cuda_path():
    _run(rs._lib.DEFAULT_LIB, 60, 20262)


def test_give_model_lit_negative_literals_documented_divergence():
    """libcoprocessorc.cc:143 maps -k to variable index k+1 (a sign slip); the product maps it to variable k (what the header
    says).  This pins both behaviours so that the divergence stays a documented, deliberate one."""
    ref = _bind(REFLIB); prod = _bind(build_mock())
    got = {}
    for name, L in (("ref", ref), ("prod", prod)):
        h = C.c_void_p(L.CPinit(None))
        L.CPresetModel(h)
        L.CPgiveModelLit(h, -3)
        n = L.CPmodelVariables(h)
        got[name] = (n, [L.CPgetFinalModelLit(h) for _ in range(n)])
        L.CPdestroy(C.byref(h))
    assert got["prod"][0] == 3 and got["prod"][1][2] == -3                 # variable 3 is false
    assert got["ref"][0] == 5 and got["ref"][1][4] == -5                   # the reference wrote variable 5
