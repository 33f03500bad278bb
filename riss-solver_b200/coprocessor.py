"""Host-side mirror of the reference's preprocessor C interface (coprocessor/libcoprocessorc.h) and of the
pass-level scan service, on top of libcp3b200.so.

`Coprocessor` has the call sequence of the reference's libcp3: init -> parse options -> add literals ->
preprocess -> read output / counters -> postprocess a model.  `ScanService` exposes the cp3g_* kernels
(K1..K6) for callers that keep their own clause data base (what Subsumption.cc / BVE.cc would bind to).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

SUBSUME, STRENGTHEN = 0, 1


class Coprocessor:
    def __init__(self, options=(), preset: str | None = None, lib_path: str | None = None):
        """`options` and `preset` both go to CPinit: like the reference's command line they are parsed BEFORE the
        preprocessor is constructed.  That matters: cp3_limited and the step limits (cp3_sub_limit, cp3_str_limit,
        cp3_call_inc, cp3_bve_limit) are read by the constructors only (libcoprocessorc.cc:35-46); given to
        parse_options() afterwards they are accepted and have no effect - in the reference and here alike."""
        self.L = _lib.load(lib_path)
        opts = options.split() if isinstance(options, str) else list(options)
        init = ":".join(x for x in ([preset] if preset else []) + ([" ".join(opts)] if opts else []))
        self.h = C.c_void_p(self.L.CPinit(init.encode() if init else None))

    # --- libcoprocessorc.h calls
    def parse_options(self, options):
        opts = options.split() if isinstance(options, str) else list(options)
        self.L.CPparseOptionString(self.h, " ".join(opts).encode())

    def add_lit(self, lit: int):
        self.L.CPaddLit(self.h, int(lit))

    def add_stream(self, stream):
        s = np.ascontiguousarray(stream, dtype=np.int32)
        self.L.CPaddLits(self.h, s.ctypes.data, s.size)

    def freeze(self, var: int):
        self.L.CPfreezeVariable(self.h, int(var))

    def preprocess(self):
        self.L.CPpreprocess(self.h)

    def ok(self) -> bool:
        return bool(self.L.CPok(self.h))

    def n_vars(self) -> int:
        return int(self.L.CPnVars(self.h))

    def n_clauses(self) -> int:
        return int(self.L.CPnClss(self.h))

    def n_lits(self) -> int:
        return int(self.L.CPnLits(self.h))

    def output_stream(self) -> np.ndarray:
        n = self.L.CPgetOutput(self.h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.L.CPgetOutput(self.h, out.ctypes.data, n)
        return out[:n]

    def output_units(self) -> int:
        return int(self.L.CPoutputUnits(self.h))

    def iter_output(self):
        """literal-by-literal like CPnextOutputLit (0 ends a clause)"""
        self.L.CPresetOutput(self.h)
        while self.L.CPhasNextOutputLit(self.h):
            yield int(self.L.CPnextOutputLit(self.h))

    def undo(self) -> np.ndarray:
        n = self.L.CPgetUndo(self.h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.L.CPgetUndo(self.h, out.ctypes.data, n)
        return out[:n]

    def stat(self, name: str) -> int:
        return int(self.L.CPstat(self.h, name.encode()))

    def extend_model(self, model_lits):
        """model_lits: satisfied DIMACS literals of the simplified formula -> full model (list of lits)"""
        self.L.CPresetModel(self.h)
        for l in model_lits:
            self.L.CPgiveModelLit(self.h, int(l))
        self.L.CPpostprocessModel(self.h)
        return [int(self.L.CPgetFinalModelLit(self.h)) for _ in range(self.L.CPmodelVariables(self.h))]

    def set_distributed(self, rank: int, world: int, unique_id: bytes) -> int:
        buf = C.create_string_buffer(unique_id, 128)
        return int(self.L.CPsetDistributed(self.h, rank, world, C.cast(buf, C.c_void_p)))

    def close(self):
        if self.h:
            self.L.CPdestroy(C.byref(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id(lib_path: str | None = None) -> bytes:
    L = _lib.load(lib_path)
    buf = C.create_string_buffer(128)
    if L.cp3g_nccl_unique_id(C.cast(buf, C.c_void_p)) != 0:
        raise RuntimeError("ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
    return buf.raw


class ScanService:
    """cp3g_*: device-resident clause data base + kernels.  Literals are Riss codes 2*var+sign."""

    def __init__(self, device: int = 0, lib_path: str | None = None):
        self.L = _lib.load(lib_path)
        self.g = C.c_void_p(self.L.cp3g_create(device))
        if not self.g:
            raise RuntimeError("cp3g_create failed: no usable CUDA device (there is no CPU fallback)")
        self.nvars = 0

    def _ck(self, r, what):
        if r != 0:
            raise RuntimeError(f"{what} failed ({r}): {self.L.cp3g_last_error(self.g).decode()}")

    def upload(self, nvars, lits, sizes, flags=None):
        lits = np.ascontiguousarray(lits, dtype=np.int32)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        start = np.zeros(sizes.size, dtype=np.uint32)
        if sizes.size:
            start[1:] = np.cumsum(sizes[:-1], dtype=np.uint64).astype(np.uint32)
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        self.nvars = int(nvars)
        self._ck(self.L.cp3g_upload(self.g, self.nvars, sizes.size, lits.ctypes.data, lits.size, start.ctypes.data,
                                    sizes.ctypes.data, None if fl is None else fl.ctypes.data), "upload")

    def build(self):
        self._ck(self.L.cp3g_build(self.g), "build")

    def download_occ(self):
        off = np.zeros(2 * self.nvars + 1, dtype=np.uint32)
        self._ck(self.L.cp3g_download_occ(self.g, off.ctypes.data, None), "download_occ")
        refs = np.zeros(max(int(off[-1]), 1), dtype=np.uint32)
        self._ck(self.L.cp3g_download_occ(self.g, off.ctypes.data, refs.ctypes.data), "download_occ")
        return off, refs[:int(off[-1])]

    def set_deleted(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._ck(self.L.cp3g_set_deleted(self.g, ids.size, ids.ctypes.data), "set_deleted")

    def scan(self, kind, self_ids, req_off=None, req_lits=None, cap=None, count_only=False):
        ids = np.ascontiguousarray(self_ids, dtype=np.uint32)
        if count_only:      # hits stay on the device; returns (number of hits, checks)
            n = C.c_uint32(0); chk = C.c_uint64(0)
            self._ck(self.L.cp3g_scan(self.g, kind, ids.size, ids.ctypes.data, None, None, None, 0, C.byref(n), C.byref(chk)), "scan")
            return int(n.value), int(chk.value)
        ro = None if req_off is None else np.ascontiguousarray(req_off, dtype=np.uint32)
        rl = None if req_lits is None else np.ascontiguousarray(req_lits, dtype=np.int32)
        cap = int(cap or max(1024, ids.size))
        while True:
            out = (_lib.cp3g_hit * cap)()
            n = C.c_uint32(0); chk = C.c_uint64(0)
            r = self.L.cp3g_scan(self.g, kind, ids.size, ids.ctypes.data, None if ro is None else ro.ctypes.data,
                                 None if rl is None else rl.ctypes.data, C.cast(out, C.c_void_p), cap, C.byref(n), C.byref(chk))
            if r == -3:
                cap = int(n.value) + 1024
                continue
            self._ck(r, "scan")
            a = np.frombuffer(out, dtype=np.dtype([("req", "<u4"), ("clause", "<u4"), ("lit", "<i4")]), count=n.value).copy()
            return a, int(chk.value)

    def scan_all(self, kind, cap=None, count_only=False):
        """full-queue bulk pass (K3/K4 bulk kernels): every live clause is a request; hit.req = clause index"""
        if count_only:
            n = C.c_uint32(0); chk = C.c_uint64(0)
            self._ck(self.L.cp3g_scan_all(self.g, kind, None, 0, C.byref(n), C.byref(chk)), "scan_all")
            return int(n.value), int(chk.value)
        cap = int(cap or 1 << 16)
        while True:
            out = (_lib.cp3g_hit * cap)()
            n = C.c_uint32(0); chk = C.c_uint64(0)
            r = self.L.cp3g_scan_all(self.g, kind, C.cast(out, C.c_void_p), cap, C.byref(n), C.byref(chk))
            if r == -3:
                cap = int(n.value) + 1024
                continue
            self._ck(r, "scan_all")
            a = np.frombuffer(out, dtype=np.dtype([("req", "<u4"), ("clause", "<u4"), ("lit", "<i4")]), count=n.value).copy()
            return a, int(chk.value)

    def sub_pass(self, max_list=4096):
        """first subsumption pass committed on the device (cp3g_sub_pass): returns None when not applicable, else a dict with
        the subsumed clause ids (sorted), literal total, step count and the post-pass lists (off, n, stale, refs)"""
        st = C.c_int(1); nd = C.c_uint32(0); nl = C.c_uint64(0); steps = C.c_uint64(0); chk = C.c_uint64(0)
        self._ck(self.L.cp3g_sub_pass(self.g, int(max_list), C.byref(st), C.byref(nd), C.byref(nl), C.byref(steps), C.byref(chk)), "sub_pass")
        if st.value != 0:
            return None
        off = np.zeros(2 * self.nvars + 1, dtype=np.uint32)
        self._ck(self.L.cp3g_download_occ(self.g, off.ctypes.data, None), "download_occ")
        dead = np.zeros(max(nd.value, 1), dtype=np.uint32); n = np.zeros(2 * self.nvars, dtype=np.uint32); stale = np.zeros(2 * self.nvars, dtype=np.uint32)
        refs = np.zeros(max(int(off[-1]), 1), dtype=np.uint32)
        self._ck(self.L.cp3g_sub_pass_fetch(self.g, dead.ctypes.data, n.ctypes.data, stale.ctypes.data, refs.ctypes.data), "sub_pass_fetch")
        return {"dead": np.sort(dead[:nd.value]), "lits": int(nl.value), "steps": int(steps.value), "checks": int(chk.value),
                "off": off, "n": n, "stale": stale, "refs": refs[:int(off[-1])]}

    def str_static(self, real_bits):
        """K10 (cp3g_str_static) behind a committed sub_pass: real_bits = uint32 words, one bit per clause.  Returns None when not
        applicable, else (inert entries, inert steps, small clauses, walked bits as uint32 words over the literals)"""
        real_bits = np.ascontiguousarray(real_bits, dtype=np.uint32)
        st = C.c_int(1); ne = C.c_uint64(0); ns = C.c_uint64(0); sm = C.c_uint64(0)
        walked = np.zeros((2 * self.nvars + 31) // 32 + 1, dtype=np.uint32)
        self._ck(self.L.cp3g_str_static(self.g, real_bits.ctypes.data, C.byref(st), C.byref(ne), C.byref(ns), C.byref(sm), walked.ctypes.data), "str_static")
        if st.value != 0:
            return None
        return int(ne.value), int(ns.value), int(sm.value), walked

    def bve_pairs(self, vars_, pos_lists, neg_lists):
        vars_ = np.ascontiguousarray(vars_, dtype=np.uint32)
        p_off = np.zeros(vars_.size + 1, dtype=np.uint32); n_off = np.zeros(vars_.size + 1, dtype=np.uint32)
        pair_off = np.zeros(vars_.size + 1, dtype=np.uint64)
        for i, (p, n) in enumerate(zip(pos_lists, neg_lists)):
            p_off[i + 1] = p_off[i] + len(p); n_off[i + 1] = n_off[i] + len(n)
            pair_off[i + 1] = pair_off[i] + len(p) * len(n)
        p_refs = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint32) for p in pos_lists]) if len(pos_lists) else np.zeros(0, np.uint32), dtype=np.uint32)
        n_refs = np.ascontiguousarray(np.concatenate([np.asarray(n, dtype=np.uint32) for n in neg_lists]) if len(neg_lists) else np.zeros(0, np.uint32), dtype=np.uint32)
        sizes = np.zeros(max(int(pair_off[-1]), 1), dtype=np.int16)
        self._ck(self.L.cp3g_bve_pairs(self.g, vars_.size, vars_.ctypes.data, p_off.ctypes.data, p_refs.ctypes.data,
                                       n_off.ctypes.data, n_refs.ctypes.data, pair_off.ctypes.data, sizes.ctypes.data), "bve_pairs")
        return pair_off, sizes[:int(pair_off[-1])]

    def bce_masks(self, vars_, pos_lists, neg_lists):
        """K9 (cp3g_bce_masks): per variable a bit matrix rows = pos clauses, columns = neg clauses; bit set iff the resolvent
        is not a tautology (BCE.cc:624-646).  Returns (word_off, masks)."""
        vars_ = np.ascontiguousarray(vars_, dtype=np.uint32)
        p_off = np.zeros(vars_.size + 1, dtype=np.uint32); n_off = np.zeros(vars_.size + 1, dtype=np.uint32)
        w_off = np.zeros(vars_.size + 1, dtype=np.uint64)
        for i, (p, n) in enumerate(zip(pos_lists, neg_lists)):
            p_off[i + 1] = p_off[i] + len(p); n_off[i + 1] = n_off[i] + len(n)
            w_off[i + 1] = w_off[i] + len(p) * ((len(n) + 31) // 32)
        cat = lambda ls: np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint32) for x in ls]) if len(ls) else np.zeros(0, np.uint32), dtype=np.uint32)
        p_refs = cat(pos_lists); n_refs = cat(neg_lists)
        masks = np.zeros(max(int(w_off[-1]), 1), dtype=np.uint32)
        self._ck(self.L.cp3g_bce_masks(self.g, vars_.size, vars_.ctypes.data, p_off.ctypes.data, p_refs.ctypes.data,
                                       n_off.ctypes.data, n_refs.ctypes.data, w_off.ctypes.data, masks.ctypes.data), "bce_masks")
        return w_off, masks[:int(w_off[-1])]

    def bve_eval(self, vars_, pos_lists, neg_lists, use_gates=True, stream=True):
        """K5'/K6' (cp3g_bve_eval / cp3g_bve_stream): per variable the record (p_limit, n_limit, pos_n, neg_n, resolvents, lit_clauses,
        steps, flags), the rewritten lists, the per-clause statistics and - for every variable without the HOST flag - its resolvents"""
        vars_ = np.ascontiguousarray(vars_, dtype=np.uint32)
        nv = vars_.size
        p_off = np.zeros(nv + 1, dtype=np.uint32); n_off = np.zeros(nv + 1, dtype=np.uint32)
        for i, (p, n) in enumerate(zip(pos_lists, neg_lists)):
            p_off[i + 1] = p_off[i] + len(p); n_off[i + 1] = n_off[i] + len(n)
        cat = lambda ls: np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint32) for x in ls]) if len(ls) else np.zeros(0, np.uint32), dtype=np.uint32)
        p_refs = np.concatenate([cat(pos_lists), np.zeros(1, np.uint32)]); n_refs = np.concatenate([cat(neg_lists), np.zeros(1, np.uint32)])
        recs = np.zeros((max(nv, 1), 8), dtype=np.int32); pst = np.zeros(p_refs.size + 1, np.uint16); nst = np.zeros(n_refs.size + 1, np.uint16)
        self._ck(self.L.cp3g_bve_eval(self.g, nv, vars_.ctypes.data, p_off.ctypes.data, p_refs.ctypes.data, n_off.ctypes.data, n_refs.ctypes.data,
                                      1 if use_gates else 0, recs.ctypes.data, pst.ctypes.data, nst.ctypes.data), "bve_eval")
        recs = recs[:nv]
        out = {"recs": recs, "p_off": p_off, "n_off": n_off, "p_refs": p_refs, "n_refs": n_refs, "p_stats": pst, "n_stats": nst}
        if stream:
            sel = np.flatnonzero((recs[:, 7] & 2) == 0).astype(np.uint32)
            res_off = np.zeros(sel.size + 1, dtype=np.uint64); lit_off = np.zeros(sel.size + 1, dtype=np.uint64)
            res_off[1:] = np.cumsum(recs[sel, 4].astype(np.uint64)); lit_off[1:] = np.cumsum(recs[sel, 5].astype(np.uint64))
            rsz = np.zeros(int(res_off[-1]) + 1, np.uint16); rl = np.zeros(int(lit_off[-1]) + 1, np.int32)
            self._ck(self.L.cp3g_bve_stream(self.g, sel.size, sel.ctypes.data, res_off.ctypes.data, lit_off.ctypes.data, rsz.ctypes.data, rl.ctypes.data), "bve_stream")
            out.update(sel=sel, res_off=res_off, lit_off=lit_off, rsz=rsz[:int(res_off[-1])], rlits=rl[:int(lit_off[-1])])
        return out

    def bve_resolve(self, vars_, p_refs, n_refs, sizes):
        vars_ = np.ascontiguousarray(vars_, dtype=np.uint32)
        p = np.ascontiguousarray(p_refs, dtype=np.uint32); n = np.ascontiguousarray(n_refs, dtype=np.uint32)
        off = np.zeros(vars_.size + 1, dtype=np.uint64)
        off[1:] = np.cumsum(np.asarray(sizes, dtype=np.uint64))
        out = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
        self._ck(self.L.cp3g_bve_resolve(self.g, vars_.size, vars_.ctypes.data, p.ctypes.data, n.ctypes.data,
                                         off.ctypes.data, out.ctypes.data, int(off[-1])), "bve_resolve")
        return off, out[:int(off[-1])]

    def ingest(self, stream):
        """K0: build the device arena from a DIMACS literal stream; returns (status, nvars, nclauses, nlits) -
        status 1 = the stream needs the sequential host ingest (unit / empty clause)"""
        st = np.ascontiguousarray(stream, dtype=np.int32)
        nv = C.c_uint32(0); nc = C.c_uint32(0); nl = C.c_size_t(0); status = C.c_int(1)
        self._ck(self.L.cp3g_ingest(self.g, st.ctypes.data, st.size, C.byref(nv), C.byref(nc), C.byref(nl), C.byref(status)), "ingest")
        return int(status.value), int(nv.value), int(nc.value), int(nl.value)

    def download_arena(self):
        """(literal pool, starts, sizes, flags) of the device data base"""
        ncl = int(self.counter("nclauses")); nl = int(self.counter("nlits"))
        lits = np.zeros(max(nl, 1), dtype=np.int32); rec = np.zeros((max(ncl, 1), 4), dtype=np.uint32)
        self._ck(self.L.cp3g_download_arena(self.g, lits.ctypes.data, rec.ctypes.data), "download_arena")
        rec = rec[:ncl]
        return lits[:nl], rec[:, 0].copy(), (rec[:, 1] & 0xffffff).astype(np.uint32), (rec[:, 1] >> 24).astype(np.uint8)

    def compact(self, download=True):
        """K8: pack the clause arena (garbageCollect / relocAll); returns (new_start[nclauses], packed literal pool)"""
        ncl = int(self.counter("nclauses")); nl = int(self.counter("nlits"))
        ns = np.zeros(max(ncl, 1), dtype=np.uint32); out = np.zeros(max(nl, 1), dtype=np.int32); n = C.c_size_t(0)
        self._ck(self.L.cp3g_compact(self.g, ns.ctypes.data, out.ctypes.data if download else None, out.size, C.byref(n)), "compact")
        return ns[:ncl], out[:int(n.value)]

    def counter(self, name):
        return int(self.L.cp3g_counter(self.g, name.encode()))

    def close(self):
        if self.g:
            self.L.cp3g_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
