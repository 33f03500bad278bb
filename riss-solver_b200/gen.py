"""Seeded synthetic CNF generators for the Coprocessor hot path.

A formula is a flat little-endian int32 stream of DIMACS literals (+-(1..n)), every clause terminated
by 0 - the same stream CPaddLit (coprocessor/libcoprocessorc.h:64) consumes one literal at a time.
Uniform random k-SAT at ratio 4.26 contains essentially no subsumed / strengthenable clauses, so all
generators can plant redundancy: exact duplicates, supersets, and one-literal-flipped copies.
"""
from __future__ import annotations

import numpy as np


def _distinct_vars(rng, n, m, k, pvar=None):
    """m rows of k distinct variables in 1..n (rows with repeats are redrawn)."""
    def draw(cnt):
        if pvar is None:
            return rng.integers(1, n + 1, size=(cnt, k), dtype=np.int64)
        return rng.choice(n, size=(cnt, k), p=pvar).astype(np.int64) + 1
    v = draw(m)
    while True:
        s = np.sort(v, axis=1)
        bad = (s[:, 1:] == s[:, :-1]).any(axis=1)
        nb = int(bad.sum())
        if nb == 0:
            return v
        v[bad] = draw(nb)


def _signed(rng, v):
    sgn = rng.integers(0, 2, size=v.shape, dtype=np.int64) * 2 - 1
    return (v * sgn).astype(np.int32)


def _rows_to_stream(rows_by_k):
    """rows_by_k: list of (m_i, k_i) int32 arrays -> list of per-clause arrays is too slow; we build
    (lits, sizes) in CSR form instead."""
    lits = np.concatenate([r.reshape(-1) for r in rows_by_k]) if rows_by_k else np.zeros(0, np.int32)
    sizes = np.concatenate([np.full(r.shape[0], r.shape[1], np.int32) for r in rows_by_k]) if rows_by_k else np.zeros(0, np.int32)
    return lits.astype(np.int32), sizes


def csr_to_stream(lits, sizes):
    """(lits, sizes) -> 0-terminated int32 stream."""
    m = sizes.shape[0]
    out = np.zeros(lits.shape[0] + m, dtype=np.int32)
    ends = np.cumsum(sizes.astype(np.int64))
    starts = ends - sizes
    # position of literal j of clause i in out = starts[i] + i + j
    clause_of = np.repeat(np.arange(m, dtype=np.int64), sizes)
    out[np.arange(lits.shape[0], dtype=np.int64) + clause_of] = lits
    return out


def stream_to_csr(stream):
    stream = np.asarray(stream, dtype=np.int32)
    z = np.flatnonzero(stream == 0)
    sizes = np.diff(np.concatenate([[-1], z])) - 1
    lits = stream[stream != 0]
    return lits, sizes.astype(np.int32)


def permute_clauses(rng, lits, sizes):
    m = sizes.shape[0]
    perm = rng.permutation(m)
    ends = np.cumsum(sizes.astype(np.int64))
    starts = ends - sizes
    nsz = sizes[perm]
    nstart = starts[perm]
    nends = np.cumsum(nsz.astype(np.int64))
    nbeg = nends - nsz
    idx = np.repeat(nstart - nbeg, nsz) + np.arange(int(nends[-1]) if m else 0, dtype=np.int64)
    return lits[idx], nsz


def _ragged_take(lits, starts, sizes, pick):
    """gather clauses `pick` -> (lits, sizes)"""
    sz = sizes[pick]
    st = starts[pick]
    ends = np.cumsum(sz.astype(np.int64))
    beg = ends - sz
    idx = np.repeat(st - beg, sz) + np.arange(int(ends[-1]) if pick.size else 0, dtype=np.int64)
    return lits[idx], sz


def plant(rng, n, lits, sizes, dup=0.02, sup=0.03, flip=0.03):
    """Plant redundancy (fractions of the base clause count), then shuffle clause order.

    dup:  exact duplicates                         -> subsumption, tie rule
    sup:  copy + 1..2 fresh extra literals         -> subsumed supersets
    flip: copy with one literal negated + 0..1 extra -> self-subsuming resolution targets
    """
    m = sizes.shape[0]
    ends = np.cumsum(sizes.astype(np.int64))
    starts = ends - sizes
    parts_l, parts_s = [lits], [sizes]

    def with_extra(pick, extra, flip_one):
        pl, ps = _ragged_take(lits, starts, sizes, pick)
        pl = pl.copy()
        pe = np.cumsum(ps.astype(np.int64))
        pb = pe - ps
        if flip_one:
            which = pb + (rng.integers(0, 1 << 30, size=pick.size) % ps)
            pl[which] = -pl[which]
        # append `extra[i]` literals over variables not in the clause (retry on clash, cheap at n >> k)
        tot = ps + extra
        out = np.zeros(int(tot.sum()), dtype=np.int32)
        oe = np.cumsum(tot.astype(np.int64))
        ob = oe - tot
        idx = np.repeat(ob - pb, ps) + np.arange(pl.shape[0], dtype=np.int64)
        out[idx] = pl
        for j in range(int(extra.max()) if extra.size else 0):
            sel = np.flatnonzero(extra > j)
            cand = rng.integers(1, n + 1, size=sel.size)
            for _ in range(8):  # avoid variables already present
                clash = np.zeros(sel.size, dtype=bool)
                for t in range(int(tot.max())):
                    has = tot[sel] > t
                    pos = np.minimum(ob[sel] + t, out.shape[0] - 1)
                    clash |= has & (np.abs(out[pos]) == cand)
                if not clash.any():
                    break
                cand[clash] = rng.integers(1, n + 1, size=int(clash.sum()))
            sg = rng.integers(0, 2, size=sel.size) * 2 - 1
            out[ob[sel] + ps[sel] + j] = (cand * sg).astype(np.int32)
        return out, tot.astype(np.int32)

    nd, ns_, nf = int(m * dup), int(m * sup), int(m * flip)
    if nd:
        pick = rng.integers(0, m, size=nd)
        pl, ps = _ragged_take(lits, starts, sizes, pick)
        parts_l.append(pl); parts_s.append(ps)
    if ns_:
        pick = rng.integers(0, m, size=ns_)
        l2, s2 = with_extra(pick, rng.integers(1, 3, size=ns_).astype(np.int32), False)
        parts_l.append(l2); parts_s.append(s2)
    if nf:
        pick = rng.integers(0, m, size=nf)
        l2, s2 = with_extra(pick, rng.integers(0, 2, size=nf).astype(np.int32), True)
        parts_l.append(l2); parts_s.append(s2)
    L = np.concatenate(parts_l).astype(np.int32)
    S = np.concatenate(parts_s).astype(np.int32)
    return permute_clauses(rng, L, S)


def random_ksat(n, m, k=3, seed=12345, planted=True, zipf_s=0.0, dup=0.02, sup=0.03, flip=0.03):
    """Uniform (or Zipf-skewed, C5) random k-SAT, optionally planted (C2). Returns (lits, sizes)."""
    rng = np.random.default_rng(seed)
    pvar = None
    if zipf_s > 0:
        w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), zipf_s)
        pvar = w / w.sum()
    v = _distinct_vars(rng, n, m, k, pvar)
    lits, sizes = _rows_to_stream([_signed(rng, v)])
    if planted:
        lits, sizes = plant(rng, n, lits, sizes, dup, sup, flip)
    return lits, sizes


def community_ksat(n, m, seed=20260001, p_intra=0.8, ks=(3, 4, 5), kfrac=(0.5, 0.3, 0.2),
                   planted=True, bve_frac=0.05):
    """Community-attachment k-SAT (C3): c=sqrt(n) communities; with probability p_intra all variables
    of a clause come from one community.  The last bve_frac*n variables are kept rare (each is used in
    about 2+2 clauses only) so bounded variable elimination has candidates."""
    rng = np.random.default_rng(seed)
    n_rare = int(n * bve_frac)
    n_main = n - n_rare
    c = max(1, int(np.sqrt(n_main)))
    csize = n_main // c
    rows = []
    for k, fr in zip(ks, kfrac):
        mk = int(m * fr)
        intra = rng.random(mk) < p_intra
        comm = rng.integers(0, c, size=mk)
        v = np.where(intra[:, None],
                     comm[:, None] * csize + rng.integers(0, csize, size=(mk, k)),
                     rng.integers(0, n_main, size=(mk, k))) + 1
        while True:
            s = np.sort(v, axis=1)
            bad = (s[:, 1:] == s[:, :-1]).any(axis=1)
            nb = int(bad.sum())
            if nb == 0:
                break
            v[bad] = rng.integers(0, n_main, size=(nb, k)) + 1
        rows.append(_signed(rng, v))
    # rare variables: replace one literal of ~4 random clauses each (2 positive, 2 negative)
    if n_rare:
        for r in rows:
            pass
        allrows = rows
        tot = sum(r.shape[0] for r in allrows)
        for pol in (1, 1, -1, -1):
            tgt = rng.integers(0, tot, size=n_rare)
            off = 0
            for r in allrows:
                sel = (tgt >= off) & (tgt < off + r.shape[0])
                idx = tgt[sel] - off
                r[idx, 0] = pol * (n_main + 1 + np.flatnonzero(sel)).astype(np.int32)
                off += r.shape[0]
    lits, sizes = _rows_to_stream(rows)
    if planted:
        lits, sizes = plant(rng, n, lits, sizes)
    else:
        lits, sizes = permute_clauses(rng, lits, sizes)
    return lits, sizes


def write_bin(path, lits, sizes):
    csr_to_stream(lits, sizes).tofile(path)


def write_dimacs(path, lits, sizes, nvars=None):
    """Slow text writer - small instances only."""
    if nvars is None:
        nvars = int(np.abs(lits).max()) if lits.size else 0
    with open(path, "w") as f:
        f.write(f"p cnf {nvars} {sizes.shape[0]}\n")
        pos = 0
        for s in sizes.tolist():
            f.write(" ".join(map(str, lits[pos:pos + s].tolist())) + " 0\n")
            pos += s
