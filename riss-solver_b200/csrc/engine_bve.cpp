// engine_bve.cpp - BVE commit of the host engine (coprocessor/techniques/BVE.cc): gates, anticipation, resolution, heap walk.
#include "engine_util.h"

namespace cp3 {

// ------------------------------------------------------------------------------------------------ BVE
// findGates, BVE.cc:1143-1267 - reads clause sizes / binary partners only (no clause-vs-clause test)
bool Engine::findGates(int v, int& p_limit, int& n_limit)
{
    if (hot[2 * v].n < 3 && hot[2 * v + 1].n < 3) { return false; }
    if (hot[2 * v].n < 2 || hot[2 * v + 1].n < 2) { return false; }
    for (int pn = 0; pn < 2; ++pn) {
        const int pLit = 2 * v + (pn != 0), nLit = 2 * v + (pn == 0);
        ListRef pList = L(pLit); ListRef nList = L(nLit);
        int& pClauses = pn == 0 ? p_limit : n_limit;
        int& nClauses = pn == 0 ? n_limit : p_limit;
        if (ma_step >= (1u << 30)) { std::fill(ma.begin(), ma.end(), 0u); ma_step = 0; }
        ++ma_step;
        for (uint32_t i = 0; i < pList.n; ++i) {
            uint32_t c = LP(pList)[i];
            if ((C[c].flag & F_DEL) || C[c].size != 2) { continue; }
            const int32_t* l = CL(c);
            int other = l[0] == pLit ? l[1] : l[0];
            ma[lneg(other)] = ma_step;
        }
        for (uint32_t i = 0; i < nList.n; ++i) {
            uint32_t c = LP(nList)[i];
            if (C[c].flag & F_DEL) { continue; }
            const int32_t* l = CL(c); uint32_t j = 0;
            for (; j < C[c].size; ++j) {
                if (l[j] == nLit) { continue; }
                if (ma[l[j]] != ma_step) { break; }
            }
            if (j == C[c].size) {
                pClauses = (int)C[c].size - 1;
                nClauses = 1;
                for (uint32_t k = 0; k < C[c].size; ++k) { ma[l[k]] = 0; }
                std::swap(LP(nList)[0], LP(nList)[i]);
                uint32_t placed = 0;
                for (uint32_t k = 0; k < pList.n; ++k) {
                    uint32_t c2 = LP(pList)[k];
                    if ((C[c2].flag & F_DEL) || C[c2].size != 2) { continue; }
                    const int32_t* l2 = CL(c2);
                    int other = l2[0] == pLit ? l2[1] : l2[0];
                    if (ma[lneg(other)] != ma_step) {
                        std::swap(LP(pList)[placed], LP(pList)[k]);
                        placed++;
                        ma[lneg(other)] = ma_step;
                    }
                }
                return true;
            }
        }
    }
    return false;
}

// K5 result for variable v on its current occurrence lists: cached while none of its clauses changed,
// else one more batch (v plus the invalid variables near the top of the heap).
Engine::PairEnt& Engine::pairsFor(int v)
{
    if (pc_slot.size() < (size_t)nvars) { pc_slot.resize((size_t)nvars, -1); }
    { PairEnt* pe0 = pairLookup(v); if (pe0 && pe0->valid && pe0->touch == var_touch[v]) { return *pe0; } }
    flushDevice();
    ++n_pair_batches;
    // batch: v plus every variable that waits in the heap with an outdated (or no) matrix
    std::vector<int> vs; vs.push_back(v);
    uint64_t budget = 0;
    size_t keep = 0;
    for (size_t i = 0; i < (pairs_single ? 0 : bve_pending.size()); ++i) {
        const int w = bve_pending[i];
        if (w == v || !inHeap(w) || assigns[w] != L_UNDEF || frozen[w] || hot[2 * w].n == 0 || hot[2 * w + 1].n == 0) { bve_inpend[w] = 0; continue; }
        const uint64_t np = (uint64_t)hot[2 * w].n * hot[2 * w + 1].n;
        if (np > 4096) { bve_inpend[w] = 0; continue; }             // rare long lists: on demand only
        // the reference's cut-off (BVE.cc:418-420) on the current counters: such variables are skipped anyway
        const bool cutoff = (hot[2 * w + 1].cnt > 10 && hot[2 * w].cnt > 10) || (dcount(w) > 15 && (hot[2 * w + 1].cnt > 5 || hot[2 * w].cnt > 5));
        if (cutoff && !opt.bve_unlimited && !opt.bve_force_gates) { bve_inpend[w] = 0; continue; }
        if (budget + np > (64u << 20)) { bve_pending[keep++] = w; continue; }   // next batch
        { PairEnt* pw = pairLookup(w); if (pw && pw->valid && pw->touch == var_touch[w]) { bve_inpend[w] = 0; continue; } }
        budget += np; bve_inpend[w] = 0;
        vs.push_back(w);
    }
    if (!pairs_single) { bve_pending.resize(keep); }
    std::vector<uint32_t> vars, p_off{0}, n_off{0}, p_refs, n_refs; std::vector<uint64_t> pair_off{0};
    for (int w : vs) {
        vars.push_back((uint32_t)w);
        ListRef P = L(2 * w); ListRef N = L(2 * w + 1);
        p_refs.insert(p_refs.end(), LP(P), LP(P) + P.n); n_refs.insert(n_refs.end(), LP(N), LP(N) + N.n);
        p_off.push_back((uint32_t)p_refs.size()); n_off.push_back((uint32_t)n_refs.size());
        pair_off.push_back(pair_off.back() + (uint64_t)P.n * N.n);
    }
    std::vector<int16_t> sizes(pair_off.back() + 1);
    int64_t t0 = now_ns();
    if (cp3g_bve_pairs(gpu, (uint32_t)vars.size(), vars.data(), p_off.data(), p_refs.data(), n_off.data(), n_refs.data(),
                       pair_off.data(), sizes.data()) != CP3G_OK) { die("bve_pairs failed"); }
    host_ns_gpuwait += now_ns() - t0;
    // K6 in the same batch for the variables the decision rule (BVE.cc:532-548) will most likely eliminate:
    // their resolvents are then already on the host when the ordered commit reaches them
    std::vector<uint32_t> rv, rp, rn; std::vector<uint64_t> roff{0}; std::vector<size_t> rfirst(vs.size() + 1, 0); std::vector<uint8_t> want(vs.size(), 0);
    for (size_t i = 0; i < vs.size(); ++i) {
        rfirst[i] = rv.size();
        const uint32_t np = p_off[i + 1] - p_off[i], nn = n_off[i + 1] - n_off[i];
        int res = 0;
        for (uint64_t k = pair_off[i]; k < pair_off[i + 1]; ++k) { res += sizes[k] > 1; }
        // the global growth allowance (bve_cgrow_t - totallyAdded, BVE.cc:532-540) is shared by ALL later
        // eliminations: taken at face value it would predict every variable of the batch as eliminable and spend
        // the prefetch budget on resolvents nobody asks for.  Only a small part of it is granted per variable; the
        // few clause-increasing eliminations beyond that fetch their resolvents on demand.
        int slack = opt.bve_cgrow;
        if (opt.bve_cgrow_t != INT32_MAX && opt.bve_cgrow_t > opt.bve_cgrow) { slack += std::min(8, std::max(0, opt.bve_cgrow_t - bve_totallyAdded)); }
        if (res > (int)(np + nn) + slack || roff.back() > (32u << 20)) { continue; }
        want[i] = 1;
        for (uint32_t a = 0; a < np; ++a) for (uint32_t b = 0; b < nn; ++b) {
            const int sz = sizes[pair_off[i] + (uint64_t)a * nn + b];
            if (sz >= 1) { rv.push_back(vars[i]); rp.push_back(p_refs[p_off[i] + a]); rn.push_back(n_refs[n_off[i] + b]); roff.push_back(roff.back() + (uint64_t)sz); }
        }
    }
    rfirst[vs.size()] = rv.size();
    std::vector<int32_t> rlits(roff.back() + 1);
    if (!rv.empty()) {
        ++n_resolve_batches;
        int64_t t1 = now_ns();
        if (cp3g_bve_resolve(gpu, (uint32_t)rv.size(), rv.data(), rp.data(), rn.data(), roff.data(), rlits.data(), roff.back()) != CP3G_OK) { die("bve_resolve failed"); }
        host_ns_gpuwait += now_ns() - t1;
    }
    // one block per batch owns the data; the per-variable entries are views into it
    pair_blocks.emplace_back(new PairBlock());
    PairBlock& blk = *pair_blocks.back();
    blk.roff.reserve(pair_off.back() + vs.size() + 1);
    std::vector<size_t> roff_first(vs.size() + 1, 0), rl_first(vs.size() + 1, 0);
    for (size_t i = 0; i < vs.size(); ++i) {
        roff_first[i] = blk.roff.size(); rl_first[i] = (size_t)roff[rfirst[i]];
        if (!want[i]) { continue; }
        // roff: offset of pair (a,b) into this variable's literals, row-major (only pairs with size >= 1 own literals)
        size_t r = rfirst[i]; const uint64_t base = roff[rfirst[i]];
        for (uint64_t k = pair_off[i]; k < pair_off[i + 1]; ++k) {
            blk.roff.push_back((uint32_t)(roff[r] - base));
            if (sizes[k] >= 1) { ++r; }
        }
        blk.roff.push_back((uint32_t)(roff[r] - base));
    }
    roff_first[vs.size()] = blk.roff.size(); rl_first[vs.size()] = (size_t)roff.back();
    blk.p_refs.swap(p_refs); blk.n_refs.swap(n_refs); blk.sizes.swap(sizes); blk.rlits.swap(rlits);
    for (size_t i = 0; i < vs.size(); ++i) {
        const int w = vs[i];
        if (pc_slot[w] < 0) { pc_slot[w] = (int32_t)pc_ent.size(); pc_ent.emplace_back(); }
        PairEnt& q = pc_ent[pc_slot[w]];
        q.touch = var_touch[w]; q.valid = true; q.has_res = want[i] != 0;
        q.pos = {blk.p_refs.data() + p_off[i], (size_t)(p_off[i + 1] - p_off[i])};
        q.neg = {blk.n_refs.data() + n_off[i], (size_t)(n_off[i + 1] - n_off[i])};
        q.sizes = {blk.sizes.data() + pair_off[i], (size_t)(pair_off[i + 1] - pair_off[i])};
        q.roff = {blk.roff.data() + roff_first[i], roff_first[i + 1] - roff_first[i]};
        q.rlits = {blk.rlits.data() + rl_first[i], (size_t)(roff[rfirst[i + 1]] - roff[rfirst[i]])};
    }
    return *pairLookup(v);
}

// K5'/K6' for variable v: cached while its clauses and lists are unchanged, else one more batch - v plus every variable that
// waits in the heap without a valid record.  The kernel evaluates gates, cleaning and anticipation per variable; for the
// variables the decision rule (BVE.cc:532-548) will most likely eliminate the resolvents come along as a literal stream.
Engine::EvalEnt* Engine::evalFor(int v)
{
    if (ev_ref.size() < (size_t)nvars) { const size_t o = ev_ref.size(); ev_ref.resize((size_t)nvars); parallelFor((size_t)nvars - o, [&](size_t b, size_t e, int) { for (size_t i = o + b; i < o + e; ++i) { ev_ref[i] = EvRef{-1, 0u, 0u, 0u}; } }); }
    auto fresh = [&](int w) -> bool { const EvRef& r = ev_ref[(size_t)w]; return r.blk >= 0 && r.touch == var_touch[w]; };
    auto view = [&](int w) -> EvalEnt* {
        const EvRef& r = ev_ref[(size_t)w]; const EvalBlock& blk = *eval_blocks[(size_t)r.blk]; const size_t i = r.idx;
        EvalEnt& q = ev_view; q.rec = blk.recs[i];
        const uint32_t* p1 = blk.p1.data() + blk.p_off[i]; const uint32_t* n1 = blk.n1.data() + blk.n_off[i];
        q.pos1 = {p1, (size_t)q.rec.pos_n}; q.neg1 = {n1, (size_t)q.rec.neg_n};
        if (r.applied) { q.pos0 = q.pos1; q.neg0 = q.neg1; }
        else { q.pos0 = {blk.p0.data() + blk.p_off[i], (size_t)(blk.p_off[i + 1] - blk.p_off[i])}; q.neg0 = {blk.n0.data() + blk.n_off[i], (size_t)(blk.n_off[i + 1] - blk.n_off[i])}; }
        q.pst = {blk.pst.data() + blk.p_off[i], (size_t)q.rec.pos_n}; q.nst = {blk.nst.data() + blk.n_off[i], (size_t)q.rec.neg_n};
        const int32_t si = blk.sidx[i];
        q.has_stream = si >= 0;
        if (si >= 0) { q.rsz = {blk.rsz.data() + blk.res_off[(size_t)si], (size_t)(blk.res_off[(size_t)si + 1] - blk.res_off[(size_t)si])};
                       q.rlits = {blk.rlits.data() + blk.lit_off[(size_t)si], (size_t)(blk.lit_off[(size_t)si + 1] - blk.lit_off[(size_t)si])}; }
        else { q.rsz = {}; q.rlits = {}; }
        return &q;
    };
    if (fresh(v)) { return view(v); }
    static const uint32_t LCAP = 64;
    if (hot[2 * v].n > LCAP || hot[2 * v + 1].n > LCAP) { return nullptr; }
    flushDevice();
    ++n_eval_batches;
    eval_blocks.emplace_back(new EvalBlock());
    EvalBlock& blk = *eval_blocks.back();
    std::vector<uint32_t>& vars = blk.vars; vars.push_back((uint32_t)v);
    uint64_t budget = 0; size_t keep = 0;
    for (size_t i = 0; i < bve_pending.size(); ++i) {
        const int w = bve_pending[i];
        if (w == v || !inHeap(w) || assigns[w] != L_UNDEF || frozen[w] || hot[2 * w].n == 0 || hot[2 * w + 1].n == 0) { bve_inpend[w] = 0; continue; }
        if (hot[2 * w].n > LCAP || hot[2 * w + 1].n > LCAP) { bve_inpend[w] = 0; continue; }
        const uint64_t np = (uint64_t)hot[2 * w].n * hot[2 * w + 1].n;
        // the reference's cut-off (BVE.cc:418-420) on the current counters: such variables are skipped anyway
        const bool cutoff = (hot[2 * w + 1].cnt > 10 && hot[2 * w].cnt > 10) || (dcount(w) > 15 && (hot[2 * w + 1].cnt > 5 || hot[2 * w].cnt > 5));
        if (cutoff && !opt.bve_unlimited && !opt.bve_force_gates) { bve_inpend[w] = 0; continue; }
        if (budget + np > (256u << 20)) { bve_pending[keep++] = w; continue; }   // next batch
        if (fresh(w)) { bve_inpend[w] = 0; continue; }
        budget += np; bve_inpend[w] = 0;
        vars.push_back((uint32_t)w);
    }
    bve_pending.resize(keep);
    const size_t nv = vars.size();
    std::vector<uint32_t>& p_off = blk.p_off; std::vector<uint32_t>& n_off = blk.n_off;
    p_off.assign(nv + 1, 0); n_off.assign(nv + 1, 0);
    for (size_t i = 0; i < nv; ++i) { p_off[i + 1] = p_off[i] + hot[2 * vars[i]].n; n_off[i + 1] = n_off[i] + hot[2 * vars[i] + 1].n; }
    blk.p0.resize(p_off[nv] + 1); blk.n0.resize(n_off[nv] + 1); blk.p1.resize(p_off[nv] + 1); blk.n1.resize(n_off[nv] + 1);
    blk.pst.resize(p_off[nv] + 1); blk.nst.resize(n_off[nv] + 1);
    parallelFor(nv, [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            const int w = (int)vars[i];
            memcpy(blk.p0.data() + p_off[i], occ_pool.data() + occ[2 * w].off, sizeof(uint32_t) * hot[2 * w].n);
            memcpy(blk.n0.data() + n_off[i], occ_pool.data() + occ[2 * w + 1].off, sizeof(uint32_t) * hot[2 * w + 1].n);
            memcpy(blk.p1.data() + p_off[i], blk.p0.data() + p_off[i], sizeof(uint32_t) * hot[2 * w].n);
            memcpy(blk.n1.data() + n_off[i], blk.n0.data() + n_off[i], sizeof(uint32_t) * hot[2 * w + 1].n);
        }
    }, 16384);
    std::vector<cp3g_bve_rec>& recs = blk.recs; recs.resize(nv);
    int64_t t0 = now_ns();
    if (cp3g_bve_eval(gpu, (uint32_t)nv, vars.data(), p_off.data(), blk.p1.data(), n_off.data(), blk.n1.data(), opt.bve_gates ? 1 : 0,
                      recs.data(), opt.bve_BCElim ? blk.pst.data() : nullptr, opt.bve_BCElim ? blk.nst.data() : nullptr) != CP3G_OK) { die("bve_eval failed"); }
    host_ns_gpuwait += now_ns() - t0;
    // resolvent streams for the likely eliminations.  The global growth allowance (bve_cgrow_t - totallyAdded, BVE.cc:532-540) is
    // shared by ALL later eliminations: only a small part of it is granted per variable here; the few clause-increasing
    // eliminations beyond that fetch their resolvents through the pair-matrix path.
    int slack = opt.bve_cgrow;
    if (opt.bve_cgrow_t != INT32_MAX && opt.bve_cgrow_t > opt.bve_cgrow) { slack += std::min(8, std::max(0, opt.bve_cgrow_t - bve_totallyAdded)); }
    std::vector<uint32_t> sel; std::vector<uint64_t>& res_off = blk.res_off; std::vector<uint64_t>& lit_off = blk.lit_off;
    res_off.assign(1, 0); lit_off.assign(1, 0); blk.sidx.assign(nv, -1);
    for (size_t i = 0; i < nv; ++i) {
        const cp3g_bve_rec& r = recs[i];
        if ((r.flags & CP3G_BVE_HOST) || r.resolvents == 0 || opt.bve_BCElim) { continue; }
        if ((int64_t)r.resolvents > (int64_t)r.pos_n + r.neg_n + slack || lit_off.back() > (64u << 20)) { continue; }
        blk.sidx[i] = (int32_t)sel.size();
        sel.push_back((uint32_t)i); res_off.push_back(res_off.back() + r.resolvents); lit_off.push_back(lit_off.back() + r.lit_clauses);
    }
    blk.rsz.resize(res_off.back() + 1); blk.rlits.resize(lit_off.back() + 1);
    if (!sel.empty()) {
        ++n_resolve_batches;
        int64_t t1 = now_ns();
        if (cp3g_bve_stream(gpu, (uint32_t)sel.size(), sel.data(), res_off.data(), lit_off.data(), blk.rsz.data(), blk.rlits.data()) != CP3G_OK) { die("bve_stream failed"); }
        host_ns_gpuwait += now_ns() - t1;
    }
    const int32_t bi = (int32_t)eval_blocks.size() - 1;
    parallelFor(nv, [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { const uint32_t w = vars[i]; ev_ref[w] = EvRef{bi, (uint32_t)i, var_touch[w], 0u}; } }, 16384);
    return fresh(v) ? view(v) : nullptr;
}


// anticipateElimination, BVE.cc:695-880; tryResolve sizes come from K5
int Engine::anticipate(int v, int p_limit, int n_limit, int& lit_clauses, int& resolvents)
{
    ListRef positive = L(2 * v); ListRef negative = L(2 * v + 1);
    const bool hasDefinition = (p_limit < (int)positive.n || n_limit < (int)negative.n);
    lit_clauses = 0;
    int clausesToUse = 0;
    if (opt.bve_early) {
        for (uint32_t i = 0; i < positive.n; ++i) { if (!(C[LP(positive)[i]].flag & F_DEL)) { clausesToUse++; } }
        for (uint32_t i = 0; i < negative.n; ++i) { if (!(C[LP(negative)[i]].flag & F_DEL)) { clausesToUse++; } }
        clausesToUse += opt.bve_cgrow;
        if (opt.bve_cgrow_t != INT32_MAX && opt.bve_cgrow_t > opt.bve_cgrow) { clausesToUse += opt.bve_cgrow_t - bve_totallyAdded; }
    }
    bool binariesOnly = false;
    PairEnt* pe = &pairsFor(v);
    std::vector<int>& col = bve_col;             // column of negative[cn] in the K5 matrix (reused: no allocation per variable)
    auto remap = [&]() { col.assign(negative.n, -1); for (uint32_t j = 0; j < negative.n; ++j) { col[j] = indexOf(pe->neg, LP(negative)[j]); } };
    remap();
    for (uint32_t cp = 0; cp < positive.n; ++cp) {
        const uint32_t p = LP(positive)[cp];
        if (C[p].flag & F_DEL) { continue; }
        if (binariesOnly && C[p].size > 2) { continue; }
        int row = indexOf(pe->pos, p);
        for (uint32_t cn = 0; cn < negative.n; ++cn) {
            if ((int)cp >= p_limit && (int)cn >= n_limit) { continue; }
            if (hasDefinition && (int)cp < p_limit && (int)cn < n_limit) { continue; }
            bve_steps++;
            const uint32_t n = LP(negative)[cn];
            if (C[n].flag & F_DEL) { continue; }
            if (binariesOnly && C[n].size > 2) { continue; }
            if (row < 0 || col[cn] < 0) { die("internal: K5 matrix misses a clause"); }
            const int newLits = pe->sizes[(size_t)row * pe->neg.size() + col[cn]];
            if (newLits > 1) {
                ++pos_stats[cp]; ++neg_stats[cn];
                lit_clauses += newLits;
                resolvents++;
                if (opt.bve_early && resolvents > clausesToUse) { binariesOnly = true; }
            } else if (newLits == 0) {
                successful(t_bve);
                ok = false;
                return L_FALSE;
            } else if (newLits == 1) {
                // the unit literal: K6 on this single pair
                uint32_t vv = (uint32_t)v; uint64_t off[2] = {0, 1}; int32_t ul = 0;
                flushDevice();
                if (cp3g_bve_resolve(gpu, 1, &vv, &p, &n, off, &ul, 1) != CP3G_OK) { die("bve_resolve failed"); }
                int status = dataEnqueue(ul);
                if (status == L_FALSE) { successful(t_bve); return L_FALSE; }
                else if (status == L_TRUE) {
                    successful(t_bve);
                    ++bve_unitsEnqueued;
                    if (propagationProcess(true, false, VAR_UNDEF) == L_FALSE) { return L_FALSE; }
                    t_bve.modified = t_bve.modified || t_up.modified;
                    // clauses changed under us: fresh sizes for the rest of the loop
                    if (positive.n == 0 || negative.n == 0) { break; }
                    pe = &pairsFor(v); remap(); row = indexOf(pe->pos, p);
                    if (C[p].flag & F_DEL) { break; }
                }
            }
        }
    }
    return binariesOnly ? L_TRUE : L_UNDEF;
}

// removeClauses, BVE.cc:644-684
void Engine::bveRemoveClauses(int x, int extLitOrNeg)
{
    const bool useHeap = opt.bve_heap_updates > 0;
    ListRef list = L(x);
    for (uint32_t i = 0; i < list.n; ++i) {
        uint32_t c = LP(list)[i];
        if (C[c].flag & F_DEL) { continue; }
        removedClause(c, useHeap, false, VAR_UNDEF);
        successful(t_bve);
        setDeleted(c);
        if (extLitOrNeg >= 0) { extClause(c, extLitOrNeg); }
        ++bve_removedClauses; bve_removedLiterals += (int)C[c].size;
    }
}
// removeBlockedClauses, BVE.cc:1053-1090
void Engine::bveRemoveBlocked(int x, const std::vector<int>& stats)
{
    const bool useHeap = opt.bve_heap_updates > 0;
    ListRef list = L(x);
    for (uint32_t i = 0; i < list.n; ++i) {
        uint32_t c = LP(list)[i];
        if (C[c].flag & F_DEL) { continue; }
        if (stats[i] == 0) {
            setDeleted(c);
            removedClause(c, useHeap, false, VAR_UNDEF);
            successful(t_bve);
            extClause(c, x);
            ++bve_removedBC;
        }
    }
}

// resolveSet, BVE.cc:892-1046 (force == false): K5 tells which pairs resolve, K6 writes the resolvents
int Engine::resolveSet(int v, int p_limit, int n_limit)
{
    const bool useHeap = opt.bve_heap_updates > 0;
    ListRef positive = L(2 * v); ListRef negative = L(2 * v + 1);
    const bool hasDefinition = (p_limit < (int)positive.n || n_limit < (int)negative.n);
    PairEnt* pe = &pairsFor(v);
    // gather the pairs in reference order and fetch their resolvents with one K6 launch
    typedef BvePr Pr;
    std::vector<Pr>& prs = bve_prs; std::vector<int32_t>& rl = bve_rl;       // reused buffers: no allocation per variable
    auto fetch = [&](uint32_t cp0, uint32_t cn0) {
        prs.clear();
        std::vector<uint32_t>& vv = bve_vv; std::vector<uint32_t>& pr = bve_pr; std::vector<uint32_t>& nr = bve_nr; std::vector<uint64_t>& off = bve_off;
        vv.clear(); pr.clear(); nr.clear(); off.clear(); off.push_back(0);
        std::vector<int>& col = bve_col; col.assign(negative.n, -1);
        for (uint32_t j = 0; j < negative.n; ++j) { col[j] = indexOf(pe->neg, LP(negative)[j]); }
        for (uint32_t cp = cp0; cp < positive.n; ++cp) {
            const uint32_t p = LP(positive)[cp];
            if (C[p].flag & F_DEL) { continue; }
            const int row = indexOf(pe->pos, p);
            for (uint32_t cn = (cp == cp0 ? cn0 : 0); cn < negative.n; ++cn) {
                if ((int)cp >= p_limit && (int)cn >= n_limit) { continue; }
                if (hasDefinition && (int)cp < p_limit && (int)cn < n_limit) { continue; }
                const uint32_t n = LP(negative)[cn];
                if (C[n].flag & F_DEL) { continue; }
                if (row < 0 || col[cn] < 0) { die("internal: K5 matrix misses a clause"); }
                const int s = pe->sizes[(size_t)row * pe->neg.size() + col[cn]];
                if (s >= 1) {
                    prs.push_back({cp, cn, s, off.back()});
                    vv.push_back((uint32_t)v); pr.push_back(p); nr.push_back(n); off.push_back(off.back() + (uint64_t)s);
                }
            }
        }
        rl.assign(off.back() + 1, 0);
        if (!prs.empty() && pe->has_res) {
            // prefetched with the K5 batch: copy out of the cache (pair (row, col) -> literals)
            uint32_t last_cp = 0xffffffffu; size_t row2 = 0;
            for (const Pr& q : prs) {
                if (q.cp != last_cp) { last_cp = q.cp; row2 = (size_t)indexOf(pe->pos, LP(positive)[q.cp]); }
                const size_t idx = row2 * pe->neg.size() + (size_t)col[q.cn];
                memcpy(rl.data() + q.off, pe->rlits.p + pe->roff[idx], sizeof(int32_t) * (size_t)q.size);
            }
        } else if (!prs.empty()) {
            ++n_resolve_batches;
            flushDevice();                       // only the on-demand fetch needs the device copy up to date
            int64_t t0 = now_ns();
            if (cp3g_bve_resolve(gpu, (uint32_t)prs.size(), vv.data(), pr.data(), nr.data(), off.data(), rl.data(), off.back()) != CP3G_OK) { die("bve_resolve failed"); }
            host_ns_gpuwait += now_ns() - t0;
        }
    };
    fetch(0, 0);
    size_t k = 0;
    for (uint32_t cp = 0; cp < positive.n; ++cp) {
        const uint32_t p = LP(positive)[cp];
        if (C[p].flag & F_DEL) { continue; }
        for (uint32_t cn = 0; cn < negative.n; ++cn) {
            if ((int)cp >= p_limit && (int)cn >= n_limit) { continue; }
            if (hasDefinition && (int)cp < p_limit && (int)cn < n_limit) { continue; }
            bve_steps++;
            const uint32_t n = LP(negative)[cn];
            if (C[n].flag & F_DEL) { continue; }
            while (k < prs.size() && (prs[k].cp < cp || (prs[k].cp == cp && prs[k].cn < cn))) { ++k; }
            if (k >= prs.size() || prs[k].cp != cp || prs[k].cn != cn) { continue; }   // tautology / empty
            const Pr& q = prs[k];
            if (q.size == 1) {
                int status = dataEnqueue(rl[q.off]);
                if (status == L_FALSE) { return L_FALSE; }
                else if (status == L_TRUE) {
                    ++bve_unitsEnqueued;
                    if (propagationProcess(true, false, VAR_UNDEF) == L_FALSE) { return L_FALSE; }
                    t_bve.modified = t_bve.modified || t_up.modified;
                    if (positive.n == 0 || negative.n == 0) { return L_UNDEF; }
                    const bool pdel = (C[p].flag & F_DEL) != 0;
                    pe = &pairsFor(v);
                    // continue behind (cp,cn); a deleted p ends its row (BVE.cc:942-944)
                    if (pdel) { fetch(cp + 1, 0); k = 0; break; }
                    fetch(cp, cn + 1); k = 0;
                }
            } else {
                uint32_t cr = allocClause(rl.data() + q.off, q.size);
                dataAddClause(cr, useHeap);
                clauses.push_back(cr);
                subInitClause(cr);
                ++bve_createdClauses; bve_createdLiterals += q.size;
            }
        }
    }
    return L_UNDEF;
}

// One step of the staged prefetcher (engine.h).  Candidates = the variables in the first levels of the heap.
void Engine::bvePrefetchStep()
{
    // retire finished / stale entries, advance the others by one dependency level
    int keep = 0;
    for (int i = 0; i < bve_pf_n; ++i) {
        BvePf e = bve_pf[i];
        const int v = e.var;
        if (!inHeap(v)) { continue; }
        switch (e.stage) {
        case 0:                                                   // counters arrived: list descriptors, K5 record slot
            {   // a variable the cut-off (BVE.cc:418-420) will skip needs nothing else
                const bool cutoff = (hot[2 * v + 1].cnt > 10 && hot[2 * v].cnt > 10) || (dcount(v) > 15 && (hot[2 * v + 1].cnt > 5 || hot[2 * v].cnt > 5));
                if ((cutoff && !opt.bve_unlimited && !opt.bve_force_gates && !opt.bve_gates) || assigns[v] != L_UNDEF || frozen[v]) { e.stage = 4; break; }
                if (cutoff && !opt.bve_unlimited && !opt.bve_force_gates && (hot[2 * v].n > 24 || hot[2 * v + 1].n > 24)) { e.stage = 4; break; }
            }
            __builtin_prefetch(&occ[2 * v]);
            if ((size_t)v < pc_slot.size()) { __builtin_prefetch(&pc_slot[v]); }
            __builtin_prefetch(&var_touch[v]);
            break;
        case 1: {                                                 // descriptors arrived: the lists themselves, the record
            if (hot[2 * v].n) { __builtin_prefetch(&occ_pool[occ[2 * v].off]); }
            if (hot[2 * v + 1].n) { __builtin_prefetch(&occ_pool[occ[2 * v + 1].off]); }
            if ((size_t)v < pc_slot.size() && pc_slot[v] >= 0) { __builtin_prefetch(&pc_ent[(size_t)pc_slot[v]]); }
            break; }
        case 2: {                                                 // lists arrived: clause headers; record arrived: its matrix
            for (int s2 = 0; s2 < 2; ++s2) {
                const uint32_t n = std::min<uint32_t>(hot[2 * v + s2].n, 12); const uint32_t* lp = occ_pool.data() + occ[2 * v + s2].off;
                for (uint32_t k = 0; k < n; ++k) { if (lp[k] < C.size()) { __builtin_prefetch(&C[lp[k]]); } }
            }
            if ((size_t)v < pc_slot.size() && pc_slot[v] >= 0) {
                const PairEnt& pe = pc_ent[(size_t)pc_slot[v]];
                if (pe.valid) { __builtin_prefetch(pe.sizes.p); __builtin_prefetch(pe.pos.p); __builtin_prefetch(pe.neg.p); }
            }
            break; }
        case 3: {                                                 // headers arrived: literals of the short clauses (gate detection reads them)
            for (int s2 = 0; s2 < 2; ++s2) {
                const uint32_t n = std::min<uint32_t>(hot[2 * v + s2].n, 12); const uint32_t* lp = occ_pool.data() + occ[2 * v + s2].off;
                for (uint32_t k = 0; k < n; ++k) { if (lp[k] < C.size() && C[lp[k]].size <= 4) { __builtin_prefetch(&pool[C[lp[k]].start]); } }
            }
            break; }
        default: break;
        }
        if (e.stage < 4) { ++e.stage; }
        bve_pf[keep++] = e;
    }
    bve_pf_n = keep;
    // the percolation of the next pops compares the counters of the first levels
    { const int lv = (int)std::min<size_t>(heap.size(), 31); for (int i = 7; i < lv; ++i) { __builtin_prefetch(&hot[2 * heap[(size_t)i]]); } }
    // new candidates: the variables that can become the root within the next two or three pops
    const int top = (int)std::min<size_t>(heap.size(), 7);
    for (int i = 0; i < top && bve_pf_n < 16; ++i) {
        const int v = heap[(size_t)i];
        bool have = false;
        for (int k = 0; k < bve_pf_n; ++k) { if (bve_pf[k].var == v) { have = true; break; } }
        if (have) { continue; }
        __builtin_prefetch(&hot[2 * v]); __builtin_prefetch(&assigns[v]); __builtin_prefetch(&frozen[v]);
        bve_pf[bve_pf_n++] = {v, 0};
    }
}

// bve_worker, BVE.cc:388-637
void Engine::bveWorker()
{
    const bool unlimited = !opt.limited;
    const bool dbg = tune.debug_par || tune.debug_bve;
    int64_t tg = 0, tc = 0, ta = 0, tr = 0, trm = 0, ts = 0, tl = dbg ? now_ns() : 0;
    auto lapt = [&](int64_t& acc) { if (dbg) { const int64_t n2 = now_ns(); acc += n2 - tl; tl = n2; } };
    struct Rep { bool on; int64_t &tg, &tc, &ta, &tr, &trm, &ts; ~Rep() { if (on) { fprintf(stderr, "  bveWorker: heap+gates %.1f ms, clean %.1f, anticipate %.1f, resolveSet %.1f, removeClauses %.1f, inner subsumption %.1f\n", tg / 1e6, tc / 1e6, ta / 1e6, tr / 1e6, trm / 1e6, ts / 1e6); } } } rep{dbg, tg, tc, ta, tr, trm, ts};
    const bool no_eval = tune.no_bve_eval;
    // (-bve_early makes the anticipation depend on the global growth budget; in sharded mode every rank evaluates the whole batch -
    //  the replicas hold the identical data base, nothing has to be exchanged)
    const bool use_eval = !no_eval && !opt.bve_early;
    pairs_single = use_eval;
    struct InWorker { bool& f; explicit InWorker(bool& x) : f(x) { f = true; } ~InWorker() { f = false; } } in_worker(in_bve_worker);
    const bool no_pf = tune.no_bve_prefetch;
    bve_pf_n = 0;
    size_t skip_run = 0, skip_check_at = 1024;
    while (!heap.empty() && (bve_steps < opt.bve_limit || unlimited)) {
        if (!no_pf) { bvePrefetchStep(); }
        const int v = heapRemoveMin();
        if (assigns[v] != L_UNDEF) { continue; }
        if (hot[2 * v].n == 0 && hot[2 * v + 1].n == 0) { continue; }
        if (frozen[v]) { continue; }
        const bool cutoff = (hot[2 * v + 1].cnt > 10 && hot[2 * v].cnt > 10) || (dcount(v) > 15 && (hot[2 * v + 1].cnt > 5 || hot[2 * v].cnt > 5));
        if (!opt.bve_force_gates && !opt.bve_unlimited && cutoff) {
            ++bve_skippedVars;
            // The heap hands out the variables with few occurrences first; behind them come millions that the cut-off
            // (BVE.cc:418-420) skips one by one - pops without any effect but this counter.  After a run of skips, check the
            // whole rest of the heap at once (all cores): if every variable left would be skipped or passed over, count them
            // and empty the heap - the same final state as popping them.
            if (++skip_run >= skip_check_at && heap.size() > 4096) {
                std::vector<int64_t> nsk((size_t)nthreads + 1, 0); bool all = true;
                parallelFor(heap.size(), [&](size_t b, size_t e, int t) {
                    int64_t c = 0; bool good = true;
                    for (size_t i = b; i < e && good; ++i) {
                        const int w = heap[i];
                        if (assigns[w] != L_UNDEF || (hot[2 * w].n == 0 && hot[2 * w + 1].n == 0) || frozen[w]) { continue; }
                        const bool co = (hot[2 * w + 1].cnt > 10 && hot[2 * w].cnt > 10) || (dcount(w) > 15 && (hot[2 * w + 1].cnt > 5 || hot[2 * w].cnt > 5));
                        if (co) { ++c; } else { good = false; }
                    }
                    nsk[(size_t)t] = c; if (!good) { all = false; }
                }, 65536);
                if (all) { for (int64_t c : nsk) { bve_skippedVars += (int)c; } heapClear(); break; }
                skip_check_at = skip_run * 2;                        // not yet: try again after twice as many skips
            }
            continue;
        }
        int p_limit = (int)hot[2 * v].n, n_limit = (int)hot[2 * v + 1].n;
        bool foundGate = false;
        // The device evaluated gates, cleaning and anticipation of this variable on its lists as they were at batch time
        // (K5', bve_eval.cu); the record stands while no clause of v changed (touch counter) and the two lists are still the
        // very same sequences (a lazy cleaning by somebody else's walk reorders a list without touching the variable).
        EvalEnt* ee = use_eval ? evalFor(v) : nullptr;
        if (ee != nullptr) {
            const bool same = !(ee->rec.flags & CP3G_BVE_HOST) && ee->pos0.n == hot[2 * v].n && ee->neg0.n == hot[2 * v + 1].n &&
                              !memcmp(ee->pos0.p, occ_pool.data() + occ[2 * v].off, sizeof(uint32_t) * ee->pos0.n) &&
                              !memcmp(ee->neg0.p, occ_pool.data() + occ[2 * v + 1].off, sizeof(uint32_t) * ee->neg0.n);
            if (!same) { if (ee->rec.flags & CP3G_BVE_HOST) { ++n_eval_host; } else { ++n_eval_stale; } ee = nullptr; }
        }
        if (ee != nullptr) {
            foundGate = (ee->rec.flags & CP3G_BVE_GATE) != 0; p_limit = ee->rec.p_limit; n_limit = ee->rec.n_limit;
            if (foundGate) { bve_foundGates++; }
        } else if (opt.bve_gates) { foundGate = findGates(v, p_limit, n_limit); if (foundGate) { bve_foundGates++; } }
        if (!foundGate && opt.bve_fdepOnly) { continue; }
        if (!opt.bve_unlimited && !foundGate && cutoff) { ++bve_skippedVars; continue; }
        ++bve_testedVars;
        lapt(tg);
        int pos_count = 0, neg_count = 0;
        if (ee != nullptr) {
            // the lists in their reordered + cleaned form; every entry that left was a deleted clause
            for (int s = 0; s < 2; ++s) {
                const int x = 2 * v + s; const Span<uint32_t>& nl = s == 0 ? ee->pos1 : ee->neg1;
                const uint32_t removed = hot[x].n - (uint32_t)nl.n;
                memcpy(occ_pool.data() + occ[x].off, nl.p, sizeof(uint32_t) * nl.n);
                hot[x].n = (uint32_t)nl.n;
                if (removed) { if (stale[x] > removed) { stale[x] -= removed; } else { staleClear(x); } }
            }
            pos_count = (int)ee->pos1.n; neg_count = (int)ee->neg1.n;
            // evaluating the lists in this form again gives the same record (the gate clauses are in front already, nothing is left to
            // clean): it stays valid for the next time v comes out of the heap untouched
            ev_ref[(size_t)v].applied = 1;
            ++n_eval_used;
        } else {
            for (int s = 0; s < 2; ++s) {
                ListRef li = L(2 * v + s);
                for (uint32_t i = 0; i < li.n; ++i) {
                    if (C[LP(li)[i]].flag & F_DEL) { occSwapRemove(2 * v + s, i); --i; continue; }
                    if (s == 0) { ++pos_count; } else { ++neg_count; }
                }
            }
        }
        if (pos_stats.size() < hot[2 * v].n) { pos_stats.resize(hot[2 * v].n, 0); }
        if (neg_stats.size() < hot[2 * v + 1].n) { neg_stats.resize(hot[2 * v + 1].n, 0); }
        int lit_clauses = 0, resolvents = 0;
        int ares = L_UNDEF;
        lapt(tc);
        if (pos_count != 0 && neg_count != 0) {
            ++bve_anticipations;
            if (ee != nullptr) {
                bve_steps += ee->rec.steps; lit_clauses = (int)ee->rec.lit_clauses; resolvents = (int)ee->rec.resolvents;
                if (opt.bve_BCElim) { for (size_t k = 0; k < ee->pst.n; ++k) { pos_stats[k] = ee->pst[k]; } for (size_t k = 0; k < ee->nst.n; ++k) { neg_stats[k] = ee->nst[k]; } }
            } else {
                ares = anticipate(v, p_limit, n_limit, lit_clauses, resolvents);
                if (ares == L_FALSE) { return; }
            }
        }
        if (ares == L_UNDEF && opt.bve_BCElim) {
            bveRemoveBlocked(2 * v, pos_stats);
            bveRemoveBlocked(2 * v + 1, neg_stats);
        }
        lapt(ta);
        int compareNumber = pos_count + neg_count + opt.bve_cgrow;
        if (opt.bve_cgrow_t != INT32_MAX && opt.bve_cgrow_t > opt.bve_cgrow) { compareNumber += opt.bve_cgrow_t - bve_totallyAdded; }
        bool reducedClss = resolvents <= compareNumber;
        if (bve_totallyAdded > opt.bve_cgrow_t) { reducedClss = false; }
        if (reducedClss) { bve_totallyAdded += resolvents - (pos_count + neg_count); }
        if (reducedClss && !opt.bce_only) {
            if (resolvents < pos_count + neg_count) { bve_nDec++; }
            else if (resolvents > pos_count + neg_count) { bve_nInc++; }
            else { bve_nKeep++; }
            lapt(tg);
            // The elimination touches, per literal of every resolvent and of every parent clause, the counters, the list tail, the
            // heap slot and the touch record of that literal's variable - dependent cache misses on a 10 M variable mirror.  They
            // are requested level by level ahead of their use (side-effect free).
            auto prefetchParents = [&](int stage) {
                for (int s2 = 0; s2 < 2; ++s2) {
                    const uint32_t* lp = occ_pool.data() + occ[2 * v + s2].off; const uint32_t n = std::min<uint32_t>(hot[2 * v + s2].n, 24);
                    for (uint32_t k = 0; k < n; ++k) {
                        const uint32_t c = lp[k];
                        if (stage == 0) { __builtin_prefetch(&C[c]); if (c < log_head.size()) { __builtin_prefetch(&log_head[c]); } }
                        else if (stage == 1) { __builtin_prefetch(&pool[C[c].start]); }
                        else {
                            const int32_t* l = pool.data() + C[c].start; const uint32_t m = std::min<uint32_t>(C[c].size, 8);
                            for (uint32_t j = 0; j < m; ++j) { const int x = l[j], w = x >> 1; __builtin_prefetch(&hot[x]); __builtin_prefetch(&hidx[w]); __builtin_prefetch(&modTimer[w]); __builtin_prefetch(&stale[x]); __builtin_prefetch(&var_touch[w]); __builtin_prefetch(&bve_inpend[w]); }
                        }
                    }
                }
            };
            prefetchParents(0);
            if (ee != nullptr && ee->has_stream && !opt.bve_BCElim) {
                // resolveSet (BVE.cc:892-1046) from the K6' stream: the resolvents in the reference's order, nothing to look up
                const bool useHeap = opt.bve_heap_updates > 0;
                bve_steps += ee->rec.steps;
                const int32_t* rl = ee->rlits.p;
                for (size_t k = 0; k < ee->rlits.n; ++k) { const int x = rl[k], w = x >> 1; __builtin_prefetch(&hot[x]); __builtin_prefetch(&occ[x]); __builtin_prefetch(&var_touch[w]); __builtin_prefetch(&hidx[w]); __builtin_prefetch(&bve_inpend[w]); }
                prefetchParents(1);
                for (size_t k = 0; k < ee->rlits.n; ++k) { const int x = rl[k]; __builtin_prefetch(&occ_pool[(size_t)occ[x].off + hot[x].n]); const int hi = hidx[x >> 1]; if (hi >= 0) { __builtin_prefetch(&heap[(size_t)hi]); } }
                prefetchParents(2);
                for (size_t k = 0; k < ee->rsz.n; ++k) {
                    const int sz = ee->rsz[k];
                    const uint32_t cr = allocClause(rl, sz);
                    dataAddClause(cr, useHeap);
                    clauses.push_back(cr);
                    subInitClause(cr);
                    ++bve_createdClauses; bve_createdLiterals += sz;
                    rl += sz;
                }
                ++n_stream_used;
            } else { prefetchParents(1); if (resolveSet(v, p_limit, n_limit) == L_FALSE) { return; } prefetchParents(2); }
            lapt(tr);
            ++bve_eliminatedVars;
            bveRemoveClauses(2 * v, -1);
            if (assigns[v] == L_UNDEF) { bveRemoveClauses(2 * v + 1, 2 * v + 1); extLit(2 * v); }
            else { bveRemoveClauses(2 * v + 1, -1); }
            occRelease(2 * v); occRelease(2 * v + 1);
            lapt(trm);
            if (opt.bve_heap_updates > 0) { subsumptionProcess(opt.bve_strength, true, v); }
            else { subsumptionProcess(opt.bve_strength, false, VAR_UNDEF); }
            t_bve.modified = t_bve.modified || t_sub.modified;
            lapt(ts);
            if (!ok) { return; }
        }
        std::fill(pos_stats.begin(), pos_stats.end(), 0); std::fill(neg_stats.begin(), neg_stats.end(), 0);
    }
}

uint32_t Engine::nextModStep()
{
    if (modStep >= (1u << 30)) { std::fill(modTimer.begin(), modTimer.end(), 0u); modStep = 0; }
    return ++modStep;
}

// addClausesToSubsumption, BVE.cc:1129-1141
void Engine::addClausesToSubsumption(int x)
{
    ListRef li = L(x);
    for (uint32_t j = 0; j < li.n; ++j) {
        uint32_t d = LP(li)[j];
        if (!(C[d].flag & F_DEL) && !(C[d].flag & F_SUB)) { C[d].flag |= F_SUB | F_STR; subInitClause(d); }
    }
}

// addClausesToSubsumption (BVE.cc:1129-1141) over both lists of every touched variable, in order: a live clause that is not in
// the subsumption queue yet is queued (both queues) where the walk meets it FIRST.  For many touched variables the first
// occurrences are found on all cores: every list entry has a rank in the sequential walk, the smallest rank per clause wins.
void Engine::addTouchedToSubsumption()
{
    const size_t nt = touched.size();
    std::vector<uint64_t> off(2 * nt + 1, 0);
    for (size_t i = 0; i < nt; ++i) { off[2 * i + 1] = off[2 * i] + hot[2 * touched[i]].n; off[2 * i + 2] = off[2 * i + 1] + hot[2 * touched[i] + 1].n; }
    const uint64_t total = off[2 * nt];
    const uint64_t par_min = tune.touched_min;
    if (total < par_min || total >= 0xffffffffull || nthreads < 2) {
        for (int v : touched) { addClausesToSubsumption(2 * v); addClausesToSubsumption(2 * v + 1); }
        return;
    }
    pgrow(first_rank, C.size() + 1024, 0xffffffffu);
    auto forEntries = [&](size_t b, size_t e, auto&& f) {              // lists b..e-1 of the walk (list 2i = positive list of touched[i])
        for (size_t li = b; li < e; ++li) {
            const int x = 2 * touched[li >> 1] + (int)(li & 1);
            const uint32_t* p = occ_pool.data() + occ[x].off; const uint32_t n = hot[x].n;
            for (uint32_t j = 0; j < n; ++j) { f(p[j], (uint32_t)(off[li] + j)); }
        }
    };
    parallelFor(2 * nt, [&](size_t b, size_t e, int) {
        forEntries(b, e, [&](uint32_t d, uint32_t r) {
            if ((C[d].flag & F_DEL) || (C[d].flag & F_SUB)) { return; }
            uint32_t o = __atomic_load_n(&first_rank[d], __ATOMIC_RELAXED);
            while (r < o && !__atomic_compare_exchange_n(&first_rank[d], &o, r, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
        });
    }, 4096);
    std::vector<size_t> cnt((size_t)nthreads + 2, 0);
    parallelFor(2 * nt, [&](size_t b, size_t e, int t) {
        size_t c = 0;
        forEntries(b, e, [&](uint32_t d, uint32_t r) { c += first_rank[d] == r ? 1 : 0; });
        cnt[(size_t)t + 1] = c;
    });
    for (size_t t = 1; t < cnt.size(); ++t) { cnt[t] += cnt[t - 1]; }
    const size_t s0 = subq.size(), q0 = strq.size(), add = cnt.back();
    subq.resize(s0 + add); strq.resize(q0 + add);
    parallelFor(2 * nt, [&](size_t b, size_t e, int t) {
        size_t o = cnt[(size_t)t];
        forEntries(b, e, [&](uint32_t d, uint32_t r) { if (first_rank[d] == r) { subq[s0 + o] = d; strq[q0 + o] = d; ++o; } });
    });
    parallelFor(add, [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { const uint32_t d = subq[s0 + i]; C[d].flag |= F_SUB | F_STR; first_rank[d] = 0xffffffffu; } });
}

// sequentiellBVE, BVE.cc:316-381
void Engine::sequentiellBVE()
{
    const bool unlimited = !opt.limited;
    const bool dbg = tune.debug_par || tune.debug_bve;
    int64_t tl = dbg ? now_ns() : 0;
    auto lapb = [&](const char* w) { if (dbg) { const int64_t n2 = now_ns(); fprintf(stderr, "  seqBVE %s %.1f ms\n", w, (n2 - tl) / 1e6); tl = n2; } };
    subsumptionProcess(opt.bve_strength, false, VAR_UNDEF);
    t_bve.modified = t_bve.modified || t_sub.modified;
    if (!ok) { return; }
    lapb("first subsumption");
    touched.clear();
    while (!heap.empty() && (bve_steps < opt.bve_limit || unlimited)) {
        bve_modtime = nextModStep();
        for (int w : heap) { if (!bve_inpend[w]) { bve_inpend[w] = 1; bve_pending.push_back(w); } }
        lapb("pending");
        bveWorker();
        lapb("worker");
        if (!ok) { return; }
        if (hasToPropagate()) {
            if (propagationProcess(true, false, VAR_UNDEF) == L_FALSE) { return; }
            t_bve.modified = t_bve.modified || t_up.modified;
        }
        checkGarbage();
        lapb("propagate+gc");
        {   // getActiveVariables (CoprocessorTypes.h:1169-1187): stable parallel filter over the variables
            std::vector<size_t> cntv((size_t)nthreads + 2, 0);
            parallelFor((size_t)nvars, [&](size_t b, size_t e, int t) { size_t c = 0; for (size_t v = b; v < e; ++v) { c += modTimer[v] >= bve_modtime ? 1 : 0; } cntv[(size_t)t + 1] = c; });
            for (size_t t = 1; t < cntv.size(); ++t) { cntv[t] += cntv[t - 1]; }
            touched.resize(cntv.back());
            parallelFor((size_t)nvars, [&](size_t b, size_t e, int t) { size_t o = cntv[(size_t)t]; for (size_t v = b; v < e; ++v) { if (modTimer[v] >= bve_modtime) { touched[o++] = (int)v; } } });
        }
        addTouchedToSubsumption();
        lapb("touched");
        subsumptionProcess(opt.bve_strength, false, VAR_UNDEF);
        t_bve.modified = t_bve.modified || t_sub.modified;
        lapb("subsumption");
        heapClear();
        for (int v : touched) { heapInsert(v); }
        touched.clear();
        lapb("heap");
    }
}

// BoundedVariableElimination::process, BVE.cc:170-273
int Engine::bveProcess()
{
    if (!performSimplification(t_bve)) { return L_UNDEF; }
    ensureHostLists();
    for (int v = 0; v < nvars; ++v) { if (modTimer[v] >= bve_modtime && !inHeap(v)) { heapInsert(v); } }
    if (propagationProcess(true, false, VAR_UNDEF) == L_FALSE) { return L_FALSE; }
    const bool propagatedSomething = t_up.modified;
    sequentiellBVE();
    pair_blocks.clear(); pc_ent.clear(); std::fill(pc_slot.begin(), pc_slot.end(), -1);
    eval_blocks.clear(); ev_ref.clear();
    if (!t_bve.modified) { unsuccessful(t_bve); }
    if (t_bve.modified || propagatedSomething) { successful(t_bve); }
    return ok ? L_UNDEF : L_FALSE;
}

} // namespace cp3
