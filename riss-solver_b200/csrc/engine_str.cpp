// engine_str.cpp - strengthening commit of the host engine (static plan, walk plans, ordered walk, naive variant) and
// Subsumption::process.
#include "engine_util.h"
#include <atomic>
#include <unordered_map>
#include <unordered_set>

namespace cp3 {

// updateOccurrences, Subsumption.cc:1530-1544
void Engine::updateOccurrences(bool useHeap, int ignore)
{
    // from here on nothing is pending any more (setDeleted() consults pend_rm)
    std::vector<OccUpd> ups; ups.swap(occ_upd);
    ++pend_epoch;
    if (useHeap || in_bve_worker || ups.size() < std::max<size_t>(tune.par_min, 1) || nthreads < 2) {
        for (const OccUpd& u : ups) {
            if (removeClauseFrom(u.cr, u.l)) { removedLiteral(u.l, 1, useHeap, ignore != 0, VAR_UNDEF); }
        }
        return;
    }
    // Long update lists (the first strengthening pass of a large formula): an update touches nothing but the list, the counter and
    // the stale count of its own literal, and the order matters only among updates of the same list.  Literal buckets (64-aligned:
    // no two share a word of the stale bitmap) in update order, one core each.
    const size_t nlit = (size_t)2 * nvars; const int NB = nthreads;
    auto bucketLo = [&](int bk) -> size_t { return std::min(nlit, ((nlit * (size_t)bk / (size_t)NB) + 63) / 64 * 64); };
    const size_t bwidth = std::max<size_t>(64, (nlit / (size_t)NB) / 64 * 64);
    auto bucketOf = [&](int x) -> int { int bk = (int)std::min<size_t>((size_t)NB - 1, (size_t)x / bwidth); while (bk > 0 && (size_t)x < bucketLo(bk)) { --bk; } while (bk + 1 < NB && (size_t)x >= bucketLo(bk + 1)) { ++bk; } return bk; };
    std::vector<std::vector<OccUpd>> parts(((size_t)nthreads + 1) * (size_t)NB);
    parallelFor(ups.size(), [&](size_t b, size_t e, int t) {
        std::vector<OccUpd>* out = &parts[(size_t)t * (size_t)NB];
        for (size_t i = b; i < e; ++i) { out[bucketOf(ups[i].l)].push_back(ups[i]); }
    }, 4096);
    std::vector<int64_t> removed((size_t)NB, 0);
    parallelFor((size_t)NB, [&](size_t bb, size_t be, int) {
        for (size_t bk = bb; bk < be; ++bk) {
            int64_t r = 0;
            for (size_t t2 = 0; t2 <= (size_t)nthreads; ++t2) {                     // producer chunks are contiguous ranges of `ups`, in order
                for (const OccUpd& u : parts[t2 * (size_t)NB + bk]) {
                    const int x = u.l; uint32_t* lp = occ_pool.data() + occ[x].off; uint32_t n = hot[x].n;
                    for (uint32_t i = 0; i < n; ++i) {
                        if (lp[i] != u.cr) { continue; }
                        if (isDel(lp[i])) { staleDec(x); }
                        lp[i] = lp[n - 1]; hot[x].n = n - 1;
                        deletedVar(x >> 1); hot[x].cnt -= 1; ++r;                  // removedLiteral(l, 1) without a heap
                        break;
                    }
                }
            }
            removed[bk] = r;
        }
    }, 1);
    for (int64_t r : removed) { ca_wasted += (double)r; }
}

// the hit branch of both passes, Subsumption.cc:1385-1413 / 1474-1508
void Engine::strengthenHit(std::vector<uint32_t>& localQueue, uint32_t other, int pos)
{
    ++removedLiterals;
    if (static_lazy && lazyPending(other)) { C[other].flag |= F_STR; }      // still waiting in the queue
    if (!(C[other].flag & F_SUB)) { C[other].flag |= F_SUB; subq.push_back(other); }
    if (!(C[other].flag & F_STR)) { localQueue.push_back(other); C[other].flag |= F_STR; }
    successful(t_sub);
    const int lit = CL(other)[pos];
    occ_upd.push_back({other, lit});
    removePosSorted(other, pos);
    noteShrunk(other, lit, true);
    unpredictedEdit(other);
    dirty_str.push_back(other);
    if (C[other].size == 1) {
        dataEnqueue(CL(other)[0]);
        setDeleted(other);
    }
}

// strengthening hits of cr for its current content: cached, or one more K4 batch (cr + the other pending
// strengtheners that were edited after their scan)
const Engine::CacheEnt* Engine::strHits(uint32_t cr)
{
    const CacheEnt* e = cached(c_str, cr);
    if (e) { if (e->ver == UINT32_MAX) { ++n_spec_used; } return e; }
    ++n_ondemand;
    if (tune.debug_ondemand && n_ondemand < 12) {
        fprintf(stderr, "ondemand cr=%u size=%u ver=%u stamp=%u flag=%u slot=%d bulk_all=%d bulk_scan=%u bulk_ncl=%u qtime=%d\n", cr, (unsigned)C[cr].size, C[cr].ver, C[cr].stamp,
                (unsigned)C[cr].flag, cr < c_str.slot.size() ? c_str.slot[cr] : -9, (int)c_str.bulk_all, c_str.bulk_scan, c_str.bulk_ncl, cr < qtime.size() ? qtime[cr] : -9);
        for (int32_t i = cr < c_str.slot.size() ? c_str.slot[cr] : -1; i >= 0; i = c_str.ent[i].next) {
            fprintf(stderr, "   ent ver=%u n=%u time=%f hits=%u\n", c_str.ent[i].ver, c_str.ent[i].lit_n, c_str.ent[i].time, c_str.ent[i].he - c_str.ent[i].hb);
        }
    }
    std::vector<uint32_t> ids;
    pgrow(c_str.slot, C.size(), -1);
    for (uint32_t c : dirty_str) {
        if (c == cr || (C[c].flag & F_DEL) || !(C[c].flag & F_STR) || C[c].size < 2) { continue; }
        if (cached(c_str, c)) { continue; }
        ids.push_back(c);
    }
    dirty_str.clear();
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    ids.insert(ids.begin(), cr);
    scanClauses(CP3G_STRENGTHEN, ids, c_str);
    e = cached(c_str, cr);
    if (!e) { die("internal: strengthener without scan result"); }
    return e;
}

// Put one list that was cleaned up front back into its lazy state (nobody has walked it yet).
void Engine::restoreOneList(int x)
{
    const int32_t i = pre_index[x];
    if (i < 0 || scanned_dyn[x] || (static_lazy && last_static_walk[x] > (int64_t)cur_end)) { return; }
    const PreList& pl = pre_lists[i];
    memcpy(occ_pool.data() + occ[x].off, pre_snap.data() + pl.snap_off, sizeof(uint32_t) * pl.old_n);
    hot[x].n = pl.old_n;
    for (uint32_t k = 0; k < pl.extra; ++k) { staleInc(x); }
    str_steps -= pl.extra;
    pre_index[x] = -1;
}
// A unit is about to be propagated: clauses will die, and the reference would clean the lists nobody has walked
// yet together with those new deletions (swap-with-last cleaning is not associative).  Put every list that was
// cleaned up front but not walked so far back into its lazy state; from here on the pass runs the plain lazy way.
void Engine::rollbackPrecompact()
{
    if (!pre_active) { return; }
    for (const PreList& pl : pre_lists) { restoreOneList(pl.lit); }
    pre_active = false;
}
bool Engine::lazyPending(uint32_t c) const
{
    if (c >= qtime.size() || qtime[c] < 0) { return false; }
    const size_t pos = static_nq - 1 - (size_t)qtime[c];
    return pos < cur_end && pos < contrib.size() && contrib[pos] >= 0;
}
// The lazy static plan stops here: inert entries that still wait get their flag back, the lists walked by the
// ones already passed are marked as walked.  From here on the pass runs entry by entry.
void Engine::leaveLazy()
{
    if (!static_lazy) { return; }
    static_lazy = false;
    const size_t nq = static_nq, ce = std::min(cur_end + 1, nq);
    parallelFor(ce, [&](size_t b, size_t e, int) { for (size_t p = b; p < e; ++p) { if (contrib[p] >= 0) { C[strq[p]].flag |= F_STR; } } });
    if (pre_active) { for (size_t p = ce; p < nq; ++p) { if (contrib[p] >= 0 && walk_min[p] >= 0) { scanned_dyn[walk_min[p]] = 1; scanned_dyn[walk_min[p] ^ 1] = 1; } } }
}
// The static plan assumed that clause c keeps its content until its turn (it was nobody's predicted target).
// A hit that the speculative closure did not predict just changed it: c is no longer inert, and the lists it
// was going to walk lose one certain walker - a list without any is put back into its lazy state.
void Engine::unpredictedEdit(uint32_t c)
{
    if (!static_pass || c >= qtime.size() || qtime[c] < 0) { return; }
    const size_t pos = static_nq - 1 - (size_t)qtime[c];
    if (pos >= walk_min.size()) { return; }
    if (static_lazy && contrib[pos] >= 0 && pos < cur_end) { lazy_flipped.push((uint32_t)pos); }     // it is a real entry from now on
    contrib[pos] = -1;
    const int mn = walk_min[pos];
    if (mn < 0) { return; }
    walk_min[pos] = -1;
    for (int x : {mn, mn ^ 1}) {
        if (walk_count[x] > 0 && --walk_count[x] == 0 && pre_active) { restoreOneList(x); }
    }
}

// K4 over the strengthening queue + speculative closure (first half of a strengthening pass)
void Engine::prepareStrengthening(const std::vector<uint32_t>& ids, bool bulk)
{
    DbgLap lap("prepStr", tune.debug_par);
    c_str.clear();
    lap("clear");
    scanClauses(CP3G_STRENGTHEN, ids, c_str, scan_no_flush ? (bulk ? 1 : 0) : -1);
    lap("scan");
    // processing time of the queue entries: the queue is drained from the back
    if (qtime.size() < C.size()) { pgrow(qtime, C.size() + 1024, -1); }
    // (a clause queued twice keeps the time of its last entry, like the sequential fill)
    { std::atomic<bool> inc{true};
      const size_t nq = strq.size();
      parallelFor(nq, [&](size_t b, size_t e, int) { bool g = true; for (size_t i = std::max<size_t>(b, 1); i < e && g; ++i) { g = strq[i - 1] < strq[i]; } if (!g) { inc = false; } });
      if (inc) { parallelFor(nq, [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { qtime[strq[i]] = (int32_t)(nq - 1 - i); } }); }
      else { for (size_t i = 0; i < nq; ++i) { qtime[strq[i]] = (int32_t)(nq - 1 - i); } } }
    // the records of this scan belong to the clauses in c_str.used: their queue time is the time of the record
    parallelFor(c_str.ent.size(), [&](size_t b, size_t e, int) { for (size_t k = b; k < e; ++k) { c_str.ent[k].time = 0; } }, 16384);
    parallelFor(c_str.used.size(), [&](size_t b, size_t e, int) {
        for (size_t u = b; u < e; ++u) { const uint32_t c = c_str.used[u]; if (c_str.slot[c] >= 0 && qtime[c] >= 0) { c_str.ent[(size_t)c_str.slot[c]].time = (double)qtime[c]; } }
    }, 8192);
    lap("qtime");
    prefetchStrengthenClosure();
    lap("closure");
}

// strengthening_worker, Subsumption.cc:1250-1528 (all_strength_res == 0)
int Engine::strengtheningWorker(bool useHeap, int ignore)
{
    PhaseTimer pt(ph_str);
    const bool unlimited = !opt.limited;
    DbgLap lap("str", tune.debug_par);
    ensureHostLists();
    logClear(); dirty_str.clear();
    lap("host lists");
    if (str_prepared) { str_prepared = false; }            // scan + closure were done while the subsumption commit ran
    else if (opt.strength) {
        std::vector<uint32_t> ids;
        parallelFilter(strq, ids, [&](uint32_t c) { return !(C[c].flag & F_DEL) && (C[c].flag & F_STR) && C[c].size >= 2; });
        sortUniqueIds(ids);
        lap("ids");
        prepareStrengthening(ids, false);
    } else { c_str.clear(); }
    lap("prepare");
    // first pass behind a device-committed subsumption pass: the static plan comes from the device (K10), the real entries are
    // committed by component; false = nothing was touched, the host plan below takes the pass
    bool fast_done = false;
    if (dev_static_ok && opt.strength && !tune.no_devstatic && !tune.no_components && nthreads > 1 && !hasToPropagate() &&
        strq.size() >= std::max<size_t>(tune.par_min, 1)) {
        fast_done = strengthenFirstPassDevice();
        lap("device static + components");
    }
    dev_static_ok = false;
    const bool fine = lap.on && tune.debug_fine;          // per-entry timers (they cost as much as the entries)
    int64_t t0 = now_ns();
    // Statically inert queue entries.  While no unit is propagated, lit_occurrence_count and the occurrence lists
    // are frozen during this pass (occurrence updates are deferred, Subsumption.cc:1400,1523), so an entry that
    // (a) has no candidate hit, (b) is no candidate target of anybody and (c) walks two lists without deleted
    // entries only adds |occ(min)| - 1 + |occ(~min)| steps - independent of everything else.  Those are
    // evaluated on all host cores; the ordered walk below just adds the number up when it passes them.
    if (!fast_done) { pfill(contrib, strq.size(), -1); pfill(contrib_t, strq.size(), -1); tmin.resize(strq.size()); plan_ver.resize(strq.size()); }
    const size_t static_min = tune.static_min;
    bool static_valid = opt.strength && strq.size() > static_min && !hasToPropagate() && !fast_done;
    if (static_valid) {
        is_target.assign((C.size() >> 6) + 1, 0ull);
        for (const Hit& h : c_str.hits) { is_target[h.clause >> 6] |= 1ull << (h.clause & 63); }
        const size_t nq = strq.size();
        // the counters are frozen while this plan is valid: a 2-byte copy per variable (2 MB at 1 M variables) keeps
        // the min-count choices of the two passes below in the cores' caches; out-of-range values read the original
        std::vector<int16_t> dc16((size_t)std::max(nvars, 1));
        parallelFor((size_t)nvars, [&](size_t b, size_t e, int) {
            for (size_t v = b; v < e; ++v) { const int d = dcount((int)v); dc16[v] = (d > 32000 || d < -32000) ? INT16_MIN : (int16_t)d; }
        });
        auto dcq = [&](int v) -> int { const int16_t q = dc16[(size_t)v]; return q == INT16_MIN ? dcount(v) : (int)q; };
        // (1) lists that are certainly walked in this pass: the two lists of every pending clause whose content
        //     cannot change before its turn (it is nobody's candidate target)
        std::vector<uint8_t> walked((size_t)2 * nvars, 0);
        std::vector<int64_t> bound((size_t)nthreads + 1, 0);
        pfill(walk_min, nq, -1); pfill(walk_count, (size_t)2 * nvars, 0u);
        static_pass = true; static_nq = nq;
        parallelFor(nq, [&](size_t b, size_t e, int t) {
            int64_t bs = 0;
            for (size_t p2 = b; p2 < e; ++p2) {
                const uint32_t c = strq[p2];
                const ClauseInfo ci = C[c];
                if ((ci.flag & F_DEL) || !(ci.flag & F_STR) || ci.size < 2) { continue; }
                if (qtime[c] != (int32_t)(nq - 1 - p2)) { continue; }                      // duplicate queue entry
                if ((is_target[c >> 6] >> (c & 63)) & 1) { continue; }
                const int32_t* s2 = pool.data() + ci.start; int minT = s2[0]; int best = dcq(minT >> 1);
                for (uint32_t j = 1; j < ci.size; ++j) { const int dj = dcq(s2[j] >> 1); if (best > dj) { best = dj; minT = s2[j]; } }
                // (only lists that hold deleted entries can be pre-compacted: the walker count and the mark matter for them alone)
                walk_min[p2] = minT;
                if (hasStale(minT)) { walked[minT] = 1; __atomic_fetch_add(&walk_count[minT], 1u, __ATOMIC_RELAXED); }
                if (hasStale(minT ^ 1)) { walked[minT ^ 1] = 1; __atomic_fetch_add(&walk_count[minT ^ 1], 1u, __ATOMIC_RELAXED); }
                bs += (int64_t)hot[minT].n + hot[minT ^ 1].n;
            }
            bound[t] = bs;
        });
        int64_t total_bound = 0; for (int64_t x : bound) { total_bound += x; }
        lap("static1");
        // (2) the first walk of a list removes its deleted entries by swap-with-last (Subsumption.cc:1332-1335,
        //     1432-1435); the resulting order depends only on the list and on which entries are deleted, and no
        //     clause dies during this pass unless a unit shows up.  So the certainly-walked lists are cleaned
        //     up front, on all cores, each costing one step per removed entry like the lazy walk would.
        pre_lists.clear(); pre_snap.clear(); pre_active = false;
        const bool no_pre = tune.no_precompact;
        if (!no_pre && (unlimited || str_steps + 2 * total_bound < str_lim)) {
            const size_t nlit = (size_t)2 * nvars;
            const size_t words = (nlit + 63) / 64;
            std::vector<std::vector<PreList>> tl((size_t)nthreads + 1); std::vector<std::vector<uint32_t>> ts((size_t)nthreads + 1);
            parallelForG(words, 512, [&](size_t wb, size_t we, int t) {  // 64-literal granules: stalebits words are not shared
                for (size_t x = wb * 64; x < std::min(nlit, we * 64); ++x) {
                    if (!walked[x] || !hasStale((int)x)) { continue; }
                    uint32_t* p = occ_pool.data() + occ[x].off; uint32_t n = hot[x].n;
                    PreList pl; pl.lit = (int)x; pl.snap_off = (uint32_t)ts[t].size(); pl.old_n = n; pl.extra = 0;
                    ts[t].insert(ts[t].end(), p, p + n);                 // keep the lazy state: a unit may force us to take this back
                    for (uint32_t j = 0; j < n;) { if (isDel(p[j])) { p[j] = p[n - 1]; --n; ++pl.extra; } else { ++j; } }
                    hot[x].n = n; staleClear((int)x);
                    tl[t].push_back(pl);
                }
            });
            for (size_t t = 0; t < tl.size(); ++t) {
                const uint32_t base = (uint32_t)pre_snap.size();
                pre_snap.insert(pre_snap.end(), ts[t].begin(), ts[t].end());
                for (PreList pl : tl[t]) { pl.snap_off += base; str_steps += pl.extra; pre_lists.push_back(pl); }
            }
            pre_active = !pre_lists.empty();
            if (pre_active) {
                scanned_dyn.assign(nlit, 0); pre_index.assign(nlit, -1);
                for (size_t i = 0; i < pre_lists.size(); ++i) { pre_index[pre_lists[i].lit] = (int32_t)i; }
            }
        }
        lap("precompact");
        // (3) step contribution of the inert entries.  Far from the step limit their whole effect is applied here
        //     (flag cleared, walked lists recorded) and the ordered walk below only adds the number up.
        const bool no_lazy = tune.no_lazy;
        const bool lazy = !no_lazy && (unlimited || str_steps + 2 * total_bound < str_lim);
        static_lazy = lazy; n_lazy = 0; cur_end = nq;
        if (lazy) { last_static_walk.assign((size_t)2 * nvars, -1); }
        auto amax = [](int32_t* a, int32_t v) { int32_t o = __atomic_load_n(a, __ATOMIC_RELAXED); while (o < v && !__atomic_compare_exchange_n(a, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} };
        parallelFor(nq, [&](size_t b, size_t e, int) {
            for (size_t p2 = b; p2 < e; ++p2) {
                const uint32_t c = strq[p2];
                if (walk_min[p2] >= 0) {
                    // pass (1) has seen this entry: live, flagged, nobody's candidate target - and it has chosen its literal
                    if (c_str.slot[c] >= 0) { continue; }                   // a record exists: own hits
                    const int mn = walk_min[p2], nm = mn ^ 1;
                    if (hasStale(mn) || hasStale(nm)) { continue; }
                    contrib[p2] = (int32_t)((hot[mn].n ? hot[mn].n - 1 : 0) + hot[nm].n);
                    if (lazy) { C[c].flag &= ~F_STR; amax(&last_static_walk[mn], (int32_t)p2); amax(&last_static_walk[nm], (int32_t)p2); }
                    continue;
                }
                if (!((is_target[c >> 6] >> (c & 63)) & 1)) { continue; }   // not a target and not chosen in pass (1): dead, unflagged, duplicate
                const ClauseInfo ci = C[c];
                if ((ci.flag & F_DEL) || !(ci.flag & F_STR) || ci.size < 2) { continue; }
                if (qtime[c] != (int32_t)(nq - 1 - p2)) { continue; }
                const bool target = true;
                if (c_str.slot[c] >= 0) {
                    // a record exists: own hits, or - for a candidate target - only the predicted shorter versions
                    if (!target) { continue; }
                    const CacheEnt* e0 = cached(c_str, c);
                    if (!e0 || e0->hb != e0->he) { continue; }
                }
                const int32_t* s2 = pool.data() + ci.start; int minT = s2[0]; int best = dcq(minT >> 1);
                for (uint32_t j = 1; j < ci.size; ++j) { const int dj = dcq(s2[j] >> 1); if (best > dj) { best = dj; minT = s2[j]; } }
                const int mn = minT, nm = minT ^ 1;
                if (hasStale(mn) || hasStale(nm)) { continue; }
                if (target) {
                    // somebody's candidate target without hits of its own: inert as well IF its content is still the
                    // planned one when its turn comes (checked there by the content version)
                    contrib_t[p2] = (int32_t)((hot[mn].n ? hot[mn].n - 1 : 0) + hot[nm].n); tmin[p2] = mn; plan_ver[p2] = ci.ver;
                    continue;
                }
                contrib[p2] = (int32_t)((hot[mn].n ? hot[mn].n - 1 : 0) + hot[nm].n);
                if (lazy) { C[c].flag &= ~F_STR; amax(&last_static_walk[mn], (int32_t)p2); amax(&last_static_walk[nm], (int32_t)p2); }
            }
        });
    }
    lap("contrib");
    // In lazy mode the walk visits only the entries that are not inert ("real": own hits, candidate targets, deleted or
    // unflagged clauses, lists with deleted entries); the inert runs in between are added up from their contributions.
    std::vector<uint32_t> reals; lazy_flipped = std::priority_queue<uint32_t>();
    if (static_lazy) {
        const size_t nq = strq.size();
        std::vector<size_t> cntp((size_t)nthreads + 2, 0);
        parallelFor(nq, [&](size_t b, size_t e, int t) { size_t c = 0; for (size_t i = b; i < e; ++i) { c += contrib[i] < 0 ? 1 : 0; } cntp[(size_t)t + 1] = c; });
        for (size_t t = 1; t < cntp.size(); ++t) { cntp[t] += cntp[t - 1]; }
        reals.resize(cntp.back());
        parallelFor(nq, [&](size_t b, size_t e, int t) { size_t o = cntp[(size_t)t]; for (size_t i = b; i < e; ++i) { if (contrib[i] < 0) { reals[o++] = (uint32_t)i; } } });
    }
    size_t ri = reals.size();                                       // reals[0..ri) are still ahead (the walk goes down)
    lap("reals");
    // the real entries by dependency component on all cores; false = nothing was touched, the ordered walk below takes the pass
    bool par_done = fast_done;
    if (!fast_done && static_lazy && !tune.no_components && nthreads > 1 && reals.size() >= std::max<size_t>(tune.par_min, 1) && !hasToPropagate()) {
        par_done = strengthenComponents(reals);
        if (!par_done) { ++n_compar_fallback; }
        lap("components");
    }
    // Walk plan of the real entries.  While the plan is valid (no unit, frozen counters, clean lists) the two list walks of an
    // entry are a pure function of its content: the step count is the length of the two lists and the candidates that can
    // hit are the cached hits in list order.  Both are evaluated here on all cores for the content the entry has NOW; the
    // ordered walk below uses the record when the entry's turn comes and its content version is still the planned one, and
    // then only validates and applies the candidate hits (the order-dependent part).  Everything else takes the walk itself.
    struct PlanEnt { uint32_t ver, scan; int32_t min; uint32_t steps1, steps2, cb, c1, c2; uint32_t t; uint32_t lit_off, lit_n; int32_t alt; };
    RawVec<PlanEnt> plan; std::vector<std::vector<Hit>> plan_c((size_t)nthreads + 1); std::vector<std::vector<PlanEnt>> plan_alt((size_t)nthreads + 1);
    const bool no_plan = tune.no_plan;
    const bool planned = static_lazy && !no_plan && !reals.empty() && !par_done;
    if (planned) {
        const size_t nq = strq.size();
        plan.resize(reals.size());
        parallelFor(reals.size(), [&](size_t b, size_t e, int t) {
            std::vector<Hit>& pc = plan_c[(size_t)t]; std::vector<PlanEnt>& pa = plan_alt[(size_t)t];
            // the record of clause c with the content s2[0..n) whose cached hits are e0
            auto build = [&](PlanEnt& pe, uint32_t c, const int32_t* s2, uint32_t n, const CacheEnt* e0) {
                int minT = s2[0]; int best = dcount(minT >> 1);
                for (uint32_t j = 1; j < n; ++j) { const int dj = dcount(s2[j] >> 1); if (best > dj) { best = dj; minT = s2[j]; } }
                const int mn = minT, nm = minT ^ 1;
                if (hasStale(mn) || hasStale(nm)) { return; }
                pe.scan = e0->scan; pe.t = (uint32_t)t;
                pe.steps1 = hot[mn].n ? hot[mn].n - 1 : 0; pe.steps2 = hot[nm].n;
                pe.cb = pe.c1 = pe.c2 = (uint32_t)pc.size();
                if (e0->hb != e0->he) {
                    const uint32_t* l1 = occ_pool.data() + occ[mn].off; const uint32_t n1 = hot[mn].n;
                    for (uint32_t j = 0; j < n1; ++j) {
                        if (l1[j] == c) { continue; }
                        const Hit* h = findHit(c_str.hits, e0->hb, e0->he, l1[j]);
                        if (h && h->lit != nm) { pc.push_back(*h); }
                    }
                    pe.c1 = (uint32_t)pc.size();
                    const uint32_t* l2 = occ_pool.data() + occ[nm].off; const uint32_t n2 = hot[nm].n;
                    for (uint32_t j = 0; j < n2; ++j) {
                        const Hit* h = findHit(c_str.hits, e0->hb, e0->he, l2[j]);
                        if (h && h->lit == nm) { pc.push_back(*h); }
                    }
                    pe.c2 = (uint32_t)pc.size();
                }
                pe.min = mn;
            };
            for (size_t i = b; i < e; ++i) {
                PlanEnt& pe = plan[i]; pe.min = -1; pe.alt = -1; pe.lit_n = 0; pe.t = (uint32_t)t;
                const size_t p2 = reals[i];
                const uint32_t c = strq[p2];
                const ClauseInfo ci = C[c];
                pe.ver = ci.ver;
                if ((ci.flag & (F_DEL | F_STR)) != F_STR || ci.size < 2) { continue; }
                if (qtime[c] != (int32_t)(nq - 1 - p2)) { continue; }
                const CacheEnt* e0 = cached(c_str, c);
                if (e0) { build(pe, c, pool.data() + ci.start, ci.size, e0); }
                // a candidate target: also the shortened contents the closure predicted (records keyed by content)
                if (((is_target[c >> 6] >> (c & 63)) & 1) && c < c_str.slot.size()) {
                    for (int32_t k = c_str.slot[c]; k >= 0; k = c_str.ent[(size_t)k].next) {
                        const CacheEnt& ce = c_str.ent[(size_t)k];
                        if (ce.ver != UINT32_MAX || ce.lit_n < 2) { continue; }
                        PlanEnt a; a.min = -1; a.ver = 0; a.alt = pe.alt; a.lit_off = ce.lit_off; a.lit_n = ce.lit_n;
                        build(a, c, c_str.lits.data() + ce.lit_off, ce.lit_n, &ce);
                        if (a.min >= 0) { pe.alt = (int32_t)pa.size(); pa.push_back(a); }
                    }
                }
            }
        }, 2048);
    }
    long cur_plan = -1;                                             // plan record of the entry being processed, if any
    lap("plan");
    std::vector<uint32_t> localQueue;
    size_t end = par_done ? 0 : strq.size();
    int ret = L_UNDEF; bool early = false;
    for (; end > 0 && (unlimited || str_steps < str_lim);) {
        uint32_t cr;
        cur_plan = -1;
        const int64_t tq0 = fine ? now_ns() : 0;
        if (localQueue.empty()) {
            if (static_lazy && hasToPropagate()) { cur_end = end; leaveLazy(); }
            if (static_lazy) {
                // next real entry below `end`: from the plan, or one that an unpredicted hit turned real meanwhile
                while (ri > 0 && reals[ri - 1] >= end) { --ri; }
                while (!lazy_flipped.empty() && lazy_flipped.top() >= end) { lazy_flipped.pop(); }
                long nxt = ri > 0 ? (long)reals[ri - 1] : -1;
                if (!lazy_flipped.empty() && (long)lazy_flipped.top() > nxt) { nxt = (long)lazy_flipped.top(); }
                int64_t run = 0;
                for (size_t p = (size_t)(nxt + 1); p < end; ++p) { run += contrib[p]; }      // all inert: contributions >= 0
                str_steps += run; n_lazy += (int64_t)(end - (size_t)(nxt + 1));
                if (nxt < 0) { end = 0; cur_end = 0; break; }
                end = (size_t)nxt + 1;                               // the shared code below steps onto it
                if (planned && ri > 0 && (long)reals[ri - 1] == nxt) {
                    cur_plan = (long)ri - 1;
                    if (ri > 5) {                                    // look ahead: the clauses the next planned entries will edit
                        const PlanEnt& pf = plan[ri - 5];
                        if (pf.min >= 0) {
                            const Hit* hc = plan_c[pf.t].data();
                            for (uint32_t k = pf.cb; k < pf.c2; ++k) { const uint32_t o = hc[k].clause; __builtin_prefetch(&C[o], 1); if (o < log_head.size()) { __builtin_prefetch(&log_head[o], 1); } if (o < qtime.size()) { __builtin_prefetch(&qtime[o]); } if (o < shrunk_mark.size()) { __builtin_prefetch(&shrunk_mark[o], 1); } }
                        }
                        // an entry with predicted contents compares its literals with them when its turn comes
                        if (pf.alt >= 0) { __builtin_prefetch(pool.data() + C[strq[reals[ri - 5]]].start); __builtin_prefetch(c_str.lits.data() + plan_alt[pf.t][(size_t)pf.alt].lit_off); }
                    }
                    if (ri > 2) {
                        const PlanEnt& pf = plan[ri - 2];
                        if (pf.min >= 0) { const Hit* hc = plan_c[pf.t].data(); for (uint32_t k = pf.cb; k < pf.c2; ++k) {
                            const uint32_t o = hc[k].clause; __builtin_prefetch(pool.data() + C[o].start, 1);
                            if (o < qtime.size() && qtime[o] >= 0) { const size_t po = static_nq - 1 - (size_t)qtime[o]; if (po < contrib.size()) { __builtin_prefetch(&contrib[po], 1); __builtin_prefetch(&walk_min[po]); } } } }
                    }
                }
                if (planned) {                                       // the record of a planned entry needs its clause header only
                    if (ri > 10) { const uint32_t c2 = strq[reals[ri - 10]]; __builtin_prefetch(&C[c2]); }
                } else if (ri > 8) {                                 // look ahead: the clause, its counters and its scan record
                    const uint32_t c2 = strq[reals[ri - 8]]; const ClauseInfo& ci2 = C[c2];
                    __builtin_prefetch(&ci2);
                    if (ri > 4) { const ClauseInfo& c3 = C[strq[reals[ri - 4]]]; const int32_t* l3 = pool.data() + c3.start; const uint32_t n3 = std::min<uint32_t>(c3.size, 6); for (uint32_t k = 0; k < n3; ++k) { __builtin_prefetch(&hot[l3[k] & ~1]); } }
                    if (c2 < c_str.slot.size()) { __builtin_prefetch(&c_str.slot[c2]); }
                }
            }
            --end; cur_end = end;
            if (!static_lazy && end >= 24 && contrib[end - 24] < 0) {           // look ahead past the inert entries: the next real one
                const ClauseInfo& c2 = C[strq[end - 24]]; const int32_t* l2 = pool.data() + c2.start;
                const uint32_t n2 = std::min<uint32_t>(c2.size, 6);
                for (uint32_t k = 0; k < n2; ++k) { __builtin_prefetch(&hot[l2[k] & ~1]); }
                if (strq[end - 24] < c_str.slot.size()) { const int32_t sl = c_str.slot[strq[end - 24]]; if (sl >= 0) { __builtin_prefetch(&c_str.ent[(size_t)sl]); } }
            }
            if (static_lazy) {
                if (contrib[end] >= 0) {
                    if (!hasToPropagate()) { str_steps += contrib[end]; ++n_lazy; continue; }
                    leaveLazy();
                }
            }
            cr = strq[end];
            if (static_valid && contrib[end] >= 0 && !hasToPropagate() && (unlimited || str_steps + contrib[end] < str_lim)) {
                str_steps += contrib[end];
                C[cr].flag &= ~F_STR;
                ++n_fast; ++n_static;
                if (pre_active) { const int mn = walk_min[end]; scanned_dyn[mn] = 1; scanned_dyn[mn ^ 1] = 1; }
                continue;
            }
            if (static_valid && contrib_t[end] >= 0 && !hasToPropagate() && C[cr].ver == plan_ver[end] &&
                (C[cr].flag & (F_STR | F_DEL)) == F_STR && opt.strength && (unlimited || str_steps + contrib_t[end] + 1 < str_lim)) {
                // unchanged candidate target without own hits: what the walk below would do is add up two clean lists
                str_steps += contrib_t[end];
                C[cr].flag &= ~F_STR;
                ++n_fast; ++n_tfast;
                if (pre_active) { scanned_dyn[tmin[end]] = 1; scanned_dyn[tmin[end] ^ 1] = 1; }
                continue;
            }
            if (!static_lazy && end >= 16) {
                const uint32_t c2 = strq[end - 16];
                const int32_t* l2 = CL(c2); const uint32_t n2 = std::min<uint32_t>(C[c2].size, 6);
                for (uint32_t k = 0; k < n2; ++k) { __builtin_prefetch(&hot[l2[k] & ~1]); }
            }
        }
        else {
            cr = localQueue.back(); localQueue.pop_back();
            // a clause that was strengthened after its own turn: the records of its predicted contents hang off its queue entry
            if (planned && static_lazy && cr < qtime.size() && qtime[cr] >= 0) {
                const uint32_t pos = (uint32_t)(static_nq - 1 - (size_t)qtime[cr]);
                const auto it = std::lower_bound(reals.begin(), reals.end(), pos);
                if (it != reals.end() && *it == pos) { cur_plan = (long)(it - reals.begin()); }
            }
        }
        if ((C[cr].flag & F_DEL) || !(C[cr].flag & F_STR)) { continue; }
        if (!opt.strength) { C[cr].flag &= ~F_STR; continue; }
        if (C[cr].size < 2) {
            if (C[cr].size == 1) { if (dataEnqueue(CL(cr)[0]) == L_FALSE) { break; } else { continue; } }
            else { ok = false; break; }
        }
        // stage: how much of the two list walks the plan record has already done (0 none, 1 occ(min), 2 both)
        int stage = 0; int minT = -1;
        const int64_t tp0 = fine ? now_ns() : 0;
        if (fine) { dbg_step_ns += tp0 - tq0; }
        struct PlanT { int64_t t0; int64_t* acc; bool on; ~PlanT() { if (on) { *acc += now_ns() - t0; } } } plant{tp0, &dbg_plan_ns, fine && cur_plan >= 0};
        if (cur_plan >= 0) {
            const PlanEnt* pp = &plan[(size_t)cur_plan];
            if (C[cr].ver != pp->ver || pp->min < 0) {             // not the planned content: one of the predicted versions?
                const std::vector<PlanEnt>& pa = plan_alt[pp->t]; const PlanEnt* f = nullptr;
                for (int32_t a = pp->alt; a >= 0; a = pa[(size_t)a].alt) {
                    const PlanEnt& x = pa[(size_t)a];
                    if (x.lit_n == C[cr].size && !memcmp(c_str.lits.data() + x.lit_off, CL(cr), sizeof(int32_t) * x.lit_n)) { f = &x; break; }
                }
                pp = f;
                if (f) { ++n_spec_used; }
            }
            static const PlanEnt none{0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, -1};
            const PlanEnt& pe = pp ? *pp : none;
            if (pe.min >= 0 && static_valid && static_lazy && !hasToPropagate() && !hasStale(pe.min) && !hasStale(pe.min ^ 1) &&
                (unlimited || str_steps + (int64_t)pe.steps1 + (int64_t)pe.steps2 + 1 < str_lim)) {
                const Hit* hc = plan_c[pe.t].data();
                minT = pe.min;
                if (pe.cb == pe.c2) { ++n_fast; } else { ++n_slow; }
                if (pre_active) { scanned_dyn[pe.min] = 1; }
                str_steps += pe.steps1;
                for (uint32_t k = pe.cb; k < pe.c1; ++k) {
                    const uint32_t other = hc[k].clause;
                    if (isDel(other) || !strHitValid(cr, other, hc[k].lit, pe.scan)) { continue; }
                    const int32_t* t = CL(other); int pos = -1;
                    for (uint32_t k2 = 0; k2 < C[other].size; ++k2) { if (t[k2] == hc[k].lit) { pos = (int)k2; break; } }
                    if (pos >= 0) { strengthenHit(localQueue, other, pos); }
                }
                stage = 1;
                if (!hasToPropagate()) {
                    if (pre_active) { scanned_dyn[pe.min ^ 1] = 1; }
                    str_steps += pe.steps2;
                    for (uint32_t k = pe.c1; k < pe.c2; ++k) {
                        const uint32_t other = hc[k].clause;
                        if (isDel(other) || !strHitValid(cr, other, hc[k].lit, pe.scan)) { continue; }
                        const int32_t* t = CL(other); int pos = -1;
                        for (uint32_t k2 = 0; k2 < C[other].size; ++k2) { if (t[k2] == hc[k].lit) { pos = (int)k2; break; } }
                        if (pos >= 0) { strengthenHit(localQueue, other, pos); }
                    }
                    stage = 2;
                    if (!hasToPropagate()) { C[cr].flag &= ~F_STR; continue; }
                }
            }
        }
        if (stage == 0) { const int32_t* s = CL(cr); minT = s[0];
          for (uint32_t j = 1; j < C[cr].size; ++j) { if (dcount(minT >> 1) > dcount(s[j] >> 1)) { minT = s[j]; } } }
        const int min = minT, nmin = lneg(minT);
        const int64_t th0 = fine ? now_ns() : 0;
        const CacheEnt* e = strHits(cr);
        if (fine) { dbg_hits_ns += now_ns() - th0; }
        if (stage == 0 && e->hb == e->he && !hasStale(min) && !hasStale(nmin) && !hasToPropagate() &&
            (unlimited || str_steps + (int64_t)hot[min].n + (int64_t)hot[nmin].n < str_lim)) {
            // no candidate can hit and no list needs lazy cleaning: only the step counters move
            str_steps += (hot[min].n ? (int64_t)hot[min].n - 1 : 0) + (int64_t)hot[nmin].n;
            C[cr].flag &= ~F_STR;
            ++n_fast;
            if (pre_active) { scanned_dyn[min] = 1; scanned_dyn[nmin] = 1; }    // walked (found nothing to do)
            continue;
        }
        if (stage == 0) { ++n_slow; }
        const int64_t ts0 = fine ? now_ns() : 0;
        struct SlowT { int64_t t0; int64_t* acc; bool on; ~SlowT() { if (on) { *acc += now_ns() - t0; } } } slowt{ts0, &dbg_slow_ns, fine};
        if (pre_active) { scanned_dyn[min] = 1; }
        // pass 1: occ(min) - the complementary literal is any literal but min
        if (stage < 1) {
            ListRef list = L(min);
            for (uint32_t j = 0; j < list.n && (unlimited || str_steps < str_lim); ++j) {
                const uint32_t other = LP(list)[j];
                if (other == cr) { continue; }
                str_steps++;
                // (Subsumption.cc:1331 breaks on an empty clause; one can only exist after ok became false, and
                //  every path that creates it returns before this loop is reached again)
                if (isDel(other)) { occSwapRemove(min, j); --j; continue; }
                if (e->hb == e->he) { continue; }
                const Hit* h = findHit(c_str.hits, e->hb, e->he, other);
                if (h && h->lit != nmin && strHitValid(cr, other, h->lit, e->scan)) {
                    const int32_t* t = CL(other); int pos = -1;
                    for (uint32_t k = 0; k < C[other].size; ++k) { if (t[k] == h->lit) { pos = (int)k; break; } }
                    if (pos >= 0) { strengthenHit(localQueue, other, pos); }
                }
            }
        }
        if (stage < 2 && hasToPropagate()) {
            leaveLazy();
            static_valid = false;            // counts and lists move now: the precomputed entries are void
            rollbackPrecompact();
            if (propagationProcess(true, useHeap, ignore) == L_FALSE) { ret = L_FALSE; early = true; break; }
            e = strHits(cr);                 // cr itself may have been shortened by the propagation
        }
        if (pre_active) { scanned_dyn[nmin] = 1; }
        // pass 2: occ(~min) - the complementary literal must be ~min
        if (stage < 2) {
            ListRef list_neg = L(nmin);
            for (uint32_t j = 0; j < list_neg.n && (unlimited || str_steps < str_lim); ++j) {
                const uint32_t other = LP(list_neg)[j];
                str_steps++;
                if (isDel(other)) { occSwapRemove(nmin, j); --j; continue; }
                if (e->hb == e->he) { continue; }
                const Hit* h = findHit(c_str.hits, e->hb, e->he, other);
                if (h && h->lit == nmin && strHitValid(cr, other, h->lit, e->scan)) {
                    const int32_t* t = CL(other); int pos = -1;
                    for (uint32_t k = 0; k < C[other].size; ++k) { if (t[k] == h->lit) { pos = (int)k; break; } }
                    if (pos >= 0) { strengthenHit(localQueue, other, pos); }
                }
            }
        }
        if (hasToPropagate()) {
            leaveLazy();
            static_valid = false;
            rollbackPrecompact();
            if (propagationProcess(true, useHeap, ignore) == L_FALSE) { ret = L_FALSE; early = true; break; }
            t_sub.modified = t_sub.modified || t_up.modified;
        }
        C[cr].flag &= ~F_STR;
    }
    lap("ordered");
    if (fine) { fprintf(stderr, "  str slow entries %.1f ms, strHits %.1f ms, stepping %.1f ms, planned entries (whole) %.1f ms\n", dbg_slow_ns / 1e6, dbg_hits_ns / 1e6, dbg_step_ns / 1e6, dbg_plan_ns / 1e6); dbg_slow_ns = 0; dbg_hits_ns = 0; dbg_step_ns = 0; dbg_plan_ns = 0; }
    if (!early) {
        updateOccurrences(useHeap, ignore);
        ret = ok ? L_UNDEF : L_FALSE;
    }
    lap("updateOcc");
    parallelFor(strq.size(), [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { const uint32_t c = strq[i]; if (c < qtime.size()) { qtime[c] = -1; } } });
    if (lap.on) { fprintf(stderr, "  str lazy entries %lld, lazy still on %d\n", (long long)n_lazy, (int)static_lazy); }
    static_pass = false; static_lazy = false; n_fast += n_lazy; n_static += n_lazy; n_lazy = 0;
    if (tune.debug_pre) { fprintf(stderr, "strpass q=%zu steps=%lld removed=%d pre=%zu active=%d static=%lld\n", strq.size(), (long long)str_steps, removedLiterals, pre_lists.size(), (int)pre_active, (long long)n_static); }
    pre_active = false;
    host_ns_commit += now_ns() - t0;
    return ret;
}

// fullStrengthening with -naive_strength, Subsumption.cc:1056-1197: the queue runs front to back, every literal of the
// strengthener is flipped in turn and the walk goes over occ(min) - or occ(~min) when the flipped literal is min itself.  The
// candidates a walk can hit are the K4 hits of the strengthener whose literal is the flipped one (K4 finds every clause that
// contains the strengthener with exactly one literal complemented, whichever list pair it scanned).
int Engine::naiveStrengthening(bool useHeap, int ignore)
{
    PhaseTimer pt(ph_str);
    const bool unlimited = !opt.limited;
    ensureHostLists();
    logClear(); dirty_str.clear();
    if (str_prepared) { str_prepared = false; }
    else if (opt.strength) {
        std::vector<uint32_t> ids;
        for (uint32_t c : strq) { if (!(C[c].flag & F_DEL) && (C[c].flag & F_STR) && C[c].size >= 2) { ids.push_back(c); } }
        sortUniqueIds(ids);
        c_str.clear();
        scanClauses(CP3G_STRENGTHEN, ids, c_str);
    } else { c_str.clear(); }
    const int64_t t0 = now_ns();
    std::vector<uint32_t> localQueue;
    int ret = L_UNDEF; bool early = false;
    for (size_t i = 0; i < strq.size();) {
        uint32_t cr;
        if (localQueue.empty()) { cr = strq[i]; ++i; }
        else { cr = localQueue.back(); localQueue.pop_back(); }
        if ((C[cr].flag & F_DEL) || !(C[cr].flag & F_STR)) { continue; }
        if (!opt.strength) { C[cr].flag &= ~F_STR; continue; }
        const uint32_t cn = C[cr].size;
        int min = CL(cr)[0];
        for (uint32_t j = 1; j < cn; ++j) {
            const int x = CL(cr)[j];
            if ((int64_t)hot[min].cnt * (cn - 1) + hot[min ^ 1].cnt > (int64_t)hot[x].cnt * (cn - 1) + hot[x ^ 1].cnt) { min = x; }
        }
        const CacheEnt* e = cn >= 2 ? strHits(cr) : nullptr;
        const uint32_t hb = e ? e->hb : 0, he = e ? e->he : 0, escan = e ? e->scan : 0;
        for (uint32_t j = 0; j < cn && (unlimited || str_steps < str_lim); ++j) {
            const int neg_lit = CL(cr)[j] ^ 1;
            const int lx = (neg_lit == (min ^ 1)) ? neg_lit : min;
            for (uint32_t k = 0; k < hot[lx].n && (unlimited || str_steps < str_lim); ++k) {
                const uint32_t other = occ_pool[occ[lx].off + k];
                if (other == cr) { continue; }
                str_steps++;
                if (isDel(other)) { occSwapRemove(lx, k); --k; continue; }
                if (hb == he) { continue; }
                const Hit* h = findHit(c_str.hits, hb, he, other);
                if (!h || h->lit != neg_lit || !strHitValid(cr, other, h->lit, escan)) { continue; }
                const int32_t* t = CL(other); int pos = -1;
                for (uint32_t k2 = 0; k2 < C[other].size; ++k2) { if (t[k2] == neg_lit) { pos = (int)k2; break; } }
                if (pos < 0) { continue; }
                ++removedLiterals;
                removePosSorted(other, pos);
                noteShrunk(other, neg_lit, true);
                dirty_str.push_back(other);
                t_sub.modified = true;                               // modifiedFormula = true: the penalty peak stays as it is
                if (C[other].size == 1) {
                    setDeleted(other);
                    dataEnqueue(CL(other)[0]);
                } else {
                    if (!(C[other].flag & F_SUB)) { C[other].flag |= F_SUB; subq.push_back(other); }
                    if (!(C[other].flag & F_STR)) { localQueue.push_back(other); C[other].flag |= F_STR; }
                    occ_upd.push_back({other, neg_lit});
                }
            }
        }
        if (hasToPropagate()) {
            if (propagationProcess(true, false, VAR_UNDEF) == L_FALSE) { ok = false; ret = L_FALSE; early = true; break; }
        }
        C[cr].flag &= ~F_STR;
    }
    if (!early) { updateOccurrences(useHeap, ignore); }
    host_ns_commit += now_ns() - t0;
    return ret;
}

// Subsumption::process, Subsumption.cc:85-180
bool Engine::subsumptionProcess(bool doStrengthen, bool useHeap, int ignore)
{
    t_sub.modified = false; str_prepared = false;
    if (!ok) { return false; }
    if (!performSimplification(t_sub)) {
        clearQueue(subq, F_SUB); clearQueue(strq, F_STR);
        return false;
    }
    // (the technique-level size gates - Subsumption.cc:102, BVE.cc:180-184: cp3_susi_* / cp3_bve_* - can never fire: they need
    //  nTotLits() > limit, and cleanSolver() (Coprocessor.cc:1395-1406) zeroes clauses_literals before any technique runs;
    //  verified against the compiled reference with all limits set to 10.  The options are accepted and have no effect.)
    if (!(sub_steps + opt.call_inc < sub_lim)) { sub_lim += opt.call_inc; limitIncreases++; }
    if (!(str_steps + opt.call_inc < str_lim)) { str_lim += opt.call_inc; limitIncreases++; }
    DbgLap lap("susi", tune.debug_par);
    while (ok && (!subq.empty() || !strq.empty())) {
        if (!subq.empty()) {
            overlap_str = doStrengthen && !strq.empty();
            subsumptionWorker(useHeap, ignore);
            overlap_str = false;
            lap("subsumption worker");
            clearQueue(subq, F_SUB);
            lap("clear queue");
        }
        if (!strq.empty()) {
            if (!doStrengthen) { clearQueue(strq, F_STR); continue; }
            if (opt.naive_strength) { naiveStrengthening(useHeap, ignore); } else { strengtheningWorker(useHeap, ignore); }
            lap("strengthening worker");
            clearQueue(strq, F_STR);
            lap("clear queue");
        }
        str_prepared = false;
    }
    t_sub.modified = t_sub.modified || t_up.modified;   // propagation.appliedSomething(): its last call
    if (!t_sub.modified) { unsuccessful(t_sub); }
    return t_sub.modified;
}

} // namespace cp3

namespace cp3 {

// ------------------------------------------------------------------------------------------------ component commit
// The real entries of a strengthening pass, committed on all cores.
//
// While no unit shows up, a strengthening pass changes nothing but clause contents and flags: the occurrence lists and counters
// are frozen (updates are deferred, Subsumption.cc:1400,1523), lazily cleaned lists end in the same state whoever walks them
// first, and the step counter is a sum.  Two queue entries can therefore only influence each other through a clause both of them
// touch - as strengthener or as target of a cached hit.  The transitive closure of "touches the same clause" (union-find over all
// cached hits of all content versions) cuts the real entries into dependency components; inside a component the reference's
// order is replayed exactly - queue position descending, re-queued clauses right behind the entry that hit them (the LIFO
// localQueue) - on PRIVATE copies of the clauses it edits, components side by side on the cores.  Nothing of the engine's state
// is written before every component has finished: a component that meets a content version without a scan record reports it,
// the missing versions are scanned in one batch and the simulation is repeated; a unit clause, a hit on a statically inert
// entry or the step limit in reach hands the pass - still untouched - to the ordered walk.  The commit then applies the final
// clause contents, and the queue pushes / occurrence updates in the order of their (queue time, sequence) keys.
bool Engine::strengthenComponents(const std::vector<uint32_t>& reals, const DevStatic* ds)
{
    const size_t nq = strq.size(), ncl = C.size(), nlit = (size_t)2 * nvars;
    const bool unlimited = !opt.limited;
    const bool dbg = tune.debug_par; int64_t tt = now_ns();
    auto lap = [&](const char* w) { if (dbg) { const int64_t n2 = now_ns(); fprintf(stderr, "  compar %s %.1f ms\n", w, (n2 - tt) / 1e6); tt = n2; } };
    const int T = std::max(1, nthreads);
    struct Ent { uint32_t root, pos; };
    struct Ev { uint64_t key; uint32_t clause; int32_t lit; uint32_t push_sub; };
    struct Final { uint32_t clause, nrem; bool sub, str; std::vector<int32_t> lits; };
    struct Missing { uint32_t clause; std::vector<int32_t> lits; };
    // result of one simulated component: slices of its producer's event / final-state / cleaned-list arrays
    struct Comp { uint32_t root; uint32_t ev0, ev1, fin0, fin1, st0, st1; int64_t steps; uint32_t nfast, nslow; bool valid; };
    struct Out { std::vector<Ev> evs; std::vector<Final> fin; std::vector<int32_t> stale_lists; std::vector<Comp> comps; std::vector<Missing> miss; };
    std::vector<Out> outs;                                       // one per (round, bucket); kept until the commit
    // ---- components: a strengthener and every target of every cached content version of it (lock-free union-find, smaller root wins)
    RawVec<uint32_t> uf; uf.resize(ncl);
    parallelFor(ncl, [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { uf[i] = (uint32_t)i; } });
    auto find = [&](uint32_t x) -> uint32_t {
        for (;;) {
            const uint32_t p = __atomic_load_n(&uf[x], __ATOMIC_RELAXED);
            if (p == x) { return x; }
            const uint32_t g = __atomic_load_n(&uf[p], __ATOMIC_RELAXED);
            if (g != p) { __atomic_store_n(&uf[x], g, __ATOMIC_RELAXED); }           // path halving
            x = p;
        }
    };
    auto unite = [&](uint32_t a, uint32_t b2) {
        for (;;) {
            a = find(a); b2 = find(b2);
            if (a == b2) { return; }
            if (a < b2) { std::swap(a, b2); }                                         // a > b2: a's tree goes under b2
            uint32_t expect = a;
            if (__atomic_compare_exchange_n(&uf[a], &expect, b2, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { return; }
        }
    };
    auto uniteEntries = [&](size_t first_ent) {
        const size_t ne = c_str.ent.size();
        // owner clause of a cache entry: walk the slot chains of the clauses that own one
        parallelFor(c_str.used.size(), [&](size_t b, size_t e, int) {
            for (size_t u = b; u < e; ++u) {
                const uint32_t sc = c_str.used[u];
                if (sc >= c_str.slot.size()) { continue; }
                for (int32_t k = c_str.slot[sc]; k >= 0; k = c_str.ent[(size_t)k].next) {
                    if ((size_t)k < first_ent) { continue; }
                    const CacheEnt& ce = c_str.ent[(size_t)k];
                    for (uint32_t h = ce.hb; h < ce.he; ++h) { unite(sc, c_str.hits[h].clause); }
                }
            }
        }, 2048);
        (void)ne;
    };
    uniteEntries(0);
    lap("union-find");
    // ---- the entries that will be processed
    std::vector<std::vector<uint32_t>> tpos((size_t)T + 1); std::atomic<bool> unfit{false};
    parallelFor(reals.size(), [&](size_t b, size_t e, int t) {
        std::vector<uint32_t>& out = tpos[(size_t)t];
        for (size_t i = b; i < e; ++i) {
            const uint32_t p = reals[i]; const uint32_t c = strq[p]; const ClauseInfo ci = C[c];
            if ((ci.flag & F_DEL) || !(ci.flag & F_STR)) { continue; }
            if (ci.size < 2 || qtime[c] != (int32_t)(nq - 1 - p)) { unfit = true; continue; }       // a unit in the queue, a clause queued twice
            out.push_back(p);
        }
    }, 4096);
    if (unfit) { return false; }
    std::unordered_set<uint32_t> dirty;                          // roots to simulate again (empty in the first round: everything)
    for (int round = 0; round < 6; ++round) {
        ++n_compar_rounds;
        // bucket the entries by root range so that a component lands in exactly one bucket; sort by (root, position descending)
        std::vector<std::vector<Ent>> bucket((size_t)T);
        {
            // root ranges with about the same number of entries each (roots are the smallest clause index of their component:
            // low indices are over-represented): 4096 coarse bins, cut where the running count passes k/T of the total
            std::vector<std::vector<Ent>> parts(((size_t)T + 1) * (size_t)T);
            std::vector<std::vector<Ent>> rooted((size_t)T + 1);
            const uint64_t NBIN = 4096, bw = ((uint64_t)ncl + NBIN - 1) / NBIN;
            std::vector<std::vector<uint32_t>> hist((size_t)T + 1, std::vector<uint32_t>(NBIN, 0u));
            parallelFor((size_t)T + 1, [&](size_t b, size_t e, int) {
                for (size_t t = b; t < e; ++t) {
                    for (const uint32_t p : tpos[t]) {
                        const uint32_t r = find(strq[p]);
                        if (round > 0 && !dirty.count(r)) { continue; }
                        rooted[t].push_back(Ent{r, p}); ++hist[t][(size_t)(r / bw)];
                    }
                }
            }, 1);
            std::vector<uint32_t> bin_bucket(NBIN, 0u);
            { uint64_t tot = 0; for (size_t t = 0; t <= (size_t)T; ++t) { for (uint64_t k = 0; k < NBIN; ++k) { tot += hist[t][k]; } }
              uint64_t run = 0;
              for (uint64_t k = 0; k < NBIN; ++k) {
                  bin_bucket[k] = (uint32_t)std::min<uint64_t>((uint64_t)T - 1, tot ? run * (uint64_t)T / tot : 0);
                  for (size_t t = 0; t <= (size_t)T; ++t) { run += hist[t][k]; }
              } }
            parallelFor((size_t)T + 1, [&](size_t b, size_t e, int) {
                for (size_t t = b; t < e; ++t) { for (const Ent& x : rooted[t]) { parts[t * (size_t)T + bin_bucket[(size_t)(x.root / bw)]].push_back(x); } }
            }, 1);
            parallelFor((size_t)T, [&](size_t b, size_t e, int) {
                for (size_t bk = b; bk < e; ++bk) {
                    std::vector<Ent>& v = bucket[bk];
                    for (size_t t = 0; t <= (size_t)T; ++t) { const std::vector<Ent>& src = parts[t * (size_t)T + bk]; v.insert(v.end(), src.begin(), src.end()); }
                    std::sort(v.begin(), v.end(), [](const Ent& a, const Ent& b2) { return a.root != b2.root ? a.root < b2.root : a.pos > b2.pos; });
                }
            }, 1);
        }
        lap("entries");
        // ---- simulate the components of this round
        const size_t out0 = outs.size(); outs.resize(out0 + (size_t)T);
        std::atomic<bool> abort_units{false};
        parallelFor((size_t)T, [&](size_t bb, size_t be, int) {
            for (size_t bk = bb; bk < be; ++bk) {
                Out& out = outs[out0 + bk];
                const std::vector<Ent>& ents = bucket[bk];
                struct Rec { uint32_t c; uint8_t flag; bool changed; std::vector<int32_t> lits, removed; };
                std::vector<Rec> recs; std::unordered_map<uint32_t, uint32_t> index;
                std::vector<uint32_t> lq, clean1, clean2;
                auto recOf = [&](uint32_t c) -> uint32_t {
                    if (recs.size() <= 8) { for (uint32_t i = 0; i < recs.size(); ++i) { if (recs[i].c == c) { return i; } } }
                    else { auto it = index.find(c); if (it != index.end()) { return it->second; } }
                    recs.push_back(Rec{c, (uint8_t)C[c].flag, false, {}, {}});
                    if (recs.size() == 9) { for (uint32_t i = 0; i < 8; ++i) { index[recs[i].c] = i; } }
                    if (recs.size() > 8) { index[c] = (uint32_t)recs.size() - 1; }
                    return (uint32_t)recs.size() - 1;
                };
                // occurrence list of x as a walk sees it: the list itself, or - deleted entries still listed - its lazily cleaned form
                auto view = [&](int x, std::vector<uint32_t>& buf, const uint32_t*& p, uint32_t& n) {
                    p = occ_pool.data() + occ[x].off; n = hot[x].n;
                    if (!hasStale(x)) { return; }
                    buf.assign(p, p + n);
                    for (uint32_t j = 0; j < n;) { if (isDel(buf[j])) { buf[j] = buf[n - 1]; --n; } else { ++j; } }
                    p = buf.data(); out.stale_lists.push_back(x);
                };
                size_t i = 0;
                while (i < ents.size() && !abort_units) {
                    size_t j = i; while (j < ents.size() && ents[j].root == ents[i].root) { ++j; }
                    // one component: entries [i, j)
                    recs.clear(); index.clear(); lq.clear();
                    Comp cp{ents[i].root, (uint32_t)out.evs.size(), 0, (uint32_t)out.fin.size(), 0, (uint32_t)out.stale_lists.size(), 0, 0, 0, 0, true};
                    bool deferred = false;
                    uint64_t key = 0;
                    auto process = [&](uint32_t cr) {
                        const uint32_t ri = recOf(cr);
                        if ((recs[ri].flag & F_DEL) || !(recs[ri].flag & F_STR)) { return; }
                        const int32_t* s = recs[ri].changed ? recs[ri].lits.data() : CL(cr);
                        const uint32_t sn = recs[ri].changed ? (uint32_t)recs[ri].lits.size() : (uint32_t)C[cr].size;
                        if (sn < 2) { abort_units = true; return; }
                        int minT = s[0];
                        for (uint32_t k = 1; k < sn; ++k) { if (dcount(minT >> 1) > dcount(s[k] >> 1)) { minT = s[k]; } }
                        const int mn = minT, nm = minT ^ 1;
                        // scan record of this very content
                        const CacheEnt* e = nullptr;
                        if (!recs[ri].changed) { e = cached(c_str, cr); }
                        else if (cr < c_str.slot.size()) {
                            for (int32_t k = c_str.slot[cr]; k >= 0; k = c_str.ent[(size_t)k].next) {
                                const CacheEnt& ce = c_str.ent[(size_t)k];
                                if (ce.ver == UINT32_MAX && ce.lit_n == sn && !memcmp(c_str.lits.data() + ce.lit_off, s, sizeof(int32_t) * sn)) { e = &ce; break; }
                            }
                        }
                        if (!e) { out.miss.push_back(Missing{cr, std::vector<int32_t>(s, s + sn)}); deferred = true; return; }
                        const uint32_t* l1; uint32_t n1; const uint32_t* l2; uint32_t n2;
                        view(mn, clean1, l1, n1); view(nm, clean2, l2, n2);
                        cp.steps += (n1 ? (int64_t)n1 - 1 : 0) + (int64_t)n2;
                        if (e->hb == e->he) { ++cp.nfast; }
                        else {
                            ++cp.nslow;
                            for (int pass = 0; pass < 2 && !abort_units; ++pass) {
                                const uint32_t* l = pass == 0 ? l1 : l2; const uint32_t n = pass == 0 ? n1 : n2;
                                for (uint32_t q = 0; q < n; ++q) {
                                    const uint32_t other = l[q];
                                    if (pass == 0 && other == cr) { continue; }
                                    const Hit* h = findHit(c_str.hits, e->hb, e->he, other);
                                    if (!h || (pass == 0 ? h->lit == nm : h->lit != nm)) { continue; }
                                    const uint32_t di = recOf(other);
                                    // (recOf may have moved the records: the strengthener's literals are re-read from its record)
                                    const int32_t* s2 = recs[ri].changed ? recs[ri].lits.data() : CL(cr);
                                    bool valid = true;                       // the target lost the literal itself or a literal of the strengthener
                                    for (int32_t x : recs[di].removed) { if (x == h->lit || containsLit(s2, sn, x)) { valid = false; break; } }
                                    if (!valid) { continue; }
                                    Rec& d = recs[di];
                                    if (!d.changed) { d.lits.assign(CL(other), CL(other) + C[other].size); d.changed = true; }
                                    int pos = -1;
                                    for (size_t k2 = 0; k2 < d.lits.size(); ++k2) { if (d.lits[k2] == h->lit) { pos = (int)k2; break; } }
                                    if (pos < 0) { continue; }
                                    d.lits.erase(d.lits.begin() + pos); d.removed.push_back(h->lit);
                                    out.evs.push_back(Ev{key++, other, h->lit, (d.flag & F_SUB) ? 0u : 1u});
                                    d.flag |= F_SUB;
                                    if (!(d.flag & F_STR)) { lq.push_back(other); d.flag |= F_STR; }
                                    if (d.lits.size() == 1) { abort_units = true; break; }
                                }
                            }
                        }
                        recs[ri].flag &= (uint8_t)~F_STR;
                    };
                    for (size_t k = i; k < j && !deferred && !abort_units; ++k) {
                        key = ((uint64_t)(nq - 1 - ents[k].pos) << 24);
                        process(strq[ents[k].pos]);
                        // QUIRK: the worker's loop ends with the entry at queue position 0 (`end > start`, Subsumption.cc:1276): clauses
                        // that this very last entry re-queues are never processed and keep their flag
                        if (ents[k].pos == 0) { lq.clear(); }
                        while (!lq.empty() && !deferred && !abort_units) { const uint32_t c2 = lq.back(); lq.pop_back(); process(c2); }
                    }
                    if (deferred) { out.evs.resize(cp.ev0); out.stale_lists.resize(cp.st0); }
                    else if (!abort_units) {
                        for (Rec& r : recs) { if (r.changed) { out.fin.push_back(Final{r.c, (uint32_t)r.removed.size(), (r.flag & F_SUB) != 0, (r.flag & F_STR) != 0, std::move(r.lits)}); } }
                        cp.ev1 = (uint32_t)out.evs.size(); cp.fin1 = (uint32_t)out.fin.size(); cp.st1 = (uint32_t)out.stale_lists.size();
                        out.comps.push_back(cp);
                    }
                    i = j;
                }
            }
        }, 1);
        lap("simulate");
        if (abort_units) { return false; }
        size_t nmiss = 0; for (size_t o = out0; o < outs.size(); ++o) { nmiss += outs[o].miss.size(); }
        if (nmiss == 0) { break; }
        if (round == 5) { return false; }
        // ---- scan the content versions nobody predicted (what the ordered walk would fetch one batch at a time), merge the components
        //      their hits connect, and simulate those - and only those - again
        std::vector<SpecReq> reqs; spec_lits.clear();
        std::unordered_set<uint64_t> seen; dirty.clear();
        for (size_t o = out0; o < outs.size(); ++o) {
            for (const Missing& m : outs[o].miss) {
                dirty.insert(find(m.clause));
                uint64_t hsh = 1469598103934665603ull ^ m.clause; for (int32_t x : m.lits) { hsh = (hsh ^ (uint32_t)x) * 1099511628211ull; }
                if (!seen.insert(hsh).second) { continue; }
                reqs.push_back({m.clause, (uint32_t)spec_lits.size(), (uint32_t)m.lits.size(), 0.0});
                spec_lits.insert(spec_lits.end(), m.lits.begin(), m.lits.end());
            }
        }
        const size_t h0 = c_str.hits.size(), e0 = c_str.ent.size();
        ++n_ondemand;
        scanSpeculative(CP3G_STRENGTHEN, reqs, c_str);
        // a new hit on a statically inert entry would make it a real one: leave that to the ordered walk
        for (size_t h = h0; h < c_str.hits.size(); ++h) {
            const uint32_t x = c_str.hits[h].clause;
            if (ds) { if (x < ncl && !((ds->real_bits[x >> 5] >> (x & 31)) & 1u) && !(C[x].flag & F_DEL)) { return false; } }
            else if (x < qtime.size() && qtime[x] >= 0) { const size_t pos = nq - 1 - (size_t)qtime[x]; if (pos < contrib.size() && contrib[pos] >= 0) { return false; } }
        }
        // components that the new hits connect: both sides are simulated again under their common root
        for (size_t k = e0; k < c_str.ent.size(); ++k) {
            const CacheEnt& ce = c_str.ent[k];
            const uint32_t sc = reqs[k - e0].clause;                                  // scanSpeculative appends one record per request, in order
            for (uint32_t h = ce.hb; h < ce.he; ++h) { dirty.insert(find(sc)); dirty.insert(find(c_str.hits[h].clause)); unite(sc, c_str.hits[h].clause); }
        }
        { std::unordered_set<uint32_t> nd; for (uint32_t r : dirty) { nd.insert(find(r)); } dirty.swap(nd); }
        for (Out& o : outs) { for (Comp& cp : o.comps) { if (cp.valid && dirty.count(find(cp.root))) { cp.valid = false; } } }
        lap("missing versions");
    }
    // ---- commit: nothing of the engine's state was written so far
    int64_t steps = 0, nfast = 0, nslow = 0; size_t nev = 0, nfin = 0;
    std::vector<uint8_t> walked_stale(nlit, 0);
    for (const Out& o : outs) {
        for (const Comp& cp : o.comps) {
            if (!cp.valid) { continue; }
            steps += cp.steps; nfast += cp.nfast; nslow += cp.nslow; nev += cp.ev1 - cp.ev0; nfin += cp.fin1 - cp.fin0;
            for (uint32_t k = cp.st0; k < cp.st1; ++k) { walked_stale[(size_t)o.stale_lists[k]] = 1; }
        }
    }
    std::vector<int64_t> part((size_t)T + 1, 0), ninert((size_t)T + 1, 0), pstale((size_t)T + 1, 0);
    if (ds) {
        // the inert entries were evaluated by K10; the lists they walk are cleaned (and paid for) together with the simulated walks'
        part[0] = (int64_t)ds->inert_steps; ninert[0] = (int64_t)ds->inert_n;
        parallelFor(nlit, [&](size_t b, size_t e, int) { for (size_t x = b; x < e; ++x) { if (((ds->walked_bits[x >> 5] >> (x & 31)) & 1u) && hasStale((int)x)) { walked_stale[x] = 1; } } });
    } else {
        parallelFor(nq, [&](size_t b, size_t e, int t) { int64_t a = 0, c2 = 0; for (size_t p = b; p < e; ++p) { if (contrib[p] >= 0) { a += contrib[p]; ++c2; } } part[(size_t)t] = a; ninert[(size_t)t] = c2; });
    }
    // the first walk of a list pays one step per deleted entry it removes: counted ahead of the commit when the step limit
    // can stop the pass, otherwise while the lists are cleaned below
    if (!unlimited) {
        parallelFor(nlit, [&](size_t b, size_t e, int t) {
            int64_t a = 0;
            for (size_t x = b; x < e; ++x) { if (walked_stale[x]) { const uint32_t* p = occ_pool.data() + occ[x].off; const uint32_t n = hot[x].n; for (uint32_t j = 0; j < n; ++j) { a += isDel(p[j]) ? 1 : 0; } } }
            pstale[(size_t)t] = a;
        });
    }
    int64_t inert_steps = 0, inert_n = 0, stale_steps = 0;
    for (int t = 0; t <= T; ++t) { inert_steps += part[(size_t)t]; inert_n += ninert[(size_t)t]; stale_steps += pstale[(size_t)t]; }
    int64_t total = steps + inert_steps + stale_steps;
    if (!unlimited && !(str_steps + total + 1 < str_lim)) { return false; }
    // clause contents and flags; the real entries are done (the inert ones lost their flag up front), re-queued clauses are
    // processed too - or left over by the last entry of the queue
    parallelFor(reals.size(), [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { C[strq[reals[i]]].flag &= ~F_STR; } }, 4096);
    std::vector<const Final*> fins; fins.reserve(nfin);
    std::vector<Ev> evs; evs.reserve(nev);
    for (const Out& o : outs) {
        for (const Comp& cp : o.comps) {
            if (!cp.valid) { continue; }
            for (uint32_t k = cp.fin0; k < cp.fin1; ++k) { fins.push_back(&o.fin[k]); }
            evs.insert(evs.end(), o.evs.begin() + cp.ev0, o.evs.begin() + cp.ev1);
        }
    }
    parallelFor(fins.size(), [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            const Final& f = *fins[i]; ClauseInfo& ci = C[f.clause];
            memcpy(pool.data() + ci.start, f.lits.data(), sizeof(int32_t) * f.lits.size());
            ci.size = (uint32_t)f.lits.size(); ci.ver += f.nrem; ci.stamp = scan_id;
            if (f.sub) { ci.flag |= F_SUB; }
            if (f.str) { ci.flag |= F_STR; } else { ci.flag &= ~F_STR; }
        }
    }, 4096);
    for (const Final* f : fins) {
        const uint32_t c = f->clause;
        if (dev_uploaded && c < shrunk_mark.size() && !shrunk_mark[c]) { shrunk_mark[c] = 1; pend_shrunk.push_back(c); }
    }
    // queue pushes and occurrence updates in reference order: sorted by (queue time, sequence) - ranges of the key per core
    {
        const uint64_t kmax = ((uint64_t)nq << 24) + 1, kw = kmax / (uint64_t)T + 1;
        std::vector<std::vector<Ev>> kb((size_t)T);
        parallelFor((size_t)T, [&](size_t b, size_t e, int) {
            for (size_t bk = b; bk < e; ++bk) {
                std::vector<Ev>& v = kb[bk];
                for (const Ev& ev : evs) { if (ev.key / kw == bk) { v.push_back(ev); } }
                std::sort(v.begin(), v.end(), [](const Ev& a, const Ev& b2) { return a.key < b2.key; });
            }
        }, 1);
        // (with an empty BVE heap touching a variable is a counter increment: on all cores, like the deaths of a subsumption pass)
        const bool touch_seq = opt.bve && !heap.empty();
        size_t nsub = 0; std::vector<size_t> base((size_t)T + 1, 0), sbase((size_t)T + 1, 0);
        for (size_t bk = 0; bk < (size_t)T; ++bk) { size_t c2 = 0; for (const Ev& ev : kb[bk]) { c2 += ev.push_sub ? 1 : 0; } base[bk + 1] = base[bk] + kb[bk].size(); sbase[bk + 1] = sbase[bk] + c2; nsub += c2; }
        const size_t s0 = subq.size(), u0 = occ_upd.size();
        subq.resize(s0 + nsub); occ_upd.resize(u0 + base[(size_t)T]);
        parallelFor((size_t)T, [&](size_t b, size_t e, int) {
            for (size_t bk = b; bk < e; ++bk) {
                size_t si = s0 + sbase[bk], ui = u0 + base[bk];
                for (const Ev& ev : kb[bk]) {
                    if (ev.push_sub) { subq[si++] = ev.clause; }
                    occ_upd[ui++] = OccUpd{ev.clause, ev.lit};
                    if (opt.bve && !touch_seq) {
                        const int32_t* l = CL(ev.clause);
                        for (uint32_t k = 0; k < C[ev.clause].size; ++k) { __atomic_fetch_add(&var_touch[(size_t)(l[k] >> 1)], 1u, __ATOMIC_RELAXED); }
                        __atomic_fetch_add(&var_touch[(size_t)(ev.lit >> 1)], 1u, __ATOMIC_RELAXED);
                    }
                }
            }
        }, 1);
        if (touch_seq) { for (const std::vector<Ev>& v : kb) { for (const Ev& ev : v) { touchClauseVars(ev.clause); bveTouch(ev.lit >> 1); } } }
    }
    removedLiterals += (int)evs.size();
    if (!evs.empty()) { successful(t_sub); }
    // lists that a walk cleaned
    std::vector<int64_t> premoved((size_t)T + 1, 0);
    parallelForG((nlit + 63) / 64, 512, [&](size_t wb, size_t we, int t) {
        int64_t removed = 0;
        for (size_t x = wb * 64; x < std::min(nlit, we * 64); ++x) {
            if (!walked_stale[x]) { continue; }
            uint32_t* p = occ_pool.data() + occ[x].off; uint32_t n = hot[x].n;
            for (uint32_t j = 0; j < n;) { if (isDel(p[j])) { p[j] = p[n - 1]; --n; ++removed; } else { ++j; } }
            hot[x].n = n; staleClear((int)x);
        }
        premoved[(size_t)t] += removed;
    });
    if (unlimited) { for (int64_t r : premoved) { total += r; } }
    str_steps += total; n_fast += nfast; n_slow += nslow; n_lazy = inert_n;
    ++n_compar;
    lap("commit");
    return true;
}

// The first strengthening pass of a run, planned by the device.  The queue holds every clause in clause order and the device still
// holds the tables of the subsumption pass it has just committed (list lengths, deleted entries still listed, deaths per literal):
// K10 evaluates every queue entry that no cached hit touches - literal choice, the two list lengths, the lists it walks - in one
// launch; the host only needs the entries hits do touch (O(#hits)) and commits them by component.
bool Engine::strengthenFirstPassDevice()
{
    const size_t ncl = C.size(), nq = strq.size();
    if (nq != ncl || ncl == 0 || c_str.slot.size() < ncl) { return false; }
    std::atomic<bool> fit{true};
    parallelFor(nq, [&](size_t b, size_t e, int) {
        bool good = true;
        for (size_t p = b; p < e && good; ++p) { const ClauseInfo& ci = C[p]; good = strq[p] == (uint32_t)p && ((ci.flag & F_DEL) || (ci.flag & F_STR)); }
        if (!good) { fit = false; }
    });
    if (!fit) { return false; }
    DevStatic ds;
    ds.real_bits.assign((ncl + 31) / 32, 0u);
    // real entries: clauses with a scan record (own hits, predicted contents) and the targets of all cached hits
    parallelFor(c_str.used.size(), [&](size_t b, size_t e, int) { for (size_t u = b; u < e; ++u) { const uint32_t c = c_str.used[u]; __atomic_fetch_or(&ds.real_bits[c >> 5], 1u << (c & 31), __ATOMIC_RELAXED); } }, 8192);
    parallelFor(c_str.hits.size(), [&](size_t b, size_t e, int) { for (size_t h = b; h < e; ++h) { const uint32_t c = c_str.hits[h].clause; __atomic_fetch_or(&ds.real_bits[c >> 5], 1u << (c & 31), __ATOMIC_RELAXED); } }, 8192);
    ds.walked_bits.assign(((size_t)2 * nvars + 31) / 32 + 1, 0u);
    int status = 1; uint64_t small = 0;
    const int64_t t0 = now_ns();
    if (cp3g_str_static(gpu, ds.real_bits.data(), &status, &ds.inert_n, &ds.inert_steps, &small, ds.walked_bits.data()) != CP3G_OK) { die("device static plan failed"); }
    host_ns_gpuwait += now_ns() - t0;
    if (status != 0 || small != 0) { return false; }
    // queue positions of the real entries = their clause indices, ascending
    std::vector<uint32_t> reals;
    {
        const size_t words = ds.real_bits.size();
        std::vector<size_t> cnt((size_t)nthreads + 2, 0);
        parallelFor(words, [&](size_t b, size_t e, int t) { size_t c = 0; for (size_t w = b; w < e; ++w) { c += (size_t)__builtin_popcount(ds.real_bits[w]); } cnt[(size_t)t + 1] = c; }, 4096);
        for (size_t t = 1; t < cnt.size(); ++t) { cnt[t] += cnt[t - 1]; }
        reals.resize(cnt.back());
        parallelFor(words, [&](size_t b, size_t e, int t) {
            size_t o = cnt[(size_t)t];
            for (size_t w = b; w < e; ++w) { uint32_t x = ds.real_bits[w]; while (x) { reals[o++] = (uint32_t)(w * 32 + (size_t)__builtin_ctz(x)); x &= x - 1; } }
        }, 4096);
    }
    if (!strengthenComponents(reals, &ds)) { return false; }
    ++n_devstatic;
    return true;
}

} // namespace cp3
