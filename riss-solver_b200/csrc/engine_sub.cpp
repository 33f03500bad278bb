// engine_sub.cpp - subsumption commit of the host engine: time-indexed parallel form, device-side first pass, ordered walk.
#include "engine_util.h"

namespace cp3 {

// ------------------------------------------------------------------------------------------------ Subsumption
void Engine::clearQueue(RawVec<uint32_t>& q, uint8_t flag)
{
    // a queue in clause order (first pass) names every record once: plain stores, chunk by chunk; otherwise a clause may be queued
    // twice and two chunks may meet in one record
    if (q.size() >= 65536 && nthreads > 1) {
        if (isIncreasing(q)) { parallelFor(q.size(), [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { C[q[i]].flag &= ~flag; } }); }
        else { parallelFor(q.size(), [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { __atomic_fetch_and((uint32_t*)&C[q[i]] + 1, ~((uint32_t)flag << 24), __ATOMIC_RELAXED); } }); }
    }
    else { for (uint32_t c : q) { C[c].flag &= ~flag; } }
    q.clear();
}


// Time-indexed form of the subsumption queue walk (subsumption_worker, Subsumption.cc:244-297) for large queues.
//
// The ordered walk is sequential only through three kinds of state, all of which are functions of *when the
// subsumed clauses die*: (1) whether a subsumer is still alive at its turn, (2) lit_occurrence_count, which picks
// the list a subsumer walks (removedClause decrements it at the moment of deletion), (3) which entries a walk
// finds deleted and removes by swap-with-last.  Death times follow from the K3 hit lists alone - clause contents
// do not change during this pass, so every hit is valid and a subsumer reaches every live superset through any of
// its lists - by visiting ONLY the queue entries that have hits, in order (step 1, sequential, O(#hits)).
// Everything else is then evaluated independently: the list choice of every queue entry against the counters
// as of its own time (step 3, parallel over entries) and the lazy cleaning of every list that sees a deletion,
// replayed over that list's walks in time order (step 4, parallel over lists).  The result - deleted set, final
// list orders, step counter, doubly decremented counters - is identical to the ordered walk.
bool Engine::subsumptionCommitParallel()
{
    const size_t nq = subq.size();
    const int32_t INF = INT32_MAX;
    const bool dbg = tune.debug_par; int64_t tt = now_ns();
    auto lap = [&](const char* w) { if (dbg) { int64_t n2 = now_ns(); fprintf(stderr, "  parsub %s %.1f ms\n", w, (n2 - tt) / 1e6); tt = n2; } };
    if (sub_dtime.size() < C.size()) { pgrow(sub_dtime, C.size() + 1024, INF); pgrow(sub_pos, C.size() + 1024, -1); }
    // step 0: queue position of every clause; a clause queued twice falls back to the ordered walk
    bool dups = false;
    bool increasing = true;                              // a queue in clause order (the first round) cannot hold duplicates
    const size_t qgrain = 4096;                          // queues of later rounds are short: small chunks keep all cores busy
    parallelFor(nq, [&](size_t b, size_t e, int) {
        bool inc = true;
        for (size_t p = std::max<size_t>(b, 1); p < e && inc; ++p) { inc = subq[p - 1] < subq[p]; }
        if (!inc) { increasing = false; }
    }, qgrain);
    if (increasing) {
        parallelFor(nq, [&](size_t b, size_t e, int) { for (size_t p = b; p < e; ++p) { sub_pos[subq[p]] = (int32_t)p; } }, qgrain);
    } else {
        parallelFor(nq, [&](size_t b, size_t e, int) {
            for (size_t p = b; p < e; ++p) {
                const int32_t old = __atomic_exchange_n(&sub_pos[subq[p]], (int32_t)p, __ATOMIC_RELAXED);
                if (old >= 0) { dups = true; }
            }
        }, qgrain);
    }
    auto reset_pos = [&]() { parallelFor(nq, [&](size_t b, size_t e, int) { for (size_t p = b; p < e; ++p) { sub_pos[subq[p]] = -1; } }, qgrain); };
    if (dups) { reset_pos(); return false; }
    lap("step0");
    // step 1: death times.  Entries with hits, in processing order (time = nq-1-position).
    std::vector<std::pair<int32_t, uint32_t>> hitters;           // (time, cache entry)
    // the cache is keyed by clause: walk the clauses that own an entry
    if (c_sub.used.size() >= std::max<size_t>(tune.par_min, 1) && nthreads > 1) {
        // long lists: collected per time range on all cores, each range sorted by its owner, ranges concatenated
        const size_t NBk = (size_t)nthreads;
        std::vector<std::vector<std::pair<int32_t, uint32_t>>> parts(((size_t)nthreads + 1) * NBk);
        parallelFor(c_sub.used.size(), [&](size_t b, size_t e, int t) {
            std::vector<std::pair<int32_t, uint32_t>>* out = &parts[(size_t)t * NBk];
            for (size_t u = b; u < e; ++u) {
                const uint32_t c = c_sub.used[u]; const int32_t slot = c_sub.slot[c];
                if (slot < 0 || c_sub.ent[slot].hb == c_sub.ent[slot].he || sub_pos[c] < 0 || (C[c].flag & F_DEL)) { continue; }
                const int32_t tm = (int32_t)(nq - 1 - (size_t)sub_pos[c]);
                out[(size_t)tm * NBk / nq].push_back({tm, (uint32_t)slot});
            }
        }, 4096);
        std::vector<std::vector<std::pair<int32_t, uint32_t>>> sorted(NBk);
        parallelFor(NBk, [&](size_t b, size_t e, int) {
            for (size_t bk = b; bk < e; ++bk) {
                std::vector<std::pair<int32_t, uint32_t>>& v = sorted[bk];
                for (size_t t2 = 0; t2 <= (size_t)nthreads; ++t2) { const auto& src = parts[t2 * NBk + bk]; v.insert(v.end(), src.begin(), src.end()); }
                std::sort(v.begin(), v.end());
            }
        }, 1);
        for (const auto& v : sorted) { hitters.insert(hitters.end(), v.begin(), v.end()); }
    } else {
    for (uint32_t c : c_sub.used) {
        const int32_t slot = c_sub.slot[c];
        if (slot < 0 || c_sub.ent[slot].hb == c_sub.ent[slot].he || sub_pos[c] < 0 || (C[c].flag & F_DEL)) { continue; }
        hitters.push_back({(int32_t)(nq - 1 - (size_t)sub_pos[c]), (uint32_t)slot});
    }
    std::sort(hitters.begin(), hitters.end());
    }
    std::vector<uint32_t> dead;
    for (const auto& h : hitters) {
        const CacheEnt& e = c_sub.ent[h.second];
        const uint32_t cr = subq[nq - 1 - (size_t)h.first];
        if (sub_dtime[cr] < h.first) { continue; }                 // subsumed before its turn
        for (uint32_t k = e.hb; k < e.he; ++k) {
            const uint32_t d = c_sub.hits[k].clause;
            if ((C[d].flag & F_DEL) || sub_dtime[d] != INF) { continue; }
            sub_dtime[d] = h.first; dead.push_back(d);
        }
    }
    lap("step1");
    if (dbg) { fprintf(stderr, "  parsub nq %zu hitters %zu dead %zu\n", nq, hitters.size(), dead.size()); }
    // step 2: per literal, the sorted death times of the clauses that contain it (bucketed by literal range,
    //         64-aligned so that no two buckets share a word of the stale bitmap)
    struct LT { int32_t lit, time; };
    const size_t nlit = (size_t)2 * nvars;
    const int NB = std::max(1, nthreads);
    auto bucketLo = [&](int bk) -> int32_t { return (int32_t)std::min(nlit, ((nlit * (size_t)bk / NB) + 63) / 64 * 64); };
    const size_t bwidth = std::max<size_t>(64, (nlit / NB) / 64 * 64);
    auto bucketOf = [&](int x) -> int { int bk = (int)std::min<size_t>((size_t)NB - 1, (size_t)x / bwidth); while (bk > 0 && x < bucketLo(bk)) { --bk; } while (bk + 1 < NB && x >= bucketLo(bk + 1)) { ++bk; } return bk; };
    std::vector<std::vector<LT>> dlb((size_t)NB);
    if (dl_start.size() < nlit) { pfill(dl_start, nlit, -1); pfill(dl_cnt, nlit, 0u); dl_any.assign((nlit >> 6) + 1, 0ull); }
    // (the dead clauses are read once, each producer files their literals per bucket; the bucket owners then merge and sort)
    std::vector<std::vector<LT>> dparts(((size_t)nthreads + 1) * (size_t)NB);
    parallelFor(dead.size(), [&](size_t b, size_t e, int t) {
        std::vector<LT>* out = &dparts[(size_t)t * (size_t)NB];
        for (size_t i = b; i < e; ++i) { const uint32_t d = dead[i]; const int32_t* l = CL(d); const int32_t dt = sub_dtime[d]; for (uint32_t k = 0; k < C[d].size; ++k) { out[bucketOf(l[k])].push_back({l[k], dt}); } }
    }, 2048);
    parallelFor((size_t)NB, [&](size_t bb, size_t be, int) {
        for (size_t bk = bb; bk < be; ++bk) {
            std::vector<LT>& v = dlb[bk];
            for (size_t t2 = 0; t2 <= (size_t)nthreads; ++t2) { const std::vector<LT>& src = dparts[t2 * (size_t)NB + bk]; v.insert(v.end(), src.begin(), src.end()); }
            std::sort(v.begin(), v.end(), [](const LT& a2, const LT& b2) { return a2.lit != b2.lit ? a2.lit < b2.lit : a2.time < b2.time; });
            for (size_t i = 0; i < v.size(); ++i) { if (dl_start[v[i].lit] < 0) { dl_start[v[i].lit] = (int32_t)i; dl_any[(uint32_t)v[i].lit >> 6] |= 1ull << (v[i].lit & 63); } ++dl_cnt[v[i].lit]; }
        }
    }, 1);
    // 2-byte copy of the counters as of the pass start (cache resident); out-of-range values read the original
    RawVec<int16_t> cnt16; cnt16.resize(nlit);                  // filled right below, on all cores
    parallelFor(nlit, [&](size_t b, size_t e, int) {
        for (size_t x = b; x < e; ++x) { const int32_t c = hot[x].cnt; cnt16[x] = (c > 32000 || c < -32000) ? INT16_MIN : (int16_t)c; }
    });
    lap("step2");
    auto countAt = [&](int x, int32_t T) -> int32_t {               // lit_occurrence_count[x] as of time T
        int32_t c = cnt16[(size_t)x]; if (c == INT16_MIN) { c = hot[x].cnt; }
        if ((dl_any[(uint32_t)x >> 6] >> (x & 63)) & 1ull) { const LT* q = dlb[bucketOf(x)].data() + dl_start[x]; for (uint32_t k = 0; k < dl_cnt[x] && q[k].time < T; ++k) { --c; } }
        return c;
    };
    // step 3: every queue entry picks its list with the counters of its own time.  Walks of lists that never hold
    // a deleted clause just add |list|-1 steps; the others become events for step 4.
    struct Ev { int32_t lit, time; uint32_t pos; };
    // events are filed per (producing thread, literal bucket), so that step 4 reads only what belongs to its bucket
    std::vector<std::vector<Ev>> tev(((size_t)nthreads + 1) * (size_t)NB); std::vector<int64_t> tsteps((size_t)nthreads + 1, 0);
    parallelFor(nq, [&](size_t b, size_t e, int t) {
        int64_t st = 0; std::vector<Ev>* ev = &tev[(size_t)t * (size_t)NB];
        for (int k = 0; k < NB; ++k) { ev[k].reserve((e - b) / (size_t)NB / 2 + 64); }     // about half of the entries walk a list with deaths
        for (size_t p = b; p < e; ++p) {
            const uint32_t cr = subq[p]; const ClauseInfo ci = C[cr];
            if (ci.flag & F_DEL) { continue; }
            const int32_t T = (int32_t)(nq - 1 - p);
            if (sub_dtime[cr] < T) { continue; }
            const int32_t* c = pool.data() + ci.start;
            int mn = c[0]; int32_t best = countAt(mn, T);
            for (uint32_t l = 1; l < ci.size; ++l) { const int32_t v = countAt(c[l], T); if (v < best) { best = v; mn = c[l]; } }
            if (!((dl_any[(uint32_t)mn >> 6] >> (mn & 63)) & 1ull) && !hasStale(mn)) { st += hot[mn].n ? (int64_t)hot[mn].n - 1 : 0; }
            else { ev[bucketOf(mn)].push_back({mn, T, (uint32_t)p}); }
        }
        tsteps[t] = st;
    }, qgrain);
    int64_t steps = 0; for (int64_t x : tsteps) { steps += x; }
    lap("step3");
    // step 4: replay the walks of every list that holds a deleted clause, list by list, in time order
    std::vector<int64_t> bsteps((size_t)NB, 0);
    std::vector<std::vector<int32_t>> bwalked((size_t)NB);
    parallelFor((size_t)NB, [&](size_t bb, size_t be, int) {
        for (size_t bk = bb; bk < be; ++bk) {
            const int32_t lo = bucketLo((int)bk), hi = (bk + 1 == (size_t)NB) ? (int32_t)nlit : bucketLo((int)bk + 1);
            std::vector<Ev> mine;
            { size_t tot = 0; for (size_t t2 = 0; t2 <= (size_t)nthreads; ++t2) { tot += tev[t2 * (size_t)NB + bk].size(); } mine.reserve(tot); }
            for (size_t t2 = 0; t2 <= (size_t)nthreads; ++t2) { const std::vector<Ev>& v = tev[t2 * (size_t)NB + bk]; mine.insert(mine.end(), v.begin(), v.end()); }
            (void)lo; (void)hi;
            std::sort(mine.begin(), mine.end(), [](const Ev& a, const Ev& b) { return a.lit != b.lit ? a.lit < b.lit : a.time < b.time; });
            int64_t st = 0;
            for (size_t i = 0; i < mine.size();) {
                const int x = mine[i].lit;
                uint32_t* lp = occ_pool.data() + occ[x].off; uint32_t n = hot[x].n; uint32_t removed = 0;
                for (; i < mine.size() && mine[i].lit == x; ++i) {
                    const uint32_t cr = subq[mine[i].pos]; const uint32_t cn = C[cr].size; const int32_t T = mine[i].time;
                    for (uint32_t j = 0; j < n; ++j) {
                        const uint32_t d = lp[j];
                        if (d == cr) { continue; }
                        ++st;
                        if ((C[d].flag & F_DEL) || sub_dtime[d] < T) {          // found deleted: size filter first, then lazy removal
                            if (C[d].size < cn) { continue; }
                            lp[j] = lp[n - 1]; --n; --j; ++removed;
                        }
                    }
                }
                hot[x].n = n;
                // deleted clauses still listed here afterwards = those before + the ones that died now - the removed
                stale[x] = stale[x] + dl_cnt[x] - removed;
                bwalked[bk].push_back(x);
            }
            // lists that hold a newly deleted clause but were not walked keep it: count it as stale.  Then the
            // counters (removedClause twice per clause - Subsumption.cc:284,305) and the stale bitmap of this bucket.
            const std::vector<LT>& v = dlb[bk]; size_t w = 0; const std::vector<int32_t>& wl = bwalked[bk];
            for (size_t i = 0; i < v.size(); i += dl_cnt[v[i].lit]) {
                const int x = v[i].lit;
                while (w < wl.size() && wl[w] < x) { ++w; }
                if (!(w < wl.size() && wl[w] == x)) { stale[x] += dl_cnt[x]; }
                hot[x].cnt -= 2 * (int32_t)dl_cnt[x];
                if (stale[x]) { stalebits[(uint32_t)x >> 6] |= 1ull << (x & 63); } else { stalebits[(uint32_t)x >> 6] &= ~(1ull << (x & 63)); }
            }
            for (int32_t x : wl) { if (stale[x]) { stalebits[(uint32_t)x >> 6] |= 1ull << (x & 63); } else { stalebits[(uint32_t)x >> 6] &= ~(1ull << (x & 63)); } }
            bsteps[bk] = st;
        }
    }, 1);
    for (int64_t x : bsteps) { steps += x; }
    lap("step4");
    // step 5: commit the deletions themselves
    {
        std::vector<int64_t> tlits((size_t)nthreads + 1, 0);
        parallelFor(dead.size(), [&](size_t b2, size_t e2, int t) {
            int64_t nl = 0;
            for (size_t i = b2; i < e2; ++i) {
                const uint32_t d = dead[i];
                nl += C[d].size;
                C[d].flag |= F_DEL;
                __atomic_fetch_or(&delbits[d >> 6], 1ull << (d & 63), __ATOMIC_RELAXED);
                const int32_t* l = CL(d);
                for (uint32_t k = 0; k < C[d].size; ++k) { deletedVar(l[k] >> 1); }
            }
            tlits[t] = nl;
        }, 8192);
        int64_t nl = 0; for (int64_t x : tlits) { nl += x; }
        subsumedClauses += (int)dead.size(); subsumedLiterals += (int)nl;
        n_live -= (uint32_t)dead.size();
        numberOfCls -= 2 * (int)dead.size(); ca_wasted += 2 * (2.0 * (double)dead.size() + (double)nl);
        if (dev_uploaded) { pend_del.insert(pend_del.end(), dead.begin(), dead.end()); }
        touchDeadClauses(dead);
    }
    if (!dead.empty()) { successful(t_sub); }
    sub_steps += steps;
    // sparse resets
    for (uint32_t d : dead) { sub_dtime[d] = INF; }
    for (const auto& v : dlb) { for (const LT& x : v) { dl_start[x.lit] = -1; dl_cnt[x.lit] = 0; dl_any[(uint32_t)x.lit >> 6] = 0ull; } }
    reset_pos();
    lap("step5");
    ++n_parsub;
    return true;
}

// The lists K2 built (clause order, CoprocessorTypes.h:812-833) come to the host when somebody first needs them.
void Engine::ensureHostLists()
{
    if (!lists_on_device) { return; }
    lists_on_device = false;
    const int n2 = 2 * nvars;
    int64_t t0 = now_ns();
    std::vector<uint32_t> off((size_t)n2 + 1, 0);
    if (cp3g_download_occ(gpu, off.data(), occ_pool.data()) != CP3G_OK) { die("download_occ failed"); }
    ph_download += now_ns() - t0; host_ns_gpuwait += now_ns() - t0;
}

// First subsumption pass, committed by the device (cp3g_sub_pass, commit.cu): the queue holds every live clause in clause
// order and the lists are pristine.  The host receives the subsumed set, the step counter and the lists in their post-pass
// order and does the O(#subsumed) bookkeeping of subsumptionCommitParallel's last step.
bool Engine::subsumptionCommitDevice()
{
    const bool off_env = tune.no_devcommit;
    const uint32_t max_list = tune.devcommit_maxlist;
    if (off_env || !lists_on_device || !dev_uploaded || !pend_del.empty() || !pend_shrunk.empty() || dev_ncl != C.size()) { return false; }
    const size_t nq = subq.size();
    if (nq != n_live || nq == 0) { return false; }
    bool ok_q = true;                                    // every entry live and flagged, strictly increasing clause index
    parallelFor(nq, [&](size_t b, size_t e, int) {
        bool good = true;
        for (size_t p = b; p < e && good; ++p) { const ClauseInfo& ci = C[subq[p]]; good = !(ci.flag & F_DEL) && (ci.flag & F_SUB) && (p == 0 || subq[p - 1] < subq[p]); }
        if (!good) { ok_q = false; }
    });
    if (!ok_q) { return false; }
    DbgLap lap("devsub", tune.debug_par);
    int status = 1; uint32_t ndead = 0; uint64_t nlits = 0, steps = 0, checks = 0;
    int64_t t0 = now_ns();
    if (cp3g_sub_pass(gpu, max_list, &status, &ndead, &nlits, &steps, &checks) != CP3G_OK) { die("device subsumption pass failed"); }
    host_ns_gpuwait += now_ns() - t0;
    lap("cp3g_sub_pass");
    if (status != 0) { return false; }
    ++scan_id; ++n_scan_batches; ++n_devsub;
    const size_t nlit = (size_t)2 * nvars;
    std::vector<uint32_t> dead(ndead); RawVec<uint32_t> ln, ls; ln.resize(nlit); ls.resize(nlit);
    t0 = now_ns();
    if (cp3g_sub_pass_fetch(gpu, dead.data(), ln.data(), ls.data(), occ_pool.data()) != CP3G_OK) { die("device subsumption pass: fetch failed"); }
    host_ns_gpuwait += now_ns() - t0; ph_download += now_ns() - t0;
    lists_on_device = false;
    lap("fetch");
    parallelForG((nlit + 63) / 64, 512, [&](size_t wb, size_t we, int) {      // 64-literal granules: stalebits words are not shared
        for (size_t x = wb * 64; x < std::min(nlit, we * 64); ++x) {
            hot[x].n = ln[x]; stale[x] = ls[x];
            if (ls[x]) { stalebits[x >> 6] |= 1ull << (x & 63); }
        }
    });
    parallelFor(dead.size(), [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            const uint32_t d = dead[i];
            C[d].flag |= F_DEL;
            __atomic_fetch_or(&delbits[d >> 6], 1ull << (d & 63), __ATOMIC_RELAXED);
            const int32_t* l = CL(d);
            // removedClause twice per subsumed clause (Subsumption.cc:284,305)
            for (uint32_t k = 0; k < C[d].size; ++k) { deletedVar(l[k] >> 1); __atomic_fetch_sub(&hot[l[k]].cnt, 2, __ATOMIC_RELAXED); }
        }
    }, 8192);
    subsumedClauses += (int)ndead; subsumedLiterals += (int)nlits;
    n_live -= ndead;
    numberOfCls -= 2 * (int)ndead; ca_wasted += 2 * (2.0 * (double)ndead + (double)nlits);
    touchDeadClauses(dead);
    if (ndead) { successful(t_sub); }
    sub_steps += (int64_t)steps;
    dev_static_ok = true;                              // the device holds the post-pass tables: K10 can plan the strengthening pass
    lap("bookkeeping");                                // (the queue flags are cleared by the caller's clearQueue)
    return true;
}

// subsumption_worker, Subsumption.cc:220-314.  One bulk K3 scan answers every ordered_subsumes() of the
// queue; the loop below replays the queue back to front.
void Engine::subsumptionWorker(bool useHeap, int ignore)
{
    PhaseTimer pt(ph_sub);
    const bool unlimited = !opt.limited;
    DbgLap lap0("sub", tune.debug_par);
    c_sub.clear(); logClear();
    lap0("clear");
    if (lists_on_device) {
        const size_t parsub_min0 = tune.parsub_min;
        bool far = unlimited;
        if (!far && !useHeap && subq.size() >= parsub_min0) {      // the device commit cannot stop in the middle of the queue: only far from the step limit
            std::vector<int64_t> part((size_t)nthreads + 1, 0);
            parallelFor(subq.size(), [&](size_t b, size_t e, int t) {
                int64_t a = 0;
                for (size_t p = b; p < e; ++p) { const ClauseInfo& ci = C[subq[p]]; if (ci.flag & F_DEL) { continue; } uint32_t mx = 0; const int32_t* l = pool.data() + ci.start; for (uint32_t k = 0; k < ci.size; ++k) { mx = std::max(mx, hot[l[k]].n); } a += mx; }
                part[(size_t)t] = a;
            });
            int64_t bound = 0; for (int64_t a : part) { bound += a; }
            far = sub_steps + bound < sub_lim;
        }
        if (!useHeap && subq.size() >= parsub_min0 && far) {
            const int64_t t00 = now_ns();
            if (subsumptionCommitDevice()) { host_ns_commit += now_ns() - t00; lap0("device commit"); return; }
        }
        ensureHostLists();
    }
    {
        std::vector<uint32_t> ids;
        parallelFilter(subq, ids, [&](uint32_t c) { return !(C[c].flag & F_DEL); });
        sortUniqueIds(ids);
        lap0("ids");
        scanClauses(CP3G_SUBSUME, ids, c_sub);
    }
    DbgLap lap("sub", tune.debug_par); lap("start(after scan)");
    int64_t t0 = now_ns();
    const size_t parsub_min = tune.parsub_min;
    if (!useHeap && subq.size() >= parsub_min) {
        // step limit: the data-parallel form cannot stop in the middle of the queue; only use it far from the limit
        int64_t bound = 0;
        if (!unlimited) { for (uint32_t c : subq) { if (!(C[c].flag & F_DEL)) { int mx = 0; const int32_t* l = CL(c); for (uint32_t k = 0; k < C[c].size; ++k) { mx = std::max<int>(mx, (int)hot[l[k]].n); } bound += mx; } } }
        if (unlimited || sub_steps + bound < sub_lim) {
            // the strengthening pass that follows needs nothing from this commit but the deleted flags (a hit on a
            // clause that dies meanwhile is dropped at commit time): run its scan + closure concurrently
            const bool no_overlap = tune.no_overlap;
            std::thread helper; std::vector<uint32_t> sids; bool sbulk = false;
            const bool ov = overlap_str && !no_overlap && nthreads > 1 && opt.strength && strq.size() >= parsub_min && !hasToPropagate();
            if (ov) {
                for (uint32_t c : strq) { if (!(C[c].flag & F_DEL) && (C[c].flag & F_STR) && C[c].size >= 2) { sids.push_back(c); } }
                sortUniqueIds(sids);
                const uint32_t bulk_min2 = tune.bulk_min;
                sbulk = sids.size() == n_live && n_live > bulk_min2;
                flushDevice();
                scan_no_flush = true;
                helper = std::thread([this, &sids, sbulk]() { prepareStrengthening(sids, sbulk); });
            }
            const bool done = subsumptionCommitParallel();
            lap("commitParallel");
            if (ov) { helper.join(); lap("helper join"); scan_no_flush = false; str_prepared = true; ++n_overlap; }
            if (done) {
                for (uint32_t c : subq) { C[c].flag &= ~F_SUB; }
                host_ns_commit += now_ns() - t0;
                return;
            }
        }
    }
    size_t end = subq.size();
    sub_occ_updates.clear();
    const size_t PF = 16;                               // the walk is bound by cache misses on occ[lit]
    for (; end > 0 && (unlimited || sub_steps < sub_lim);) {
        --end;
        if (end >= PF) {
            const uint32_t c2 = subq[end - PF];
            const int32_t* l2 = CL(c2); const uint32_t n2 = std::min<uint32_t>(C[c2].size, 6);
            for (uint32_t k = 0; k < n2; ++k) { __builtin_prefetch(&hot[l2[k]]); }
        }
        const uint32_t cr = subq[end];
        if (C[cr].flag & F_DEL) { continue; }
        const int32_t* c = CL(cr); const uint32_t cn = C[cr].size;
        int min = c[0];
        for (uint32_t l = 1; l < cn; ++l) { if (hot[c[l]].cnt < hot[min].cnt) { min = c[l]; } }
        const CacheEnt* e = cached(c_sub, cr);
        if (!e) { die("internal: subsumer without scan result"); }
        ListRef list = L(min);
        if (e->hb == e->he && !hasStale(min) && (unlimited || sub_steps + (int64_t)list.n < sub_lim)) {
            sub_steps += list.n ? (int64_t)list.n - 1 : 0;          // every entry but cr itself is one check
            ++n_fast;
        } else {
            ++n_slow;
            for (uint32_t i = 0; i < list.n && (unlimited || sub_steps < sub_lim); ++i) {
                const uint32_t d = LP(list)[i];
                if (d == cr) { continue; }
                sub_steps++;
                // reference order: size filter, then lazy removal of deleted entries, then the subsumption test.
                // For a live clause the size filter is implied by the test, so only deleted entries need C[d].
                if (isDel(d)) { if (C[d].size < cn) { continue; } occSwapRemove(min, i); --i; continue; }
                if (e->hb != e->he && findHit(c_sub.hits, e->hb, e->he, d) && subHitValid(cr, d, e->scan)) {
                    ++subsumedClauses; subsumedLiterals += (int)C[d].size;
                    setDeleted(d);
                    removedClause(d, false, false, VAR_UNDEF);
                    successful(t_sub);
                    sub_occ_updates.push_back(d);
                }
            }
        }
        C[cr].flag &= ~F_SUB;
    }
    // QUIRK: removedClause a second time per subsumed clause (Subsumption.cc:302-313); `ignore` lands in `update`
    for (uint32_t d : sub_occ_updates) { removedClause(d, useHeap, ignore != 0, VAR_UNDEF); }
    sub_occ_updates.clear();
    lap("ordered walk");
    host_ns_commit += now_ns() - t0;
}

} // namespace cp3
