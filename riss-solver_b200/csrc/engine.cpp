// engine.cpp - ordered-commit host engine (see engine.h): solver side, device replica and scan cache, occurrence lists, heap,
// CoprocessorData bookkeeping, unit propagation, Dense and the preprocess() driver.  The techniques live in engine_sub.cpp
// (subsumption), engine_str.cpp (strengthening), engine_bve.cpp (BVE), engine_bce.cpp (BCE).  Reference lines are cited per function.
#include "engine_util.h"
#include <atomic>
#include <sys/mman.h>
#include <condition_variable>
#include <mutex>
#include <new>

namespace cp3 {

// ------------------------------------------------------------------------------------------------ worker pool
struct Engine::WorkerPool {
    std::vector<std::thread> th;
    std::mutex m, user;                                  // m guards the job state, user serialises callers
    std::condition_variable cv, done;
    const std::function<void(int)>* job = nullptr;
    int ntasks = 0, next = 0, pending = 0; uint64_t gen = 0; bool stop = false;
    explicit WorkerPool(int workers)
    {
        for (int w = 0; w < workers; ++w) {
            th.emplace_back([this]() {
                uint64_t seen = 0;
                std::unique_lock<std::mutex> lk(m);
                for (;;) {
                    cv.wait(lk, [&]() { return stop || gen != seen; });
                    if (stop) { return; }
                    seen = gen;
                    while (next < ntasks) {
                        const int t = next++;
                        const std::function<void(int)>* j = job;
                        lk.unlock(); (*j)(t); lk.lock();
                        if (--pending == 0) { done.notify_all(); }
                    }
                }
            });
        }
    }
    ~WorkerPool()
    {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        for (auto& x : th) { x.join(); }
    }
};

void Engine::poolRun(int tasks, const std::function<void(int)>& job)
{
    if (tasks <= 0) { return; }
    if (tasks == 1) { job(0); return; }
    if (!pool_) { pool_ = new WorkerPool(std::max(1, nthreads - 1)); }
    WorkerPool& P = *pool_;
    if (!P.user.try_lock()) {                            // the pool is busy with another caller's loop: plain threads for this one
        std::vector<std::thread> th;
        for (int t = 0; t < tasks; ++t) { th.emplace_back([&job, t]() { job(t); }); }
        for (auto& x : th) { x.join(); }
        return;
    }
    {
        std::unique_lock<std::mutex> lk(P.m);
        P.job = &job; P.ntasks = tasks; P.next = 0; P.pending = tasks; ++P.gen;
        P.cv.notify_all();
        while (P.next < P.ntasks) {                      // the caller works too
            const int t = P.next++;
            lk.unlock(); job(t); lk.lock();
            if (--P.pending == 0) { P.done.notify_all(); }
        }
        P.done.wait(lk, [&]() { return P.pending == 0; });
        P.job = nullptr;
    }
    P.user.unlock();
}

void* hugeAlloc(size_t bytes)
{
    void* p;
    if (bytes >= ((size_t)4 << 20)) {
        const size_t al = (size_t)2 << 20, sz = (bytes + al - 1) & ~(al - 1);
        p = aligned_alloc(al, sz);
        if (p) { madvise(p, sz, MADV_HUGEPAGE); }
    } else { p = malloc(bytes ? bytes : 1); }
    if (!p) { throw std::bad_alloc(); }
    return p;
}

// strictly increasing (sorted, no duplicates): the identity queue of a first pass, checked on all cores
bool Engine::isIncreasing(const uint32_t* v, size_t n)
{
    if (n < std::max<size_t>(tune.par_min, 2) || nthreads < 2) {
        for (size_t i = 1; i < n; ++i) { if (!(v[i - 1] < v[i])) { return false; } }
        return true;
    }
    std::atomic<bool> inc{true};
    parallelFor(n, [&](size_t b, size_t e, int) {
        bool g = true;
        for (size_t i = std::max<size_t>(b, 1); i < e && g; ++i) { g = v[i - 1] < v[i]; }
        if (!g) { inc.store(false, std::memory_order_relaxed); }
    }, 65536);
    return inc.load();
}

void Engine::sortUniqueIds(std::vector<uint32_t>& ids)
{
    if (ids.size() < std::max<size_t>(tune.par_min, 2) || nthreads < 2) { sortUnique(ids); return; }
    if (isIncreasing(ids)) { return; }
    // mark, count per chunk of words, extract in order
    const size_t words = (C.size() + 63) / 64;
    std::vector<uint64_t> bm(words, 0ull);
    parallelFor(ids.size(), [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { __atomic_fetch_or(&bm[ids[i] >> 6], 1ull << (ids[i] & 63), __ATOMIC_RELAXED); } }, 8192);
    std::vector<size_t> cnt((size_t)nthreads + 2, 0);
    parallelFor(words, [&](size_t b, size_t e, int t) { size_t c = 0; for (size_t w = b; w < e; ++w) { c += (size_t)__builtin_popcountll(bm[w]); } cnt[(size_t)t + 1] = c; }, 4096);
    for (size_t t = 1; t < cnt.size(); ++t) { cnt[t] += cnt[t - 1]; }
    ids.resize(cnt.back());
    parallelFor(words, [&](size_t b, size_t e, int t) {
        size_t o = cnt[(size_t)t];
        for (size_t w = b; w < e; ++w) { uint64_t x = bm[w]; while (x) { ids[o++] = (uint32_t)(w * 64 + (size_t)__builtin_ctzll(x)); x &= x - 1; } }
    }, 4096);
}

void Engine::ScanCache::clear()
{
    for (uint32_t c : used) { slot[c] = -1; }
    used.clear(); ent.clear(); hits.clear(); lits.clear(); bulk_all = false;
}

Engine::Engine()
{
    sub_lim = opt.sub_limit; str_lim = opt.str_limit;
    nthreads = tune.threads;
    if (nthreads < 1) { nthreads = 1; }
}
Engine::~Engine() { delete pool_; if (gpu) { cp3g_destroy(gpu); } }

void Engine::die(const char* msg) const
{
    fprintf(stderr, "cp3b200: fatal: %s%s%s\n", msg, gpu ? " - " : "", gpu ? cp3g_last_error(gpu) : "");
    abort();
}

// ------------------------------------------------------------------------------------------------ solver side
void Engine::newVar()
{
    assigns.push_back(L_UNDEF); frozen.push_back(0);
    if (ing_built) { watches.emplace_back(); watches.emplace_back(); }
    ++nvars;
}

uint32_t Engine::allocClause(const int* lits, int n)
{
    uint32_t c = (uint32_t)C.size();
    ClauseInfo ci; ci.start = (uint32_t)pool.size(); ci.size = (uint32_t)n;
    ci.flag = F_SUB | F_STR;                        // fresh clause: can_subsume, can_strengthen
    ci.stamp = scan_id; ci.ver = 0;
    C.push_back(ci);
    if ((c >> 6) >= delbits.size()) { delbits.resize((c >> 6) + 1024, 0); }
    ++n_live;
    pool.insert(pool.end(), lits, lits + n);
    ca_size += 2 + n;                               // ClauseAllocator::clauseWord32Size, SolverTypes.h:641-644
    return c;
}

void Engine::uncheckedEnqueue(int x) { assigns[x >> 1] = (uint8_t)(x & 1); trail.push_back(x); }

// CoprocessorData::enqueue, CoprocessorTypes.h:767-783
int Engine::dataEnqueue(int x)
{
    int v = value(x);
    if (v == L_FALSE) { ok = false; return L_FALSE; }
    if (v == L_UNDEF) { uncheckedEnqueue(x); return L_TRUE; }
    return L_UNDEF;
}

// Solver::attachClause, riss/core/Solver.cc:692-708
void Engine::attachClause(uint32_t c)
{
    const int32_t* l = CL(c); const uint32_t bin = C[c].size == 2 ? 0x80000000u : 0u;
    watches[lneg(l[0])].push_back({c | bin, l[1]});
    watches[lneg(l[1])].push_back({c | bin, l[0]});
}
// Preprocessor::sortClauses (Coprocessor.cc:118,1447-1463): addClause_ sorts every clause, only those propagate() permuted need it
void Engine::sortPermuted()
{
    for (uint32_t c : permuted) { std::sort(CL(c), CL(c) + C[c].size); perm_mark[c] = 0; }
    permuted.clear();
}

// Solver::propagate(true) at level 0 while clauses are added, riss/core/Solver.cc:1809-1937.  The trail order of implied
// literals is a function of the watch-list order, the blocker literals and the way propagate() permutes the clauses it
// visits (c[0]/c[1] swap, new watch moved to c[1]); all of it is restated here, so inputs with unit clauses give the
// reference's trail (and with it the reference's strengthening / BVE order).
bool Engine::ingestPropagate()
{
    if (!ing_built) {
        watches.assign(2 * (size_t)nvars, {});
        ing_built = true;
        for (uint32_t c : clauses) { attachClause(c); }
    }
    auto notePermuted = [&](uint32_t c) {
        if (perm_mark.size() <= c) { perm_mark.resize(C.size() + 1024, 0); }
        if (!perm_mark[c]) { perm_mark[c] = 1; permuted.push_back(c); }
    };
    bool confl = false;
    while (qhead < (int)trail.size()) {
        const int p = trail[qhead++];
        size_t i = 0, j = 0; const size_t end = watches[p].size();
        while (i != end) {
            std::vector<Watcher>& ws = watches[p];             // (pushes below go to other lists; the outer vector does not grow here)
            const Watcher wi = ws[i];
            if (wi.cref & 0x80000000u) {
                const int vi = value(wi.blocker);
                if (vi == L_FALSE) {                            // opt_long_conflict is off: stop at the first conflict
                    while (i < end) { ws[j++] = ws[i++]; }
                    ws.resize(j);
                    return false;
                } else if (vi == L_UNDEF) { uncheckedEnqueue(wi.blocker); }
                ws[j++] = ws[i++];
                continue;
            }
            if (value(wi.blocker) == L_TRUE) { ws[j++] = ws[i++]; continue; }
            const uint32_t c = wi.cref; int32_t* l = CL(c); const uint32_t n = C[c].size;
            const int false_lit = lneg(p);
            if (l[0] == false_lit) { l[0] = l[1]; l[1] = false_lit; notePermuted(c); }
            i++;
            const int first = l[0];
            const Watcher w{c, first};
            if (first != wi.blocker && value(first) == L_TRUE) { ws[j++] = w; continue; }
            bool moved = false;
            for (uint32_t k = 2; k < n; ++k) {
                if (value(l[k]) != L_FALSE) {
                    l[1] = l[k]; l[k] = false_lit; notePermuted(c);
                    watches[lneg(l[1])].push_back(w);          // never watches[p]: l[1] != ~p
                    moved = true; break;
                }
            }
            if (moved) { continue; }
            ws[j++] = w;
            if (value(first) == L_FALSE) {
                confl = true;
                qhead = (int)trail.size();
                while (i < end) { ws[j++] = ws[i++]; }
            } else { uncheckedEnqueue(first); }
        }
        watches[p].resize(j);
    }
    return !confl;
}

// Solver::addClause_, riss/core/Solver.cc:409-499
void Engine::addClauseIngest(std::vector<int>& ps)
{
    if (!ok) { return; }
    std::sort(ps.begin(), ps.end());
    int j = 0, p = -2;
    for (size_t i = 0; i < ps.size(); ++i) {
        int v = value(ps[i]);
        if (v == L_TRUE || ps[i] == lneg(p)) { return; }
        if (v != L_FALSE && ps[i] != p) { ps[j++] = p = ps[i]; }
    }
    if (j == 0) { ok = false; return; }
    if (j == 1) { uncheckedEnqueue(ps[0]); ok = ingestPropagate(); return; }
    uint32_t c = allocClause(ps.data(), j);
    clauses.push_back(c);
    if (ing_built) { attachClause(c); }
}

void Engine::addLit(int x)
{
    if (x == 0) { addClauseIngest(cur); cur.clear(); return; }
    int v = x < 0 ? -x : x;
    while (nvars < v) { newVar(); }
    cur.push_back(2 * (v - 1) + (x < 0));
}
// Bulk ingest.  A long stream into an empty solver goes to the device (K0 cp3g_ingest: per-clause sort, duplicate
// literals, tautologies - Solver::addClause_ without assignments); the arena comes back for the host mirror and stays
// on the device for the preprocessing that follows.  A stream with unit or empty clauses (unit propagation while
// adding is order dependent), short streams and later additions take the literal-by-literal path.
void Engine::addStream(const int32_t* s, size_t n)
{
    const size_t bulk_min = tune.ingest_min;
    if (n >= bulk_min && C.empty() && cur.empty() && trail.empty() && ok && !ing_built && clauses.empty() && cp3g_device_count() > 0) {
        size_t m = n;
        while (m > 0 && s[m - 1] != 0) { --m; }                    // complete clauses only
        {   // ... and only up to the first clause with fewer than two literals: a unit starts unit propagation, whose effect on
            // the clauses behind it is order dependent (two-watched-literal lists) - that part goes literal by literal
            std::vector<size_t> first((size_t)nthreads + 1, m);
            parallelFor(m, [&](size_t b, size_t e, int t) {
                for (size_t i = b; i < e; ++i) {
                    if (s[i] != 0) { continue; }
                    if (i == 0 || s[i - 1] == 0 || i == 1 || s[i - 2] == 0) { first[(size_t)t] = (i == 0 || s[i - 1] == 0) ? i : i - 1; break; }
                }
            }, 1u << 20);
            for (size_t f : first) { m = std::min(m, f); }
            if (m < bulk_min) { m = 0; }
        }
        DbgLap lap("ingest", tune.debug_par);
        lap("first unit");
        if (m) {
            ensureGpu();
            lap("device handle");
            uint32_t nv = 0, nc = 0; size_t nl = 0; int st = 1;
            int64_t t0 = now_ns();
            if (cp3g_ingest(gpu, s, m, &nv, &nc, &nl, &st) != CP3G_OK) { die("bulk ingest failed"); }
            lap("cp3g_ingest");
            if (st == 0) {
                if ((int)nv > nvars) { assigns.resize((size_t)nv, (uint8_t)L_UNDEF); frozen.resize((size_t)nv, 0); nvars = (int)nv; }
                pool.resize(nl); C.resize(nc);
                if (cp3g_download_arena(gpu, pool.data(), C.data()) != CP3G_OK) { die("arena download failed"); }
                lap("arena download");
                clauses.resize(nc);
                parallelFor(nc, [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { clauses[i] = (uint32_t)i; C[i].stamp = scan_id; } });
                lap("clause list");
                n_live = nc; ca_size += 2.0 * (double)nc + (double)nl;
                if (((size_t)nc >> 6) >= delbits.size()) { delbits.resize(((size_t)nc >> 6) + 1024, 0); }
                dev_arena_ncl = nc; dev_arena_valid = true; ++n_bulk_ingests;
                s += m; n -= m;
            }
            host_ns_gpuwait += now_ns() - t0;
        }
    }
    for (size_t i = 0; i < n; ++i) { addLit(s[i]); }
}
void Engine::freeze(int dv) { if (dv <= 0) { return; } while (nvars < dv) { newVar(); } frozen[dv - 1] = 1; }   // freezeExtern(0) is a no-op
int Engine::litValue(int d) const
{
    int v = d < 0 ? -d : d;
    if (v > nvars) { return L_UNDEF; }
    return value(2 * (v - 1) + (d < 0));
}

// Solver::simplify at level 0, riss/core/Solver.cc:2212-2268
bool Engine::solverSimplify()
{
    if (!ok) { return false; }
    if (qhead < (int)trail.size() && !ingestPropagate()) { ok = false; return false; }
    if ((int)trail.size() == simpDB_assigns || simpDB_props > 0) { return true; }
    size_t j = 0; int64_t lits = 0;
    if (trail.empty()) {                          // nothing is assigned: no clause can be satisfied, only the literal count is needed
        std::vector<int64_t> part((size_t)nthreads + 1, 0);
        parallelFor(clauses.size(), [&](size_t b, size_t e, int t) { int64_t a = 0; for (size_t i = b; i < e; ++i) { a += C[clauses[i]].size; } part[(size_t)t] = a; });
        for (int64_t a : part) { lits += a; }
        if (ca_wasted > ca_size * opt.gc_frac) { ca_size -= ca_wasted; ca_wasted = 0; }
        simpDB_assigns = 0; simpDB_props = lits;
        return true;
    }
    for (size_t i = 0; i < clauses.size(); ++i) {
        uint32_t c = clauses[i]; const int32_t* l = CL(c); bool sat = false;
        for (uint32_t k = 0; k < C[c].size; ++k) { if (value(l[k]) == L_TRUE) { sat = true; break; } }
        if (sat) { caFree(c); C[c].flag |= F_DEL; delbits[c >> 6] |= 1ull << (c & 63); --n_live; }
        else { clauses[j++] = c; lits += C[c].size; }
    }
    clauses.resize(j);
    if (ca_wasted > ca_size * opt.gc_frac) { ca_size -= ca_wasted; ca_wasted = 0; }
    simpDB_assigns = (int)trail.size(); simpDB_props = lits;
    return true;
}

// ------------------------------------------------------------------------------------------------ device
void Engine::ensureGpu()
{
    if (gpu) { return; }
    gpu = cp3g_create(opt.device);
    if (!gpu) { die("no CUDA device for the Coprocessor GPU path (there is no CPU fallback)"); }
}

int Engine::setDistributed(int rank, int world, const void* id)
{
    ensureGpu();
    sharded = world > 1;
    return cp3g_nccl_init(gpu, rank, world, id);
}

// upload the flattened arena and let the GPU build signatures + occurrence CSR (K1, K2)
void Engine::uploadAll()
{
    ensureGpu();
    // occurrence entries carry the clause index in 30 bits (two class bits share the word): the reference's CRef reaches 2^31
    // words, this path stops earlier - and says so before anything is uploaded
    if (C.size() > 0x3fffffffull || (size_t)nvars >= (1ull << 30) || pool.size() >= 0x7fffffffull) {
        die("data base too large for the GPU path: more than 2^30 clauses / variables or 2^31 literals (30-bit clause references)");
    }
    const uint32_t ncl = (uint32_t)C.size();
    static_assert(sizeof(ClauseInfo) == 16, "ClauseInfo is uploaded as a packed 16-byte record");
    int64_t t0 = now_ns();
    // the arena is still on the device from the bulk ingest unless something was added, deleted or assigned since
    const bool resident = dev_arena_valid && dev_arena_ncl == ncl && n_live == ncl && trail.empty() &&
                          cp3g_counter(gpu, "nclauses") == (int64_t)ncl && cp3g_counter(gpu, "nlits") == (int64_t)pool.size() &&
                          cp3g_counter(gpu, "nvars") == (int64_t)nvars;
    dev_arena_valid = false;
    if (!resident && cp3g_upload_packed(gpu, (uint32_t)nvars, ncl, pool.data(), pool.size(), C.data()) != CP3G_OK) { die("upload failed"); }
    if (cp3g_build(gpu) != CP3G_OK) { die("occurrence build failed"); }
    host_ns_gpuwait += now_ns() - t0;
    dev_ncl = ncl; dev_uploaded = true; dev_static_ok = false;
    pend_del.clear(); pend_shrunk.clear();
    pfill(shrunk_mark, (size_t)ncl, (uint8_t)0);
}

// push the edits committed since the last scan (new resolvents, shrunk clauses, deletions)
void Engine::flushDevice()
{
    const uint32_t ncl = (uint32_t)C.size();
    int64_t t0 = now_ns();
    const uint32_t old_ncl = dev_ncl;
    if (ncl > dev_ncl || !pend_shrunk.empty() || !pend_del.empty()) { dev_static_ok = false; }      // the device tables of the last pass commit are history
    if (ncl > dev_ncl) {
        std::vector<uint32_t> sz; std::vector<int32_t> li;
        for (uint32_t c = dev_ncl; c < ncl; ++c) {
            sz.push_back(C[c].size);
            li.insert(li.end(), CL(c), CL(c) + C[c].size);
            if (C[c].flag & F_DEL) { pend_del.push_back(c); }
        }
        if (cp3g_append(gpu, ncl - dev_ncl, sz.data(), li.data(), li.size()) != CP3G_OK) { die("append failed"); }
        dev_ncl = ncl;
        shrunk_mark.resize(ncl, 0);
    }
    if (!pend_shrunk.empty()) {
        std::vector<uint32_t> ids, st, sz; std::vector<int32_t> li;
        if (pend_shrunk.size() >= std::max<size_t>(tune.par_min, 1) && nthreads > 1) {
            // long edit lists (a strengthening pass over a large formula): sizes per chunk, prefix, fill - on all cores
            const size_t np = pend_shrunk.size();
            std::vector<size_t> cn((size_t)nthreads + 2, 0), cl((size_t)nthreads + 2, 0);
            parallelFor(np, [&](size_t b, size_t e, int t) {
                size_t a = 0, l = 0;
                for (size_t i = b; i < e; ++i) { const uint32_t c = pend_shrunk[i]; shrunk_mark[c] = 0; if (c < old_ncl) { ++a; l += C[c].size; } }
                cn[(size_t)t + 1] = a; cl[(size_t)t + 1] = l;
            }, 4096);
            for (size_t t = 1; t < cn.size(); ++t) { cn[t] += cn[t - 1]; cl[t] += cl[t - 1]; }
            ids.resize(cn.back()); st.resize(cn.back()); sz.resize(cn.back()); li.resize(cl.back());
            parallelFor(np, [&](size_t b, size_t e, int t) {
                size_t a = cn[(size_t)t], l = cl[(size_t)t];
                for (size_t i = b; i < e; ++i) {
                    const uint32_t c = pend_shrunk[i];
                    if (c >= old_ncl) { continue; }
                    ids[a] = c; st[a] = (uint32_t)l; sz[a] = C[c].size; ++a;
                    memcpy(li.data() + l, CL(c), sizeof(int32_t) * C[c].size); l += C[c].size;
                }
            }, 4096);
        } else {
        for (uint32_t c : pend_shrunk) {
            shrunk_mark[c] = 0;
            if (c >= old_ncl) { continue; }          // appended above with its current content
            ids.push_back(c); st.push_back((uint32_t)li.size()); sz.push_back(C[c].size);
            li.insert(li.end(), CL(c), CL(c) + C[c].size);
        }
        }
        pend_shrunk.clear();
        if (!ids.empty() && cp3g_shrink(gpu, (uint32_t)ids.size(), ids.data(), st.data(), sz.data(), li.data(), li.size()) != CP3G_OK) { die("shrink failed"); }
    }
    if (!pend_del.empty()) {
        if (cp3g_set_deleted(gpu, (uint32_t)pend_del.size(), pend_del.data()) != CP3G_OK) { die("set_deleted failed"); }
        pend_del.clear();
    }
    host_ns_gpuwait += now_ns() - t0;
}

// One GPU scan over the clauses `ids`; results land in `cache` keyed by clause and content version.
void Engine::scanClauses(int kind, const std::vector<uint32_t>& ids, ScanCache& cache, int forceBulk)
{
    if (ids.empty()) { return; }
    DbgLap lap(kind == CP3G_SUBSUME ? "scanClauses<sub>" : "scanClauses<str>", tune.debug_par);
    if (!scan_no_flush) { flushDevice(); }
    lap("flush");
    ++scan_id; ++n_scan_batches;
    pgrow(cache.slot, C.size(), -1);
    // deleted clauses that still act as strengthener (Subsumption.cc:1438) must be sent by content
    const uint32_t bulk_min = tune.bulk_min;
    bool bulk = forceBulk == 1;
    if (forceBulk < 0 && ids.size() == n_live && n_live > bulk_min) {
        bool any_del = false;
        parallelFor(ids.size(), [&](size_t b, size_t e, int) { bool d = false; for (size_t i = b; i < e && !d; ++i) { d = (C[ids[i]].flag & F_DEL) != 0; } if (d) { any_del = true; } });
        bulk = !any_del;
    }
    std::vector<uint32_t> live, dead;
    if (!bulk) { for (uint32_t c : ids) { ((C[c].flag & F_DEL) ? dead : live).push_back(c); } }
    lap("split");
    int64_t t0 = now_ns();
    if (bulk) {
        // the queue holds every live clause: full-queue bulk kernel (one thread per occurrence entry)
        uint32_t cap = (uint32_t)std::max<size_t>(hitbuf.size(), std::max<size_t>(4096, ids.size() / 8));
        uint32_t n = 0; uint64_t chk = 0;
        for (;;) {
            hitbuf.resize(cap);
            int r = cp3g_scan_all(gpu, kind, hitbuf.data(), cap, &n, &chk);
            if (r == CP3G_OK) { break; }
            if (r == CP3G_ERR_OVERFLOW) { cap = n + n / 8 + 1024; continue; }
            die("bulk scan failed");
        }
        lap("scan_all");
        sortHits(hitbuf, n);
        lap("sort hits");
        cache.bulk_all = true; cache.bulk_scan = scan_id; cache.bulk_ncl = (uint32_t)C.size();
        cache.empty.ver = 0; cache.empty.scan = scan_id; cache.empty.hb = cache.empty.he = 0; cache.empty.next = -1;
        cache.empty.lit_off = cache.empty.lit_n = 0; cache.empty.time = 0;
        if (n >= std::max<size_t>(tune.par_min, 1) && nthreads > 1 && cache.hits.empty() && cache.ent.empty()) {
            // one record per run of equal `req` in the sorted hit list, built on all cores: run starts counted per chunk, prefix, fill
            cache.hits.resize(n);
            std::vector<size_t> cnt((size_t)nthreads + 2, 0);
            parallelFor(n, [&](size_t b, size_t e, int t) {
                size_t c2 = 0;
                for (size_t h = b; h < e; ++h) { cache.hits[h] = Hit{hitbuf[h].clause, hitbuf[h].lit}; c2 += (h == 0 || hitbuf[h].req != hitbuf[h - 1].req) ? 1 : 0; }
                cnt[(size_t)t + 1] = c2;
            }, 8192);
            for (size_t t = 1; t < cnt.size(); ++t) { cnt[t] += cnt[t - 1]; }
            const size_t ne = cnt.back();
            cache.ent.resize(ne); const size_t used0 = cache.used.size(); cache.used.resize(used0 + ne);
            parallelFor(n, [&](size_t b, size_t e, int t) {
                size_t k = cnt[(size_t)t];
                for (size_t h = b; h < e; ++h) {
                    if (!(h == 0 || hitbuf[h].req != hitbuf[h - 1].req)) { continue; }
                    const uint32_t c = hitbuf[h].req;
                    size_t he = h + 1; while (he < n && hitbuf[he].req == c) { ++he; }
                    CacheEnt& en = cache.ent[k];
                    en.ver = C[c].ver; en.scan = scan_id; en.hb = (uint32_t)h; en.he = (uint32_t)he; en.next = -1; en.lit_off = en.lit_n = 0; en.time = 0;
                    cache.slot[c] = (int32_t)k; cache.used[used0 + k] = c;           // (the cache was empty: no slot of this scan is set yet)
                    ++k;
                }
            }, 8192);
        } else {
        for (uint32_t h = 0; h < n;) {
            const uint32_t c = hitbuf[h].req;
            CacheEnt e; e.ver = C[c].ver; e.scan = scan_id; e.hb = (uint32_t)cache.hits.size();
            while (h < n && hitbuf[h].req == c) { cache.hits.push_back({hitbuf[h].clause, hitbuf[h].lit}); ++h; }
            e.he = (uint32_t)cache.hits.size();
            e.next = cache.slot[c]; e.lit_off = e.lit_n = 0; e.time = 0;
            cache.setSlot(c, (int32_t)cache.ent.size());
            cache.ent.push_back(e);
        }
        }
        lap("cache");
        host_ns_gpuwait += now_ns() - t0;
        return;
    }
    for (int part = 0; part < 2; ++part) {
        const std::vector<uint32_t>& v = part == 0 ? live : dead;
        if (v.empty()) { continue; }
        std::vector<uint32_t> roff; std::vector<int32_t> rl;
        if (part == 1) {
            roff.push_back(0);
            for (uint32_t c : v) { rl.insert(rl.end(), CL(c), CL(c) + C[c].size); roff.push_back((uint32_t)rl.size()); }
        }
        uint32_t cap = (uint32_t)std::max<size_t>(hitbuf.size(), std::max<size_t>(4096, v.size() / 8));
        uint32_t n = 0; uint64_t chk = 0;
        for (;;) {
            hitbuf.resize(cap);
            int r = cp3g_scan(gpu, kind, (uint32_t)v.size(), v.data(), part ? roff.data() : nullptr, part ? rl.data() : nullptr,
                              hitbuf.data(), cap, &n, &chk);
            if (r == CP3G_OK) { break; }
            if (r == CP3G_ERR_OVERFLOW) { cap = n + n / 8 + 1024; continue; }
            die("scan failed");
        }
        sortHits(hitbuf, n);
        uint32_t h = 0;
        for (uint32_t i = 0; i < v.size(); ++i) {
            CacheEnt e; e.ver = C[v[i]].ver; e.scan = scan_id; e.hb = (uint32_t)cache.hits.size();
            while (h < n && hitbuf[h].req == i) { cache.hits.push_back({hitbuf[h].clause, hitbuf[h].lit}); ++h; }
            e.he = (uint32_t)cache.hits.size();
            e.next = cache.slot[v[i]]; e.lit_off = e.lit_n = 0; e.time = 0;
            cache.setSlot(v[i], (int32_t)cache.ent.size());
            cache.ent.push_back(e);
        }
    }
    lap("request scan");
    host_ns_gpuwait += now_ns() - t0;
}

const Engine::CacheEnt* Engine::cached(const ScanCache& cache, uint32_t c) const
{
    if (c >= cache.slot.size()) { return nullptr; }
    for (int32_t i = cache.slot[c]; i >= 0; i = cache.ent[i].next) {
        const CacheEnt* e = &cache.ent[i];
        if (e->ver != UINT32_MAX) { if (e->ver == C[c].ver) { return e; } continue; }
        if (e->lit_n == C[c].size && !memcmp(cache.lits.data() + e->lit_off, CL(c), sizeof(int32_t) * e->lit_n)) { return e; }
    }
    // no record: a clause that is unchanged since the full-queue scan had no hit there
    return (cache.bulk_all && c < cache.bulk_ncl && C[c].stamp < cache.bulk_scan) ? &cache.empty : nullptr;
}

// Scan explicit literal contents (predicted future versions of clauses) against the current device state.
void Engine::scanSpeculative(int kind, const std::vector<SpecReq>& reqs, ScanCache& cache, bool grouped)
{
    if (reqs.empty()) { return; }
    ++scan_id; ++n_scan_batches; ++n_spec_rounds; n_spec_requests += (int64_t)reqs.size();
    pgrow(cache.slot, C.size(), -1);
    const size_t nr = reqs.size();
    std::vector<uint32_t> self(nr), roff(nr + 1);
    roff[0] = 0;
    for (size_t i = 0; i < nr; ++i) { roff[i + 1] = roff[i] + reqs[i].n; }
    RawVec<int32_t> rl; rl.resize((size_t)roff[nr] + 1);
    parallelFor(nr, [&](size_t b2, size_t e2, int) {
        for (size_t i = b2; i < e2; ++i) { self[i] = reqs[i].clause; memcpy(rl.data() + roff[i], spec_lits.data() + reqs[i].off, sizeof(int32_t) * reqs[i].n); }
    }, 8192);
    int64_t t0 = now_ns();
    uint32_t cap = (uint32_t)std::max<size_t>(hitbuf.size(), std::max<size_t>(4096, nr / 4));
    uint32_t n = 0; uint64_t chk = 0;
    for (;;) {
        hitbuf.resize(cap);
        int r = cp3g_scan(gpu, kind, (uint32_t)nr, self.data(), roff.data(), rl.data(), hitbuf.data(), cap, &n, &chk);
        if (r == CP3G_OK) { break; }
        if (r == CP3G_ERR_OVERFLOW) { cap = n + n / 8 + 1024; continue; }
        die("speculative scan failed");
    }
    host_ns_gpuwait += now_ns() - t0;
    if (tune.debug_par) { fprintf(stderr, "  scanSpeculative gpu call %.1f ms, %zu reqs, %u hits\n", (now_ns() - t0) / 1e6, reqs.size(), n); }
    sortHits(hitbuf, n);
    // one record per request, in request order: hit ranges by counting, literals as one block, fields on all cores
    const size_t e0 = cache.ent.size(), l0 = cache.lits.size(), h0 = cache.hits.size();
    std::vector<uint32_t> hstart(nr + 1, 0u);
    for (uint32_t k = 0; k < n; ++k) { ++hstart[(size_t)hitbuf[k].req + 1]; }
    for (size_t i = 0; i < nr; ++i) { hstart[i + 1] += hstart[i]; }
    cache.hits.resize(h0 + n);
    for (uint32_t k = 0; k < n; ++k) { cache.hits[h0 + k] = Hit{hitbuf[k].clause, hitbuf[k].lit}; }
    cache.lits.resize(l0 + roff[nr]);
    cache.ent.resize(e0 + nr);
    parallelFor(nr, [&](size_t b2, size_t e2, int) {
        memcpy(cache.lits.data() + l0 + roff[b2], rl.data() + roff[b2], sizeof(int32_t) * (size_t)(roff[e2] - roff[b2]));
        for (size_t i = b2; i < e2; ++i) {
            CacheEnt& en = cache.ent[e0 + i];
            en.ver = UINT32_MAX; en.scan = scan_id; en.hb = (uint32_t)(h0 + hstart[i]); en.he = (uint32_t)(h0 + hstart[i + 1]);
            en.lit_off = (uint32_t)(l0 + roff[i]); en.lit_n = reqs[i].n; en.time = reqs[i].time; en.next = -1;
        }
    }, 8192);
    // chains: the newest record of a clause first
    if (grouped && nr >= std::max<size_t>(tune.par_min, 1) && nthreads > 1) {
        std::vector<std::vector<uint32_t>> fresh((size_t)nthreads + 1);
        parallelFor(nr, [&](size_t b2, size_t e2, int t) {
            for (size_t i = b2; i < e2; ++i) {
                if (i > 0 && reqs[i - 1].clause == reqs[i].clause) { continue; }          // a run belongs to the chunk it starts in
                const uint32_t c = reqs[i].clause;
                int32_t prev = cache.slot[c];
                if (prev < 0) { fresh[(size_t)t].push_back(c); }
                size_t j = i;
                for (; j < nr && reqs[j].clause == c; ++j) { cache.ent[e0 + j].next = prev; prev = (int32_t)(e0 + j); }
                cache.slot[c] = prev;
            }
        }, 8192);
        for (const auto& v : fresh) { cache.used.insert(cache.used.end(), v.begin(), v.end()); }
    } else {
        for (size_t i = 0; i < nr; ++i) { cache.ent[e0 + i].next = cache.slot[reqs[i].clause]; cache.setSlot(reqs[i].clause, (int32_t)(e0 + i)); }
    }
}

// Speculative closure of a strengthening pass.  The bulk K4 scan of the queue tells which clauses will lose
// which literal; a clause that loses a literal is processed again as a strengthener in its shortened form
// (at its own queue turn, or right away through the LIFO localQueue, Subsumption.cc:1280-1287,1394-1397).
// Instead of discovering those versions one by one during the ordered commit (one GPU round trip each),
// predict them from the hit lists, scan them in bulk, and repeat on the hits they produce.  A wrong
// prediction costs nothing but an unused cache entry; a missing one falls back to an on-demand scan.
void Engine::prefetchStrengthenClosure()
{
    PhaseTimer pt(ph_prefetch);
    struct Ev { uint32_t clause; int32_t lit; double time; };
    // hash of (clause, content) already scanned: open addressing (the node-based set cost 40 % of the request pass).
    // One set per worker; a clause always belongs to the same worker (fixed clause-index ranges), so no set is shared.
    struct HashSet {
        std::vector<uint64_t> t; size_t n = 0;
        HashSet() : t(1u << 12, 0ull) {}
        bool insert(uint64_t h) {                           // true if new
            if (h == 0) { h = 0x9E3779B97F4A7C15ull; }
            if (2 * (n + 1) > t.size()) {
                std::vector<uint64_t> o; o.swap(t); t.assign(o.size() * 4, 0ull); n = 0;
                for (uint64_t x : o) { if (x) { insert(x); } }
            }
            size_t i = (size_t)(h * 0x9E3779B97F4A7C15ull >> 20) & (t.size() - 1);
            while (t[i] && t[i] != h) { i = (i + 1) & (t.size() - 1); }
            if (t[i] == h) { return false; }
            t[i] = h; ++n; return true;
        }
    };
    const int W = std::max(1, nthreads);
    const uint32_t ncl = (uint32_t)C.size();
    const uint32_t width = ncl / (uint32_t)W + 1;               // worker w owns the clauses [w*width, (w+1)*width)
    std::vector<HashSet> requested((size_t)W);
    std::vector<std::vector<Ev>> evs((size_t)W);                 // per worker: predicted removals of its clauses, sorted by (clause, time)
    size_t first_ent = 0;
    DbgLap lap("closure", tune.debug_par);
    // Speculation is only worth its scans while they are cheap next to the full-queue pass: a request costs the
    // length of its shortest list pair (long on skewed inputs), and after a unit shows up most predictions are void.
    // Budget: four times the occurrence count of the data base; what is not prefetched is scanned on demand.
    int64_t spec_budget = 4 * (int64_t)pool.size();
    bool budget_left = true;
    struct Req { uint32_t clause, off, n, walk; double time; };
    for (int round = 0; round < 8 && budget_left; ++round) {
        // (1) the hits of the cache entries that are new since the last round become events of the target's owner
        const size_t ne = c_str.ent.size();
        if (ne == first_ent) { break; }
        std::vector<std::vector<Ev>> fresh((size_t)W * (size_t)W);          // [producer][owner]
        parallelFor(ne - first_ent, [&](size_t b2, size_t e2, int t) {
            std::vector<Ev>* out = &fresh[(size_t)t * (size_t)W];
            for (size_t ei = first_ent + b2; ei < first_ent + e2; ++ei) {
                const CacheEnt& e = c_str.ent[ei];
                for (uint32_t h = e.hb; h < e.he; ++h) { const Hit& x = c_str.hits[h]; out[x.clause / width].push_back({x.clause, x.lit, e.time}); }
            }
        }, 4096);
        first_ent = ne;
        size_t nfresh = 0; for (const auto& v : fresh) { nfresh += v.size(); }
        if (nfresh == 0) { break; }
        // (2)+(3) per owner: merge, keep the earliest prediction per (clause, literal), order by (clause, time), then build the
        //         requests of the clauses that got a new event: current content minus the first k predicted removals
        std::vector<std::vector<Req>> wreq((size_t)W); std::vector<std::vector<int32_t>> wlits((size_t)W);
        parallelFor((size_t)W, [&](size_t wb, size_t we, int) {
            for (size_t w = wb; w < we; ++w) {
                std::vector<Ev>& ev = evs[w];
                std::vector<uint32_t> changed;
                for (int t = 0; t < W; ++t) { const std::vector<Ev>& f = fresh[(size_t)t * (size_t)W + w]; for (const Ev& x : f) { changed.push_back(x.clause); } ev.insert(ev.end(), f.begin(), f.end()); }
                if (changed.empty()) { continue; }
                std::sort(ev.begin(), ev.end(), [](const Ev& a2, const Ev& b2) {
                    return a2.clause != b2.clause ? a2.clause < b2.clause : (a2.lit != b2.lit ? a2.lit < b2.lit : a2.time < b2.time); });
                ev.erase(std::unique(ev.begin(), ev.end(), [](const Ev& a2, const Ev& b2) { return a2.clause == b2.clause && a2.lit == b2.lit; }), ev.end());
                std::sort(ev.begin(), ev.end(), [](const Ev& a2, const Ev& b2) { return a2.clause != b2.clause ? a2.clause < b2.clause : a2.time < b2.time; });
                std::sort(changed.begin(), changed.end());
                changed.erase(std::unique(changed.begin(), changed.end()), changed.end());
                std::vector<Req>& reqs = wreq[w]; std::vector<int32_t>& sl = wlits[w]; HashSet& seen = requested[w];
                size_t p = 0;
                for (uint32_t d : changed) {
                    while (p < ev.size() && ev[p].clause < d) { ++p; }
                    size_t q = p;
                    while (q < ev.size() && ev[q].clause == d) { ++q; }
                    const Ev* e1 = ev.data() + p; const size_t m = q - p;
                    p = q;
                    if ((C[d].flag & F_DEL) || m == 0) { continue; }
                    const bool pending = (C[d].flag & F_STR) && d < qtime.size() && qtime[d] >= 0;
                    const double own = pending ? (double)qtime[d] : -1.0;
                    size_t k0 = 0;
                    if (pending) { while (k0 < m && e1[k0].time < own) { ++k0; } }
                    for (size_t k = std::max<size_t>(k0, 1); k <= m && k <= 6; ++k) {
                        if (C[d].size < k + 2) { break; }          // would be a unit: handled by propagation, never scanned
                        // content = current literals minus the first k predicted removals
                        uint64_t hsh = 1469598103934665603ull ^ d;
                        const uint32_t off = (uint32_t)sl.size();
                        const int32_t* l = CL(d);
                        for (uint32_t i = 0; i < C[d].size; ++i) {
                            bool rm = false;
                            for (size_t j = 0; j < k; ++j) { if (e1[j].lit == l[i]) { rm = true; break; } }
                            if (!rm) { sl.push_back(l[i]); hsh = (hsh ^ (uint32_t)l[i]) * 1099511628211ull; }
                        }
                        const uint32_t nl = (uint32_t)sl.size() - off;
                        if (nl != C[d].size - k || !seen.insert(hsh)) { sl.resize(off); continue; }
                        const double t = (pending && k == k0) ? own : std::max(own, e1[k - 1].time) + 1e-3 * (round + 1);
                        uint32_t walk = 0xffffffffu;
                        for (uint32_t i = 0; i < nl; ++i) { const int x = sl[off + i]; walk = std::min(walk, hot[x].n + hot[x ^ 1].n); }
                        reqs.push_back({d, off, nl, walk, t});
                    }
                }
            }
        }, 1);
        lap("events+requests");
        // (4) one request list in clause order, cut where the budget ends
        std::vector<SpecReq> reqs; spec_lits.clear();
        for (int w = 0; w < W && budget_left; ++w) {
            const uint32_t base = (uint32_t)spec_lits.size(); size_t used = 0;
            for (const Req& r : wreq[(size_t)w]) {
                spec_budget -= r.walk;
                if (spec_budget < 0) { budget_left = false; break; }
                reqs.push_back({r.clause, base + r.off, r.n, r.time}); used = (size_t)r.off + r.n;
            }
            spec_lits.insert(spec_lits.end(), wlits[(size_t)w].begin(), wlits[(size_t)w].begin() + (std::ptrdiff_t)used);
        }
        lap("merge");
        if (reqs.empty()) { break; }
        scanSpeculative(CP3G_STRENGTHEN, reqs, c_str, true);
        lap("scanSpeculative");
    }
}


void Engine::logClear()
{
    for (uint32_t c : log_touched) { log_head[c] = -1; }
    log_touched.clear(); log_nodes.clear();
}
void Engine::logPush(uint32_t c, int lit, bool pendingOcc)
{
    if (c >= log_head.size()) { log_head.resize(C.size() + 1024, -1); }
    if (log_head[c] < 0) { log_touched.push_back(c); }
    log_nodes.push_back({scan_id, lit, log_head[c], pendingOcc ? pend_epoch : 0u});
    log_head[c] = (int32_t)log_nodes.size() - 1;
}

// A snapshot hit "C subsumes D" stays true unless D lost a literal of C afterwards.
bool Engine::subHitValid(uint32_t c, uint32_t d, uint32_t scan) const
{
    if (C[d].stamp < scan || d >= log_head.size()) { return true; }
    for (int32_t i = log_head[d]; i >= 0; i = log_nodes[i].next) {
        if (log_nodes[i].stamp >= scan && containsLit(CL(c), C[c].size, log_nodes[i].lit)) { return false; }
    }
    return true;
}
// A snapshot hit "S strengthens D on literal f" stays true unless D lost f or a literal of S afterwards.
bool Engine::strHitValid(uint32_t s, uint32_t d, int f, uint32_t scan) const
{
    if (C[d].stamp < scan || d >= log_head.size()) { return true; }
    for (int32_t i = log_head[d]; i >= 0; i = log_nodes[i].next) {
        if (log_nodes[i].stamp < scan) { continue; }
        if (log_nodes[i].lit == f || containsLit(CL(s), C[s].size, log_nodes[i].lit)) { return false; }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ occ lists
void Engine::occPush(int x, uint32_t c)
{
    ListRef l = L(x);
    if (l.n == l.cap) {
        uint32_t ncap = l.cap ? l.cap * 2 : 4;
        uint32_t noff = (uint32_t)occ_pool.size();
        occ_pool.resize(occ_pool.size() + ncap);
        if (l.n) { memcpy(occ_pool.data() + noff, occ_pool.data() + l.off, sizeof(uint32_t) * l.n); }
        l.off = noff; l.cap = ncap;
    }
    occ_pool[l.off + l.n++] = c;
}
void Engine::occSwapRemove(int x, uint32_t i)
{
    ListRef l = L(x); uint32_t* p = LP(l);
    if (isDel(p[i])) { staleDec(x); }
    p[i] = p[l.n - 1]; l.n--;
    // a walk that lazily cleans a list reorders it without changing any clause of the variable: its K5' record (evaluated on
    // the old order) is void, let it ride along with the next batch instead of being found stale when the variable is popped
    if (in_bve_worker) { bveTouch(x >> 1); }
}
void Engine::occRelease(int x) { hot[x].n = 0; staleClear(x); }

// the variables of many deleted clauses at once: while the BVE heap is empty (the passes in front of BVE) touching a variable is
// just a counter increment - done on all cores; with a heap the pending list is filled in order
void Engine::touchDeadClauses(const std::vector<uint32_t>& dead)
{
    if (!opt.bve) { return; }
    if (!heap.empty() || dead.size() < 4096) { for (uint32_t d : dead) { touchClauseVars(d); } return; }
    parallelFor(dead.size(), [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) { const uint32_t d = dead[i]; const int32_t* l = CL(d); for (uint32_t k = 0; k < C[d].size; ++k) { __atomic_fetch_add(&var_touch[(size_t)(l[k] >> 1)], 1u, __ATOMIC_RELAXED); } }
    }, 8192);
}
void Engine::touchClauseVars(uint32_t c)
{
    if (!opt.bve) { return; }                       // only the BVE pair cache looks at var_touch
    const int32_t* l = CL(c);
    for (uint32_t k = 0; k < C[c].size; ++k) { bveTouch(l[k] >> 1); }
}

void Engine::setDeleted(uint32_t c)
{
    if (C[c].flag & F_DEL) { return; }
    C[c].flag |= F_DEL; delbits[c >> 6] |= 1ull << (c & 63); --n_live;
    const int32_t* l = CL(c);
    for (uint32_t k = 0; k < C[c].size; ++k) { staleInc(l[k]); }
    touchClauseVars(c);
    // literals removed by strengthening whose occurrence entry is still pending (Subsumption.cc:1400,1523)
    if (c < log_head.size()) {
        for (int32_t i = log_head[c]; i >= 0; i = log_nodes[i].next) { if (log_nodes[i].pending == pend_epoch) { staleInc(log_nodes[i].lit); } }
    }
    if (dev_uploaded) { pend_del.push_back(c); }
}

void Engine::noteShrunk(uint32_t c, int removedLit, bool pendingOcc)
{
    logPush(c, removedLit, pendingOcc);
    ++C[c].ver; C[c].stamp = scan_id;
    if (dev_uploaded && c < shrunk_mark.size() && !shrunk_mark[c]) { shrunk_mark[c] = 1; pend_shrunk.push_back(c); }
    else if (dev_uploaded && c >= shrunk_mark.size()) { /* not yet on the device: appended with current content */ }
    touchClauseVars(c); if (opt.bve) { bveTouch(removedLit >> 1); }
}

void Engine::removePosSorted(uint32_t c, int pos)
{
    int32_t* l = CL(c); uint32_t n = C[c].size;
    for (uint32_t j = (uint32_t)pos; j + 1 < n; ++j) { l[j] = l[j + 1]; }
    C[c].size = n - 1;
}
void Engine::removePosUnsorted(uint32_t c, int pos)
{
    int32_t* l = CL(c); uint32_t n = C[c].size;
    l[pos] = l[n - 1]; C[c].size = n - 1;
}

// ------------------------------------------------------------------------------------------------ heap (riss/mtl/Heap.h:37-183)
void Engine::heapUp(int i)
{
    int x = heap[i], p = (i - 1) >> 1;
    while (i != 0 && heapLt(x, heap[p])) { heap[i] = heap[p]; hidx[heap[p]] = i; i = p; p = (p - 1) >> 1; }
    heap[i] = x; hidx[x] = i;
}
void Engine::heapDown(int i)
{
    int x = heap[i]; const int n = (int)heap.size();
    while (2 * i + 1 < n) {
        int l = 2 * i + 1, r = 2 * i + 2;
        int child = (r < n && heapLt(heap[r], heap[l])) ? r : l;
        if (!heapLt(heap[child], x)) { break; }
        heap[i] = heap[child]; hidx[heap[i]] = i; i = child;
    }
    heap[i] = x; hidx[x] = i;
}
void Engine::heapInsert(int v)
{
    if (v >= (int)hidx.size()) { hidx.resize(v + 1, -1); }
    hidx[v] = (int)heap.size(); heap.push_back(v);
    heapUp(hidx[v]);
}
void Engine::heapUpdate(int v) { if (!inHeap(v)) { heapInsert(v); } else { heapUp(hidx[v]); heapDown(hidx[v]); } }
int Engine::heapRemoveMin()
{
    int x = heap[0];
    heap[0] = heap.back(); hidx[heap[0]] = 0; hidx[x] = -1; heap.pop_back();
    if (heap.size() > 1) { heapDown(0); }
    return x;
}
void Engine::heapClear() { for (int v : heap) { hidx[v] = -1; } heap.clear(); }

// ------------------------------------------------------------------------------------------------ data bookkeeping
// removedClause(cr, heap, update, ignore), CoprocessorTypes.h:1310-1350
void Engine::removedClause(uint32_t c, bool useHeap, bool update, int ignore)
{
    const int32_t* l = CL(c);
    for (uint32_t i = 0; i < C[c].size; ++i) {
        int v = l[i] >> 1;
        deletedVar(v);
        --hot[l[i]].cnt;
        if (useHeap) {
            if (inHeap(v)) { heapUp(hidx[v]); }
            else if (update && v != ignore) { heapUpdate(v); }
        }
    }
    numberOfCls--;
    caFree(c);
}
// removedLiteral(l, diff, heap, update, ignore), CoprocessorTypes.h:1243-1277
void Engine::removedLiteral(int x, int diff, bool useHeap, bool update, int ignore)
{
    int v = x >> 1;
    deletedVar(v);
    hot[x].cnt -= diff;
    ca_wasted += 1;
    if (useHeap) {
        if (inHeap(v)) { heapUp(hidx[v]); }
        else if (update && v != ignore) { heapUpdate(v); }
    }
}
// addClause(cr, heap), CoprocessorTypes.h:835-864
void Engine::dataAddClause(uint32_t c, bool useHeap)
{
    if (C[c].flag & F_DEL) { return; }
    const int32_t* l = CL(c);
    for (uint32_t i = 0; i < C[c].size; ++i) {
        occPush(l[i], c);
        hot[l[i]].cnt += 1;
        bveTouch(l[i] >> 1);
        if (useHeap && inHeap(l[i] >> 1)) { heapDown(hidx[l[i] >> 1]); }
    }
    numberOfCls++;
}
// removeClauseFrom, CoprocessorTypes.h:867-878
bool Engine::removeClauseFrom(uint32_t c, int x)
{
    ListRef l = L(x); uint32_t* p = LP(l);
    for (uint32_t i = 0; i < l.n; ++i) { if (p[i] == c) { occSwapRemove(x, i); return true; } }
    return false;
}
void Engine::subInitClause(uint32_t c)
{
    if (C[c].flag & F_DEL) { return; }
    if (C[c].flag & F_SUB) { subq.push_back(c); }
    if (C[c].flag & F_STR) { strq.push_back(c); }
}
void Engine::addSubStrength(uint32_t c)
{
    if (!(C[c].flag & F_STR)) { C[c].flag |= F_STR; strq.push_back(c); }
    if (!(C[c].flag & F_SUB)) { C[c].flag |= F_SUB; subq.push_back(c); }
}
void Engine::extClause(uint32_t c, int x)
{
    undo.push_back(0); undo.push_back(toDimacs(x));
    const int32_t* l = CL(c);
    for (uint32_t i = 0; i < C[c].size; ++i) { if (l[i] != x) { undo.push_back(toDimacs(l[i])); } }
}
void Engine::extLit(int x) { undo.push_back(0); undo.push_back(toDimacs(x)); }

size_t Engine::filterDeletedRange(uint32_t* a, size_t n)
{
    if (n >= std::max<size_t>(4 * tune.par_min, 2) && nthreads > 1) {
        // in place: every chunk compacts itself on its own core, then the chunks are moved down in order (no second array to fault in)
        const int T = nthreads; const size_t per = (n + (size_t)T - 1) / (size_t)T;
        std::vector<size_t> kept((size_t)T, 0);
        parallelFor((size_t)T, [&](size_t tb, size_t te, int) {
            for (size_t t = tb; t < te; ++t) {
                const size_t b = per * t, e = std::min(n, per * (t + 1)); size_t j = b;
                for (size_t i = b; i < e; ++i) { if (!(C[a[i]].flag & F_DEL)) { a[j++] = a[i]; } }
                kept[t] = j > b ? j - b : 0;
            }
        }, 1);
        size_t o = 0;
        for (int t = 0; t < T; ++t) { const size_t b = per * (size_t)t; if (b >= n) { break; } if (o != b && kept[(size_t)t]) { memmove(a + o, a + b, sizeof(uint32_t) * kept[(size_t)t]); } o += kept[(size_t)t]; }
        return o;
    }
    size_t j = 0;
    for (size_t i = 0; i < n; ++i) { if (!(C[a[i]].flag & F_DEL)) { a[j++] = a[i]; } }
    return j;
}
// garbageCollect / relocAll, CoprocessorTypes.h:1396-1550 (stable removal of deleted clauses everywhere)
void Engine::checkGarbage()
{
    if (!(ca_wasted > ca_size * opt.gc_frac)) { return; }
    ++n_gc;
    ensureHostLists();
    filterDeleted(subq); filterDeleted(strq);
    for (int x = 0; x < 2 * nvars; ++x) {
        ListRef l = L(x); uint32_t* p = LP(l); uint32_t j = 0;
        for (uint32_t i = 0; i < l.n; ++i) { if (!(C[p[i]].flag & F_DEL)) { p[j++] = p[i]; } }
        l.n = j; staleClear(x);
    }
    filterDeleted(clauses);
    double sz = 0;
    for (uint32_t c : clauses) { sz += 2 + C[c].size; }
    ca_size = sz; ca_wasted = 0;
    // relocAll: the arena itself.  The device packs its copy (K8: scan of the live sizes + gather); the host mirror is
    // packed by the same rule - clause order, live sizes - so both keep identical offsets without a transfer.
    if (dev_uploaded) {
        flushDevice();
        RawVec<int32_t> np; np.reserve(pool.size());
        for (size_t c = 0; c < C.size(); ++c) {
            const uint32_t ns = (uint32_t)np.size();
            if (C[c].flag & F_DEL) { C[c].size = 0; }
            else { np.insert(np.end(), pool.begin() + C[c].start, pool.begin() + C[c].start + C[c].size); }
            C[c].start = ns;
        }
        pool.swap(np);
        size_t nl = 0;
        int64_t t0 = now_ns();
        if (cp3g_compact(gpu, nullptr, nullptr, 0, &nl) != CP3G_OK) { die("arena compaction failed"); }
        host_ns_gpuwait += now_ns() - t0;
        if (nl != pool.size()) { die("arena compaction: host and device pools differ"); }
        ++n_compactions;
    }
}

// ------------------------------------------------------------------------------------------------ Propagation::process (Propagation.cc:32-131)
int Engine::propagationProcess(bool sort, bool useHeap, int ignore)
{
    t_up.modified = false;
    if (!ok) { return L_FALSE; }
    if (up_last < (int)trail.size()) { ensureHostLists(); }
    for (; up_last < (int)trail.size(); up_last++) {
        const int l = trail[up_last];
        {
            ListRef positive = L(l);
            for (uint32_t i = 0; i < positive.n; ++i) {
                uint32_t c = LP(positive)[i];
                if (C[c].flag & F_DEL) { continue; }
                ++up_removedClauses;
                setDeleted(c);
                t_up.modified = true;
                removedClause(c, useHeap, false, ignore);
            }
            occRelease(l);
        }
        const int nl = lneg(l);
        int count = 0;
        ListRef negative = L(nl);
        for (uint32_t i = 0; i < negative.n; ++i) {
            uint32_t c = LP(negative)[i];
            if (C[c].flag & F_DEL) { continue; }
            int32_t* lits = CL(c);
            for (uint32_t j = 0; j < C[c].size; ++j) {
                if (lits[j] == nl) {
                    if (!sort) { removePosUnsorted(c, (int)j); } else { removePosSorted(c, (int)j); }
                    noteShrunk(c, nl, false);
                    unpredictedEdit(c);
                    t_up.modified = true;
                    count++;
                    break;
                }
            }
            addSubStrength(c);
            dirty_str.push_back(c);
            if (C[c].size == 0) { setDeleted(c); ok = false; return L_FALSE; }
            else if (C[c].size == 1) {
                setDeleted(c);
                int u = CL(c)[0];
                if (value(u) == L_UNDEF) { uncheckedEnqueue(u); }
                else if (value(u) == L_FALSE) { ok = false; }
            }
        }
        occRelease(nl);
        // QUIRK: removedLiteral(nl, count, heap, ignore) - `ignore` lands in the `update` slot (Propagation.cc:114)
        removedLiteral(nl, count, useHeap, ignore != 0, VAR_UNDEF);
        up_removedLiterals += count;
    }
    qhead = up_last;
    if (!t_up.modified) { unsuccessful(t_up); }
    return ok ? L_UNDEF : L_FALSE;
}

// ------------------------------------------------------------------------------------------------ driver
// data.init + initializePreprocessor + sortClauses (Coprocessor.cc:100-118,1243-1283,1447-1463).
// The occurrence lists are built on the GPU (K2) and copied back - they start in clause order.
void Engine::initData()
{
    PhaseTimer pt(ph_init);
    const int n2 = 2 * nvars;
    // (per-literal / per-variable tables: unzeroed storage, filled on all cores)
    pfill(occ, (size_t)n2, OccList()); pfill(hot, (size_t)n2, LitHot()); pfill(stale, (size_t)n2, 0u);
    stalebits.assign(((size_t)n2 >> 6) + 1, 0ull);
    pfill(modTimer, (size_t)std::max(nvars, 1), 0u); pfill(ma, (size_t)std::max(n2, 1), 0u);
    pfill(var_touch, (size_t)std::max(nvars, 1), 0u); pfill(bve_inpend, (size_t)std::max(nvars, 1), (uint8_t)0); bve_pending.clear();
    modStep = 0; numberOfCls = 0; qhead = 0;
    subq.clear(); strq.clear();
    int64_t ti0 = now_ns();
    uploadAll();
    ph_upload += now_ns() - ti0; ti0 = now_ns();
    RawVec<uint32_t> off; off.resize((size_t)n2 + 1);
    if (cp3g_download_occ(gpu, off.data(), nullptr) != CP3G_OK) { die("download_occ failed"); }
    occ_pool.resize((size_t)off[n2] + 16);                 // filled by ensureHostLists() / the device-side commit of the first pass
    lists_on_device = true;
    parallelFor((size_t)n2, [&](size_t b, size_t e, int) {
        for (size_t x = b; x < e; ++x) {
            occ[x].off = off[x]; hot[x].n = occ[x].cap = off[x + 1] - off[x];
            hot[x].cnt = (int32_t)hot[x].n;
        }
    });
    ph_download += now_ns() - ti0;
    DbgLap lap("init"); 
    {   // every live clause: count it, queue it per its flags (Subsumption::initClause), drop the deleted ones.
        // Right after the ingest every clause is live and carries both flags: then the queues are copies.
        std::vector<size_t> part((size_t)nthreads + 1, 0);
        parallelFor(clauses.size(), [&](size_t b, size_t e, int t) {
            size_t a = 0;
            for (size_t i = b; i < e; ++i) { const uint8_t f = C[clauses[i]].flag; a += (!(f & F_DEL) && (f & F_SUB) && (f & F_STR)) ? 1 : 0; }
            part[(size_t)t] = a;
        });
        size_t full = 0; for (size_t a : part) { full += a; }
        if (full == clauses.size()) {
            numberOfCls += (int)clauses.size();
            const size_t nc = clauses.size(), s0 = subq.size(), t0q = strq.size();
            subq.resize(s0 + nc); strq.resize(t0q + nc);
            parallelFor(nc, [&](size_t b, size_t e, int) { memcpy(subq.data() + s0 + b, clauses.data() + b, sizeof(uint32_t) * (e - b)); memcpy(strq.data() + t0q + b, clauses.data() + b, sizeof(uint32_t) * (e - b)); });
        } else {
            size_t kept = 0;
            for (size_t i = 0; i < clauses.size(); ++i) {
                uint32_t c = clauses[i];
                if (C[c].flag & F_DEL) { continue; }
                numberOfCls++;
                subInitClause(c);
                clauses[kept++] = c;
            }
            clauses.resize(kept);
        }
    }
    lap("initClause loop");
    data_ready = true; data_nvars = nvars;
}

// Dense::compress, coprocessor/techniques/Dense.cc:26-238.  The data-parallel part - which variables still occur,
// the dense renumbering (a prefix sum) and the rewrite of every clause literal - runs on the device (K7,
// cp3g_dense); the host moves the per-variable solver state and the trail like CoprocessorData::moveVar
// (CoprocessorTypes.h:693-729) and keeps the map for Dense::decompress.
void Engine::denseCompress()
{
    if (propagationProcess(false, false, VAR_UNDEF) == L_FALSE) { ok = false; }      // Dense.cc:28
    if (!ok || nvars == 0) { return; }
    flushDevice();
    const int n = nvars;
    std::vector<uint8_t> keep((size_t)n, 0);
    for (int v = 0; v < n; ++v) { keep[v] = frozen[v] || (opt.dense_keep_set && assigns[v] != L_UNDEF); }
    std::vector<uint32_t> mapping((size_t)n, 0);
    uint32_t nv_new = (uint32_t)n; int compressed = 0;
    int64_t t0 = now_ns();
    if (cp3g_dense(gpu, keep.data(), opt.dense_frag, mapping.data(), &nv_new, &compressed, pool.data(), pool.size()) != CP3G_OK) { die("dense failed"); }
    host_ns_gpuwait += now_ns() - t0;
    if (!compressed) { return; }                                                     // isCompact, Dense.h:124-127
    comp_map.assign((size_t)n, -1); comp_val.assign((size_t)n, (uint8_t)L_UNDEF);
    for (int v = 0; v < n; ++v) {
        if (mapping[v] != 0xffffffffu) { comp_map[v] = (int32_t)mapping[v]; }
        else if (assigns[v] != L_UNDEF) { comp_val[v] = assigns[v]; }
    }
    // trail (Dense.cc:121-150): literals of removed variables leave it - their value lives in the compression
    size_t j = 0;
    for (size_t i = 0; i < trail.size(); ++i) {
        const int x = trail[i], v = x >> 1;
        if (comp_map[v] >= 0) { trail[j++] = 2 * comp_map[v] + (x & 1); }
        else { assigns[v] = L_UNDEF; }
    }
    trail.resize(j); qhead = (int)j;                                                 // propagation.reset
    for (int v = 0; v < n; ++v) {                                                    // moveVar
        const int to = comp_map[v];
        if (to >= 0 && to != v) { assigns[to] = assigns[v]; assigns[v] = L_UNDEF; frozen[to] = frozen[v]; frozen[v] = 0; }
    }
    nvars = (int)nv_new; data_nvars = nvars;                                         // CoprocessorTypes.h:725
    assigns.resize((size_t)nvars); frozen.resize((size_t)nvars);
    data_ready = false;                                                              // occurrence data is void (data.destroy follows)
}

int Engine::importLit(int d) const
{
    const int v = (d < 0 ? -d : d) - 1;
    if (comp_map.empty()) { return d; }
    if (v >= (int)comp_map.size() || comp_map[v] < 0) { return 0; }
    return d < 0 ? -(comp_map[v] + 1) : comp_map[v] + 1;
}

void Engine::preprocess()
{
    // Coprocessor.cc:1002-1006, the size gates.  data.nVars() is CoprocessorData::numberOfVars (0 before the first
    // data.init()); nTotLits() = clauses_literals + learnts_literals (riss/core/Solver.h:1867): the literals of all attached
    // clauses (Solver.cc:706-707), i.e. of every stored clause - there are no learnt clauses on this path
    if (opt.limited) {
        std::vector<uint64_t> part((size_t)nthreads + 1, 0);
        parallelFor(clauses.size(), [&](size_t b, size_t e, int t) { uint64_t a = 0; for (size_t i = b; i < e; ++i) { a += C[clauses[i]].size; } part[(size_t)t] = a; });
        uint64_t totlits = 0; for (uint64_t a : part) { totlits += a; }
        if ((uint32_t)data_nvars > (uint32_t)opt.cp3_vars || (int64_t)clauses.size() > (int64_t)opt.cp3_cls || (uint32_t)totlits > (uint32_t)opt.cp3_lits) { return; }
    }
    if (!opt.enabled) { return; }
    PhaseTimer ptot(ph_total);
    ensureGpu();                                  // fail before touching anything when there is no device
    sub_lim = opt.sub_limit; str_lim = opt.str_limit;
    DbgLap plap("preprocess");
    { PhaseTimer ps(ph_simplify); if (!solverSimplify()) { return; } }
    sortPermuted();
    plap("simplify + sort");
    initData();
    plap("initData");
    int status = L_UNDEF;
    for (int it = 0; it < opt.iters; ++it) {
        if (opt.up) { if (status == L_UNDEF) { status = propagationProcess(true, false, VAR_UNDEF); } }
        checkGarbage();                               // unconditional checkpoints behind the UP / XOR / ENT blocks, Coprocessor.cc:156,171
        if (!ok) { break; }
        if (opt.subsimp) {
            if (status == L_UNDEF) { subsumptionProcess(true, false, VAR_UNDEF); }
            if (!ok) { status = L_FALSE; }
            checkGarbage();
        }
        if (!ok) { break; }
        if (opt.bve) {
            if (status == L_UNDEF) { status = bveProcess(); }
            checkGarbage();
        }
        if (!ok) { break; }
        if (opt.bce) {                                // Coprocessor.cc:345-352
            if (status == L_UNDEF) { bceProcess(); }
            checkGarbage();
        }
        if (!ok) { break; }
    }
    plap("techniques");
    if (ok && hasToPropagate()) { propagationProcess(false, false, VAR_UNDEF); }
    if (opt.dense) { denseCompress(); }           // very last step, Coprocessor.cc:492-496
    if (ok) { filterDeleted(clauses); }           // reSetupSolver, CoprocessorTypes.h:937-1008
    plap("tail");
    if (opt.stats) {
        fprintf(stderr, "c [STAT] SuSi(1) %d cls,  with %d lits, %d strengthed\n", subsumedClauses, subsumedLiterals, removedLiterals);
        fprintf(stderr, "c [STAT] SuSi(2) %lld subs-steps, %lld strenght-steps, %d increases\n", (long long)sub_steps, (long long)str_steps, limitIncreases);
        fprintf(stderr, "c [STAT] BVE(3) %d eliminated, %d created, %d removed, %lld checks\n", bve_eliminatedVars, bve_createdClauses, bve_removedClauses, (long long)bve_steps);
        fprintf(stderr, "c [STAT] GPU %lld launches, %lld scan batches (%lld on demand), %lld fast / %lld slow queue entries\n",
                (long long)cp3g_counter(gpu, "launches"), (long long)n_scan_batches, (long long)n_ondemand, (long long)n_fast, (long long)n_slow);
    }
}

size_t Engine::output(int32_t* out, size_t cap) const
{
    // trail units first, then solver->clauses in order (libcoprocessorc.cc:80-108); sizes per chunk, prefix, parallel fill
    Engine* self = const_cast<Engine*>(this);
    const size_t nc = clauses.size();
    const size_t T = (size_t)std::max(1, nthreads);
    std::vector<size_t> part(T + 1, 0);
    self->parallelFor(nc, [&](size_t b, size_t e, int t) {
        size_t a = 0;
        for (size_t i = b; i < e; ++i) { const ClauseInfo& ci = C[clauses[i]]; if (!(ci.flag & F_DEL)) { a += (size_t)ci.size + 1; } }
        part[(size_t)t + 1] = a;
    });
    for (size_t t = 1; t <= T; ++t) { part[t] += part[t - 1]; }      // (the second parallelFor below splits [0,nc) the same way)
    const size_t total = 2 * trail.size() + part[T];
    if (out == nullptr || cap < total) {
        if (out != nullptr) {                        // short buffer: fill what fits, literal by literal
            size_t k = 0;
            auto emit = [&](int32_t x) { if (k < cap) { out[k] = x; } ++k; };
            for (int x : trail) { emit(toDimacs(x)); emit(0); }
            for (uint32_t c : clauses) {
                if (C[c].flag & F_DEL) { continue; }
                const int32_t* l = CL(c);
                for (uint32_t j = 0; j < C[c].size; ++j) { emit(toDimacs(l[j])); }
                emit(0);
            }
        }
        return total;
    }
    size_t k = 0;
    for (int x : trail) { out[k++] = toDimacs(x); out[k++] = 0; }
    const size_t base = k;
    self->parallelFor(nc, [&](size_t b, size_t e, int t) {
        size_t o = base + part[(size_t)t];
        for (size_t i = b; i < e; ++i) {
            const ClauseInfo& ci = C[clauses[i]];
            if (ci.flag & F_DEL) { continue; }
            const int32_t* l = pool.data() + ci.start;
            for (uint32_t j = 0; j < ci.size; ++j) { out[o++] = toDimacs(l[j]); }
            out[o++] = 0;
        }
    });
    return total;
}
size_t Engine::nLiveClauses() const { size_t n = 0; for (uint32_t c : clauses) { n += !(C[c].flag & F_DEL); } return n; }
size_t Engine::nLiveLits() const { size_t n = 0; for (uint32_t c : clauses) { if (!(C[c].flag & F_DEL)) { n += C[c].size; } } return n; }
size_t Engine::undoStack(int32_t* out, size_t cap) const
{
    for (size_t i = 0; i < undo.size() && i < cap; ++i) { out[i] = undo[i]; }
    return undo.size();
}

// CoprocessorData::extendModel, CoprocessorTypes.h:1665-1764 (without Dense / EE): walk the undo stack
// backwards; an entry [blocking literal, rest...] that is not satisfied flips its blocking literal.
void Engine::extendModel(std::vector<uint8_t>& model) const
{
    if (!comp_map.empty()) {
        // Dense::decompress, Dense.cc:240-318: backwards, because names only grow on the way back
        const size_t n0 = comp_map.size();
        if (model.size() < n0) { model.resize(n0, (uint8_t)L_FALSE); }
        for (long v = (long)n0 - 1; v >= 0; --v) {
            const int c = comp_map[(size_t)v];
            if (c < 0) { if (comp_val[(size_t)v] != L_UNDEF) { model[(size_t)v] = comp_val[(size_t)v]; } }
            else if (c != v) { model[(size_t)v] = model[(size_t)c]; }
        }
        for (auto& m : model) { if (m == L_UNDEF) { m = L_TRUE; } }
        // the trail is in compressed names; its values were copied above with the variables they belong to
        const std::vector<int32_t>& cm = comp_map;
        std::vector<int> inv((size_t)nvars, -1);
        for (size_t v = 0; v < n0; ++v) { if (cm[v] >= 0) { inv[(size_t)cm[v]] = (int)v; } }
        for (int x : trail) { if (inv[(size_t)(x >> 1)] >= 0) { model[(size_t)inv[(size_t)(x >> 1)]] = (uint8_t)(x & 1); } }
        extendByUndo(model);
        return;
    }
    if (model.size() < (size_t)nvars) { model.resize((size_t)nvars, L_TRUE); }
    for (auto& m : model) { if (m == L_UNDEF) { m = L_TRUE; } }
    for (int x : trail) { model[x >> 1] = (uint8_t)(x & 1); }
    extendByUndo(model);
}

void Engine::extendByUndo(std::vector<uint8_t>& model) const
{
    long i = (long)undo.size() - 1;
    while (i >= 0) {
        long j = i;
        while (j >= 0 && undo[j] != 0) { --j; }
        // entry = undo[j+1 .. i]
        bool sat = false;
        for (long k = j + 1; k <= i; ++k) {
            int d = undo[k]; int v = (d < 0 ? -d : d) - 1;
            if ((model[v] == L_TRUE) == (d > 0)) { sat = true; break; }
        }
        if (!sat && j + 1 <= i) {
            int d = undo[j + 1]; int v = (d < 0 ? -d : d) - 1;
            model[v] = d > 0 ? L_TRUE : L_FALSE;
        }
        i = j - 1;
    }
}

int64_t Engine::stat(const char* s) const
{
#define ST(name, val) if (!strcmp(s, name)) { return (int64_t)(val); }
    ST("sub_subsumedClauses", subsumedClauses) ST("sub_subsumedLiterals", subsumedLiterals)
    ST("sub_removedLiterals", removedLiterals) ST("sub_subsSteps", sub_steps) ST("sub_strengthSteps", str_steps)
    ST("bve_eliminatedVars", bve_eliminatedVars) ST("bve_createdClauses", bve_createdClauses)
    ST("bve_removedClauses", bve_removedClauses) ST("bve_testedVars", bve_testedVars)
    ST("bve_anticipations", bve_anticipations) ST("bve_skippedVars", bve_skippedVars)
    ST("bve_unitsEnqueued", bve_unitsEnqueued) ST("bve_removedBC", bve_removedBC)
    ST("bve_totallyAddedClauses", bve_totallyAdded) ST("bve_steps", bve_steps)
    ST("up_removedClauses", up_removedClauses) ST("up_removedLiterals", up_removedLiterals)
    ST("bve_foundGates", bve_foundGates)
    ST("bce_steps", bce_steps) ST("bce_testedLits", bce_testedLits) ST("bce_remBCE", bce_remBCE) ST("bce_batches", n_bce_batches) ST("bce_rows_ondemand", n_bce_rows)
    ST("scan_batches", n_scan_batches) ST("scan_ondemand", n_ondemand) ST("queue_fast", n_fast) ST("queue_slow", n_slow)
    ST("eval_batches", n_eval_batches) ST("eval_used", n_eval_used) ST("eval_host", n_eval_host) ST("eval_stale", n_eval_stale) ST("stream_used", n_stream_used)
    ST("pair_batches", n_pair_batches) ST("resolve_batches", n_resolve_batches) ST("arena_compactions", n_compactions)
    ST("bulk_ingests", n_bulk_ingests) ST("queue_target_fast", n_tfast)
    ST("ph_upload_ns", ph_upload) ST("ph_download_ns", ph_download) ST("ph_simplify_ns", ph_simplify)
    ST("ph_init_ns", ph_init) ST("ph_sub_ns", ph_sub) ST("ph_str_ns", ph_str) ST("ph_prefetch_ns", ph_prefetch) ST("ph_total_ns", ph_total)
    ST("queue_static", n_static) ST("parallel_sub_passes", n_parsub) ST("device_sub_passes", n_devsub) ST("overlapped_str_scans", n_overlap)
    ST("device_static_passes", n_devstatic) ST("component_passes", n_compar) ST("component_fallbacks", n_compar_fallback) ST("component_rounds", n_compar_rounds)
    ST("spec_requests", n_spec_requests) ST("spec_rounds", n_spec_rounds) ST("spec_used", n_spec_used)
    ST("host_commit_ns", host_ns_commit) ST("host_gpuwait_ns", host_ns_gpuwait)
    if (!strncmp(s, "gpu_", 4)) { return gpu ? cp3g_counter(gpu, s + 4) : 0; }
#undef ST
    return -1;
}


} // namespace cp3
