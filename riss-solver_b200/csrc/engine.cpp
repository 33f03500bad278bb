// engine.cpp - ordered-commit host engine (see engine.h).  Reference lines are cited per function.
#include "engine.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_set>

namespace cp3 {

struct PhaseTimer {
    int64_t& acc; int64_t t0;
    explicit PhaseTimer(int64_t& a);
    ~PhaseTimer();
};
static inline int64_t now_ns()
{
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

PhaseTimer::PhaseTimer(int64_t& a) : acc(a), t0(now_ns()) {}
PhaseTimer::~PhaseTimer() { acc += now_ns() - t0; }

// ------------------------------------------------------------------------------------------------ options
int Options::parse(const char* o)
{
    if (!o || o[0] != '-') { return 0; }
    const char* s = o + 1;
    bool neg = false;
    if (!strncmp(s, "no-", 3)) { neg = true; s += 3; }
    const char* eq = strchr(s, '=');
    std::string name = eq ? std::string(s, eq - s) : std::string(s);
    long long val = eq ? atoll(eq + 1) : 0;
#define BOPT(n, f) if (name == n) { f = !neg; return 1; }
#define IOPT(n, f) if (name == n && eq) { f = val; return 1; }
    BOPT("enabled_cp3", enabled) BOPT("up", up) BOPT("subsimp", subsimp) BOPT("bve", bve)
    BOPT("cp3_limited", limited) BOPT("cp3_strength", strength) BOPT("cp3_stats", stats)
    BOPT("bve_unlimited", bve_unlimited) BOPT("bve_strength", bve_strength) BOPT("bve_gates", bve_gates)
    BOPT("bve_force_gates", bve_force_gates) BOPT("bve_fdepOnly", bve_fdepOnly) BOPT("bve_BCElim", bve_BCElim)
    BOPT("bve_early", bve_early) BOPT("bce_only", bce_only) BOPT("dense", dense) BOPT("cp3_keep_set", dense_keep_set)
    IOPT("cp3_dense_frag", dense_frag)
    BOPT("bce", bce) BOPT("bce-bce", bce_bce) BOPT("bce-compl", bce_compl) BOPT("bce-bin", bce_binary) IOPT("bce-limit", bce_limit)
    IOPT("cp3_iters", iters) IOPT("cp3_vars", cp3_vars) IOPT("cp3_cls", cp3_cls) IOPT("cp3_lits", cp3_lits)
    IOPT("cp3_sub_limit", sub_limit) IOPT("cp3_str_limit", str_limit) IOPT("cp3_call_inc", call_inc)
    IOPT("cp3_bve_limit", bve_limit) IOPT("bve_cgrow", bve_cgrow) IOPT("bve_cgrow_t", bve_cgrow_t)
    IOPT("bve_heap_updates", bve_heap_updates) IOPT("cp3_gpu_device", device)
    if (name == "cp3_bve_heap" && eq) { if (val < 0 || val > 10 || val == 2) { return -1; } bve_heap = (int)val; return 1; }   // 2 = random order, no heap
    // options of the path whose non-default values select code the GPU path does not provide
    if (name == "cp3_threads" && eq) { threads = (int)val; return 1; }      // pthread pool is replaced by the GPU
    BOPT("naive_strength", naive_strength)
    if (name == "all_strength_res" && eq) { return val == 0 ? 1 : -1; }
    if (name == "cp3_inpPrefL" || name == "cp3_lock_stats" || name == "bve_progress" || name == "bve_totalG") { return 1; }
    if ((name == "cp3_par_strength" || name == "cp3_par_subs" || name == "par_subs_counts" || name == "susi_chunk_size" ||
         name == "par_str_minCls" || name == "cp3_par_bve" || name == "par_bve_th" || name == "postp_lockd_neighb" ||
         name == "cp3_sub_inpInc" || name == "cp3_bve_inpInc" || name == "cp3_bve_verbose" || name == "cp3_bve_learnt_growth" ||
         name == "cp3_bve_resolve_learnts" || name == "cp3_verbose" || name == "cp3_susi_vars" || name == "cp3_susi_cls" ||
         name == "cp3_susi_lits" || name == "cp3_bve_vars" || name == "cp3_bve_cls" || name == "cp3_bve_lits" || name == "cp3_bce_vars" ||
         name == "cp3_bce_cls" || name == "cp3_bce_lits" || name == "bce-incInp" || name == "bce-verbose") && eq) { return 1; }
    // BCE variants the GPU path does not provide: covered literal elimination / addition, blocked clause minimisation
    if (name == "bce-cle" || name == "bce-cla" || name == "bce-bcm" || name == "bce-cle-cons") { return neg ? 1 : -1; }
    if (name == "par_bve_min_upd" || name == "inprocess" || name == "dense_inp") { return neg ? 1 : -1; }
    if (name == "cp3_dense_debug" && eq) { return 1; }
#undef BOPT
#undef IOPT
    return 0;
}

// Presets (Config::setPreset, riss/utils/Config.h:97-560) that reach this path.  A preset is an option string; the options
// of the path are applied, the techniques the GPU drop-in does not provide are named on stderr (once per preset), so that
// nobody mistakes the result for the reference's full pipeline.  plain_SUSI :377, plain_BVE :381, BVEEARLY :383,
// Riss427 :317, Riss427nd :315, Riss427-NoCLE :319, 505 / 505-M :542.
void Options::preset(const char* name)
{
    if (!name) { return; }
    struct P { const char* name; const char* path; const char* rest; };
    static const P table[] = {
        {"plain_SUSI", "-enabled_cp3 -cp3_stats -subsimp", ""},
        {"plain_BVE", "-enabled_cp3 -cp3_stats -bve", ""},
        {"plain_UP", "-enabled_cp3 -cp3_stats -up", ""},
        {"BVEEARLY", "-bve_early", ""},
        {"Riss427", "-enabled_cp3 -cp3_stats -bve -dense", "-fm -unhide -bce (-bce-cle)"},
        {"Riss427nd", "-enabled_cp3 -cp3_stats -bve", "-fm -unhide -bce (-bce-cle)"},
        {"RissND427", "-enabled_cp3 -cp3_stats -bve", "-fm -unhide -bce (-bce-cle)"},
        {"Riss427-NoCLE", "-enabled_cp3 -cp3_stats -bve -dense", "-fm -unhide"},
        {"505", "-enabled_cp3 -cp3_stats -bve -dense -cp3_iters=2 -bve_early", "-fm -unhide -bce (-bce-cle) -xor -ee and the search options"},
        {"505-M", "-enabled_cp3 -cp3_stats -bve -dense -cp3_iters=2 -bve_early", "-fm -unhide -bce (-bce-cle) -xor -ee and the search options"},
    };
    auto parseString = [&](const std::string& opts) {
        size_t q = 0;
        while (q < opts.size()) { size_t e = opts.find(' ', q); if (e == std::string::npos) { e = opts.size(); } if (e > q) { parse(opts.substr(q, e - q).c_str()); } q = e + 1; }
    };
    std::string n(name);
    size_t pos = 0;
    while (pos <= n.size()) {                              // Config::setPreset: tokens separated by ':'
        size_t c = n.find(':', pos);
        std::string one = n.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
        bool known = false;
        for (const P& p : table) {
            if (one != p.name) { continue; }
            known = true;
            parseString(p.path);
            if (p.rest[0]) { fprintf(stderr, "cp3b200: preset %s: applied \"%s\"; techniques of this preset outside the GPU path are NOT run: %s\n", p.name, p.path, p.rest); }
        }
        // Config::addPreset, riss/utils/Config.h:736-739: a token that names no preset is parsed as an option string
        if (!known && !one.empty()) {
            if (one.find('-') != std::string::npos) { parseString(one); }
            else { fprintf(stderr, "cp3b200: preset %s is not known to the GPU path and has no effect\n", one.c_str()); }
        }
        if (c == std::string::npos) { break; }
        pos = c + 1;
    }
}

// sort + unique of a clause id list; queues that were filled in clause order (the first round) are already sorted
static void sortUnique(std::vector<uint32_t>& v)
{
    bool strictly = true;
    for (size_t i = 1; i < v.size(); ++i) { if (v[i - 1] >= v[i]) { strictly = false; break; } }
    if (strictly) { return; }
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

// large hit lists arrive sorted by (request, clause) from the device (gpu_db.cu sort_hits); small or gathered ones are sorted here
static void sortHits(std::vector<cp3g_hit>& h, uint32_t n)
{
    auto less = [](const cp3g_hit& a, const cp3g_hit& b) { return a.req != b.req ? a.req < b.req : a.clause < b.clause; };
    if (!std::is_sorted(h.begin(), h.begin() + n, less)) { std::sort(h.begin(), h.begin() + n, less); }
}

struct DbgLap {
    bool on; int64_t t; const char* who;
    explicit DbgLap(const char* w) : on(getenv("CP3_DEBUG_PAR") != nullptr), t(on ? now_ns() : 0), who(w) {}
    void operator()(const char* what) { if (on) { int64_t n2 = now_ns(); fprintf(stderr, "  %s %s %.1f ms\n", who, what, (n2 - t) / 1e6); t = n2; } }
};

void Engine::ScanCache::clear()
{
    for (uint32_t c : used) { slot[c] = -1; }
    used.clear(); ent.clear(); hits.clear(); lits.clear(); bulk_all = false;
}

Engine::Engine()
{
    sub_lim = opt.sub_limit; str_lim = opt.str_limit;
    const char* e = getenv("CP3_THREADS");
    nthreads = e ? atoi(e) : (int)std::min(24u, std::max(1u, std::thread::hardware_concurrency()));
    if (nthreads < 1) { nthreads = 1; }
}
Engine::~Engine() { if (gpu) { cp3g_destroy(gpu); } }

void Engine::die(const char* msg) const
{
    fprintf(stderr, "cp3b200: fatal: %s%s%s\n", msg, gpu ? " - " : "", gpu ? cp3g_last_error(gpu) : "");
    abort();
}

// contiguous chunks of [0,n) on the host cores (the statically inert part of the queue walks)
template <class F> void Engine::parallelFor(size_t n, F f, size_t grain)
{
    const int T = (int)std::min<size_t>((size_t)nthreads, (n + grain - 1) / grain);
    if (T <= 1) { f((size_t)0, n, 0); return; }
    std::vector<std::thread> th;
    const size_t per = (n + T - 1) / T;
    for (int t = 0; t < T; ++t) {
        const size_t b = per * t, e = std::min(n, per * (t + 1));
        if (b >= e) { break; }
        th.emplace_back([=]() { f(b, e, t); });
    }
    for (auto& x : th) { x.join(); }
}

template <class P> void Engine::parallelFilter(const std::vector<uint32_t>& src, std::vector<uint32_t>& out, P pred)
{
    const size_t n = src.size();
    std::vector<size_t> cnt((size_t)nthreads + 2, 0);
    parallelFor(n, [&](size_t b, size_t e, int t) { size_t c = 0; for (size_t i = b; i < e; ++i) { c += pred(src[i]) ? 1 : 0; } cnt[(size_t)t + 1] = c; });
    for (size_t t = 1; t < cnt.size(); ++t) { cnt[t] += cnt[t - 1]; }
    out.resize(cnt.back());
    parallelFor(n, [&](size_t b, size_t e, int t) { size_t o = cnt[(size_t)t]; for (size_t i = b; i < e; ++i) { if (pred(src[i])) { out[o++] = src[i]; } } });
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
    static const size_t bulk_min = getenv("CP3_INGEST_MIN") ? (size_t)atoll(getenv("CP3_INGEST_MIN")) : (size_t)262144;
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
        if (m) {
            ensureGpu();
            uint32_t nv = 0, nc = 0; size_t nl = 0; int st = 1;
            int64_t t0 = now_ns();
            if (cp3g_ingest(gpu, s, m, &nv, &nc, &nl, &st) != CP3G_OK) { die("bulk ingest failed"); }
            if (st == 0) {
                if ((int)nv > nvars) { assigns.resize((size_t)nv, (uint8_t)L_UNDEF); frozen.resize((size_t)nv, 0); nvars = (int)nv; }
                pool.resize(nl); C.resize(nc);
                if (cp3g_download_arena(gpu, pool.data(), C.data()) != CP3G_OK) { die("arena download failed"); }
                clauses.resize(nc);
                parallelFor(nc, [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { clauses[i] = (uint32_t)i; C[i].stamp = scan_id; } });
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
    dev_ncl = ncl; dev_uploaded = true;
    pend_del.clear(); pend_shrunk.clear();
    shrunk_mark.assign(ncl, 0);
}

// push the edits committed since the last scan (new resolvents, shrunk clauses, deletions)
void Engine::flushDevice()
{
    const uint32_t ncl = (uint32_t)C.size();
    int64_t t0 = now_ns();
    const uint32_t old_ncl = dev_ncl;
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
        for (uint32_t c : pend_shrunk) {
            shrunk_mark[c] = 0;
            if (c >= old_ncl) { continue; }          // appended above with its current content
            ids.push_back(c); st.push_back((uint32_t)li.size()); sz.push_back(C[c].size);
            li.insert(li.end(), CL(c), CL(c) + C[c].size);
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
    DbgLap lap(kind == CP3G_SUBSUME ? "scanClauses<sub>" : "scanClauses<str>");
    if (!scan_no_flush) { flushDevice(); }
    lap("flush");
    ++scan_id; ++n_scan_batches;
    pgrow(cache.slot, C.size(), -1);
    // deleted clauses that still act as strengthener (Subsumption.cc:1438) must be sent by content
    static const uint32_t bulk_min = getenv("CP3_BULK_MIN") ? (uint32_t)atoi(getenv("CP3_BULK_MIN")) : 4096u;
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
        for (uint32_t h = 0; h < n;) {
            const uint32_t c = hitbuf[h].req;
            CacheEnt e; e.ver = C[c].ver; e.scan = scan_id; e.hb = (uint32_t)cache.hits.size();
            while (h < n && hitbuf[h].req == c) { cache.hits.push_back({hitbuf[h].clause, hitbuf[h].lit}); ++h; }
            e.he = (uint32_t)cache.hits.size();
            e.next = cache.slot[c]; e.lit_off = e.lit_n = 0; e.time = 0;
            cache.setSlot(c, (int32_t)cache.ent.size());
            cache.ent.push_back(e);
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
void Engine::scanSpeculative(int kind, const std::vector<SpecReq>& reqs, ScanCache& cache)
{
    if (reqs.empty()) { return; }
    ++scan_id; ++n_scan_batches; ++n_spec_rounds; n_spec_requests += (int64_t)reqs.size();
    pgrow(cache.slot, C.size(), -1);
    std::vector<uint32_t> self(reqs.size()), roff(reqs.size() + 1); std::vector<int32_t> rl;
    roff[0] = 0;
    for (size_t i = 0; i < reqs.size(); ++i) {
        self[i] = reqs[i].clause;
        rl.insert(rl.end(), spec_lits.begin() + reqs[i].off, spec_lits.begin() + reqs[i].off + reqs[i].n);
        roff[i + 1] = (uint32_t)rl.size();
    }
    int64_t t0 = now_ns();
    uint32_t cap = (uint32_t)std::max<size_t>(hitbuf.size(), std::max<size_t>(4096, reqs.size() / 4));
    uint32_t n = 0; uint64_t chk = 0;
    for (;;) {
        hitbuf.resize(cap);
        int r = cp3g_scan(gpu, kind, (uint32_t)reqs.size(), self.data(), roff.data(), rl.data(), hitbuf.data(), cap, &n, &chk);
        if (r == CP3G_OK) { break; }
        if (r == CP3G_ERR_OVERFLOW) { cap = n + n / 8 + 1024; continue; }
        die("speculative scan failed");
    }
    host_ns_gpuwait += now_ns() - t0;
    if (getenv("CP3_DEBUG_PAR")) { fprintf(stderr, "  scanSpeculative gpu call %.1f ms, %zu reqs, %u hits\n", (now_ns() - t0) / 1e6, reqs.size(), n); }
    sortHits(hitbuf, n);
    uint32_t h = 0;
    for (uint32_t i = 0; i < reqs.size(); ++i) {
        CacheEnt e; e.ver = UINT32_MAX; e.scan = scan_id; e.hb = (uint32_t)cache.hits.size();
        while (h < n && hitbuf[h].req == i) { cache.hits.push_back({hitbuf[h].clause, hitbuf[h].lit}); ++h; }
        e.he = (uint32_t)cache.hits.size();
        e.lit_off = (uint32_t)cache.lits.size(); e.lit_n = reqs[i].n; e.time = reqs[i].time;
        cache.lits.insert(cache.lits.end(), spec_lits.begin() + reqs[i].off, spec_lits.begin() + reqs[i].off + reqs[i].n);
        e.next = cache.slot[reqs[i].clause];
        cache.setSlot(reqs[i].clause, (int32_t)cache.ent.size());
        cache.ent.push_back(e);
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
    DbgLap lap("closure");
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
        scanSpeculative(CP3G_STRENGTHEN, reqs, c_str);
        lap("scanSpeculative");
    }
}

static inline bool containsLit(const int32_t* l, uint32_t n, int x)
{
    return std::binary_search(l, l + n, x);
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

void Engine::filterDeleted(std::vector<uint32_t>& a)
{
    size_t j = 0;
    for (size_t i = 0; i < a.size(); ++i) { if (!(C[a[i]].flag & F_DEL)) { a[j++] = a[i]; } }
    a.resize(j);
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

// ------------------------------------------------------------------------------------------------ Subsumption
void Engine::clearQueue(std::vector<uint32_t>& q, uint8_t flag)
{
    for (uint32_t c : q) { C[c].flag &= ~flag; }
    q.clear();
}

static inline const Hit* findHit(const std::vector<Hit>& hits, uint32_t hb, uint32_t he, uint32_t clause)
{
    const Hit* b = hits.data() + hb; const Hit* e = hits.data() + he;
    const Hit* it = std::lower_bound(b, e, clause, [](const Hit& h, uint32_t c) { return h.clause < c; });
    return (it != e && it->clause == clause) ? it : nullptr;
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
    const bool dbg = getenv("CP3_DEBUG_PAR") != nullptr; int64_t tt = now_ns();
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
    for (uint32_t c : c_sub.used) {
        const int32_t slot = c_sub.slot[c];
        if (slot < 0 || c_sub.ent[slot].hb == c_sub.ent[slot].he || sub_pos[c] < 0 || (C[c].flag & F_DEL)) { continue; }
        hitters.push_back({(int32_t)(nq - 1 - (size_t)sub_pos[c]), (uint32_t)slot});
    }
    std::sort(hitters.begin(), hitters.end());
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
    std::vector<int16_t> cnt16(nlit);
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
    static const bool off_env = getenv("CP3_NO_DEVCOMMIT") != nullptr;
    static const uint32_t max_list = getenv("CP3_DEVCOMMIT_MAXLIST") ? (uint32_t)atoi(getenv("CP3_DEVCOMMIT_MAXLIST")) : 4096u;
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
    DbgLap lap("devsub");
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
    parallelFor(nq, [&](size_t b, size_t e, int) { for (size_t p = b; p < e; ++p) { C[subq[p]].flag &= ~F_SUB; } });
    lap("bookkeeping");
    return true;
}

// subsumption_worker, Subsumption.cc:220-314.  One bulk K3 scan answers every ordered_subsumes() of the
// queue; the loop below replays the queue back to front.
void Engine::subsumptionWorker(bool useHeap, int ignore)
{
    PhaseTimer pt(ph_sub);
    const bool unlimited = !opt.limited;
    DbgLap lap0("sub"); 
    c_sub.clear(); logClear();
    lap0("clear");
    if (lists_on_device) {
        static const size_t parsub_min0 = getenv("CP3_PARSUB_MIN") ? (size_t)atoi(getenv("CP3_PARSUB_MIN")) : 65536;
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
        sortUnique(ids);
        lap0("ids");
        scanClauses(CP3G_SUBSUME, ids, c_sub);
    }
    DbgLap lap("sub"); lap("start(after scan)");
    int64_t t0 = now_ns();
    static const size_t parsub_min = getenv("CP3_PARSUB_MIN") ? (size_t)atoi(getenv("CP3_PARSUB_MIN")) : 65536;
    if (!useHeap && subq.size() >= parsub_min) {
        // step limit: the data-parallel form cannot stop in the middle of the queue; only use it far from the limit
        int64_t bound = 0;
        if (!unlimited) { for (uint32_t c : subq) { if (!(C[c].flag & F_DEL)) { int mx = 0; const int32_t* l = CL(c); for (uint32_t k = 0; k < C[c].size; ++k) { mx = std::max<int>(mx, (int)hot[l[k]].n); } bound += mx; } } }
        if (unlimited || sub_steps + bound < sub_lim) {
            // the strengthening pass that follows needs nothing from this commit but the deleted flags (a hit on a
            // clause that dies meanwhile is dropped at commit time): run its scan + closure concurrently
            static const bool no_overlap = getenv("CP3_NO_OVERLAP") != nullptr;
            std::thread helper; std::vector<uint32_t> sids; bool sbulk = false;
            const bool ov = overlap_str && !no_overlap && nthreads > 1 && opt.strength && strq.size() >= parsub_min && !hasToPropagate();
            if (ov) {
                for (uint32_t c : strq) { if (!(C[c].flag & F_DEL) && (C[c].flag & F_STR) && C[c].size >= 2) { sids.push_back(c); } }
                sortUnique(sids);
                static const uint32_t bulk_min2 = getenv("CP3_BULK_MIN") ? (uint32_t)atoi(getenv("CP3_BULK_MIN")) : 4096u;
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

// updateOccurrences, Subsumption.cc:1530-1544
void Engine::updateOccurrences(bool useHeap, int ignore)
{
    // from here on nothing is pending any more (setDeleted() consults pend_rm)
    std::vector<OccUpd> ups; ups.swap(occ_upd);
    ++pend_epoch;
    for (const OccUpd& u : ups) {
        if (removeClauseFrom(u.cr, u.l)) { removedLiteral(u.l, 1, useHeap, ignore != 0, VAR_UNDEF); }
    }
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
    if (getenv("CP3_DEBUG_ONDEMAND") && n_ondemand < 12) {
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
    DbgLap lap("prepStr");
    c_str.clear();
    lap("clear");
    scanClauses(CP3G_STRENGTHEN, ids, c_str, scan_no_flush ? (bulk ? 1 : 0) : -1);
    lap("scan");
    // processing time of the queue entries: the queue is drained from the back
    if (qtime.size() < C.size()) { pgrow(qtime, C.size() + 1024, -1); }
    // (a clause queued twice keeps the time of its last entry, like the sequential fill)
    { bool inc = true; for (size_t i = 1; i < strq.size(); ++i) { if (strq[i - 1] >= strq[i]) { inc = false; break; } }
      const size_t nq = strq.size();
      if (inc) { parallelFor(nq, [&](size_t b, size_t e, int) { for (size_t i = b; i < e; ++i) { qtime[strq[i]] = (int32_t)(nq - 1 - i); } }); }
      else { for (size_t i = 0; i < nq; ++i) { qtime[strq[i]] = (int32_t)(nq - 1 - i); } } }
    for (size_t k = 0; k < c_str.ent.size(); ++k) { c_str.ent[k].time = 0; }
    for (uint32_t c : ids) { if (c_str.slot[c] >= 0) { c_str.ent[c_str.slot[c]].time = (double)qtime[c]; } }
    lap("qtime");
    prefetchStrengthenClosure();
    lap("closure");
}

// strengthening_worker, Subsumption.cc:1250-1528 (all_strength_res == 0)
int Engine::strengtheningWorker(bool useHeap, int ignore)
{
    PhaseTimer pt(ph_str);
    const bool unlimited = !opt.limited;
    ensureHostLists();
    logClear(); dirty_str.clear();
    if (str_prepared) { str_prepared = false; }            // scan + closure were done while the subsumption commit ran
    else if (opt.strength) {
        std::vector<uint32_t> ids;
        for (uint32_t c : strq) {
            if ((C[c].flag & F_DEL) || !(C[c].flag & F_STR) || C[c].size < 2) { continue; }
            ids.push_back(c);
        }
        sortUnique(ids);
        prepareStrengthening(ids, false);
    } else { c_str.clear(); }
    DbgLap lap("str"); lap("prepare");
    const bool fine = lap.on && getenv("CP3_DEBUG_FINE") != nullptr;          // per-entry timers (they cost as much as the entries)
    int64_t t0 = now_ns();
    // Statically inert queue entries.  While no unit is propagated, lit_occurrence_count and the occurrence lists
    // are frozen during this pass (occurrence updates are deferred, Subsumption.cc:1400,1523), so an entry that
    // (a) has no candidate hit, (b) is no candidate target of anybody and (c) walks two lists without deleted
    // entries only adds |occ(min)| - 1 + |occ(~min)| steps - independent of everything else.  Those are
    // evaluated on all host cores; the ordered walk below just adds the number up when it passes them.
    pfill(contrib, strq.size(), -1); pfill(contrib_t, strq.size(), -1); tmin.resize(strq.size()); plan_ver.resize(strq.size());
    static const size_t static_min = getenv("CP3_STATIC_MIN") ? (size_t)atoi(getenv("CP3_STATIC_MIN")) : 4096;
    bool static_valid = opt.strength && strq.size() > static_min && !hasToPropagate();
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
        static const bool no_pre = getenv("CP3_NO_PRECOMPACT") != nullptr;
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
        static const bool no_lazy = getenv("CP3_NO_LAZY") != nullptr;
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
    // Walk plan of the real entries.  While the plan is valid (no unit, frozen counters, clean lists) the two list walks of an
    // entry are a pure function of its content: the step count is the length of the two lists and the candidates that can
    // hit are the cached hits in list order.  Both are evaluated here on all cores for the content the entry has NOW; the
    // ordered walk below uses the record when the entry's turn comes and its content version is still the planned one, and
    // then only validates and applies the candidate hits (the order-dependent part).  Everything else takes the walk itself.
    struct PlanEnt { uint32_t ver, scan; int32_t min; uint32_t steps1, steps2, cb, c1, c2; uint32_t t; uint32_t lit_off, lit_n; int32_t alt; };
    RawVec<PlanEnt> plan; std::vector<std::vector<Hit>> plan_c((size_t)nthreads + 1); std::vector<std::vector<PlanEnt>> plan_alt((size_t)nthreads + 1);
    static const bool no_plan = getenv("CP3_NO_PLAN") != nullptr;
    const bool planned = static_lazy && !no_plan && !reals.empty();
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
    size_t end = strq.size();
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
                            for (uint32_t k = pf.cb; k < pf.c2; ++k) { const uint32_t o = hc[k].clause; __builtin_prefetch(&C[o]); if (o < log_head.size()) { __builtin_prefetch(&log_head[o]); } if (o < qtime.size()) { __builtin_prefetch(&qtime[o]); } if (o < shrunk_mark.size()) { __builtin_prefetch(&shrunk_mark[o]); } }
                        }
                    }
                    if (ri > 2) {
                        const PlanEnt& pf = plan[ri - 2];
                        if (pf.min >= 0) { const Hit* hc = plan_c[pf.t].data(); for (uint32_t k = pf.cb; k < pf.c2; ++k) {
                            const uint32_t o = hc[k].clause; __builtin_prefetch(pool.data() + C[o].start);
                            if (o < qtime.size() && qtime[o] >= 0) { const size_t po = static_nq - 1 - (size_t)qtime[o]; if (po < contrib.size()) { __builtin_prefetch(&contrib[po]); __builtin_prefetch(&walk_min[po]); } } } }
                    }
                }
                if (ri > 8) {                                        // look ahead: the clause, its counters and its scan record
                    const uint32_t c2 = strq[reals[ri - 8]]; const ClauseInfo& ci2 = C[c2];
                    __builtin_prefetch(&ci2);
                    if (ri > 4) { const ClauseInfo& c3 = C[strq[reals[ri - 4]]]; const int32_t* l3 = pool.data() + c3.start; const uint32_t n3 = std::min<uint32_t>(c3.size, 6); for (uint32_t k = 0; k < n3; ++k) { __builtin_prefetch(&hot[l3[k] & ~1]); } }
                    if (c2 < c_str.slot.size()) { __builtin_prefetch(&c_str.slot[c2]); }
                }
            }
            --end; cur_end = end;
            if (end >= 24 && contrib[end - 24] < 0) {           // look ahead past the inert entries: the next real one
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
            if (end >= 16) {
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
    for (uint32_t c : strq) { if (c < qtime.size()) { qtime[c] = -1; } }
    if (lap.on) { fprintf(stderr, "  str lazy entries %lld, lazy still on %d\n", (long long)n_lazy, (int)static_lazy); }
    static_pass = false; static_lazy = false; n_fast += n_lazy; n_static += n_lazy; n_lazy = 0;
    if (getenv("CP3_DEBUG_PRE")) { fprintf(stderr, "strpass q=%zu steps=%lld removed=%d pre=%zu active=%d static=%lld\n", strq.size(), (long long)str_steps, removedLiterals, pre_lists.size(), (int)pre_active, (long long)n_static); }
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
        sortUnique(ids);
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
    while (ok && (!subq.empty() || !strq.empty())) {
        if (!subq.empty()) {
            overlap_str = doStrengthen && !strq.empty();
            subsumptionWorker(useHeap, ignore);
            overlap_str = false;
            clearQueue(subq, F_SUB);
        }
        if (!strq.empty()) {
            if (!doStrengthen) { clearQueue(strq, F_STR); continue; }
            if (opt.naive_strength) { naiveStrengthening(useHeap, ignore); } else { strengtheningWorker(useHeap, ignore); }
            clearQueue(strq, F_STR);
        }
        str_prepared = false;
    }
    t_sub.modified = t_sub.modified || t_up.modified;   // propagation.appliedSomething(): its last call
    if (!t_sub.modified) { unsuccessful(t_sub); }
    return t_sub.modified;
}

// ------------------------------------------------------------------------------------------------ stand-alone BCE
// K9 over every variable that can be tested: masks[bce_woff[v] + a*W + (b>>5)] bit (b&31) = the resolvent of the a-th clause
// of occ(+v) with the b-th clause of occ(~v) is not a tautology (tautologicResolvent, BCE.cc:624-646).
void Engine::bceBuildMasks()
{
    ensureHostLists();
    flushDevice();
    const size_t nv = (size_t)nvars;
    bce_woff.resize(nv + 1); bce_row_var = -1;
    std::vector<uint32_t> vars, p_off{0}, n_off{0}; std::vector<uint64_t> w_off{0};
    static const uint64_t row_cap = getenv("CP3_BCE_MAXWORDS") ? (uint64_t)atoll(getenv("CP3_BCE_MAXWORDS")) : (1ull << 22);   // words per variable
    for (size_t v = 0; v < nv; ++v) {
        bce_woff[v] = UINT64_MAX;
        const uint32_t np = hot[2 * v].n, nn = hot[2 * v + 1].n;
        if (frozen[v] || assigns[v] != L_UNDEF || np == 0 || nn == 0) { continue; }
        const uint64_t words = (uint64_t)np * ((nn + 31) / 32);
        if (words > row_cap) { continue; }                          // very long lists: rows on demand
        bce_woff[v] = w_off.back();
        vars.push_back((uint32_t)v); p_off.push_back(p_off.back() + np); n_off.push_back(n_off.back() + nn); w_off.push_back(w_off.back() + words);
    }
    bce_masks.resize((size_t)w_off.back() + 1);
    if (vars.empty()) { return; }
    RawVec<uint32_t> p_refs, n_refs; p_refs.resize(p_off.back() + 1); n_refs.resize(n_off.back() + 1);
    parallelFor(vars.size(), [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            const uint32_t v = vars[i];
            memcpy(p_refs.data() + p_off[i], occ_pool.data() + occ[2 * v].off, sizeof(uint32_t) * hot[2 * v].n);
            memcpy(n_refs.data() + n_off[i], occ_pool.data() + occ[2 * v + 1].off, sizeof(uint32_t) * hot[2 * v + 1].n);
        }
    }, 16384);
    ++n_bce_batches;
    int64_t t0 = now_ns();
    if (cp3g_bce_masks(gpu, (uint32_t)vars.size(), vars.data(), p_off.data(), p_refs.data(), n_off.data(), n_refs.data(), w_off.data(), bce_masks.data()) != CP3G_OK) { die("bce_masks failed"); }
    host_ns_gpuwait += now_ns() - t0;
}
const uint32_t* Engine::bceRow(int v, uint32_t a)
{
    const uint32_t W = (hot[2 * v + 1].n + 31) / 32;
    if (bce_woff[(size_t)v] != UINT64_MAX) { return bce_masks.data() + bce_woff[(size_t)v] + (uint64_t)a * W; }
    if (bce_row_var == v && bce_row_idx == a) { return bce_row.data(); }
    // a variable without a mask block (lists too long for the batch): this one row
    ++n_bce_rows;
    const uint32_t vv = (uint32_t)v, po[2] = {0, 1}, no[2] = {0, hot[2 * v + 1].n}; const uint64_t wo[2] = {0, W};
    const uint32_t pr = occ_pool[occ[2 * v].off + a];
    bce_row.assign((size_t)W + 1, 0u);
    int64_t t0 = now_ns();
    if (cp3g_bce_masks(gpu, 1, &vv, po, &pr, no, occ_pool.data() + occ[2 * v + 1].off, wo, bce_row.data()) != CP3G_OK) { die("bce_masks failed"); }
    host_ns_gpuwait += now_ns() - t0;
    bce_row_var = v; bce_row_idx = a;
    return bce_row.data();
}
bool Engine::bceNonTaut(int right, uint32_t i, uint32_t j)
{
    const int v = right >> 1;
    const uint32_t a = (right & 1) ? j : i, b = (right & 1) ? i : j;          // row = clause of occ(+v), column = clause of occ(~v)
    return (bceRow(v, a)[b >> 5] >> (b & 31)) & 1u;
}

// blockedClauseElimination, BCE.cc:239-601, with -bce-bce and without CLE / CLA / BCM: the literal heap, the rounds and the
// step counter are the reference's; every tautologicResolvent() answer is a bit of the K9 masks
void Engine::blockedClauseElimination()
{
    const int nl = 2 * nvars;
    // Riss::Heap<LitOrderBCEHeapLt> (riss/mtl/Heap.h:37-183, BCE.h:27-36)
    std::vector<int> H, idx((size_t)nl + 1, -1);
    auto lt = [&](int x, int y) { return opt.bce_compl ? hot[x ^ 1].cnt < hot[y ^ 1].cnt : hot[x].cnt < hot[y].cnt; };
    auto up = [&](int i) { const int x = H[(size_t)i]; int p = (i - 1) >> 1; while (i != 0 && lt(x, H[(size_t)p])) { H[(size_t)i] = H[(size_t)p]; idx[(size_t)H[(size_t)p]] = i; i = p; p = (p - 1) >> 1; } H[(size_t)i] = x; idx[(size_t)x] = i; };
    auto down = [&](int i) {
        const int x = H[(size_t)i]; const int n = (int)H.size();
        while (2 * i + 1 < n) {
            const int l = 2 * i + 1, r = 2 * i + 2; const int child = (r < n && lt(H[(size_t)r], H[(size_t)l])) ? r : l;
            if (!lt(H[(size_t)child], x)) { break; }
            H[(size_t)i] = H[(size_t)child]; idx[(size_t)H[(size_t)i]] = i; i = child;
        }
        H[(size_t)i] = x; idx[(size_t)x] = i;
    };
    auto insert = [&](int x) { idx[(size_t)x] = (int)H.size(); H.push_back(x); up(idx[(size_t)x]); };
    auto removeMin = [&]() { const int x = H[0]; H[0] = H.back(); idx[(size_t)H[0]] = 0; idx[(size_t)x] = -1; H.pop_back(); if (H.size() > 1) { down(0); } return x; };
    std::vector<uint32_t> nextRound((size_t)nl + 1, 0u); uint32_t nr_step = 0;     // MarkArray: every literal is "current" at step 0
    std::vector<int> nextRoundLits;
    for (int v = 0; v < nvars; ++v) {
        if (frozen[v]) { continue; }
        if (hot[2 * v].cnt > 0) { nextRoundLits.push_back(2 * v); }
        if (hot[2 * v + 1].cnt > 0) { nextRoundLits.push_back(2 * v + 1); }
    }
    ++ma_step;                                                     // data.ma.nextStep(), BCE.cc:269
    bceBuildMasks();
    const bool unlimited = !opt.limited;
    do {
        for (int l : nextRoundLits) { if (nextRound[(size_t)l] == nr_step) { insert(l); } }
        nextRoundLits.clear();
        ++nr_step;
        while (!H.empty() && (unlimited || opt.bce_limit > bce_steps)) {
            const int right = removeMin();
            if (frozen[right >> 1] || value(right) != L_UNDEF) { continue; }
            bce_testedLits++;
            const int left = right ^ 1;
            for (uint32_t i = 0; i < hot[right].n; ++i) {
                if (hasToPropagate()) { break; }
                const uint32_t c = occ_pool[occ[right].off + i];
                if ((C[c].flag & F_DEL) || (!opt.bce_binary && C[c].size == 2)) { continue; }
                bool isBlocked = true;
                const uint32_t nleft = hot[left].n; const uint32_t* ll = occ_pool.data() + occ[left].off;
                for (uint32_t j = 0; j < nleft; ++j) {
                    if (isDel(ll[j])) { continue; }
                    bce_steps++;
                    if (bceNonTaut(right, i, j)) { isBlocked = false; break; }
                }
                if (isBlocked) {
                    setDeleted(c);
                    removedClause(c, false, false, VAR_UNDEF);
                    bce_remBCE++;
                    const int32_t* l = CL(c);
                    for (uint32_t k = 0; k < C[c].size; ++k) {
                        const int nk = l[k] ^ 1;
                        if (idx[(size_t)nk] >= 0) { up(idx[(size_t)nk]); down(idx[(size_t)nk]); }
                        else if (nextRound[(size_t)nk] != nr_step) { nextRoundLits.push_back(nk); nextRound[(size_t)nk] = nr_step; }
                    }
                    successful(t_bce);
                    extClause(c, right);
                }
            }
            // (BCE.cc:565-590: units can only come from CLE / BCM, which this path refuses; nothing to propagate here)
            if (hasToPropagate()) { die("internal: unit during blocked clause elimination"); }
        }
        const int64_t gc0 = n_gc;
        checkGarbage();
        // a collection filters the deleted entries out of the lists (CoprocessorTypes.h:1396-1550): new list positions, new masks
        if (n_gc != gc0 && !nextRoundLits.empty() && (unlimited || opt.bce_limit > bce_steps)) { bceBuildMasks(); }
    } while (!nextRoundLits.empty() && (unlimited || opt.bce_limit > bce_steps));
    bce_masks.clear(); bce_masks.shrink_to_fit();
}

// BlockedClauseElimination::process, BCE.cc:604-622
bool Engine::bceProcess()
{
    if (!opt.bce_bce) { return false; }
    if (!performSimplification(t_bce)) { return false; }
    t_bce.modified = false;
    // (the size gate needs nTotLits() > cp3_bce_lits, which is 0 after cleanSolver: it never fires, like cp3_susi_* / cp3_bve_*)
    propagationProcess(true, false, VAR_UNDEF);
    t_bce.modified = t_bce.modified || t_up.modified;
    if (!ok) { return t_bce.modified; }
    blockedClauseElimination();
    return t_bce.modified;
}

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

template <class A> static inline int indexOf(const A& a, uint32_t x)
{
    for (size_t i = 0; i < a.size(); ++i) { if (a[i] == x) { return (int)i; } }
    return -1;
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
    static const bool dbg = getenv("CP3_DEBUG_PAR") != nullptr || getenv("CP3_DEBUG_BVE") != nullptr;
    int64_t tg = 0, tc = 0, ta = 0, tr = 0, trm = 0, ts = 0, tl = dbg ? now_ns() : 0;
    auto lapt = [&](int64_t& acc) { if (dbg) { const int64_t n2 = now_ns(); acc += n2 - tl; tl = n2; } };
    struct Rep { bool on; int64_t &tg, &tc, &ta, &tr, &trm, &ts; ~Rep() { if (on) { fprintf(stderr, "  bveWorker: heap+gates %.1f ms, clean %.1f, anticipate %.1f, resolveSet %.1f, removeClauses %.1f, inner subsumption %.1f\n", tg / 1e6, tc / 1e6, ta / 1e6, tr / 1e6, trm / 1e6, ts / 1e6); } } } rep{dbg, tg, tc, ta, tr, trm, ts};
    static const bool no_eval = getenv("CP3_NO_BVE_EVAL") != nullptr;
    // (-bve_early makes the anticipation depend on the global growth budget; in sharded mode every rank evaluates the whole batch -
    //  the replicas hold the identical data base, nothing has to be exchanged)
    const bool use_eval = !no_eval && !opt.bve_early;
    pairs_single = use_eval;
    struct InWorker { bool& f; explicit InWorker(bool& x) : f(x) { f = true; } ~InWorker() { f = false; } } in_worker(in_bve_worker);
    static const bool no_pf = getenv("CP3_NO_BVE_PREFETCH") != nullptr;
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
    static const uint64_t par_min = getenv("CP3_TOUCHED_MIN") ? (uint64_t)atoll(getenv("CP3_TOUCHED_MIN")) : 262144;
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
    static const bool dbg = getenv("CP3_DEBUG_PAR") != nullptr || getenv("CP3_DEBUG_BVE") != nullptr;
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

// ------------------------------------------------------------------------------------------------ driver
// data.init + initializePreprocessor + sortClauses (Coprocessor.cc:100-118,1243-1283,1447-1463).
// The occurrence lists are built on the GPU (K2) and copied back - they start in clause order.
void Engine::initData()
{
    PhaseTimer pt(ph_init);
    const int n2 = 2 * nvars;
    pfill(occ, (size_t)n2, OccList()); pfill(hot, (size_t)n2, LitHot()); stale.assign((size_t)n2, 0u);
    stalebits.assign(((size_t)n2 >> 6) + 1, 0ull);
    modTimer.assign((size_t)std::max(nvars, 1), 0); ma.assign((size_t)std::max(n2, 1), 0);
    var_touch.assign((size_t)std::max(nvars, 1), 0); bve_inpend.assign((size_t)std::max(nvars, 1), 0); bve_pending.clear();
    modStep = 0; numberOfCls = 0; qhead = 0;
    subq.clear(); strq.clear();
    int64_t ti0 = now_ns();
    uploadAll();
    ph_upload += now_ns() - ti0; ti0 = now_ns();
    std::vector<uint32_t> off((size_t)n2 + 1, 0);
    if (cp3g_download_occ(gpu, off.data(), nullptr) != CP3G_OK) { die("download_occ failed"); }
    occ_pool.resize((size_t)off[n2] + 16);                 // filled by ensureHostLists() / the device-side commit of the first pass
    lists_on_device = true;
    for (int x = 0; x < n2; ++x) {
        occ[x].off = off[x]; hot[x].n = occ[x].cap = off[x + 1] - off[x];
        hot[x].cnt = (int32_t)hot[x].n;
    }
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
            subq.insert(subq.end(), clauses.begin(), clauses.end()); strq.insert(strq.end(), clauses.begin(), clauses.end());
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
    { PhaseTimer ps(ph_simplify); if (!solverSimplify()) { return; } }
    sortPermuted();
    initData();
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
    if (ok && hasToPropagate()) { propagationProcess(false, false, VAR_UNDEF); }
    if (opt.dense) { denseCompress(); }           // very last step, Coprocessor.cc:492-496
    if (ok) { filterDeleted(clauses); }           // reSetupSolver, CoprocessorTypes.h:937-1008
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
    ST("spec_requests", n_spec_requests) ST("spec_rounds", n_spec_rounds) ST("spec_used", n_spec_used)
    ST("host_commit_ns", host_ns_commit) ST("host_gpuwait_ns", host_ns_gpuwait)
    if (!strncmp(s, "gpu_", 4)) { return gpu ? cp3g_counter(gpu, s + 4) : 0; }
#undef ST
    return -1;
}

} // namespace cp3
