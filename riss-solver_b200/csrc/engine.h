// engine.h - host side of the drop-in: Coprocessor's occurrence-list clause data base and the
// *ordered commit* of subsumption / strengthening / unit propagation / BVE.
//
// Division of labour: every clause-against-clause test (subsumption merge, self-subsuming
// resolution merge, resolvent size, resolvent generation) is computed by a CUDA kernel behind the cp3g_*
// scan service, in bulk, on a snapshot of the data base.  The host never merges two clauses.  What the host
// does is what makes the result identical to the sequential reference: it walks the reference's own work
// queues / occurrence lists / variable heap in the reference's order, consumes the precomputed hits,
// validates them against the edits committed since the snapshot (clauses only ever shrink or die, so a
// snapshot hit can only be invalidated, never missed) and maintains the reference's counters, flags and
// quirks (Subsumption.cc:85-180,220-314,1250-1544; Propagation.cc:32-131; BVE.cc:170-637).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <unordered_map>
#include <vector>
#include <queue>
#include <thread>
#include <memory>
#include <functional>
#include "../../include/cp3b200.h"

namespace cp3 {

// literal code like riss/core/SolverTypes.h:65-98
static inline int lvar(int x) { return x >> 1; }
static inline int lneg(int x) { return x ^ 1; }
static inline int toDimacs(int x) { return (x & 1) ? -((x >> 1) + 1) : ((x >> 1) + 1); }

enum : uint8_t { F_DEL = 1, F_SUB = 2, F_STR = 4 };
enum { L_TRUE = 0, L_FALSE = 1, L_UNDEF = 2 };
static const int VAR_UNDEF = -1;

// CP3Config subset (coprocessor/CP3Config.cc:56-112,196-224,560-577): same names, same defaults
struct Options {
    bool enabled = false, up = false, subsimp = false, bve = false, limited = true;
    int iters = 1;
    int cp3_vars = 5000000, cp3_cls = 30000000, cp3_lits = 50000000;
    bool strength = true, naive_strength = false;      // -cp3_strength, -naive_strength (CP3Config.cc:560-561)
    int64_t sub_limit = 300000000, str_limit = 300000000; int call_inc = 200;
    int64_t bve_limit = 25000000;
    bool bve_unlimited = false, bve_strength = true, bve_gates = true, bve_force_gates = false,
         bve_fdepOnly = false, bve_BCElim = false, bve_early = false, bce_only = false;
    int bve_heap = 0, bve_cgrow = 0, bve_cgrow_t = 1000, bve_heap_updates = 1;
    int threads = 0;
    bool stats = false;
    bool bce = false, bce_bce = true, bce_compl = true, bce_binary = false; int64_t bce_limit = 100000000;   // -bce, -bce-bce, -bce-compl, -bce-bin, -bce-limit (CP3Config.cc:86,266-269)
    bool dense = false, dense_keep_set = false; int dense_frag = 0;   // -dense, -cp3_keep_set, -cp3_dense_frag (CP3Config.cc:100,344,345)
    double gc_frac = 0.20;
    int device = 0;              // -cp3_gpu_device=N
    // returns 1 known & applied, 0 unknown, -1 known but unsupported value on the GPU path
    int parse(const char* opt);
    void preset(const char* name);
};

struct Tech { int penalty = 0, peak = 0; bool modified = false; };

// Size thresholds and switches of the host engine: where a data-parallel form takes over from the plain ordered walk.  One place,
// read from the environment once per engine (CP3_<NAME>); the tests force the large-queue forms onto tiny formulas with them.
struct Tuning {
    int threads;                 // CP3_THREADS: host cores of the parallel phases (default: all, at most 24)
    size_t ingest_min;           // CP3_INGEST_MIN: stream length from which CPaddLits goes through the device bulk ingest K0
    uint32_t bulk_min;           // CP3_BULK_MIN: live clauses from which a queue of all of them is scanned by the full-queue kernel
    size_t parsub_min;           // CP3_PARSUB_MIN: queue length from which the subsumption commit is time-indexed (device or all cores)
    size_t static_min;           // CP3_STATIC_MIN: queue length from which a strengthening pass gets the static plan
    uint64_t touched_min;        // CP3_TOUCHED_MIN: occurrences from which the touched-variable queueing of BVE runs on all cores
    uint32_t devcommit_maxlist;  // CP3_DEVCOMMIT_MAXLIST: longest list with deaths that one device thread may replay
    size_t par_min;              // CP3_PAR_MIN: list length from which id sorting, scan-cache build and occurrence updates run on all cores
    uint64_t bce_maxwords;       // CP3_BCE_MAXWORDS: K9 mask words per variable in the batch (longer: rows on demand)
    bool no_devcommit, no_overlap, no_precompact, no_lazy, no_plan, no_components, no_devstatic, no_bve_eval, no_bve_prefetch;   // CP3_NO_<...>=1: switch one form off
    bool debug_par, debug_bve, debug_fine, debug_ondemand, debug_pre;                                // CP3_DEBUG_<...>: phase timers on stderr
    Tuning();
};

// std::vector whose resize() leaves trivially constructible elements uninitialised: the big work arrays are filled by a
// device download or by a parallel loop right after they are sized, a serial zero fill first would only cost bandwidth
// Blocks of 4 MB and more are 2 MB aligned and marked MADV_HUGEPAGE: the queue walks and per-literal gathers are random accesses
// over arrays of 50-900 MB, where every access misses a 4 KB-page TLB (and the first touch faults page by page).  Only the engine's
// own blocks are touched - no process-wide policy.
void* hugeAlloc(size_t bytes);
template <class T> struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { typedef NoInitAlloc<U> other; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
    T* allocate(size_t n) { return (T*)hugeAlloc(n * sizeof(T)); }
    void deallocate(T* p, size_t) { free(p); }
    template <class U> void construct(U* p) { (void)p; }
    template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
};
template <class T> using RawVec = std::vector<T, NoInitAlloc<T>>;
// the same blocks for vectors that keep std::vector's value initialisation (per-clause / per-variable tables that are read at random)
template <class T> struct HugeAlloc : std::allocator<T> {
    template <class U> struct rebind { typedef HugeAlloc<U> other; };
    HugeAlloc() = default;
    template <class U> HugeAlloc(const HugeAlloc<U>&) {}
    T* allocate(size_t n) { return (T*)hugeAlloc(n * sizeof(T)); }
    void deallocate(T* p, size_t) { free(p); }
};
template <class T> using HVec = std::vector<T, HugeAlloc<T>>;

struct Hit { uint32_t clause; int32_t lit; };

// occurrence list: slice of one pooled array; grows by relocation to the pool end
// (the hot per-literal fields - lit_occurrence_count and the list length - live in a separate 8-byte record so
// that the queue walks, which are bound by cache misses on them, touch 16 MB instead of 64 MB at 1M variables)
struct OccList { uint32_t off = 0, cap = 0; };
struct LitHot { int32_t cnt = 0; uint32_t n = 0; };
struct ListRef { uint32_t& off; uint32_t& cap; uint32_t& n; };

class Engine {
public:
    Engine();
    ~Engine();
    Options opt;
    const Tuning tune;

    // --- Solver-side interface used by the CP* layer
    void addLit(int dimacsLit);
    void addStream(const int32_t* s, size_t n);
    void freeze(int dimacsVar);
    void preprocess();
    // Dense (coprocessor/techniques/Dense.cc): DIMACS literal of the compressed formula for a literal of the original
    // one (Preprocessor::importLit, Coprocessor.cc:1174-1187; 0 = no replacement)
    int importLit(int dimacsLit) const;
    bool okay() const { return ok; }
    int nVars() const { return nvars; }
    size_t output(int32_t* out, size_t cap) const;
    size_t nTrail() const { return trail.size(); }
    size_t nLiveClauses() const;
    size_t nLiveLits() const;
    size_t undoStack(int32_t* out, size_t cap) const;
    int64_t stat(const char* name) const;
    int litValue(int dimacsLit) const;
    void extendModel(std::vector<uint8_t>& model) const;   // CoprocessorData::extendModel, CoprocessorTypes.h:1665
    int setDistributed(int rank, int world, const void* id);

private:
    // ---------------- solver state
    int nvars = 0; bool ok = true; int qhead = 0;
    // Riss::Compression (riss/utils/Compression.h): mapping[original var] = compressed var or -1, and the values
    // of the removed assigned variables; empty until Dense compressed the formula
    std::vector<int32_t> comp_map; std::vector<uint8_t> comp_val;
    void denseCompress();
    void extendByUndo(std::vector<uint8_t>& model) const;
    std::vector<uint8_t> assigns, frozen;
    std::vector<int> trail;
    RawVec<int32_t> pool;                            // literal pool (codes)
    // per clause: pool offset, size + flags, edit counter, scan count at the last edit - one 16-byte record
    struct ClauseInfo { uint32_t start; uint32_t size : 24; uint32_t flag : 8; uint32_t ver; uint32_t stamp; };
    RawVec<ClauseInfo> C;
    std::vector<uint32_t> clauses;                   // solver->clauses
    double ca_size = 0, ca_wasted = 0;
    int simpDB_assigns = -1; int64_t simpDB_props = 0;
    std::vector<int> cur;
    // two-watched-literal lists while clauses are added (Solver::attachClause / propagate, riss/core/Solver.cc:692-708,1809-1937):
    // watches[p] = clauses that watch ~p; built at the first unit.  propagate() permutes the literals of the clauses it
    // visits; those are re-sorted where the reference does it (Preprocessor::sortClauses, Coprocessor.cc:118)
    struct Watcher { uint32_t cref; int32_t blocker; };          // top bit of cref: binary clause
    std::vector<std::vector<Watcher>> watches; bool ing_built = false;
    std::vector<uint32_t> permuted; std::vector<uint8_t> perm_mark;
    void attachClause(uint32_t c);
    void sortPermuted();

    // ---------------- CoprocessorData mirror
    bool data_ready = false; int data_nvars = 0;      // CoprocessorData::numberOfVars: 0 until the first data.init()
    RawVec<OccList> occ; RawVec<LitHot> hot; RawVec<uint32_t> occ_pool;
    RawVec<uint32_t> stale; std::vector<uint64_t> stalebits;   // deleted clauses still listed in occ[l] (+ bitmap of >0)
    inline ListRef L(int x) { return ListRef{occ[x].off, occ[x].cap, hot[x].n}; }
    inline bool hasStale(int x) const { return (stalebits[(uint32_t)x >> 6] >> (x & 63)) & 1; }
    inline void staleInc(int x) { if (stale[x]++ == 0) { stalebits[(uint32_t)x >> 6] |= 1ull << (x & 63); } }
    inline void staleDec(int x) { if (stale[x] > 0 && --stale[x] == 0) { stalebits[(uint32_t)x >> 6] &= ~(1ull << (x & 63)); } }
    inline void staleClear(int x) { stale[x] = 0; stalebits[(uint32_t)x >> 6] &= ~(1ull << (x & 63)); }
    int numberOfCls = 0;
    RawVec<uint32_t> subq, strq; std::vector<int32_t> undo;      // (no zero fill when initData sizes the queues)
    RawVec<uint32_t> modTimer; uint32_t modStep = 0;
    RawVec<uint32_t> ma; uint32_t ma_step = 0;

    // ---------------- technique state
    Tech t_sub, t_up, t_bve, t_bce;
    int bce_steps = 0, bce_testedLits = 0, bce_remBCE = 0; int64_t n_bce_batches = 0, n_bce_rows = 0;
    int64_t sub_steps = 0, str_steps = 0, sub_lim = 0, str_lim = 0; int limitIncreases = 0;
    int subsumedClauses = 0, subsumedLiterals = 0, removedLiterals = 0;
    int up_last = 0, up_removedClauses = 0, up_removedLiterals = 0;
    uint32_t bve_modtime = 0; int64_t bve_steps = 0;
    HVec<int> heap, hidx;
    int bve_removedClauses = 0, bve_removedLiterals = 0, bve_createdClauses = 0, bve_createdLiterals = 0,
        bve_testedVars = 0, bve_anticipations = 0, bve_eliminatedVars = 0, bve_removedBC = 0, bve_skippedVars = 0,
        bve_unitsEnqueued = 0, bve_foundGates = 0, bve_nInc = 0, bve_nDec = 0, bve_nKeep = 0, bve_totallyAdded = 0;
    std::vector<int> pos_stats, neg_stats, touched;
    struct OccUpd { uint32_t cr; int l; };
    std::vector<OccUpd> occ_upd; std::vector<uint32_t> sub_occ_updates;

    // ---------------- device replica + scan cache
    cp3g* gpu = nullptr;
    bool dev_uploaded = false;
    int64_t n_gc = 0, n_compactions = 0, n_bulk_ingests = 0, n_tfast = 0;
    uint32_t dev_arena_ncl = 0; bool dev_arena_valid = false;       // the device still holds the arena of the bulk ingest
    struct BvePr { uint32_t cp, cn; int size; uint64_t off; };
    std::vector<BvePr> bve_prs; std::vector<int32_t> bve_rl; std::vector<uint32_t> bve_vv, bve_pr, bve_nr; std::vector<uint64_t> bve_off; std::vector<int> bve_col;
    std::vector<uint32_t> pend_del, pend_shrunk; uint32_t dev_ncl = 0;    // edits not yet on the device
    RawVec<uint8_t> shrunk_mark;
    uint32_t scan_id = 0;                            // number of scans issued so far
    // literals removed per clause during the current pass: per-clause chains in one flat node array
    struct LogNode { uint32_t stamp; int32_t lit; int32_t next; uint32_t pending; };
    HVec<LogNode> log_nodes; HVec<int32_t> log_head; std::vector<uint32_t> log_touched;
    uint32_t pend_epoch = 1;                         // LogNode.pending == pend_epoch: occ entry removal still pending
    void logClear();
    void logPush(uint32_t c, int lit, bool pendingOcc);
    // one scanned content version of a clause: keyed by the edit counter (ver) when the device copy was
    // scanned by id, or by the literal content itself (ver == UINT32_MAX) for speculative versions
    struct CacheEnt { uint32_t ver, scan, hb, he; int32_t next; uint32_t lit_off, lit_n; double time; };
    struct ScanCache {
        RawVec<int32_t> slot; HVec<CacheEnt> ent; HVec<Hit> hits; HVec<int32_t> lits;
        // a full-queue bulk scan covers every clause that was live at that time: clauses without an entry had
        // no hit (no per-request record is kept for them)
        bool bulk_all = false; uint32_t bulk_scan = 0, bulk_ncl = 0; CacheEnt empty;
        std::vector<uint32_t> used;                  // clauses whose slot is set (sparse reset: inner BVE calls are tiny)
        void clear();
        inline void setSlot(uint32_t c, int32_t v) { if (slot[c] < 0) { used.push_back(c); } slot[c] = v; }
    };
    uint32_t n_live = 0;                             // clauses without F_DEL
    std::vector<uint64_t> delbits;                   // F_DEL as a cache-resident bitmap for the list walks
    inline bool isDel(uint32_t c) const { return (delbits[c >> 6] >> (c & 63)) & 1; }
    struct SpecReq { uint32_t clause; uint32_t off, n; double time; };   // literals in spec_lits[off..off+n)
    std::vector<int32_t> spec_lits;
    // grouped: the requests of one clause are consecutive (the closure builds them clause by clause): records are linked on all cores
    void scanSpeculative(int kind, const std::vector<SpecReq>& reqs, ScanCache& cache, bool grouped = false);
    void prefetchStrengthenClosure();
    RawVec<int32_t> qtime;                           // strengthening queue: processing time of a pending clause
    int64_t n_spec_requests = 0, n_spec_rounds = 0, n_spec_used = 0, n_static = 0;
    RawVec<int32_t> contrib;                         // per queue position: step contribution of a statically inert entry, or -1
    RawVec<int32_t> contrib_t, tmin; RawVec<uint32_t> plan_ver;   // same for candidate targets, valid while the content version is the planned one
    std::vector<uint64_t> is_target;                 // clauses that some cached hit may modify
    int nthreads = 1;
    // up-front cleaning of certainly-walked lists (strengtheningWorker) and what is needed to take it back
    struct PreList { int lit; uint32_t snap_off, old_n, extra; };
    std::vector<PreList> pre_lists; std::vector<uint32_t> pre_snap; std::vector<uint8_t> scanned_dyn;
    std::vector<int32_t> pre_index; RawVec<int32_t> walk_min; RawVec<uint32_t> walk_count;   // per literal / per queue position
    bool pre_active = false, static_pass = false; size_t static_nq = 0;
    // lazy form of the static plan: the inert entries' F_STR flags are cleared up front (in parallel) and the ordered
    // walk only adds their step contribution; cur_end = queue position being processed (everything above is done)
    bool static_lazy = false; size_t cur_end = 0; int64_t n_lazy = 0;
    std::priority_queue<uint32_t> lazy_flipped;      // inert entries below the walk that an unpredicted hit turned into real ones
    std::vector<int32_t> last_static_walk;           // per literal: highest queue position of an inert entry that walks its list
    void leaveLazy();                                // make the deferred state explicit (a unit is about to be propagated)
    bool lazyPending(uint32_t c) const;              // inert entry whose turn has not come yet (its flag is logically still set)
    void rollbackPrecompact();
    void restoreOneList(int x);
    void unpredictedEdit(uint32_t c);               // a clause the static plan relied on was modified
    template <class F> void parallelFor(size_t n, F f, size_t grain = 65536);
    // persistent workers behind parallelFor (a fixpoint makes ~60 such calls; starting 16 threads for each costs more than many of
    // the loops themselves).  One job at a time: a second caller - the scan helper thread - starts its own threads instead.
    struct WorkerPool;
    WorkerPool* pool_ = nullptr;
    void poolRun(int tasks, const std::function<void(int)>& job);
    template <class F> void parallelForG(size_t n, size_t grain, F f) { parallelFor(n, f, grain); }
    // stable parallel filter: out = [x in src | pred(x)] (count per chunk, prefix, write)
    // size a work array and fill it on all cores
    template <class V, class T> void pfill(V& v, size_t n, T val) { v.resize(n); parallelFor(n, [&](size_t b, size_t e, int) { std::fill(v.begin() + (std::ptrdiff_t)b, v.begin() + (std::ptrdiff_t)e, val); }); }
    // grow a work array to n elements, new elements = val (existing ones are kept)
    template <class V, class T> void pgrow(V& v, size_t n, T val) { const size_t o = v.size(); if (n <= o) { return; } v.resize(n); parallelFor(n - o, [&](size_t b, size_t e, int) { std::fill(v.begin() + (std::ptrdiff_t)(o + b), v.begin() + (std::ptrdiff_t)(o + e), val); }); }
    template <class V, class P> void parallelFilter(const V& src, std::vector<uint32_t>& out, P pred);
    void sortUniqueIds(std::vector<uint32_t>& ids);      // sort + unique of a clause id list; long lists through a bitmap on all cores
    bool isIncreasing(const uint32_t* v, size_t n);
    template <class V> bool isIncreasing(const V& v) { return isIncreasing(v.data(), v.size()); }
    ScanCache c_sub, c_str;
    std::vector<uint32_t> dirty_str;                 // pending strengtheners edited after their scan
    std::vector<cp3g_hit> hitbuf;
    // BVE pair cache
    RawVec<uint32_t> var_touch; RawVec<uint8_t> bve_inpend; std::vector<int> bve_pending;
    inline void bveTouch(int v) { ++var_touch[v]; if (!bve_inpend[v] && inHeap(v)) { bve_inpend[v] = 1; bve_pending.push_back(v); } }
    // K5/K6 results per variable: views into per-batch blocks (one allocation per batch, not per variable)
    template <class T> struct Span { const T* p = nullptr; size_t n = 0; size_t size() const { return n; } const T& operator[](size_t i) const { return p[i]; } };
    struct PairEnt { uint32_t touch = 0; bool valid = false, has_res = false;
                     Span<uint32_t> pos, neg; Span<int16_t> sizes; Span<uint32_t> roff; Span<int32_t> rlits; };
    struct PairBlock { std::vector<uint32_t> p_refs, n_refs, roff; std::vector<int16_t> sizes; std::vector<int32_t> rlits; };
    std::vector<std::unique_ptr<PairBlock>> pair_blocks;
    std::vector<int32_t> pc_slot; std::vector<PairEnt> pc_ent;
    inline PairEnt* pairLookup(int v) { return (v < (int)pc_slot.size() && pc_slot[v] >= 0) ? &pc_ent[pc_slot[v]] : nullptr; }
    std::vector<int> scratch_col;
    // K5'/K6' results per variable (cp3g_bve_eval / cp3g_bve_stream): the whole bve_worker evaluation of a variable on the
    // lists as they were at batch time.  Valid while the touch counter and the two lists themselves are unchanged.
    struct EvalEnt { bool has_stream = false; cp3g_bve_rec rec;
                     Span<uint32_t> pos0, neg0, pos1, neg1; Span<uint16_t> pst, nst, rsz; Span<int32_t> rlits; };   // a view, built on demand
    struct EvalBlock { std::vector<uint32_t> vars, p_off, n_off; std::vector<cp3g_bve_rec> recs; RawVec<uint32_t> p0, n0, p1, n1; RawVec<uint16_t> pst, nst;
                       std::vector<int32_t> sidx; std::vector<uint64_t> res_off, lit_off; RawVec<uint16_t> rsz; RawVec<int32_t> rlits; };
    // per variable: where its record lives (16 bytes; the records themselves stay in the batch blocks).  applied: the lists
    // already have the rewritten form (the record was used once), compare against that form from now on
    struct EvRef { int32_t blk; uint32_t idx, touch, applied; };
    std::vector<std::unique_ptr<EvalBlock>> eval_blocks; RawVec<EvRef> ev_ref; EvalEnt ev_view;
    EvalEnt* evalFor(int v);                            // cached record of v, or one more batch (v + the waiting heap variables)
    bool in_bve_worker = false;
    bool sharded = false;                               // scan service split over several ranks (CPsetDistributed)
    bool pairs_single = false;                          // the pair-matrix path serves single variables only (fallback of the evaluation)
    int64_t n_eval_batches = 0, n_eval_used = 0, n_eval_host = 0, n_eval_stale = 0, n_stream_used = 0;
    int64_t n_scan_batches = 0, n_ondemand = 0, n_fast = 0, n_slow = 0, n_pair_batches = 0, n_resolve_batches = 0;
    int64_t host_ns_commit = 0, host_ns_gpuwait = 0, dbg_slow_ns = 0, dbg_hits_ns = 0, dbg_step_ns = 0, dbg_plan_ns = 0;
    int64_t ph_init = 0, ph_sub = 0, ph_str = 0, ph_prefetch = 0, ph_total = 0, ph_upload = 0, ph_download = 0, ph_simplify = 0;

    // ---------------- basics
    inline int value(int x) const { uint8_t a = assigns[x >> 1]; return a == L_UNDEF ? L_UNDEF : (a ^ (x & 1)); }
    inline int32_t* CL(uint32_t c) { return pool.data() + C[c].start; }
    inline const int32_t* CL(uint32_t c) const { return pool.data() + C[c].start; }
    inline int dcount(int v) const { return hot[2 * v].cnt + hot[2 * v + 1].cnt; }
    void newVar();
    uint32_t allocClause(const int* lits, int n);
    void caFree(uint32_t c) { ca_wasted += 2 + C[c].size; }
    void uncheckedEnqueue(int x);
    int dataEnqueue(int x);
    void addClauseIngest(std::vector<int>& ps);
    bool ingestPropagate();
    bool solverSimplify();
    void initData();
    void die(const char* msg) const;

    // occurrence lists
    inline uint32_t* LP(const ListRef& l) { return occ_pool.data() + l.off; }
    void occPush(int x, uint32_t c);
    void occSwapRemove(int x, uint32_t i);            // list[i] = back; pop  (+ stale bookkeeping)
    void occRelease(int x);
    void setDeleted(uint32_t c);                      // mark + stale + device edit
    void noteShrunk(uint32_t c, int removedLit, bool pendingOcc);      // bookkeeping after a literal left clause c

    // CoprocessorData bookkeeping (CoprocessorTypes.h:812-878,1243-1350,1615-1655,1815-1831)
    void deletedVar(int v) { modTimer[v] = modStep; }
    void removedClause(uint32_t c, bool useHeap, bool update, int ignore);
    void removedLiteral(int x, int diff, bool useHeap, bool update, int ignore);
    void dataAddClause(uint32_t c, bool useHeap);
    bool removeClauseFrom(uint32_t c, int x);
    void subInitClause(uint32_t c);
    void addSubStrength(uint32_t c);
    void extClause(uint32_t c, int x);
    void extLit(int x);
    size_t filterDeletedRange(uint32_t* a, size_t n);
    template <class V> void filterDeleted(V& a) { a.resize(filterDeletedRange(a.data(), a.size())); }
    void checkGarbage();
    void removePosSorted(uint32_t c, int pos);
    void removePosUnsorted(uint32_t c, int pos);

    // heap (riss/mtl/Heap.h)
    // VarOrderBVEHeapLt, CoprocessorTypes.h:493-547: -cp3_bve_heap=0 fewest occurrences first, 1 most first, 3-10 the polarity
    // ratio of the variable combined with its occurrence count (2 = random order without a heap: not provided)
    // (the default order stays a two-load inline compare: the heap walks call this ~10^8 times at 10 M variables)
    bool heapLt(int x, int y) const { return opt.bve_heap == 0 ? dcount(x) < dcount(y) : heapLtOption(x, y); }
    bool heapLtOption(int x, int y) const __attribute__((noinline))
    {
        const int h = opt.bve_heap;
        if (h == 1) { return dcount(x) > dcount(y); }
        const double xp = hot[2 * x].cnt, xn = hot[2 * x + 1].cnt, yp = hot[2 * y].cnt, yn = hot[2 * y + 1].cnt;
        double rx = 0, ry = 0;
        if (xp != 0 || xn != 0) { rx = xp > xn ? (xn != 0 ? xp / xn : xp * 1000) : (xp != 0 ? xn / xp : xn * 1000); }
        if (yp != 0 || yn != 0) { ry = yp > yn ? (yn != 0 ? yp / yn : yp * 1000) : (yp != 0 ? yn / yp : yn * 1000); }
        const int dx = dcount(x), dy = dcount(y);
        switch (h) {
        case 3: return (rx < ry) || (rx == ry && dx < dy);
        case 4: return (rx < ry) || (rx == ry && dx > dy);
        case 5: return (rx > ry) || (rx == ry && dx < dy);
        case 6: return (rx > ry) || (rx == ry && dx > dy);
        case 7: return (dx < dy) || (rx < ry && dx == dy);
        case 8: return (dx > dy) || (rx < ry && dx == dy);
        case 9: return (dx < dy) || (rx > ry && dx == dy);
        default: return (dx > dy) || (rx > ry && dx == dy);
        }
    }
    bool inHeap(int v) const { return v < (int)hidx.size() && hidx[v] >= 0; }
    void heapUp(int i); void heapDown(int i); void heapInsert(int v); void heapUpdate(int v);
    int heapRemoveMin(); void heapClear();

    // techniques
    static bool performSimplification(Tech& t) { if (t.penalty > 0) { --t.penalty; return false; } return true; }
    static void unsuccessful(Tech& t) { t.peak += 1; t.penalty = t.peak; }
    static void successful(Tech& t) { t.modified = true; t.peak = 0; }
    int propagationProcess(bool sort, bool useHeap, int ignore);
    bool hasToPropagate() const { return (int)trail.size() != qhead; }
    void subsumptionWorker(bool useHeap, int ignore);
    bool subsumptionCommitParallel();
    // the occurrence lists as K2 built them are still on the device only (nobody on the host has looked at them yet): the first
    // subsumption pass can then be committed by the device (cp3g_sub_pass), which hands back the lists in their post-pass order
    bool lists_on_device = false; int64_t n_devsub = 0;
    void ensureHostLists();
    bool subsumptionCommitDevice();
    // overlap: the strengthening scan + speculative closure of this fixpoint round run on a helper thread (GPU +
    // one core) while the subsumption commit occupies the other cores
    // (the helper only READS what the committing threads write - deleted flags, list lengths - through plain loads of aligned
    //  32-bit words: it may see a clause as live that dies a moment later, which monotonicity makes harmless for the RESULT - the
    //  hit is dropped at commit time - but the speculation budget and request set, hence the counters spec_requests /
    //  scan_ondemand / gpu_launches, are not reproducible from run to run while the overlap is on; CP3_NO_OVERLAP=1 pins them)
    bool overlap_str = false, str_prepared = false, scan_no_flush = false;
    void prepareStrengthening(const std::vector<uint32_t>& ids, bool bulk);
    int64_t n_overlap = 0;                // time-indexed, data-parallel form of the queue walk
    RawVec<int32_t> sub_dtime, sub_pos; RawVec<int32_t> dl_start; RawVec<uint32_t> dl_cnt; std::vector<uint64_t> dl_any;   // dl_any: bitmap of dl_cnt != 0 (cache resident)
    int64_t n_parsub = 0;
    int strengtheningWorker(bool useHeap, int ignore);
    int naiveStrengthening(bool useHeap, int ignore);
    // the real entries of a strengthening pass, committed by dependency component on all cores (engine_str.cpp)
    // static part of the pass as the device computed it (K10, cp3g_str_static): the inert entries' number and steps, the lists they walk
    struct DevStatic { uint64_t inert_n = 0, inert_steps = 0; std::vector<uint32_t> walked_bits, real_bits; };
    bool strengthenComponents(const std::vector<uint32_t>& reals, const DevStatic* ds = nullptr);
    // first strengthening pass behind a device-committed subsumption pass: K10 replaces the static plan of the host
    bool dev_static_ok = false; int64_t n_devstatic = 0;
    bool strengthenFirstPassDevice();
    int64_t n_compar = 0, n_compar_fallback = 0, n_compar_rounds = 0;
    void strengthenHit(std::vector<uint32_t>& localQueue, uint32_t other, int pos);
    void updateOccurrences(bool useHeap, int ignore);
    bool subsumptionProcess(bool doStrengthen, bool useHeap, int ignore);
    void clearQueue(RawVec<uint32_t>& q, uint8_t flag);
    bool findGates(int v, int& p_limit, int& n_limit);
    int anticipate(int v, int p_limit, int n_limit, int& lit_clauses, int& resolvents);
    void bveRemoveClauses(int x, int extLitOrNeg);
    void bveRemoveBlocked(int x, const std::vector<int>& stats);
    int resolveSet(int v, int p_limit, int n_limit);
    void bveWorker();
    // The BVE commit is bound by cache misses of the host mirror (one variable = its counters, two lists, the headers of
    // their clauses, the K5 record - some 25 dependent lines at 10 M variables).  The next variables are the ones near the
    // top of the heap: a small staged prefetcher walks them one dependency level per loop iteration.  Side-effect free.
    struct BvePf { int var; int stage; };
    BvePf bve_pf[16]; int bve_pf_n = 0;
    void bvePrefetchStep();
    void sequentiellBVE();
    int bveProcess();
    // stand-alone blocked clause elimination (coprocessor/techniques/BCE.cc).  K9 (cp3g_bce_masks) answers every
    // tautologicResolvent() of the run at once: one bit per (clause with v, clause with ~v); clause contents and list orders
    // do not change during BCE, so the masks of one batch serve all rounds.
    RawVec<uint32_t> bce_masks; RawVec<uint64_t> bce_woff;      // per variable: first mask word, or UINT64_MAX (no mask: row on demand)
    std::vector<uint32_t> bce_row; int bce_row_var = -1; uint32_t bce_row_idx = 0;
    void bceBuildMasks();
    const uint32_t* bceRow(int v, uint32_t posIdx);        // mask row of the posIdx-th clause of occ(+v) over occ(~v)
    bool bceNonTaut(int right, uint32_t i, uint32_t j);    // resolvent of occ(right)[i] with occ(~right)[j] on right is not a tautology
    void blockedClauseElimination();
    bool bceProcess();
    uint32_t nextModStep();
    void addClausesToSubsumption(int x);
    void addTouchedToSubsumption();
    RawVec<uint32_t> first_rank;                     // per clause: rank of its first occurrence in the walk over the touched lists

    // device interaction
    void ensureGpu();
    void uploadAll();
    void flushDevice();
    // scan `ids` (clause indices, device content) and store the result in `cache`
    void scanClauses(int kind, const std::vector<uint32_t>& ids, ScanCache& cache, int forceBulk = -1);
    const CacheEnt* cached(const ScanCache& cache, uint32_t c) const;
    bool subHitValid(uint32_t c, uint32_t d, uint32_t scan) const;
    bool strHitValid(uint32_t s, uint32_t d, int f, uint32_t scan) const;
    const CacheEnt* strHits(uint32_t cr);             // cached or on-demand strengthening hits of cr
    PairEnt& pairsFor(int v);
    void touchClauseVars(uint32_t c);
    void touchDeadClauses(const std::vector<uint32_t>& dead);
};

} // namespace cp3
