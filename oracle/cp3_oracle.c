/* cp3_oracle.c - plain-C restatement of Coprocessor's sequential simplification hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see cp3_oracle.h).  Parity status: PINNED against the compiled reference
 * (oracle/_ref) by tests/test_oracle_vs_ref.py and the golden vectors under tests/golden/.
 *
 * This is NOT reference source: it is a restatement with its own data structures (flat literal pool,
 * per-literal int vectors) of the *order-exact behaviour* of
 *   Subsumption::process / subsumption_worker / strengthening_worker   coprocessor/techniques/Subsumption.cc
 *   Propagation::process                                               coprocessor/techniques/Propagation.cc
 *   BoundedVariableElimination::process / bve_worker / anticipate...   coprocessor/techniques/BVE.cc
 *   CoprocessorData bookkeeping                                        coprocessor/CoprocessorTypes.h
 *   Heap<VarOrderBVEHeapLt>                                            riss/mtl/Heap.h
 * including the reference's quirks that leak into results (each is marked QUIRK below).
 * Scope: preprocessing of original clauses (no learnt clauses, no proofs, cp3_threads=0).
 */
#include "cp3_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small containers */
typedef struct { int* v; int n, cap; } ivec;
typedef struct { int cref, blocker, bin; } watcher_t;     /* Riss::Watcher: clause, blocking literal, binary flag */
typedef struct { watcher_t* v; int n, cap; } wvec;
static void iv_push(ivec* a, int x)
{
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 4;
        a->v = (int*)realloc(a->v, sizeof(int) * (size_t)a->cap);
        if (!a->v) { fprintf(stderr, "cp3_oracle: out of memory\n"); abort(); }
    }
    a->v[a->n++] = x;
}
static void iv_release(ivec* a) { free(a->v); a->v = NULL; a->n = a->cap = 0; }

/* literal code like riss/core/SolverTypes.h:65-98: x = 2*var + sign, ~x = x^1 */
#define LVAR(x) ((x) >> 1)
#define LNEG(x) ((x) ^ 1)
#define F_DEL 1u
#define F_SUB 2u
#define F_STR 4u
#define F_PERM 8u   /* literal order changed by Solver::propagate, see ingest_propagate */
#define L_TRUE 0
#define L_FALSE 1
#define L_UNDEF 2
#define VAR_UNDEF (-1)

typedef struct { int penalty, peak; int modified; } tech_t; /* coprocessor/Technique.h:29-31,58 */

struct cp3o {
    /* options (coprocessor/CP3Config.cc:56-112,196-224,560-577) */
    int enabled, opt_up, opt_subsimp, opt_bve, limited, iters;
    int cp3_vars, cp3_cls, cp3_lits;
    int strength, all_strength_res, naive_strength;
    int64_t sub_limit, str_limit; int call_inc;
    int64_t bve_limit;
    int bve_unlimited, bve_strength, bve_gates, bve_force_gates, bve_fdepOnly, bve_heap;
    int bve_cgrow, bve_cgrow_t, bve_BCElim, bve_heap_updates, bve_early, bce_only;
    int opt_dense, dense_frag, dense_keep_set;   /* CP3Config.cc:100,344,345 */
    int opt_bce, bce_compl, bce_binary; int64_t bce_limit;   /* -bce, -bce-compl, -bce-bin, -bce-limit (CP3Config.cc:86,266-268) */
    double gc_frac;
    /* solver side (riss/core/Solver.h): assignment, trail, clauses vector, arena accounting */
    int nvars, ok, qhead;
    unsigned char* assigns; unsigned char* frozen;
    ivec trail;
    int* pool; size_t pool_n, pool_cap;         /* literal pool */
    size_t* cstart; int* csize; unsigned char* cflag; int ncl, ncl_cap; /* clause table, id = creation order */
    ivec clauses;                               /* solver->clauses */
    double ca_size, ca_wasted;                  /* RegionAllocator size()/wasted() in 32-bit words */
    int simpDB_assigns; int64_t simpDB_props;
    wvec* watches; int ing_built;               /* two-watched-literal lists while clauses are added (see ingest_propagate) */
    ivec permuted;                              /* clauses whose literal order propagate() changed; re-sorted by sortClauses */
    ivec cur;                                   /* clause under construction */
    /* CoprocessorData (coprocessor/CoprocessorTypes.h:59-412) */
    int data_ready; int data_nvars;            /* CoprocessorData::numberOfVars: 0 until the first data.init() */
    ivec* occ; int* cnt; int numberOfCls;
    ivec subq, strq, undo;
    uint32_t* modTimer; uint32_t modStep;
    /* techniques */
    tech_t t_sub, t_up, t_bve, t_bce;
    int bce_steps, bce_testedLits, bce_remBCE;  /* BCE.h:39-41 */
    int64_t sub_steps, str_steps, sub_lim, str_lim; int limitIncreases;
    int subsumedClauses, subsumedLiterals, removedLiterals;
    int up_last, up_removedClauses, up_removedLiterals;
    uint32_t bve_modtime;
    int64_t bve_steps;
    int *heap, heap_n, heap_cap, *hidx, hidx_n; /* riss/mtl/Heap.h */
    int bve_removedClauses, bve_removedLiterals, bve_createdClauses, bve_createdLiterals, bve_testedVars,
        bve_anticipations, bve_eliminatedVars, bve_removedBC, bve_skippedVars, bve_unitsEnqueued,
        bve_foundGates, bve_nInc, bve_nDec, bve_nKeep, bve_totallyAdded;
    ivec pos_stats, neg_stats, resolvent, touched;
    uint32_t* ma; uint32_t ma_step;             /* data.ma, MarkArray SolverTypes.h:1120 */
    struct { int cr, l; } *occ_upd; int occ_upd_n, occ_upd_cap; /* strength_occ_updates */
    ivec sub_occ_updates;
    /* Dense / Riss::Compression (riss/utils/Compression.h): mapping[orig var] = compressed var or -1 (removed),
     * value of the removed assigned variables; empty until a compression happened */
    int comp_nvars; int* comp_map; unsigned char* comp_val;
};

/* ------------------------------------------------------------------ basics */
static inline int value_lit(const cp3o* o, int x)
{
    unsigned char a = o->assigns[LVAR(x)];
    return a == L_UNDEF ? L_UNDEF : (a ^ (x & 1));
}
static inline int* CL(cp3o* o, int c) { return o->pool + o->cstart[c]; }

static void new_var(cp3o* o)
{
    int n = o->nvars + 1;
    o->assigns = (unsigned char*)realloc(o->assigns, (size_t)n);
    o->frozen = (unsigned char*)realloc(o->frozen, (size_t)n);
    o->assigns[n - 1] = L_UNDEF; o->frozen[n - 1] = 0;
    if (o->ing_built) {
        o->watches = (wvec*)realloc(o->watches, sizeof(wvec) * 2 * (size_t)n);
        memset(&o->watches[2 * (n - 1)], 0, 2 * sizeof(wvec));
    }
    o->nvars = n;
}

static int alloc_clause(cp3o* o, const int* lits, int n)
{
    if (o->ncl == o->ncl_cap) {
        o->ncl_cap = o->ncl_cap ? o->ncl_cap * 2 : 1024;
        o->cstart = (size_t*)realloc(o->cstart, sizeof(size_t) * (size_t)o->ncl_cap);
        o->csize = (int*)realloc(o->csize, sizeof(int) * (size_t)o->ncl_cap);
        o->cflag = (unsigned char*)realloc(o->cflag, (size_t)o->ncl_cap);
    }
    if (o->pool_n + (size_t)n > o->pool_cap) {
        while (o->pool_n + (size_t)n > o->pool_cap) { o->pool_cap = o->pool_cap ? o->pool_cap * 2 : 4096; }
        o->pool = (int*)realloc(o->pool, sizeof(int) * o->pool_cap);
        if (!o->pool) { fprintf(stderr, "cp3_oracle: out of memory\n"); abort(); }
    }
    int c = o->ncl++;
    o->cstart[c] = o->pool_n; o->csize[c] = n;
    o->cflag[c] = F_SUB | F_STR;     /* fresh clause: can_subsume, can_strengthen (SolverTypes.h:247-251) */
    memcpy(o->pool + o->pool_n, lits, sizeof(int) * (size_t)n);
    o->pool_n += (size_t)n;
    o->ca_size += 2 + n;             /* ClauseAllocator::clauseWord32Size, SolverTypes.h:641-644 */
    return c;
}
static inline void ca_free(cp3o* o, int c) { o->ca_wasted += 2 + o->csize[c]; } /* SolverTypes.h:701-705 */

static void unchecked_enqueue(cp3o* o, int x)
{
    o->assigns[LVAR(x)] = (unsigned char)(x & 1);
    iv_push(&o->trail, x);
}
/* CoprocessorData::enqueue, CoprocessorTypes.h:767-783 */
static int data_enqueue(cp3o* o, int x)
{
    int v = value_lit(o, x);
    if (v == L_FALSE) { o->ok = 0; return L_FALSE; }
    if (v == L_UNDEF) { unchecked_enqueue(o, x); return L_TRUE; }
    return L_UNDEF;
}

/* removePositionSorted, SolverTypes.h:509 */
static void remove_pos_sorted(cp3o* o, int c, int pos)
{
    int* l = CL(o, c); int n = o->csize[c];
    for (int j = pos; j < n - 1; ++j) { l[j] = l[j + 1]; }
    o->csize[c] = n - 1;
}
static void remove_pos_unsorted(cp3o* o, int c, int pos)
{
    int* l = CL(o, c); int n = o->csize[c];
    l[pos] = l[n - 1]; o->csize[c] = n - 1;
}

/* ------------------------------------------------------------------ ingest: Solver::addClause_ (Solver.cc:409-499) */
static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

/* Solver::propagate(true) at decision level 0 while clauses are being added (riss/core/Solver.cc:1809-1937) over the
 * two-watched-literal lists of Solver::attachClause (Solver.cc:692-708).  The TRAIL ORDER of implied literals is a
 * function of the watch-list order, the blocker literals and the way propagate() permutes the literals of the clauses
 * it visits (c[0]/c[1] swap, new watch moved to c[1]) - all restated here.  The permuted clauses are put back into
 * sorted order where the reference does it: Preprocessor::sortClauses (Coprocessor.cc:118,1447-1463).
 * watches[p] holds the clauses that watch ~p.  The lists are built at the first unit (until then no clause has been
 * visited, so every clause watches its first two sorted literals, attached in clause order). */
static void wv_push(wvec* a, watcher_t w)
{
    if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 4; a->v = (watcher_t*)realloc(a->v, sizeof(watcher_t) * (size_t)a->cap); }
    a->v[a->n++] = w;
}
static void attach_clause(cp3o* o, int c)
{
    const int* l = CL(o, c); const int bin = o->csize[c] == 2;
    watcher_t w0 = {c, l[1], bin}, w1 = {c, l[0], bin};
    wv_push(&o->watches[LNEG(l[0])], w0);
    wv_push(&o->watches[LNEG(l[1])], w1);
}
static void note_permuted(cp3o* o, int c)
{
    if (!(o->cflag[c] & F_PERM)) { o->cflag[c] |= F_PERM; iv_push(&o->permuted, c); }
}
static int ingest_propagate(cp3o* o)
{
    if (!o->ing_built) {
        o->watches = (wvec*)calloc((size_t)(2 * o->nvars > 0 ? 2 * o->nvars : 1), sizeof(wvec));
        o->ing_built = 1;
        for (int i = 0; i < o->clauses.n; ++i) { attach_clause(o, o->clauses.v[i]); }
    }
    wvec* W = o->watches;
    int confl = 0;
    while (o->qhead < o->trail.n) {
        const int p = o->trail.v[o->qhead++];
        wvec* ws = &W[p];
        int i = 0, j = 0; const int end = ws->n;
        while (i != end) {
            watcher_t wi = ws->v[i];
            if (wi.bin) {
                const int vi = value_lit(o, wi.blocker);
                if (vi == L_FALSE) {                        /* opt_long_conflict is off: stop at the first conflict */
                    confl = 1;
                    while (i < end) { ws->v[j++] = ws->v[i++]; }
                    ws->n = j;
                    return 0;
                } else if (vi == L_UNDEF) { unchecked_enqueue(o, wi.blocker); }
                ws->v[j++] = ws->v[i++];
                continue;
            }
            if (value_lit(o, wi.blocker) == L_TRUE) { ws->v[j++] = ws->v[i++]; continue; }
            const int c = wi.cref; int* l = CL(o, c); const int n = o->csize[c];
            const int false_lit = LNEG(p);
            if (l[0] == false_lit) { l[0] = l[1]; l[1] = false_lit; note_permuted(o, c); }
            i++;
            const int first = l[0];
            watcher_t w = {c, first, 0};
            if (first != wi.blocker && value_lit(o, first) == L_TRUE) { ws->v[j++] = w; continue; }
            int moved = 0;
            for (int k = 2; k < n; ++k) {
                if (value_lit(o, l[k]) != L_FALSE) {
                    l[1] = l[k]; l[k] = false_lit; note_permuted(o, c);
                    wv_push(&W[LNEG(l[1])], w);
                    ws = &W[p];                             /* (the push cannot touch W[p]: l[1] != ~p) */
                    moved = 1; break;
                }
            }
            if (moved) { continue; }
            ws->v[j++] = w;
            if (value_lit(o, first) == L_FALSE) {
                confl = 1;
                o->qhead = o->trail.n;
                while (i < end) { ws->v[j++] = ws->v[i++]; }
            } else { unchecked_enqueue(o, first); }
        }
        ws->n = j;
    }
    return !confl;
}

static void add_clause(cp3o* o, int* ps, int n)
{
    if (!o->ok) { return; }
    qsort(ps, (size_t)n, sizeof(int), cmp_int);
    int j = 0, p = -2;
    for (int i = 0; i < n; ++i) {
        int v = value_lit(o, ps[i]);
        if (v == L_TRUE || ps[i] == LNEG(p)) { return; }       /* satisfied or tautology */
        if (v != L_FALSE && ps[i] != p) { ps[j++] = p = ps[i]; }
    }
    n = j;
    if (n == 0) { o->ok = 0; return; }
    if (n == 1) {
        unchecked_enqueue(o, ps[0]);
        o->ok = ingest_propagate(o);
        return;
    }
    int c = alloc_clause(o, ps, n);
    iv_push(&o->clauses, c);
    if (o->ing_built) { attach_clause(o, c); }
}

void cp3o_add(cp3o* o, const int32_t* s, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        int x = s[i];
        if (x == 0) {
            add_clause(o, o->cur.v, o->cur.n);
            o->cur.n = 0;
        } else {
            int v = x < 0 ? -x : x;
            while (o->nvars < v) { new_var(o); }
            iv_push(&o->cur, 2 * (v - 1) + (x < 0));
        }
    }
}
void cp3o_freeze(cp3o* o, int dv)
{
    while (o->nvars < dv) { new_var(o); }
    o->frozen[dv - 1] = 1;
}

/* ------------------------------------------------------------------ Technique penalty (Technique.h:164-200) */
static int perform_simplification(tech_t* t) { if (t->penalty > 0) { --t->penalty; return 0; } return 1; }
static void unsuccessful(tech_t* t) { t->peak += 1; t->penalty = t->peak; }
static void successful(tech_t* t) { t->modified = 1; t->peak = 0; }

/* ------------------------------------------------------------------ Heap<VarOrderBVEHeapLt> (riss/mtl/Heap.h:37-183) */
static inline int dcount(const cp3o* o, int v) { return o->cnt[2 * v] + o->cnt[2 * v + 1]; } /* data[Var] */
static inline int heap_lt(const cp3o* o, int x, int y)
{   /* CoprocessorTypes.h:493-500; live keys */
    const int h = o->bve_heap;
    if (h == 0) { return dcount(o, x) < dcount(o, y); }
    if (h == 1) { return dcount(o, x) > dcount(o, y); }
    /* options 3-10, CoprocessorTypes.h:503-536: polarity ratio of the variable combined with its occurrence count */
    const double xp = o->cnt[2 * x], xn = o->cnt[2 * x + 1], yp = o->cnt[2 * y], yn = o->cnt[2 * y + 1];
    double rx = 0, ry = 0;
    if (xp != 0 || xn != 0) { rx = xp > xn ? (xn != 0 ? xp / xn : xp * 1000) : (xp != 0 ? xn / xp : xn * 1000); }
    if (yp != 0 || yn != 0) { ry = yp > yn ? (yn != 0 ? yp / yn : yp * 1000) : (yp != 0 ? yn / yp : yn * 1000); }
    const int dx = dcount(o, x), dy = dcount(o, y);
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
static inline int in_heap(const cp3o* o, int v) { return v < o->hidx_n && o->hidx[v] >= 0; }
static void heap_up(cp3o* o, int i)
{
    int x = o->heap[i], p = (i - 1) >> 1;
    while (i != 0 && heap_lt(o, x, o->heap[p])) {
        o->heap[i] = o->heap[p]; o->hidx[o->heap[p]] = i; i = p; p = (p - 1) >> 1;
    }
    o->heap[i] = x; o->hidx[x] = i;
}
static void heap_down(cp3o* o, int i)
{
    int x = o->heap[i];
    while (2 * i + 1 < o->heap_n) {
        int l = 2 * i + 1, r = 2 * i + 2;
        int child = (r < o->heap_n && heap_lt(o, o->heap[r], o->heap[l])) ? r : l;
        if (!heap_lt(o, o->heap[child], x)) { break; }
        o->heap[i] = o->heap[child]; o->hidx[o->heap[i]] = i; i = child;
    }
    o->heap[i] = x; o->hidx[x] = i;
}
static void heap_insert(cp3o* o, int v)
{
    if (v >= o->hidx_n) {
        o->hidx = (int*)realloc(o->hidx, sizeof(int) * (size_t)(v + 1));
        for (int i = o->hidx_n; i <= v; ++i) { o->hidx[i] = -1; }
        o->hidx_n = v + 1;
    }
    if (o->heap_n == o->heap_cap) {
        o->heap_cap = o->heap_cap ? 2 * o->heap_cap : 64;
        o->heap = (int*)realloc(o->heap, sizeof(int) * (size_t)o->heap_cap);
    }
    o->hidx[v] = o->heap_n; o->heap[o->heap_n++] = v;
    heap_up(o, o->hidx[v]);
}
static void heap_update(cp3o* o, int v)
{
    if (!in_heap(o, v)) { heap_insert(o, v); }
    else { heap_up(o, o->hidx[v]); heap_down(o, o->hidx[v]); }
}
static int heap_remove_min(cp3o* o)
{
    int x = o->heap[0];
    o->heap[0] = o->heap[o->heap_n - 1];
    o->hidx[o->heap[0]] = 0;
    o->hidx[x] = -1;
    o->heap_n--;
    if (o->heap_n > 1) { heap_down(o, 0); }
    return x;
}
static void heap_clear(cp3o* o)
{
    for (int i = 0; i < o->heap_n; ++i) { o->hidx[o->heap[i]] = -1; }
    o->heap_n = 0;
}

/* ------------------------------------------------------------------ CoprocessorData bookkeeping */
static inline void deleted_var(cp3o* o, int v) { o->modTimer[v] = o->modStep; } /* CoprocessorTypes.h:1141 */

/* removedClause(cr, heap, update, ignore), CoprocessorTypes.h:1310-1350.  use_heap==0 is the lock-free
 * branch without heap. */
static void removed_clause(cp3o* o, int c, int use_heap, int update, int ignore)
{
    int* l = CL(o, c);
    for (int i = 0; i < o->csize[c]; ++i) {
        int v = LVAR(l[i]);
        deleted_var(o, v);
        --o->cnt[l[i]];
        if (use_heap) {
            if (in_heap(o, v)) { heap_up(o, o->hidx[v]); }              /* decrease */
            else if (update && v != ignore) { heap_update(o, v); }
        }
    }
    o->numberOfCls--;
    ca_free(o, c);
}
/* removedLiteral(l, diff, heap, update, ignore), CoprocessorTypes.h:1243-1277 */
static void removed_literal(cp3o* o, int x, int diff, int use_heap, int update, int ignore)
{
    int v = LVAR(x);
    deleted_var(o, v);
    o->cnt[x] -= diff;
    o->ca_wasted += 1;                                                  /* ca.freeLit() */
    if (use_heap) {
        if (in_heap(o, v)) { heap_up(o, o->hidx[v]); }
        else if (update && v != ignore) { heap_update(o, v); }
    }
}
/* addClause(cr, heap, update=false, ignore), CoprocessorTypes.h:835-864 */
static void data_add_clause(cp3o* o, int c, int use_heap)
{
    if (o->cflag[c] & F_DEL) { return; }
    int* l = CL(o, c);
    for (int i = 0; i < o->csize[c]; ++i) {
        iv_push(&o->occ[l[i]], c);
        o->cnt[l[i]] += 1;
        if (use_heap && in_heap(o, LVAR(l[i]))) { heap_down(o, o->hidx[LVAR(l[i])]); } /* increase */
    }
    o->numberOfCls++;
}
/* removeClauseFrom(cr, l), CoprocessorTypes.h:867-878 */
static int remove_clause_from(cp3o* o, int c, int x)
{
    ivec* li = &o->occ[x];
    for (int i = 0; i < li->n; ++i) {
        if (li->v[i] == c) { li->v[i] = li->v[li->n - 1]; li->n--; return 1; }
    }
    return 0;
}
/* Subsumption::initClause, Subsumption.cc:1546-1557 */
static void sub_init_clause(cp3o* o, int c)
{
    if (o->cflag[c] & F_DEL) { return; }
    if (o->cflag[c] & F_SUB) { iv_push(&o->subq, c); }
    if (o->cflag[c] & F_STR) { iv_push(&o->strq, c); }
}
/* addSubStrengthClause, CoprocessorTypes.h:1815-1831 */
static void add_sub_strength(cp3o* o, int c)
{
    if (!(o->cflag[c] & F_STR)) { o->cflag[c] |= F_STR; iv_push(&o->strq, c); }
    if (!(o->cflag[c] & F_SUB)) { o->cflag[c] |= F_SUB; iv_push(&o->subq, c); }
}
/* addToExtension(clause, l) and (dontTouch), CoprocessorTypes.h:1615-1655; stored as DIMACS, 0 = lit_Undef */
static inline int to_dimacs(int x) { return (x & 1) ? -(LVAR(x) + 1) : (LVAR(x) + 1); }
static void ext_clause(cp3o* o, int c, int x)
{
    iv_push(&o->undo, 0);
    iv_push(&o->undo, to_dimacs(x));
    int* l = CL(o, c);
    for (int i = 0; i < o->csize[c]; ++i) { if (l[i] != x) { iv_push(&o->undo, to_dimacs(l[i])); } }
}
static void ext_lit(cp3o* o, int x) { iv_push(&o->undo, 0); iv_push(&o->undo, to_dimacs(x)); }

/* garbageCollect / relocAll, CoprocessorTypes.h:1396-1550: stable removal of deleted clauses from both
 * queues, all occurrence lists and solver->clauses; arena accounting restarts from the live clauses. */
static void filter_deleted(cp3o* o, ivec* a)
{
    int j = 0;
    for (int i = 0; i < a->n; ++i) { if (!(o->cflag[a->v[i]] & F_DEL)) { a->v[j++] = a->v[i]; } }
    a->n = j;
}
static void check_garbage(cp3o* o)
{
    if (!(o->ca_wasted > o->ca_size * o->gc_frac)) { return; }
    filter_deleted(o, &o->subq);
    filter_deleted(o, &o->strq);
    for (int x = 0; x < 2 * o->nvars; ++x) { filter_deleted(o, &o->occ[x]); }
    filter_deleted(o, &o->clauses);
    double sz = 0;
    for (int i = 0; i < o->clauses.n; ++i) { sz += 2 + o->csize[o->clauses.v[i]]; }
    o->ca_size = sz; o->ca_wasted = 0;
}

/* ------------------------------------------------------------------ Propagation::process (Propagation.cc:32-131) */
static int propagation_process(cp3o* o, int sort, int use_heap, int ignore)
{
    o->t_up.modified = 0;
    if (!o->ok) { return L_FALSE; }
    for (; o->up_last < o->trail.n; o->up_last++) {
        const int l = o->trail.v[o->up_last];
        ivec* positive = &o->occ[l];
        for (int i = 0; i < positive->n; ++i) {
            int c = positive->v[i];
            if (o->cflag[c] & F_DEL) { continue; }
            ++o->up_removedClauses;
            o->cflag[c] |= F_DEL;
            o->t_up.modified = 1;
            removed_clause(o, c, use_heap, 0, ignore);
        }
        iv_release(positive);
        const int nl = LNEG(l);
        int count = 0;
        ivec* negative = &o->occ[nl];
        for (int i = 0; i < negative->n; ++i) {
            int c = negative->v[i];
            if (o->cflag[c] & F_DEL) { continue; }
            int* lits = CL(o, c);
            for (int j = 0; j < o->csize[c]; ++j) {
                if (lits[j] == nl) {
                    if (!sort) { remove_pos_unsorted(o, c, j); } else { remove_pos_sorted(o, c, j); }
                    o->t_up.modified = 1;
                    count++;
                    break;
                }
            }
            add_sub_strength(o, c);
            if (o->csize[c] == 0) {
                o->cflag[c] |= F_DEL;
                o->ok = 0;
                return L_FALSE;
            } else if (o->csize[c] == 1) {
                o->cflag[c] |= F_DEL;
                int u = CL(o, c)[0];
                if (value_lit(o, u) == L_UNDEF) { unchecked_enqueue(o, u); }
                else if (value_lit(o, u) == L_FALSE) { o->ok = 0; }
            }
        }
        iv_release(negative);
        /* QUIRK: data.removedLiteral(nl, count, heap, ignore) passes `ignore` in the `update` slot
         * (Propagation.cc:114 vs CoprocessorTypes.h:303) */
        removed_literal(o, nl, count, use_heap, ignore != 0, VAR_UNDEF);
        o->up_removedLiterals += count;
    }
    o->qhead = o->up_last;
    if (!o->t_up.modified) { unsuccessful(&o->t_up); }
    return o->ok ? L_UNDEF : L_FALSE;
}
static inline int has_to_propagate(const cp3o* o) { return o->trail.n != o->qhead; }

/* ------------------------------------------------------------------ Subsumption */
/* Clause::ordered_subsumes, SolverTypes.h:1003-1023 */
static int ordered_subsumes(cp3o* o, int c, int d)
{
    const int* a = CL(o, c); const int* b = CL(o, d);
    int na = o->csize[c], nb = o->csize[d], i = 0, j = 0;
    while (i < na && j < nb) {
        if (a[i] == b[j]) { ++i; ++j; }
        else if (b[j] < a[i]) { ++j; }
        else { return 0; }
    }
    return i == na;
}

/* subsumption_worker, Subsumption.cc:220-314 */
static void subsumption_worker(cp3o* o, int use_heap, int ignore)
{
    const int unlimited = !o->limited;
    int end = o->subq.n;
    o->sub_occ_updates.n = 0;
    for (; end > 0 && (unlimited || o->sub_steps < o->sub_lim);) {
        --end;
        const int cr = o->subq.v[end];
        if (o->cflag[cr] & F_DEL) { continue; }
        const int* c = CL(o, cr); const int cn = o->csize[cr];
        int min = c[0];
        for (int l = 1; l < cn; ++l) { if (o->cnt[c[l]] < o->cnt[min]) { min = c[l]; } }
        ivec* list = &o->occ[min];
        for (int i = 0; i < list->n && (unlimited || o->sub_steps < o->sub_lim); ++i) {
            const int d = list->v[i];
            if (d == cr) { continue; }
            o->sub_steps++;
            if (o->csize[d] < cn) { continue; }
            if (o->cflag[d] & F_DEL) { list->v[i] = list->v[list->n - 1]; list->n--; --i; continue; }
            if (ordered_subsumes(o, cr, d)) {
                ++o->subsumedClauses;
                o->subsumedLiterals += o->csize[d];
                o->cflag[d] |= F_DEL;
                removed_clause(o, d, 0, 0, VAR_UNDEF);      /* data.removedClause(list[i]) - first time */
                successful(&o->t_sub);
                iv_push(&o->sub_occ_updates, d);
            }
        }
        o->cflag[cr] &= ~F_SUB;
    }
    /* QUIRK: removedClause a second time for every subsumed clause (Subsumption.cc:302-313), so counters,
     * numberOfCls and wasted words are decremented twice; `ignore` lands in the `update` slot. */
    for (int i = 0; i < o->sub_occ_updates.n; ++i) {
        removed_clause(o, o->sub_occ_updates.v[i], use_heap, ignore != 0, VAR_UNDEF);
    }
    o->sub_occ_updates.n = 0;
}

static void clear_queue(cp3o* o, ivec* q, unsigned flag)
{
    for (int i = 0; i < q->n; ++i) { o->cflag[q->v[i]] &= ~flag; }
    q->n = 0;
}

static void push_occ_update(cp3o* o, int cr, int l)
{
    if (o->occ_upd_n == o->occ_upd_cap) {
        o->occ_upd_cap = o->occ_upd_cap ? 2 * o->occ_upd_cap : 64;
        o->occ_upd = realloc(o->occ_upd, sizeof(*o->occ_upd) * (size_t)o->occ_upd_cap);
    }
    o->occ_upd[o->occ_upd_n].cr = cr; o->occ_upd[o->occ_upd_n].l = l; o->occ_upd_n++;
}
/* updateOccurrences, Subsumption.cc:1530-1544 */
static void update_occurrences(cp3o* o, int use_heap, int ignore)
{
    for (int i = 0; i < o->occ_upd_n; ++i) {
        if (remove_clause_from(o, o->occ_upd[i].cr, o->occ_upd[i].l)) {
            /* QUIRK: removedLiteral(l, 1, heap, ignore): ignore in the update slot */
            removed_literal(o, o->occ_upd[i].l, 1, use_heap, ignore != 0, VAR_UNDEF);
        }
    }
    o->occ_upd_n = 0;
}

/* the hit branch shared by both passes, Subsumption.cc:1385-1413 / 1474-1508 */
static void strengthen_hit(cp3o* o, ivec* localQueue, int other, int pos)
{
    ++o->removedLiterals;
    if (!(o->cflag[other] & F_SUB)) { o->cflag[other] |= F_SUB; iv_push(&o->subq, other); }
    if (!(o->cflag[other] & F_STR)) { iv_push(localQueue, other); o->cflag[other] |= F_STR; }
    successful(&o->t_sub);
    push_occ_update(o, other, CL(o, other)[pos]);
    remove_pos_sorted(o, other, pos);
    if (o->csize[other] == 1) {
        data_enqueue(o, CL(o, other)[0]);
        o->cflag[other] |= F_DEL;
    }
}

/* strengthening_worker, Subsumption.cc:1250-1528 (all_strength_res == 0) */
static int strengthening_worker(cp3o* o, int use_heap, int ignore)
{
    const int unlimited = !o->limited;
    ivec localQueue = {0, 0, 0};
    int end = o->strq.n;
    int ret = L_UNDEF;
    for (; end > 0 && (unlimited || o->str_steps < o->str_lim);) {
        int cr;
        if (localQueue.n == 0) { --end; cr = o->strq.v[end]; }
        else { cr = localQueue.v[--localQueue.n]; }
        if ((o->cflag[cr] & F_DEL) || !(o->cflag[cr] & F_STR)) { continue; }
        if (!o->strength) { o->cflag[cr] &= ~F_STR; continue; }
        if (o->csize[cr] < 2) {
            if (o->csize[cr] == 1) { if (data_enqueue(o, CL(o, cr)[0]) == L_FALSE) { break; } else { continue; } }
            else { o->ok = 0; break; }
        }
        int minT;
        { const int* s = CL(o, cr); minT = s[0];
          for (int j = 1; j < o->csize[cr]; ++j) { if (dcount(o, LVAR(minT)) > dcount(o, LVAR(s[j]))) { minT = s[j]; } } }
        const int min = minT, nmin = LNEG(minT);
        ivec* list = &o->occ[min];
        ivec* list_neg = &o->occ[nmin];
        /* pass 1: occ(min) - exactly one complementary literal anywhere */
        for (int j = 0; j < list->n && (unlimited || o->str_steps < o->str_lim); ++j) {
            const int other = list->v[j];
            if (other == cr) { continue; }
            o->str_steps++;
            if (o->csize[other] == 0) { break; }
            if (o->cflag[other] & F_DEL) { list->v[j] = list->v[list->n - 1]; list->n--; --j; continue; }
            const int* s = CL(o, cr); const int* t = CL(o, other);
            const int sn = o->csize[cr], tn = o->csize[other];
            int si = 0, so = 0, neg_pos = -1;
            while (si < sn && so < tn) {
                if (s[si] == t[so]) { ++si; ++so; }
                else if (s[si] == LNEG(t[so])) {
                    if (neg_pos == -1) { neg_pos = so; ++si; ++so; } else { break; }
                }
                else if (s[si] < t[so]) { break; }
                else { ++so; }
            }
            if (neg_pos != -1 && si == sn) { strengthen_hit(o, &localQueue, other, neg_pos); }
        }
        if (has_to_propagate(o)) {
            if (propagation_process(o, 1, use_heap, ignore) == L_FALSE) { ret = L_FALSE; goto done_noupd; }
        }
        /* pass 2: occ(~min) - the complementary literal must be ~min */
        for (int j = 0; j < list_neg->n && (unlimited || o->str_steps < o->str_lim); ++j) {
            const int other = list_neg->v[j];
            o->str_steps++;
            if (o->cflag[other] & F_DEL) { list_neg->v[j] = list_neg->v[list_neg->n - 1]; list_neg->n--; --j; continue; }
            const int* s = CL(o, cr); const int* t = CL(o, other);
            const int sn = o->csize[cr], tn = o->csize[other];
            int si = 0, so = 0, neg_pos = -1;
            while (si < sn && so < tn) {
                if (s[si] == t[so]) { ++si; ++so; }
                else if (nmin == t[so]) { neg_pos = so; ++si; ++so; }
                else if (s[si] < t[so]) { break; }
                else { ++so; }
            }
            if (si == sn && neg_pos != -1) { strengthen_hit(o, &localQueue, other, neg_pos); }
        }
        if (has_to_propagate(o)) {
            if (propagation_process(o, 1, use_heap, ignore) == L_FALSE) { ret = L_FALSE; goto done_noupd; }
            o->t_sub.modified = o->t_sub.modified || o->t_up.modified;
        }
        o->cflag[cr] &= ~F_STR;
    }
    update_occurrences(o, use_heap, ignore);
    ret = o->ok ? L_UNDEF : L_FALSE;
done_noupd:
    /* early `return l_False` paths (Subsumption.cc:1421-1423,1515-1517) leave strength_occ_updates as is */
    iv_release(&localQueue);
    return ret;
}

/* fullStrengthening with -naive_strength, Subsumption.cc:1056-1197: the queue runs front to back; every literal of the
 * strengthener is flipped in turn and the flipped clause is tested as a plain subset (the flipped literal must be found) */
static int naive_strengthening(cp3o* o, int use_heap, int ignore)
{
    const int unlimited = !o->limited;
    ivec localQueue = {0, 0, 0};
    int ret = L_UNDEF;
    for (int i = 0; i < o->strq.n;) {
        int cr;
        if (localQueue.n == 0) { cr = o->strq.v[i]; ++i; }
        else { cr = localQueue.v[--localQueue.n]; }
        if ((o->cflag[cr] & F_DEL) || !(o->cflag[cr] & F_STR)) { continue; }
        if (!o->strength) { o->cflag[cr] &= ~F_STR; continue; }
        const int cn = o->csize[cr];
        int min = CL(o, cr)[0];
        for (int j = 1; j < cn; ++j) {
            const int x = CL(o, cr)[j];
            if ((int64_t)o->cnt[min] * (cn - 1) + o->cnt[LNEG(min)] > (int64_t)o->cnt[x] * (cn - 1) + o->cnt[LNEG(x)]) { min = x; }
        }
        for (int j = 0; j < o->csize[cr] && (unlimited || o->str_steps < o->str_lim); ++j) {
            const int neg_lit = LNEG(CL(o, cr)[j]);
            CL(o, cr)[j] = neg_lit;
            ivec* list = (neg_lit == LNEG(min)) ? &o->occ[neg_lit] : &o->occ[min];
            for (int k = 0; k < list->n && (unlimited || o->str_steps < o->str_lim); ++k) {
                const int other = list->v[k];
                if (other == cr) { continue; }
                o->str_steps++;
                if (o->cflag[other] & F_DEL) { list->v[k] = list->v[list->n - 1]; list->n--; --k; continue; }
                const int* c = CL(o, cr); const int* t = CL(o, other);
                const int sn = o->csize[cr], tn = o->csize[other];
                int l1 = 0, l2 = 0, pos = -1;
                while (l1 < sn && l2 < tn) {
                    if (c[l1] == t[l2]) { if (c[l1] == neg_lit) { pos = l2; } ++l1; ++l2; }
                    else if (c[l1] < t[l2]) { break; }
                    else { ++l2; }
                }
                if (l1 == sn && pos != -1) {
                    ++o->removedLiterals;
                    remove_pos_sorted(o, other, pos);
                    o->t_sub.modified = 1;                               /* modifiedFormula = true: the penalty peak stays */
                    if (o->csize[other] == 1) {
                        o->cflag[other] |= F_DEL;
                        data_enqueue(o, CL(o, other)[0]);
                    } else {
                        if (!(o->cflag[other] & F_SUB)) { o->cflag[other] |= F_SUB; iv_push(&o->subq, other); }
                        if (!(o->cflag[other] & F_STR)) { iv_push(&localQueue, other); o->cflag[other] |= F_STR; }
                        push_occ_update(o, other, neg_lit);
                    }
                }
            }
            CL(o, cr)[j] = LNEG(neg_lit);
        }
        if (has_to_propagate(o)) {
            if (propagation_process(o, 1, 0, VAR_UNDEF) == L_FALSE) { o->ok = 0; ret = L_FALSE; goto done_noupd; }
        }
        o->cflag[cr] &= ~F_STR;
    }
    update_occurrences(o, use_heap, ignore);
done_noupd:
    iv_release(&localQueue);
    return ret;
}

/* Subsumption::process, Subsumption.cc:85-180 */
static int subsumption_process(cp3o* o, int doStrengthen, int use_heap, int ignore)
{
    o->t_sub.modified = 0;
    if (!o->ok) { return 0; }
    if (!perform_simplification(&o->t_sub)) {
        clear_queue(o, &o->subq, F_SUB);
        clear_queue(o, &o->strq, F_STR);
        return 0;
    }
    /* size gate Subsumption.cc:102 (cp3_susi_vars && cp3_susi_cls && cp3_susi_lits): nTotLits() is 0 here - cleanSolver()
     * (Coprocessor.cc:1395-1406) zeroes clauses_literals before any technique runs - so the conjunction never holds.
     * Verified against the compiled reference with all three limits set to 10. */
    if (!(o->sub_steps + o->call_inc < o->sub_lim)) { o->sub_lim += o->call_inc; o->limitIncreases++; }
    if (!(o->str_steps + o->call_inc < o->str_lim)) { o->str_lim += o->call_inc; o->limitIncreases++; }
    while (o->ok && (o->subq.n > 0 || o->strq.n > 0)) {
        if (o->subq.n > 0) {
            subsumption_worker(o, use_heap, ignore);
            clear_queue(o, &o->subq, F_SUB);
        }
        if (o->strq.n > 0) {
            if (!doStrengthen) { clear_queue(o, &o->strq, F_STR); continue; }
            if (o->naive_strength && o->all_strength_res == 0) { naive_strengthening(o, use_heap, ignore); }
            else { strengthening_worker(o, use_heap, ignore); }
            clear_queue(o, &o->strq, F_STR);
        }
    }
    o->t_sub.modified = o->t_sub.modified || o->t_up.modified; /* propagation.appliedSomething(): last call */
    if (!o->t_sub.modified) { unsuccessful(&o->t_sub); }
    return o->t_sub.modified;
}

/* ------------------------------------------------------------------ BVE */
/* tryResolve, BVE.h:248-299: resolvent size or -1 for a tautology */
static int try_resolve(cp3o* o, int p, int n, int v)
{
    const int* c = CL(o, p); const int* d = CL(o, n);
    const int cn = o->csize[p], dn = o->csize[n];
    const int pv = 2 * v, nv = 2 * v + 1;
    int i = 0, j = 0, r = 0, prev = -2;
#define UPD(x) do { int _x = (x); if (prev != -2) { if (prev == _x) { break; } if (prev == (_x ^ 1)) { return -1; } } prev = _x; ++r; } while (0)
    while (i < cn && j < dn) {
        if (c[i] == pv) { ++i; }
        else if (d[j] == nv) { ++j; }
        else if (c[i] < d[j]) { UPD(c[i]); ++i; }
        else { UPD(d[j]); ++j; }
    }
    while (i < cn) { if (c[i] == pv) { ++i; } else { UPD(c[i]); ++i; } }
    while (j < dn) { if (d[j] == nv) { ++j; } else { UPD(d[j]); ++j; } }
#undef UPD
    return r;
}
/* resolve, BVE.h:208-242: returns 1 if tautology; resolvent literals in o->resolvent */
static int resolve(cp3o* o, int p, int n, int v)
{
    const int* c = CL(o, p); const int* d = CL(o, n);
    const int cn = o->csize[p], dn = o->csize[n];
    const int pv = 2 * v, nv = 2 * v + 1;
    ivec* ps = &o->resolvent; ps->n = 0;
    int i = 0, j = 0;
#define PUSH(x) do { int _x = (x); if (ps->n > 0) { if (ps->v[ps->n - 1] == _x) { break; } if (ps->v[ps->n - 1] == (_x ^ 1)) { return 1; } } iv_push(ps, _x); } while (0)
    while (i < cn && j < dn) {
        if (c[i] == pv) { ++i; }
        else if (d[j] == nv) { ++j; }
        else if (c[i] < d[j]) { PUSH(c[i]); ++i; }
        else { PUSH(d[j]); ++j; }
    }
    while (i < cn) { if (c[i] == pv) { ++i; } else { PUSH(c[i]); ++i; } }
    while (j < dn) { if (d[j] == nv) { ++j; } else { PUSH(d[j]); ++j; } }
#undef PUSH
    return 0;
}

static void growto(ivec* a, int n) { while (a->n < n) { iv_push(a, 0); } }

/* findGates, BVE.cc:1143-1267 (AND-gate detection; reorders occ lists, narrows p_limit / n_limit) */
static int find_gates(cp3o* o, int v, int* p_limit, int* n_limit)
{
    if (o->occ[2 * v].n < 3 && o->occ[2 * v + 1].n < 3) { return 0; }
    if (o->occ[2 * v].n < 2 || o->occ[2 * v + 1].n < 2) { return 0; }
    for (int pn = 0; pn < 2; ++pn) {
        const int pLit = 2 * v + (pn != 0), nLit = 2 * v + (pn == 0);
        ivec* pList = &o->occ[pLit]; ivec* nList = &o->occ[nLit];
        int* pClauses = pn == 0 ? p_limit : n_limit;
        int* nClauses = pn == 0 ? n_limit : p_limit;
        if (o->ma_step >= (1u << 30)) { memset(o->ma, 0, sizeof(uint32_t) * 2 * (size_t)o->nvars); o->ma_step = 0; }
        ++o->ma_step;
        for (int i = 0; i < pList->n; ++i) {
            int c = pList->v[i];
            if ((o->cflag[c] & F_DEL) || o->csize[c] != 2) { continue; }
            int* l = CL(o, c);
            int other = l[0] == pLit ? l[1] : l[0];
            o->ma[LNEG(other)] = o->ma_step;
        }
        for (int i = 0; i < nList->n; ++i) {
            int c = nList->v[i];
            if (o->cflag[c] & F_DEL) { continue; }
            int* l = CL(o, c); int j = 0;
            for (; j < o->csize[c]; ++j) {
                if (l[j] == nLit) { continue; }
                if (o->ma[l[j]] != o->ma_step) { break; }
            }
            if (j == o->csize[c]) {
                *pClauses = o->csize[c] - 1;
                *nClauses = 1;
                for (int k = 0; k < o->csize[c]; ++k) { o->ma[l[k]] = 0; }
                int tmp = nList->v[0]; nList->v[0] = nList->v[i]; nList->v[i] = tmp;
                int placed = 0;
                for (int k = 0; k < pList->n; ++k) {
                    int c2 = pList->v[k];
                    if ((o->cflag[c2] & F_DEL) || o->csize[c2] != 2) { continue; }
                    int* l2 = CL(o, c2);
                    int other = l2[0] == pLit ? l2[1] : l2[0];
                    if (o->ma[LNEG(other)] != o->ma_step) {
                        int t2 = pList->v[placed]; pList->v[placed] = pList->v[k]; pList->v[k] = t2;
                        placed++;
                        o->ma[LNEG(other)] = o->ma_step;
                    }
                }
                return 1;
            }
        }
    }
    return 0;
}

/* anticipateElimination, BVE.cc:695-880.  Returns L_FALSE on conflict, L_TRUE on early abort. */
static int anticipate(cp3o* o, int v, int p_limit, int n_limit, int* lit_clauses, int* resolvents)
{
    ivec* positive = &o->occ[2 * v]; ivec* negative = &o->occ[2 * v + 1];
    const int hasDefinition = (p_limit < positive->n || n_limit < negative->n);
    *lit_clauses = 0;
    int clausesToUse = 0;
    if (o->bve_early) {
        for (int i = 0; i < positive->n; ++i) { if (!(o->cflag[positive->v[i]] & F_DEL)) { clausesToUse++; } }
        for (int i = 0; i < negative->n; ++i) { if (!(o->cflag[negative->v[i]] & F_DEL)) { clausesToUse++; } }
        clausesToUse += o->bve_cgrow;
        if (o->bve_cgrow_t != INT32_MAX && o->bve_cgrow_t > o->bve_cgrow) { clausesToUse += o->bve_cgrow_t - o->bve_totallyAdded; }
    }
    int binariesOnly = 0;
    for (int cp = 0; cp < positive->n; ++cp) {
        const int p = positive->v[cp];
        if (o->cflag[p] & F_DEL) { continue; }
        if (binariesOnly && o->csize[p] > 2) { continue; }
        for (int cn = 0; cn < negative->n; ++cn) {
            if (cp >= p_limit && cn >= n_limit) { continue; }
            if (hasDefinition && cp < p_limit && cn < n_limit) { continue; }
            o->bve_steps++;
            const int n = negative->v[cn];
            if (o->cflag[n] & F_DEL) { continue; }
            if (binariesOnly && o->csize[n] > 2) { continue; }
            const int newLits = try_resolve(o, p, n, v);
            if (newLits > 1) {
                ++o->pos_stats.v[cp]; ++o->neg_stats.v[cn];
                *lit_clauses += newLits;
                (*resolvents)++;
                if (o->bve_early && *resolvents > clausesToUse) { binariesOnly = 1; }
            } else if (newLits == 0) {
                successful(&o->t_bve);
                o->ok = 0;
                return L_FALSE;
            } else if (newLits == 1) {
                resolve(o, p, n, v);
                int status = data_enqueue(o, o->resolvent.v[0]);
                if (status == L_FALSE) { successful(&o->t_bve); return L_FALSE; }
                else if (status == L_TRUE) {
                    successful(&o->t_bve);
                    ++o->bve_unitsEnqueued;
                    if (propagation_process(o, 1, 0, VAR_UNDEF) == L_FALSE) { return L_FALSE; }
                    o->t_bve.modified = o->t_bve.modified || o->t_up.modified;
                    if (o->cflag[p] & F_DEL) { break; }
                }
            }
        }
    }
    return binariesOnly ? L_TRUE : L_UNDEF;
}

/* removeClauses, BVE.cc:644-684 */
static void bve_remove_clauses(cp3o* o, int x, int ext_lit_or_undef)
{
    const int use_heap = o->bve_heap_updates > 0;
    ivec* list = &o->occ[x];
    for (int i = 0; i < list->n; ++i) {
        int c = list->v[i];
        if (o->cflag[c] & F_DEL) { continue; }
        removed_clause(o, c, use_heap, 0, VAR_UNDEF);
        successful(&o->t_bve);
        o->cflag[c] |= F_DEL;
        if (ext_lit_or_undef >= 0) { ext_clause(o, c, ext_lit_or_undef); }
        ++o->bve_removedClauses; o->bve_removedLiterals += o->csize[c];
    }
}
/* removeBlockedClauses, BVE.cc:1053-1090 */
static void bve_remove_blocked(cp3o* o, int x, const ivec* stats)
{
    const int use_heap = o->bve_heap_updates > 0;
    ivec* list = &o->occ[x];
    for (int i = 0; i < list->n; ++i) {
        int c = list->v[i];
        if (o->cflag[c] & F_DEL) { continue; }
        if (stats->v[i] == 0) {
            o->cflag[c] |= F_DEL;
            removed_clause(o, c, use_heap, 0, VAR_UNDEF);
            successful(&o->t_bve);
            ext_clause(o, c, x);
            ++o->bve_removedBC;
        }
    }
}

/* resolveSet, BVE.cc:892-1046 (force == false, no learnt clauses) */
static int resolve_set(cp3o* o, int v, int p_limit, int n_limit)
{
    const int use_heap = o->bve_heap_updates > 0;
    ivec* positive = &o->occ[2 * v]; ivec* negative = &o->occ[2 * v + 1];
    const int hasDefinition = (p_limit < positive->n || n_limit < negative->n);
    for (int cp = 0; cp < positive->n; ++cp) {
        const int p = positive->v[cp];
        if (o->cflag[p] & F_DEL) { continue; }
        for (int cn = 0; cn < negative->n; ++cn) {
            if (cp >= p_limit && cn >= n_limit) { continue; }
            if (hasDefinition && cp < p_limit && cn < n_limit) { continue; }
            o->bve_steps++;
            const int n = negative->v[cn];
            if (o->cflag[n] & F_DEL) { continue; }
            if (!resolve(o, p, n, v)) {
                if (o->resolvent.n == 1) {
                    int status = data_enqueue(o, o->resolvent.v[0]);
                    if (status == L_FALSE) { return L_FALSE; }
                    else if (status == L_TRUE) {
                        ++o->bve_unitsEnqueued;
                        if (propagation_process(o, 1, 0, VAR_UNDEF) == L_FALSE) { return L_FALSE; }
                        o->t_bve.modified = o->t_bve.modified || o->t_up.modified;
                        if (o->cflag[p] & F_DEL) { break; }
                    }
                } else if (o->resolvent.n > 1) {
                    int cr = alloc_clause(o, o->resolvent.v, o->resolvent.n);
                    /* alloc may move the pool; positive/negative point into occ which is stable */
                    data_add_clause(o, cr, use_heap);
                    iv_push(&o->clauses, cr);
                    sub_init_clause(o, cr);
                    ++o->bve_createdClauses; o->bve_createdLiterals += o->resolvent.n;
                }
            }
        }
    }
    return L_UNDEF;
}

/* bve_worker, BVE.cc:388-637 */
static void bve_worker(cp3o* o)
{
    const int unlimited = !o->limited;
    while (o->heap_n > 0 && (o->bve_steps < o->bve_limit || unlimited)) {
        const int v = heap_remove_min(o);
        if (o->assigns[v] != L_UNDEF) { continue; }
        if (o->occ[2 * v].n == 0 && o->occ[2 * v + 1].n == 0) { continue; }
        if (o->frozen[v]) { continue; }
        const int cutoff = (o->cnt[2 * v + 1] > 10 && o->cnt[2 * v] > 10)
                           || (dcount(o, v) > 15 && (o->cnt[2 * v + 1] > 5 || o->cnt[2 * v] > 5));
        if (!o->bve_force_gates && !o->bve_unlimited && cutoff) { ++o->bve_skippedVars; continue; }
        int p_limit = o->occ[2 * v].n, n_limit = o->occ[2 * v + 1].n;
        int foundGate = 0;
        if (o->bve_gates) { foundGate = find_gates(o, v, &p_limit, &n_limit); if (foundGate) { o->bve_foundGates++; } }
        if (!foundGate && o->bve_fdepOnly) { continue; }
        if (!o->bve_unlimited && !foundGate && cutoff) { ++o->bve_skippedVars; continue; }
        ++o->bve_testedVars;
        ivec* pos = &o->occ[2 * v]; ivec* neg = &o->occ[2 * v + 1];
        int pos_count = 0, neg_count = 0;
        for (int i = 0; i < pos->n; ++i) {
            if (o->cflag[pos->v[i]] & F_DEL) { pos->v[i] = pos->v[pos->n - 1]; pos->n--; --i; continue; }
            ++pos_count;
        }
        for (int i = 0; i < neg->n; ++i) {
            if (o->cflag[neg->v[i]] & F_DEL) { neg->v[i] = neg->v[neg->n - 1]; neg->n--; --i; continue; }
            ++neg_count;
        }
        growto(&o->pos_stats, pos->n); growto(&o->neg_stats, neg->n);
        int lit_clauses = 0, resolvents = 0;
        int ares = L_UNDEF;
        if (pos_count != 0 && neg_count != 0) {
            ++o->bve_anticipations;
            ares = anticipate(o, v, p_limit, n_limit, &lit_clauses, &resolvents);
            if (ares == L_FALSE) { return; }
        }
        if (ares == L_UNDEF && o->bve_BCElim) {
            bve_remove_blocked(o, 2 * v, &o->pos_stats);
            bve_remove_blocked(o, 2 * v + 1, &o->neg_stats);
        }
        int compareNumber = pos_count + neg_count + o->bve_cgrow;
        if (o->bve_cgrow_t != INT32_MAX && o->bve_cgrow_t > o->bve_cgrow) { compareNumber += o->bve_cgrow_t - o->bve_totallyAdded; }
        int reducedClss = resolvents <= compareNumber;
        if (o->bve_totallyAdded > o->bve_cgrow_t) { reducedClss = 0; }
        if (reducedClss) { o->bve_totallyAdded += resolvents - (pos_count + neg_count); }
        if (reducedClss && !o->bce_only) {
            if (resolvents < pos_count + neg_count) { o->bve_nDec++; }
            else if (resolvents > pos_count + neg_count) { o->bve_nInc++; }
            else { o->bve_nKeep++; }
            if (resolve_set(o, v, p_limit, n_limit) == L_FALSE) { return; }
            ++o->bve_eliminatedVars;
            bve_remove_clauses(o, 2 * v, -1);
            if (o->assigns[v] == L_UNDEF) {
                bve_remove_clauses(o, 2 * v + 1, 2 * v + 1);
                ext_lit(o, 2 * v);
            } else {
                bve_remove_clauses(o, 2 * v + 1, -1);
            }
            iv_release(&o->occ[2 * v]); iv_release(&o->occ[2 * v + 1]);
            if (o->bve_heap_updates > 0) { subsumption_process(o, o->bve_strength, 1, v); }
            else { subsumption_process(o, o->bve_strength, 0, VAR_UNDEF); }
            o->t_bve.modified = o->t_bve.modified || o->t_sub.modified;
            if (!o->ok) { return; }
        }
        o->pos_stats.n = 0; o->neg_stats.n = 0;
    }
}

static uint32_t next_mod_step(cp3o* o)
{   /* MarkArray::nextStep, SolverTypes.h:1181-1185 */
    if (o->modStep >= (1u << 30)) { memset(o->modTimer, 0, sizeof(uint32_t) * (size_t)o->nvars); o->modStep = 0; }
    return ++o->modStep;
}

/* touchedVarsForSubsumption / addClausesToSubsumption, BVE.cc:1118-1141 */
static void add_clauses_to_subsumption(cp3o* o, int x)
{
    ivec* li = &o->occ[x];
    for (int j = 0; j < li->n; ++j) {
        int d = li->v[j];
        if (!(o->cflag[d] & F_DEL) && !(o->cflag[d] & F_SUB)) {
            o->cflag[d] |= F_SUB | F_STR;
            sub_init_clause(o, d);
        }
    }
}

/* sequentiellBVE, BVE.cc:316-381 */
static void sequentiell_bve(cp3o* o)
{
    const int unlimited = !o->limited;
    subsumption_process(o, o->bve_strength, 0, VAR_UNDEF);
    o->t_bve.modified = o->t_bve.modified || o->t_sub.modified;
    if (!o->ok) { return; }
    o->touched.n = 0;
    while (o->heap_n > 0 && (o->bve_steps < o->bve_limit || unlimited)) {
        o->bve_modtime = next_mod_step(o);
        bve_worker(o);
        if (!o->ok) { return; }
        if (has_to_propagate(o)) {
            if (propagation_process(o, 1, 0, VAR_UNDEF) == L_FALSE) { return; }
            o->t_bve.modified = o->t_bve.modified || o->t_up.modified;
        }
        check_garbage(o);
        for (int v = 0; v < o->nvars; ++v) { if (o->modTimer[v] >= o->bve_modtime) { iv_push(&o->touched, v); } }
        for (int i = 0; i < o->touched.n; ++i) {
            add_clauses_to_subsumption(o, 2 * o->touched.v[i]);
            add_clauses_to_subsumption(o, 2 * o->touched.v[i] + 1);
        }
        subsumption_process(o, o->bve_strength, 0, VAR_UNDEF);
        o->t_bve.modified = o->t_bve.modified || o->t_sub.modified;
        heap_clear(o);
        for (int i = 0; i < o->touched.n; ++i) { heap_insert(o, o->touched.v[i]); }
        o->touched.n = 0;
    }
}

/* BoundedVariableElimination::process, BVE.cc:170-273 */
static int bve_process(cp3o* o)
{
    if (!perform_simplification(&o->t_bve)) { return L_UNDEF; }
    /* size gate BVE.cc:180-184 never fires while preprocessing (nTotLits()==0, see subsumption_process) */
    for (int v = 0; v < o->nvars; ++v) {
        if (o->modTimer[v] >= o->bve_modtime && !in_heap(o, v)) { heap_insert(o, v); }
    }
    if (propagation_process(o, 1, 0, VAR_UNDEF) == L_FALSE) { return L_FALSE; }
    const int propagatedSomething = o->t_up.modified;
    sequentiell_bve(o);
    if (!o->t_bve.modified) { unsuccessful(&o->t_bve); }
    if (o->t_bve.modified || propagatedSomething) { successful(&o->t_bve); }
    return o->ok ? L_UNDEF : L_FALSE;
}

/* ------------------------------------------------------------------ driver */
/* Solver::simplify at level 0 (Solver.cc:2212-2268): drop satisfied clauses, keep false literals */
/* ------------------------------------------------------------------ stand-alone BCE (coprocessor/techniques/BCE.cc)
 * blocked clause elimination proper (-bce with the defaults -bce-bce -no-bce-cle -no-bce-cla -no-bce-bcm) */
/* tautologicResolvent, BCE.cc:624-646: 1 iff resolving c (holds l) with d (holds ~l) on l gives a tautology */
static int tautologic_resolvent(cp3o* o, int c, int d, int l)
{
    const int* C = CL(o, c); const int* D = CL(o, d); const int cn = o->csize[c], dn = o->csize[d];
    int i = 0, j = 0;
    while (i < cn && j < dn) {
        if (C[i] == l) { i++; }
        else if (D[j] == LNEG(l)) { j++; }
        else if (C[i] == D[j]) { i++; j++; }
        else if (C[i] == LNEG(D[j])) { return 1; }
        else if (C[i] < D[j]) { i++; }
        else { j++; }
    }
    return 0;
}
/* Riss::Heap<LitOrderBCEHeapLt> (riss/mtl/Heap.h:37-183, BCE.h:27-36): min-heap of literals keyed by the
 * occurrence count of the complement (-bce-compl) or of the literal itself */
typedef struct { int* h; int n; int* idx; } lheap_t;
static inline int lheap_lt(const cp3o* o, int x, int y) { return o->bce_compl ? o->cnt[LNEG(x)] < o->cnt[LNEG(y)] : o->cnt[x] < o->cnt[y]; }
static void lheap_up(cp3o* o, lheap_t* H, int i)
{
    int x = H->h[i], p = (i - 1) >> 1;
    while (i != 0 && lheap_lt(o, x, H->h[p])) { H->h[i] = H->h[p]; H->idx[H->h[p]] = i; i = p; p = (p - 1) >> 1; }
    H->h[i] = x; H->idx[x] = i;
}
static void lheap_down(cp3o* o, lheap_t* H, int i)
{
    int x = H->h[i];
    while (2 * i + 1 < H->n) {
        int l = 2 * i + 1, r = 2 * i + 2;
        int child = (r < H->n && lheap_lt(o, H->h[r], H->h[l])) ? r : l;
        if (!lheap_lt(o, H->h[child], x)) { break; }
        H->h[i] = H->h[child]; H->idx[H->h[i]] = i; i = child;
    }
    H->h[i] = x; H->idx[x] = i;
}
static void lheap_insert(cp3o* o, lheap_t* H, int x) { H->idx[x] = H->n; H->h[H->n++] = x; lheap_up(o, H, H->idx[x]); }
static int lheap_remove_min(cp3o* o, lheap_t* H)
{
    int x = H->h[0];
    H->h[0] = H->h[H->n - 1]; H->idx[H->h[0]] = 0; H->idx[x] = -1; H->n--;
    if (H->n > 1) { lheap_down(o, H, 0); }
    return x;
}
/* blockedClauseElimination, BCE.cc:239-601 */
static void blocked_clause_elimination(cp3o* o)
{
    const int nl = 2 * o->nvars;
    lheap_t H; H.h = (int*)malloc(sizeof(int) * (size_t)(nl + 1)); H.idx = (int*)malloc(sizeof(int) * (size_t)(nl + 1)); H.n = 0;
    for (int x = 0; x < nl; ++x) { H.idx[x] = -1; }
    uint32_t* nextRound = (uint32_t*)calloc((size_t)nl + 1, sizeof(uint32_t)); uint32_t nr_step = 0;   /* MarkArray: all "current" at step 0 */
    ivec nextRoundLits = {0, 0, 0};
    for (int v = 0; v < o->nvars; ++v) {
        if (o->frozen[v]) { continue; }
        if (o->cnt[2 * v] > 0) { iv_push(&nextRoundLits, 2 * v); }
        if (o->cnt[2 * v + 1] > 0) { iv_push(&nextRoundLits, 2 * v + 1); }
    }
    ++o->ma_step;                                           /* data.ma.nextStep(), BCE.cc:269 */
    do {
        for (int i = 0; i < nextRoundLits.n; ++i) {
            const int l = nextRoundLits.v[i];
            if (nextRound[l] != nr_step) { continue; }
            lheap_insert(o, &H, l);
        }
        nextRoundLits.n = 0;
        ++nr_step;
        while (H.n > 0 && (!o->limited || o->bce_limit > o->bce_steps)) {
            const int right = lheap_remove_min(o, &H);
            if (o->frozen[LVAR(right)] || value_lit(o, right) != L_UNDEF) { continue; }
            o->bce_testedLits++;
            const int left = LNEG(right);
            for (int i = 0; i < o->occ[right].n; ++i) {
                if (has_to_propagate(o)) { break; }
                const int c = o->occ[right].v[i];
                if ((o->cflag[c] & F_DEL) || (!o->bce_binary && o->csize[c] == 2)) { continue; }
                int isBlocked = (o->csize[c] > 2 || o->bce_binary);
                if (!isBlocked) { break; }
                for (int j = 0; j < o->occ[left].n; ++j) {
                    if (has_to_propagate(o)) { break; }
                    const int d = o->occ[left].v[j];
                    if (o->cflag[d] & F_DEL) { continue; }
                    o->bce_steps++;
                    if (!tautologic_resolvent(o, c, d, right)) { isBlocked = 0; break; }
                }
                if (isBlocked) {
                    o->cflag[c] |= F_DEL;
                    removed_clause(o, c, 0, 0, VAR_UNDEF);
                    o->bce_remBCE++;
                    const int* l = CL(o, c);
                    for (int k = 0; k < o->csize[c]; ++k) {
                        const int nk = LNEG(l[k]);
                        if (H.idx[nk] >= 0) { lheap_up(o, &H, H.idx[nk]); lheap_down(o, &H, H.idx[nk]); }
                        else if (nextRound[nk] != nr_step) { iv_push(&nextRoundLits, nk); nextRound[nk] = nr_step; }
                    }
                    successful(&o->t_bce);
                    ext_clause(o, c, right);
                }
                if (has_to_propagate(o)) {
                    int prev = o->occ[right].n;
                    propagation_process(o, 1, 0, VAR_UNDEF);
                    o->t_bce.modified = o->t_bce.modified || o->t_up.modified;
                    if (!o->ok) { goto done; }
                    if (prev != o->occ[right].n) { i = -1; }
                }
            }
            if (has_to_propagate(o)) {
                propagation_process(o, 1, 0, VAR_UNDEF);
                o->t_bce.modified = o->t_bce.modified || o->t_up.modified;
                if (!o->ok) { goto done; }
            }
        }
        check_garbage(o);
    } while (nextRoundLits.n > 0 && (!o->limited || o->bce_limit > o->bce_steps));
done:
    free(H.h); free(H.idx); free(nextRound); iv_release(&nextRoundLits);
}
/* BlockedClauseElimination::process, BCE.cc:604-622 */
static int bce_process(cp3o* o)
{
    if (!perform_simplification(&o->t_bce)) { return 0; }
    o->t_bce.modified = 0;
    /* (the size gate needs nTotLits() > cp3_bce_lits, which is 0 after cleanSolver: it never fires, like cp3_susi_* / cp3_bve_*) */
    propagation_process(o, 1, 0, VAR_UNDEF);
    o->t_bce.modified = o->t_bce.modified || o->t_up.modified;
    if (!o->ok) { return o->t_bce.modified; }
    blocked_clause_elimination(o);
    return o->t_bce.modified;
}


static int solver_simplify(cp3o* o)
{
    if (!o->ok) { return 0; }
    if (o->qhead < o->trail.n && !ingest_propagate(o)) { o->ok = 0; return 0; }
    if (o->trail.n == o->simpDB_assigns || o->simpDB_props > 0) { return 1; }
    int j = 0; int64_t lits = 0;
    for (int i = 0; i < o->clauses.n; ++i) {
        int c = o->clauses.v[i]; int* l = CL(o, c); int sat = 0;
        for (int k = 0; k < o->csize[c]; ++k) { if (value_lit(o, l[k]) == L_TRUE) { sat = 1; break; } }
        if (sat) { ca_free(o, c); o->cflag[c] |= F_DEL; }
        else { o->clauses.v[j++] = c; lits += o->csize[c]; }
    }
    o->clauses.n = j;
    if (o->ca_wasted > o->ca_size * o->gc_frac) { o->ca_size -= o->ca_wasted; o->ca_wasted = 0; } /* Solver::garbageCollect */
    o->simpDB_assigns = o->trail.n; o->simpDB_props = lits;
    return 1;
}

/* data.init + initializePreprocessor + sortClauses, Coprocessor.cc:100-118,1243-1283,1447-1463 */
static void init_data(cp3o* o)
{
    const int n2 = 2 * o->nvars;
    o->occ = (ivec*)calloc((size_t)(n2 > 0 ? n2 : 1), sizeof(ivec));
    o->cnt = (int*)calloc((size_t)(n2 > 0 ? n2 : 1), sizeof(int));
    o->modTimer = (uint32_t*)calloc((size_t)(o->nvars > 0 ? o->nvars : 1), sizeof(uint32_t));
    o->ma = (uint32_t*)calloc((size_t)(n2 > 0 ? n2 : 1), sizeof(uint32_t));
    o->modStep = 0; o->numberOfCls = 0;
    o->subq.n = o->strq.n = 0;
    o->qhead = 0;                                   /* data.resetPPhead() */
    int kept = 0;
    for (int i = 0; i < o->clauses.n; ++i) {
        int c = o->clauses.v[i];
        if (o->cflag[c] & F_DEL) { continue; }
        data_add_clause(o, c, 0);
        sub_init_clause(o, c);
        o->clauses.v[kept++] = c;
    }
    o->clauses.n = kept;
    /* sortClauses (Coprocessor.cc:118,1447-1463): addClause_ sorts (Solver.cc:415), only the clauses propagate() permuted need it */
    for (int i = 0; i < o->permuted.n; ++i) {
        int c = o->permuted.v[i];
        qsort(CL(o, c), (size_t)o->csize[c], sizeof(int), cmp_int); o->cflag[c] &= ~F_PERM;
    }
    o->permuted.n = 0;
    o->data_ready = 1; o->data_nvars = o->nvars;
}

/* ------------------------------------------------------------------ Dense::compress (techniques/Dense.cc:26-238)
 * Variables that occur in no clause (and are not frozen) are dropped, the others are renumbered densely
 * (mapping[var] = var - #dropped before it), every clause literal is rewritten, assigned dropped variables move
 * from the trail into the compression object.  Called as the very last step of preprocess() (Coprocessor.cc:492-496). */
static void dense_compress(cp3o* o)
{
    if (propagation_process(o, 0, 0, VAR_UNDEF) == L_FALSE) { o->ok = 0; }      /* Dense.cc:28 */
    if (!o->ok) { return; }
    const int n = o->nvars;
    unsigned* count = (unsigned*)calloc((size_t)(n > 0 ? n : 1), sizeof(unsigned));
    for (int i = 0; i < o->clauses.n; ++i) {                                      /* countLiterals, Dense.h:92-121 */
        const int c = o->clauses.v[i];
        if (o->cflag[c] & F_DEL) { continue; }
        const int* l = CL(o, c);
        for (int j = 0; j < o->csize[c]; ++j) { count[LVAR(l[j])]++; }
    }
    int* mapping = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    unsigned char* val = (unsigned char*)malloc((size_t)(n > 0 ? n : 1));
    int diff = 0;
    for (int v = 0; v < n; ++v) {                                                 /* Dense.cc:49-87 */
        mapping[v] = -1; val[v] = L_UNDEF;
        if (count[v] != 0 || o->frozen[v] || (o->dense_keep_set && o->assigns[v] != L_UNDEF)) { mapping[v] = v - diff; }
        else { diff++; if (o->assigns[v] != L_UNDEF) { val[v] = o->assigns[v]; } }
    }
    /* isCompact, Dense.h:124-127 (integer division as in the reference) */
    if (diff == 0 || ((double)o->dense_frag > 1.0 + (double)(((unsigned)diff * 100u) / (unsigned)n))) {
        free(count); free(mapping); free(val);
        return;
    }
    o->comp_nvars = n; o->comp_map = mapping; o->comp_val = val;                /* compression.update */
    for (int i = 0; i < o->clauses.n; ++i) {                                      /* compressClauses, Dense.cc:406-450 */
        const int c = o->clauses.v[i];
        if (o->cflag[c] & F_DEL) { continue; }
        int* l = CL(o, c);
        for (int j = 0; j < o->csize[c]; ++j) { l[j] = 2 * mapping[LVAR(l[j])] + (l[j] & 1); }
    }
    int j = 0;                                                                    /* trail, Dense.cc:121-150 */
    for (int i = 0; i < o->trail.n; ++i) {
        const int x = o->trail.v[i], v = LVAR(x);
        if (count[v] != 0 || o->frozen[v] || o->dense_keep_set) { o->trail.v[j++] = 2 * mapping[v] + (x & 1); }
        else { o->assigns[v] = L_UNDEF; }
    }
    o->trail.n = j; o->qhead = j;                                                 /* propagation.reset */
    for (int v = 0; v < n; ++v) {                                                 /* moveVar, CoprocessorTypes.h:693-729 */
        const int to = mapping[v];
        if (to >= 0 && to != v) { o->assigns[to] = o->assigns[v]; o->assigns[v] = L_UNDEF; o->frozen[to] = o->frozen[v]; o->frozen[v] = 0; }
    }
    o->nvars = n - diff; o->data_nvars = o->nvars;                              /* CoprocessorTypes.h:725 */
    free(count);
}

/* Dense::decompress (Dense.cc:240-318): model over the compressed variables -> model over the original ones.
 * model[v] in {L_TRUE, L_FALSE}; the caller provides room for cp3o_model_vars() entries. */
int cp3o_model_vars(const cp3o* o) { return o->comp_map ? o->comp_nvars : o->nvars; }
void cp3o_decompress_model(const cp3o* o, unsigned char* model, int have)
{
    if (!o->comp_map) { return; }
    for (int v = have; v < o->comp_nvars; ++v) { model[v] = L_FALSE; }            /* growTo(nvars, l_False) */
    for (int v = o->comp_nvars - 1; v >= 0; --v) {
        const int c = o->comp_map[v];
        if (c < 0) { if (o->comp_val[v] != L_UNDEF) { model[v] = o->comp_val[v]; } }
        else if (c != v) { model[v] = model[c]; }
    }
}
/* Preprocessor::extendModel (Coprocessor.cc:1214-1240): Dense::decompress, then CoprocessorData::extendModel
 * (CoprocessorTypes.h:1665-1764) - walk the undo stack backwards; an entry none of whose literals is satisfied gets
 * its first literal (the blocking literal) satisfied.  `model` has room for cp3o_model_vars() entries. */
void cp3o_extend_model(const cp3o* o, unsigned char* model, int have)
{
    const int n = cp3o_model_vars(o);
    cp3o_decompress_model(o, model, have);
    if (!o->comp_map) { for (int v = have; v < n; ++v) { model[v] = L_TRUE; } }
    for (int v = 0; v < n; ++v) { if (model[v] == L_UNDEF) { model[v] = L_TRUE; } }
    for (int i = o->undo.n - 1; i >= 0; --i) {
        const int d = o->undo.v[i];
        if (d == 0) { const int s = o->undo.v[i + 1]; model[(s < 0 ? -s : s) - 1] = s < 0 ? L_FALSE : L_TRUE; continue; }
        const int v = (d < 0 ? -d : d) - 1;
        if (model[v] == (d < 0 ? L_FALSE : L_TRUE)) { while (o->undo.v[i] != 0) { --i; } }
    }
}
int cp3o_compressed_var(const cp3o* o, int v) { return o->comp_map ? (v < o->comp_nvars ? o->comp_map[v] : -1) : v; }

void cp3o_init_only(cp3o* o)
{
    if (!solver_simplify(o)) { return; }
    init_data(o);
}

void cp3o_preprocess(cp3o* o)
{
    /* Coprocessor.cc:1002-1006: the size gates.  data.nVars() is CoprocessorData::numberOfVars - still 0 before the first
     * data.init(); nTotLits() = clauses_literals + learnts_literals (riss/core/Solver.h:1867), the literals of all attached
     * clauses (Solver::attachClause, Solver.cc:706-707): every stored clause here, there are no learnt clauses */
    if (o->limited) {
        uint32_t totlits = 0;
        for (int i = 0; i < o->clauses.n; ++i) { totlits += (uint32_t)o->csize[o->clauses.v[i]]; }
        if ((uint32_t)o->data_nvars > (uint32_t)o->cp3_vars || o->clauses.n > o->cp3_cls || totlits > (uint32_t)o->cp3_lits) { return; }
    }
    if (!o->enabled) { return; }
    if (!solver_simplify(o)) { return; }
    init_data(o);
    int status = L_UNDEF;
    for (int it = 0; it < o->iters; ++it) {
        if (o->opt_up) { if (status == L_UNDEF) { status = propagation_process(o, 1, 0, VAR_UNDEF); } }
        check_garbage(o);                       /* unconditional checkpoints behind the UP / XOR / ENT blocks, Coprocessor.cc:156,171 */
        if (!o->ok) { break; }
        if (o->opt_subsimp) {
            if (status == L_UNDEF) { subsumption_process(o, 1, 0, VAR_UNDEF); }
            if (!o->ok) { status = L_FALSE; }
            check_garbage(o);
        }
        if (!o->ok) { break; }
        if (o->opt_bve) {
            if (status == L_UNDEF) { status = bve_process(o); }
            check_garbage(o);
        }
        if (!o->ok) { break; }
        if (o->opt_bce) {                                  /* Coprocessor.cc:345-352 */
            if (status == L_UNDEF) { bce_process(o); }
            check_garbage(o);
        }
        if (!o->ok) { break; }
    }
    if (o->ok && has_to_propagate(o)) { propagation_process(o, 0, 0, VAR_UNDEF); }
    if (o->opt_dense) { dense_compress(o); }                                      /* Coprocessor.cc:492-496 */
    /* reSetupSolver, CoprocessorTypes.h:937-1008: drop deleted clauses, keep order */
    if (o->ok) { filter_deleted(o, &o->clauses); }
}

int cp3o_ok(const cp3o* o) { return o->ok; }
int cp3o_nvars(const cp3o* o) { return o->nvars; }
size_t cp3o_ntrail(const cp3o* o) { return (size_t)o->trail.n; }
size_t cp3o_nclauses(const cp3o* o) { return (size_t)o->clauses.n; }

size_t cp3o_output(const cp3o* o, int32_t* out, size_t cap)
{
    size_t k = 0;
#define EMIT(x) do { if (k < cap) { out[k] = (x); } ++k; } while (0)
    for (int i = 0; i < o->trail.n; ++i) { EMIT(to_dimacs(o->trail.v[i])); EMIT(0); }
    for (int i = 0; i < o->clauses.n; ++i) {
        int c = o->clauses.v[i];
        if (o->cflag[c] & F_DEL) { continue; }
        const int* l = o->pool + o->cstart[c];
        for (int j = 0; j < o->csize[c]; ++j) { EMIT(to_dimacs(l[j])); }
        EMIT(0);
    }
#undef EMIT
    return k;
}
size_t cp3o_undo(const cp3o* o, int32_t* out, size_t cap)
{
    for (int i = 0; i < o->undo.n && (size_t)i < cap; ++i) { out[i] = o->undo.v[i]; }
    return (size_t)o->undo.n;
}

int64_t cp3o_stat(const cp3o* o, const char* s)
{
#define ST(name, val) if (!strcmp(s, name)) { return (int64_t)(val); }
    ST("sub_subsumedClauses", o->subsumedClauses) ST("sub_subsumedLiterals", o->subsumedLiterals)
    ST("sub_removedLiterals", o->removedLiterals) ST("sub_subsSteps", o->sub_steps) ST("sub_strengthSteps", o->str_steps)
    ST("bve_eliminatedVars", o->bve_eliminatedVars) ST("bve_createdClauses", o->bve_createdClauses)
    ST("bve_removedClauses", o->bve_removedClauses) ST("bve_testedVars", o->bve_testedVars)
    ST("bve_anticipations", o->bve_anticipations) ST("bve_skippedVars", o->bve_skippedVars)
    ST("bve_unitsEnqueued", o->bve_unitsEnqueued) ST("bve_removedBC", o->bve_removedBC)
    ST("bve_totallyAddedClauses", o->bve_totallyAdded) ST("bve_steps", o->bve_steps)
    ST("up_removedClauses", o->up_removedClauses) ST("up_removedLiterals", o->up_removedLiterals)
    ST("bve_foundGates", o->bve_foundGates)
    ST("bce_steps", o->bce_steps) ST("bce_testedLits", o->bce_testedLits) ST("bce_remBCE", o->bce_remBCE)
#undef ST
    return -1;
}

/* ------------------------------------------------------------------ snapshot-level oracles */
int64_t cp3o_subsume_snapshot(cp3o* o, uint8_t* deleted)
{
    int nc = o->clauses.n;
    int* ids = (int*)malloc(sizeof(int) * (size_t)(nc > 0 ? nc : 1));
    memcpy(ids, o->clauses.v, sizeof(int) * (size_t)nc);
    int saved = o->limited; o->limited = 0;
    int64_t before = o->sub_steps;
    subsumption_worker(o, 0, VAR_UNDEF);
    clear_queue(o, &o->subq, F_SUB);
    o->limited = saved;
    for (int i = 0; i < nc; ++i) { deleted[i] = (o->cflag[ids[i]] & F_DEL) ? 1 : 0; }
    free(ids);
    return o->sub_steps - before;
}

void cp3o_anticipate_snapshot(cp3o* o, int32_t* out, int32_t* out_small)
{
    for (int v = 0; v < o->nvars; ++v) {
        ivec* pos = &o->occ[2 * v]; ivec* neg = &o->occ[2 * v + 1];
        int res = 0, lc = 0, small = 0;
        for (int i = 0; i < pos->n; ++i) {
            for (int j = 0; j < neg->n; ++j) {
                int r = try_resolve(o, pos->v[i], neg->v[j], v);
                if (r > 1) { ++res; lc += r; } else if (r >= 0) { ++small; }
            }
        }
        out[4 * v] = pos->n; out[4 * v + 1] = neg->n; out[4 * v + 2] = res; out[4 * v + 3] = lc;
        if (out_small) { out_small[v] = small; }
    }
}

/* ------------------------------------------------------------------ options / lifetime */
cp3o* cp3o_new(void)
{
    cp3o* o = (cp3o*)calloc(1, sizeof(cp3o));
    o->ok = 1;
    o->limited = 1; o->iters = 1;
    o->cp3_vars = 5000000; o->cp3_cls = 30000000; o->cp3_lits = 50000000;
    o->strength = 1; o->sub_limit = 300000000; o->str_limit = 300000000; o->call_inc = 200;
    o->bve_limit = 25000000; o->bve_strength = 1; o->bve_gates = 1; o->bve_cgrow = 0; o->bve_cgrow_t = 1000;
    o->bve_heap_updates = 1;
    o->bce_compl = 1; o->bce_limit = 100000000;
    o->gc_frac = 0.20;
    o->simpDB_assigns = -1;
    o->sub_lim = o->sub_limit; o->str_lim = o->str_limit;
    return o;
}

int cp3o_option(cp3o* o, const char* opt)
{
    if (opt[0] != '-') { return 0; }
    const char* s = opt + 1;
    int neg = 0;
    if (!strncmp(s, "no-", 3)) { neg = 1; s += 3; }
    const char* eq = strchr(s, '=');
    char name[64]; size_t nl = eq ? (size_t)(eq - s) : strlen(s);
    if (nl >= sizeof(name)) { return 0; }
    memcpy(name, s, nl); name[nl] = 0;
    long long val = eq ? atoll(eq + 1) : 0;
#define BOPT(n, f) if (!strcmp(name, n)) { o->f = !neg; return 1; }
#define IOPT(n, f) if (!strcmp(name, n) && eq) { o->f = val; return 1; }
    BOPT("enabled_cp3", enabled) BOPT("up", opt_up) BOPT("subsimp", opt_subsimp) BOPT("bve", opt_bve)
    BOPT("cp3_limited", limited) BOPT("cp3_strength", strength) BOPT("naive_strength", naive_strength)
    BOPT("bve_unlimited", bve_unlimited) BOPT("bve_strength", bve_strength) BOPT("bve_gates", bve_gates)
    BOPT("bve_force_gates", bve_force_gates) BOPT("bve_fdepOnly", bve_fdepOnly) BOPT("bve_BCElim", bve_BCElim)
    BOPT("bve_early", bve_early) BOPT("bce_only", bce_only) BOPT("dense", opt_dense) BOPT("cp3_keep_set", dense_keep_set)
    IOPT("cp3_dense_frag", dense_frag)
    BOPT("bce", opt_bce) BOPT("bce-compl", bce_compl) BOPT("bce-bin", bce_binary) IOPT("bce-limit", bce_limit)
    IOPT("cp3_iters", iters) IOPT("cp3_vars", cp3_vars) IOPT("cp3_cls", cp3_cls) IOPT("cp3_lits", cp3_lits)
    /* technique-level size gates: accepted, without effect (see subsumption_process) */
    if ((!strcmp(name, "cp3_susi_vars") || !strcmp(name, "cp3_susi_cls") || !strcmp(name, "cp3_susi_lits") ||
         !strcmp(name, "cp3_bve_vars") || !strcmp(name, "cp3_bve_cls") || !strcmp(name, "cp3_bve_lits") ||
         !strcmp(name, "cp3_bce_vars") || !strcmp(name, "cp3_bce_cls") || !strcmp(name, "cp3_bce_lits")) && eq) { return 1; }
    IOPT("cp3_call_inc", call_inc) IOPT("cp3_bve_limit", bve_limit) IOPT("cp3_bve_heap", bve_heap)
    IOPT("bve_cgrow", bve_cgrow) IOPT("bve_cgrow_t", bve_cgrow_t) IOPT("bve_heap_updates", bve_heap_updates)
    if (!strcmp(name, "cp3_sub_limit") && eq) { o->sub_limit = o->sub_lim = val; return 1; }
    if (!strcmp(name, "cp3_str_limit") && eq) { o->str_limit = o->str_lim = val; return 1; }
#undef BOPT
#undef IOPT
    return 0;
}

void cp3o_free(cp3o* o)
{
    if (!o) { return; }
    if (o->occ) { for (int x = 0; x < 2 * o->nvars; ++x) { free(o->occ[x].v); } free(o->occ); }
    if (o->watches) { for (int x = 0; x < 2 * o->nvars; ++x) { free(o->watches[x].v); } free(o->watches); }
    free(o->permuted.v);
    free(o->cnt); free(o->modTimer); free(o->ma); free(o->assigns); free(o->frozen);
    free(o->pool); free(o->cstart); free(o->csize); free(o->cflag);
    free(o->trail.v); free(o->clauses.v); free(o->cur.v); free(o->subq.v); free(o->strq.v); free(o->undo.v);
    free(o->heap); free(o->hidx); free(o->pos_stats.v); free(o->neg_stats.v); free(o->resolvent.v);
    free(o->touched.v); free(o->occ_upd); free(o->sub_occ_updates.v); free(o->comp_map); free(o->comp_val);
    free(o);
}
