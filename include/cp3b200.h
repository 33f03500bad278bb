/* cp3b200.h - C ABI of libcp3b200.so, the B200-native drop-in for Coprocessor's subsumption /
 * strengthening / BVE hot path (nmanthey/riss-solver).
 *
 * Two layers, both plain C (no torch / C++ types in any signature):
 *
 *  1. CP*   - the reference's own preprocessor C interface, coprocessor/libcoprocessorc.h:46-133, with the
 *             same names, argument meaning and error behaviour (UNSAT is reported by CPok()==0, unknown
 *             options are ignored unless `strict`).  A tool linked against the reference's libcp3 can be
 *             re-linked against libcp3b200 unchanged.  Each declaration cites the line it replaces.
 *  2. cp3g_* - the pass-level "scan service" the reference's C++ techniques bind to when they keep their
 *             own host data structures (INTEGRATION.md shows the stub): upload the flattened clause arena,
 *             then ask the GPU which (subsumer,candidate) / (strengthener,candidate,literal) pairs hit and
 *             for per-variable resolvent counts and resolvents.  These replace the inner loops of
 *             Subsumption::subsumption_worker (Subsumption.cc:244-297), strengthening_worker (:1276-1521),
 *             BoundedVariableElimination::anticipateElimination (BVE.cc:695-880) and resolveSet (:892-1046).
 *
 * There is NO CPU fallback: every entry point that needs the device fails loudly (CP* abort with a message
 * on stderr, cp3g_* return a negative error code) when no CUDA device is usable.
 */
#ifndef CP3B200_H
#define CP3B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Layer 1: libcoprocessorc.h drop-in
 * ---------------------------------------------------------------------------------------------- */
void* CPinit(const char* presetConfig);                         /* libcoprocessorc.h:49  */
void  CPpreprocess(void* preprocessor);                         /* :52  */
void  CPdestroy(void** preprocessor);                           /* :55  */
void  CPparseOptions(void* preprocessor, int* argc, char** argv, int strict); /* :58 */
void  CPparseOptionString(void* preprocessor, char* argv);      /* :59  */
void  CPsetPresetConfig(void* preprocessor, const char* presetConfig);       /* :63 */
void  CPaddLit(void* preprocessor, int lit);                    /* :66  */
float CPversion(void* preprocessor);                            /* :69  */
int   CPhasNextOutputLit(void* preprocessor);                   /* :72  */
int   CPnextOutputLit(void* preprocessor);                      /* :75  */
void  CPresetOutput(void* preprocessor);                        /* :78  */
void  CPfreezeVariable(void* preprocessor, int variable);       /* :83  */
int   CPgetReplaceLiteral(void* preprocessor, int oldLit);      /* :89  */
void  CPresetModel(void* preprocessor);                         /* :97  */
void  CPpushModelBool(void* preprocessor, int pol);             /* :102 */
void  CPgiveModelLit(void* preprocessor, int literal);          /* :105 */
void  CPpostprocessModel(void* preprocessor);                   /* :108 */
int   CPgetFinalModelLit(void* preprocessor);                   /* :111 */
int   CPmodelVariables(void* preprocessor);                     /* :114 */
int   CPnVars(void* preprocessor);                              /* :120 */
int   CPnClss(void* preprocessor);                              /* :123 */
int   CPnLits(void* preprocessor);                              /* :126 */
int   CPok(void* preprocessor);                                 /* :129 */
int   CPlitFalsified(void* preprocessor, int lit);              /* :132 */

/* Bulk extensions (not in the reference header; same semantics as the per-literal calls above). */
/* CPaddLit for a whole 0-terminated DIMACS stream */
void   CPaddLits(void* preprocessor, const int32_t* stream, size_t n);
/* CPresetOutput + CPnextOutputLit until exhausted; returns the int32 count needed, writes up to cap */
size_t CPgetOutput(void* preprocessor, int32_t* out, size_t cap);
/* number of leading unit clauses (trail) in that stream */
size_t CPoutputUnits(void* preprocessor);
/* extension (undo) stack as DIMACS literals, 0 separates entries (CoprocessorTypes.h:1615-1655) */
size_t CPgetUndo(void* preprocessor, int32_t* out, size_t cap);
/* technique counters by name: the `[STAT] SuSi` / `BVE` fields of Subsumption.cc:38-49, BVE.cc:75-108
 * ("sub_subsSteps", "sub_strengthSteps", "sub_subsumedClauses", "bve_eliminatedVars", ...) plus GPU-side
 * counters ("gpu_launches", "gpu_scan_requests", "gpu_h2d_bytes", "gpu_d2h_bytes", "gpu_kernel_ns") */
int64_t CPstat(void* preprocessor, const char* name);
/* multi-GPU: shard the scan requests of this instance over `world` ranks (rank-local GPU each);
 * `nccl_unique_id` is the 128-byte ncclUniqueId created by rank 0 (cp3g_nccl_unique_id) */
int    CPsetDistributed(void* preprocessor, int rank, int world, const void* nccl_unique_id);

/* ------------------------------------------------------------------------------------------------
 * Layer 2: pass-level GPU scan service
 * ---------------------------------------------------------------------------------------------- */
typedef struct cp3g cp3g;

/* literal code = Riss::toInt(Lit) = 2*var + sign (riss/core/SolverTypes.h:65-98); clauses are sorted by
 * that code (Coprocessor.cc:1447-1463).  Clause index = position in the uploaded table. */
typedef struct { uint32_t req; uint32_t clause; int32_t lit; } cp3g_hit;

enum { CP3G_SUBSUME = 0, CP3G_STRENGTHEN = 1 };
enum { CP3G_OK = 0, CP3G_ERR_NODEVICE = -1, CP3G_ERR_CUDA = -2, CP3G_ERR_OVERFLOW = -3, CP3G_ERR_ARG = -4 };

int  cp3g_device_count(void);
cp3g* cp3g_create(int device);                 /* NULL when the device cannot be used */
void  cp3g_destroy(cp3g* g);
const char* cp3g_last_error(const cp3g* g);

/* Upload the flattened clause table (replaces walking ClauseAllocator memory, SolverTypes.h:638-699):
 * lits = literal pool, start[i]/size[i] = clause i, flags bit0 = deleted (Clause::mark()). */
int cp3g_upload(cp3g* g, uint32_t nvars, uint32_t nclauses, const int32_t* lits, size_t nlits,
                const uint32_t* start, const uint32_t* size, const uint8_t* flags);
/* same, from 16-byte host records {start:u32, size:24|flags:8 (bit 0 = deleted), 2 words ignored} - lets a host that keeps
 * such a clause table upload it with one copy */
int cp3g_upload_packed(cp3g* g, uint32_t nvars, uint32_t nclauses, const int32_t* lits, size_t nlits, const void* records16);
/* K1 sig_build + K2 occ_build: 64-bit clause signatures and the per-literal CSR occurrence lists in
 * ascending clause order (CoprocessorData::addClause order, CoprocessorTypes.h:812-833). */
int cp3g_build(cp3g* g);
/* copy the CSR back: off[2*nvars+1], refs[off[2*nvars]] (pass NULL refs to query sizes only) */
int cp3g_download_occ(cp3g* g, uint32_t* off, uint32_t* refs);
/* incremental maintenance between scans */
int cp3g_set_deleted(cp3g* g, uint32_t n, const uint32_t* clauses);
int cp3g_shrink(cp3g* g, uint32_t n, const uint32_t* clauses, const uint32_t* new_start,
                const uint32_t* new_size, const int32_t* lits, size_t nlits);   /* rewritten literals */
int cp3g_append(cp3g* g, uint32_t n, const uint32_t* size, const int32_t* lits, size_t nlits);

/* K3/K4: one request = one subsumer / strengthener given by its literals (req_off[i]..req_off[i+1] in
 * req_lits) and the clause index to skip (self, UINT32_MAX for none).  With req_lits == NULL the device's
 * own copy of clause self[i] is used.  Hits are appended unordered to `out`:
 *   CP3G_SUBSUME:    (req, D, -1)  for every live D != self with request ⊆ D
 *   CP3G_STRENGTHEN: (req, D, l)   for every live D where exactly the literal l of D is the complement of a
 *                                  request literal and the rest of the request is contained in D
 * With out == NULL the hits are only counted (they stay on the device).
 * After cp3g_nccl_init the requests are partitioned over the ranks and the hit lists all-gathered. */
int cp3g_scan(cp3g* g, int kind, uint32_t nreq, const uint32_t* self, const uint32_t* req_off,
              const int32_t* req_lits, cp3g_hit* out, uint32_t cap, uint32_t* nout, uint64_t* checks);

/* K3/K4 bulk form: every live clause is a request (the first pass over the full queue); hit.req is the clause
 * index of the subsumer / strengthener.  One thread per occurrence entry: each list is streamed once for all
 * of its subsumers.  Rebuilds signatures + CSR first when clauses were edited since the last cp3g_build. */
int cp3g_scan_all(cp3g* g, int kind, cp3g_hit* out, uint32_t cap, uint32_t* nout, uint64_t* checks);

/* The first subsumption pass COMMITTED on the device (commit.cu): Subsumption::subsumption_worker, Subsumption.cc:244-313, for the
 * state right after Preprocessor::initializePreprocessor - every live clause in the queue in clause order, occurrence lists
 * pristine (cp3g_build, nothing edited since).  Runs K3 over the full queue and then, on the device, the death time of every
 * subsumed clause, the list choice of every queue entry with lit_occurrence_count as of its own turn, and the replay of the
 * lazily cleaned lists (swap-with-last, size filter, step counter) - the result of the reference's ordered walk.
 * *status 0: committed - *ndead clauses subsumed (*nlits literals), *steps subsumption steps, *checks K3 checks; the device
 * marked them deleted.  *status 1: not applicable (a list holding a subsumed clause is longer than max_list, sharded mode,
 * clauses appended since the build): nothing was changed, use cp3g_scan_all and commit on the host. */
int cp3g_sub_pass(cp3g* g, uint32_t max_list, int* status, uint32_t* ndead, uint64_t* nlits, uint64_t* steps, uint64_t* checks);
/* results of the committed pass: dead[ndead] clause ids (any order), list_n[2*nvars] / list_stale[2*nvars] = length of every
 * list and number of deleted clauses still listed in it, refs = the lists at the offsets of cp3g_download_occ in their
 * post-pass order (first list_n[x] entries of list x).  NULL skips an output. */
int cp3g_sub_pass_fetch(cp3g* g, uint32_t* dead, uint32_t* list_n, uint32_t* list_stale, uint32_t* refs);

/* K10: the static part of the strengthening pass that follows a committed cp3g_sub_pass (strengthening_worker,
 * Subsumption.cc:1276-1310 + the list walks of the entries that no hit touches), on the tables the device holds after that pass.
 * real_bits: one bit per clause, set = the host commits this queue entry itself (it has hits or is the target of one).  Every
 * other live clause of size >= 2 chooses the literal whose variable has the smallest occurrence count (counts after the pass, the
 * first such literal in clause order), adds live_len(min) - 1 + live_len(~min) steps and marks both lists as walked.
 * Out: number of such entries, their step total, live clauses of size < 2 outside real_bits (units in the queue: the caller must
 * take the ordered walk then), walked_bits (one bit per literal, 2*nvars bits).  *status 1: not applicable (no committed pass,
 * or the data base changed since) - nothing is returned. */
int cp3g_str_static(cp3g* g, const uint32_t* real_bits, int* status, uint64_t* inert_entries, uint64_t* inert_steps, uint64_t* small_clauses,
                    uint32_t* walked_bits);

/* K5: resolvent anticipation for candidate variables.  Variable i owns pos refs p_off[i]..p_off[i+1] and
 * neg refs n_off[i]..n_off[i+1] (host list order).  sizes[pair_off[i] + a*nneg + b] = number of literals
 * of the resolvent of pos a and neg b on vars[i], -1 for a tautology (BVE.h:248-299), -2 when a parent is
 * already deleted. */
int cp3g_bve_pairs(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, const uint32_t* p_refs,
                   const uint32_t* n_off, const uint32_t* n_refs, const uint64_t* pair_off, int16_t* sizes);
/* K6: materialise the resolvents of the listed (var, pos clause, neg clause) triples into out_lits;
 * out_off[k] is given by the caller as the prefix sum of the sizes K5 reported (BVE.h:208-242). */
int cp3g_bve_resolve(cp3g* g, uint32_t npairs, const uint32_t* vars, const uint32_t* p_refs,
                     const uint32_t* n_refs, const uint64_t* out_off, int32_t* out_lits, size_t nlits);

/* K9: stand-alone blocked clause elimination (coprocessor/techniques/BCE.cc:239-601).  Replaces every call of
 * BlockedClauseElimination::tautologicResolvent (BCE.cc:624-646) of one BCE run: variable i owns pos refs p_off[i]..p_off[i+1]
 * and neg refs n_off[i]..n_off[i+1] (host list order); row a of variable i is the W_i = ceil(nneg_i / 32) words
 * masks[word_off[i] + a*W_i ..), bit b of the row is set iff the resolvent of pos clause a and neg clause b on vars[i] is NOT a
 * tautology.  word_off[nv] = total number of words. */
int cp3g_bce_masks(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, const uint32_t* p_refs,
                   const uint32_t* n_off, const uint32_t* n_refs, const uint64_t* word_off, uint32_t* masks);

/* K5': the whole per-variable evaluation of bve_worker (BVE.cc:388-637) for a batch of candidate variables, one warp each
 * (bve_eval.cu): gate detection with list reordering (findGates, BVE.cc:1143-1267; use_gates = -bve_gates), lazy removal of the
 * deleted entries by swap-with-last (BVE.cc:474-489), resolvent anticipation over the pairs the gate limits leave
 * (anticipateElimination, BVE.cc:695-880).  Variable i owns p_refs[p_off[i]..p_off[i+1]) / n_refs[...] in the host's current
 * list ORDER; both arrays are rewritten in place with the cleaned, reordered lists (first pos_n / neg_n entries), p_stats /
 * n_stats (may be NULL) receive the per-clause resolvent counts (pos_stats / neg_stats, a clause with 0 is blocked,
 * BVE.cc:1053-1090).  CP3G_BVE_HOST in flags: a unit or empty resolvent exists or a list exceeds the kernel's capacity - the
 * record is void, evaluate this variable through cp3g_bve_pairs. */
typedef struct { int32_t p_limit, n_limit; uint32_t pos_n, neg_n, resolvents, lit_clauses, steps, flags; } cp3g_bve_rec;
enum { CP3G_BVE_GATE = 1, CP3G_BVE_HOST = 2 };
int cp3g_bve_eval(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, uint32_t* p_refs, const uint32_t* n_off, uint32_t* n_refs,
                  int use_gates, cp3g_bve_rec* recs, uint16_t* p_stats, uint16_t* n_stats);
/* K6': resolvents of the batch variables sel[0..nsel) (indices into the last cp3g_bve_eval batch) in resolveSet's order
 * (BVE.cc:892-1046).  Variable sel[k] has recs[].resolvents resolvents with recs[].lit_clauses literals in total: the caller
 * passes their prefix sums res_off / lit_off (nsel+1 entries); out_sizes[res_off[k]..) = size per resolvent, out_lits[lit_off[k]..)
 * = their literals back to back. */
int cp3g_bve_stream(cp3g* g, uint32_t nsel, const uint32_t* sel, const uint64_t* res_off, const uint64_t* lit_off, uint16_t* out_sizes, int32_t* out_lits);

/* K7: Dense::compress (coprocessor/techniques/Dense.cc:26-238, countLiterals Dense.h:92-121, compressClauses
 * Dense.cc:406-450).  A variable stays iff it occurs in a live clause or keep[v] != 0 (frozen / kept assignment);
 * the survivors are renumbered densely in order (mapping[v] = v - #removed before v, UINT32_MAX for removed).
 * When the formula is not compact (Dense.h:124-127: some variable is removed and frag_percent <= 1 + removed*100/nvars)
 * every literal of every live clause is rewritten in HBM and the literal pool (nlits ints, same offsets as uploaded /
 * appended) is copied to lits_out; *compressed tells whether that happened.  The device then holds nvars_new variables. */
int cp3g_dense(cp3g* g, const uint8_t* keep, int frag_percent, uint32_t* mapping, uint32_t* nvars_new, int* compressed,
               int32_t* lits_out, size_t nlits);

/* K0: bulk ingest of a DIMACS literal stream, 0 terminates a clause (the CPaddLit sequence, libcoprocessorc.cc:251-266 ->
 * Solver::addClause_, riss/core/Solver.cc:409-499, for a solver without assignments): per clause sort by literal code,
 * drop duplicate literals, drop tautologies; the clause arena is built on the device exactly as cp3g_upload_packed would
 * receive it from a host that ingested the stream literal by literal.  `n` must end on a terminating 0.
 * *status: 0 = done; 1 = some clause becomes unit or empty (unit propagation while clauses are added is order dependent)
 * or is longer than 4096 literals - the device data base is left empty, ingest on the host instead. */
int cp3g_ingest(cp3g* g, const int32_t* stream, size_t n, uint32_t* nvars, uint32_t* nclauses, size_t* nlits, int* status);
/* copy the arena to the host: lits[nlits] and 16-byte records {start:u32, size:24|flags:8 (bit 0 deleted, bits 1-2 the
 * can_subsume / can_strengthen flags of a fresh clause), 0, 0} - the layout cp3g_upload_packed takes */
int cp3g_download_arena(cp3g* g, int32_t* lits, void* records16);

/* K8: stream compaction of the clause arena (CoprocessorData::garbageCollect / relocAll, CoprocessorTypes.h:1396-1550;
 * ClauseAllocator::reloc).  The literals of the live clauses are packed in clause order (exclusive scan of the live
 * sizes, gather), the holes left by deleted and shortened clauses disappear.  new_start[i] (nclauses entries) is the
 * new pool offset of clause i (deleted clauses become empty), lits_out receives the packed pool, *nlits_new its
 * length; pass lits_out == NULL to compact the device copy only. */
int cp3g_compact(cp3g* g, uint32_t* new_start, int32_t* lits_out, size_t lits_cap, size_t* nlits_new);

/* multi-GPU plumbing: NCCL communicator over the ranks' devices (bootstrap id from rank 0) */
int cp3g_nccl_unique_id(void* out128);
int cp3g_nccl_init(cp3g* g, int rank, int world, const void* id128);

/* counters: "launches", "kernel_ns" (all kernels, CUDA events), "scan_ns" (K3/K4 only), "commit_ns" (device pass commit), "bve_ns" (K5'/K6'), "h2d_bytes", "d2h_bytes",
 * "scan_requests", "scan_checks", "scan_alg_bytes" (sum over checks of 4 + 8 + 4*#D) */
int64_t cp3g_counter(const cp3g* g, const char* name);

#ifdef __cplusplus
}
#endif
#endif
