/* cp3_oracle.h - CPU restatement of Coprocessor's sequential subsumption / strengthening / BVE path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (riss-solver_b200/) never does.
 *
 * Parity status: PINNED - every function here is checked against the compiled, unmodified reference
 * (oracle/_ref, built by oracle/Makefile.ref from /root/reference) by tests/test_oracle_vs_ref.py and
 * against the committed golden vectors in tests/golden/ (generated from that reference by
 * tests/golden/make_golden.py).
 */
#ifndef CP3_ORACLE_H
#define CP3_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cp3o cp3o;

cp3o*   cp3o_new(void);
void    cp3o_free(cp3o* o);
/* one option in Coprocessor syntax: "-subsimp", "-no-cp3_strength", "-bve_cgrow=3"; returns 1 if known */
int     cp3o_option(cp3o* o, const char* opt);
/* DIMACS literal stream, 0 terminates a clause (CPaddLit, libcoprocessorc.cc:251-266) */
void    cp3o_add(cp3o* o, const int32_t* stream, size_t n);
void    cp3o_freeze(cp3o* o, int dimacs_var);
/* Preprocessor::preprocess (Coprocessor.cc:996) restricted to [up] subsimp [bve] */
void    cp3o_preprocess(cp3o* o);
int     cp3o_ok(const cp3o* o);
int     cp3o_nvars(const cp3o* o);
/* output stream like CPnextOutputLit (libcoprocessorc.cc:80-108): trail units, then clauses in
 * solver->clauses order, each 0-terminated.  Returns the number of int32 needed; writes up to cap. */
size_t  cp3o_output(const cp3o* o, int32_t* out, size_t cap);
size_t  cp3o_ntrail(const cp3o* o);
/* extension (undo) stack, DIMACS literals, 0 = entry separator (CoprocessorTypes.h:1615-1655) */
size_t  cp3o_undo(const cp3o* o, int32_t* out, size_t cap);
int64_t cp3o_stat(const cp3o* o, const char* name);
/* -dense (Dense.cc): number of variables a model must cover after decompression, the decompression itself
 * (model[v] = 0 true / 1 false, `have` entries filled in compressed numbering) and the variable map (-1 = removed) */
int     cp3o_model_vars(const cp3o* o);
void    cp3o_decompress_model(const cp3o* o, unsigned char* model, int have);
int     cp3o_compressed_var(const cp3o* o, int orig_var);
/* Preprocessor::extendModel (Coprocessor.cc:1214-1240): decompress, then replay the undo stack */
void    cp3o_extend_model(const cp3o* o, unsigned char* model, int have);

/* --- kernel-level oracles on a pristine snapshot (after initializePreprocessor + sortClauses) --- */
/* prepare the occurrence-list data base without running any technique */
void    cp3o_init_only(cp3o* o);
size_t  cp3o_nclauses(const cp3o* o);
/* Full-queue backward subsumption (Subsumption.cc:220-314 on the initial queue): deleted[i]=1 iff clause
 * i (index in solver->clauses) is removed.  Returns subs-steps. */
int64_t cp3o_subsume_snapshot(cp3o* o, uint8_t* deleted);
/* anticipateElimination (BVE.cc:695) for every variable on the snapshot: out[4*v..] = pos,neg,resolvents,
 * lit_clauses (no gate detection, no side effects; unit/empty resolvents counted in out_small[v]) */
void    cp3o_anticipate_snapshot(cp3o* o, int32_t* out, int32_t* out_small);

#ifdef __cplusplus
}
#endif
#endif
