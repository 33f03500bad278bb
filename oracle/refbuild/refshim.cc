// refshim — instrumented driver around the UNMODIFIED reference objects (test infrastructure only).
//
// Ours, not reference source.  It uses Solver / CP3Config / Preprocessor the way
// coprocessor/libcoprocessorc.cc:34-64 does (CPinit / CPaddLit / CPpreprocess) and then dumps the
// internal state the parity tests need.  It is compiled with -fno-access-control (this TU only; the
// reference objects are linked unmodified) so it can reach Preprocessor::{data,subsumption,bve}.
//
// usage: refshim MODE in.cnf out.txt [coprocessor options ...]
//   MODE run        : CPpreprocess-equivalent; dump ok / trail / clauses (solver->clauses order) / undo
//                     stack / Subsumption + BVE counters
//   MODE anticipate : initialise like Preprocessor::performSimplification (Coprocessor.cc:93-118), then
//                     call BoundedVariableElimination::anticipateElimination (BVE.cc:695) for every
//                     variable on that pristine snapshot and dump (v,pos,neg,resolvents,lit_clauses)
//   MODE time       : like run, but only prints wall seconds of preprocess() and the step counters
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

#include "coprocessor/Coprocessor.h"

using namespace Riss;
using namespace Coprocessor;

static bool readCnf(const char* path, Solver& S)
{
    FILE* f = fopen(path, "rb");
    if (!f) { return false; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> buf(n + 1);
    if (fread(buf.data(), 1, n, f) != (size_t)n) { fclose(f); return false; }
    buf[n] = 0; fclose(f);
    vec<Lit> lits;
    size_t plen = strlen(path);
    if (plen > 4 && strcmp(path + plen - 4, ".bin") == 0) {
        // raw little-endian int32 stream of DIMACS literals, 0 terminates a clause
        const int32_t* q = (const int32_t*)buf.data();
        const size_t cnt = (size_t)n / 4;
        for (size_t i = 0; i < cnt; ++i) {
            const int32_t x = q[i];
            if (x == 0) { S.addClause_(lits); lits.clear(); }
            else {
                const int32_t v = x < 0 ? -x : x;
                while (S.nVars() < v) { S.newVar(); }
                lits.push(mkLit((Var)(v - 1), x < 0));
            }
        }
        return true;
    }
    const char* p = buf.data();
    while (*p) {
        while (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t') { ++p; }
        if (!*p) { break; }
        if (*p == 'c' || *p == 'p') { while (*p && *p != '\n') { ++p; } continue; }
        bool neg = false;
        if (*p == '-') { neg = true; ++p; }
        long v = 0;
        while (*p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; }
        if (v == 0) {
            // same sequence as CPaddLit (libcoprocessorc.cc:251-266): create vars, then addClause
            S.addClause_(lits);
            lits.clear();
        } else {
            while (S.nVars() < v) { S.newVar(); }
            lits.push(mkLit((Var)(v - 1), neg));
        }
    }
    return true;
}

static int toDimacs(Lit l) { return sign(l) ? -(var(l) + 1) : (var(l) + 1); }

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: refshim MODE in.cnf out.txt [options]\n"); return 2; }
    std::string mode = argv[1];
    const char* in = argv[2];
    const char* out = argv[3];

    CoreConfig* coreConfig = new CoreConfig("");
    CP3Config* cp3config = new CP3Config("");
    // options: parsed like coprocessor/Main.cc:93-98 (argv overrides)
    int oargc = argc - 3;
    char** oargv = argv + 3; // argv[3] acts as program name slot
    coreConfig->parseOptions(oargc, oargv, false);
    cp3config->parseOptions(oargc, oargv, false);

    Solver* S = new Solver(coreConfig);
    S->verbosity = 0;
    if (!readCnf(in, *S)) { fprintf(stderr, "cannot read %s\n", in); return 3; }
    Preprocessor* pp = new Preprocessor(S, *cp3config);

    FILE* fo = fopen(out, "w");
    if (!fo) { return 4; }

    if (mode == "anticipate") {
        // Coprocessor.cc:93-118
        S->simplify();
        pp->cleanSolver();
        pp->data.init(S->nVars());
        pp->data.resetPPhead();
        pp->initializePreprocessor();
        pp->sortClauses();
        pp->bve.initializeTechnique(pp->data);
        pp->data.ma.resize(pp->data.nVars() * 2);
        fprintf(fo, "vars %d\n", S->nVars());
        for (Var v = 0; v < S->nVars(); ++v) {
            std::vector<CRef>& pos = pp->data.list(mkLit(v, false));
            std::vector<CRef>& neg = pp->data.list(mkLit(v, true));
            int p_limit = pos.size(), n_limit = neg.size();
            int lit_clauses = 0, lit_learnts = 0, resolvents = 0;
            if (pos.size() == 0 || neg.size() == 0) {
                fprintf(fo, "%d %d %d %d %d\n", v + 1, (int)pos.size(), (int)neg.size(), 0, 0);
                continue;
            }
            pp->bve.pos_stats.clear(); pp->bve.neg_stats.clear();
            pp->bve.pos_stats.growTo(pos.size(), 0);
            pp->bve.neg_stats.growTo(neg.size(), 0);
            pp->bve.anticipateElimination(pp->data, pos, neg, v, p_limit, n_limit, pp->bve.pos_stats,
                                          pp->bve.neg_stats, lit_clauses, lit_learnts, resolvents, pp->bve.stepper);
            fprintf(fo, "%d %d %d %d %d", v + 1, (int)pos.size(), (int)neg.size(), resolvents, lit_clauses);
            fprintf(fo, " |");
            for (int i = 0; i < (int)pos.size(); ++i) { fprintf(fo, " %d", pp->bve.pos_stats[i]); }
            fprintf(fo, " |");
            for (int i = 0; i < (int)neg.size(); ++i) { fprintf(fo, " %d", pp->bve.neg_stats[i]); }
            fprintf(fo, "\n");
        }
        fclose(fo);
        return 0;
    }

    auto t0 = std::chrono::steady_clock::now();
    pp->preprocess();
    auto t1 = std::chrono::steady_clock::now();
    double secs = std::chrono::duration<double>(t1 - t0).count();

    fprintf(fo, "ok %d\n", S->okay() ? 1 : 0);
    fprintf(fo, "vars %d\n", S->nVars());
    fprintf(fo, "seconds %.6f\n", secs);
    fprintf(fo, "stat sub_subsumedClauses %d\n", pp->subsumption.subsumedClauses);
    fprintf(fo, "stat sub_subsumedLiterals %d\n", pp->subsumption.subsumedLiterals);
    fprintf(fo, "stat sub_removedLiterals %d\n", pp->subsumption.removedLiterals);
    fprintf(fo, "stat sub_subsSteps %lld\n", (long long)pp->subsumption.subsumptionStepper.getCurrentSteps());
    fprintf(fo, "stat sub_strengthSteps %lld\n", (long long)pp->subsumption.strengtheningStepper.getCurrentSteps());
    {   // pthread workers keep their own counters (Subsumption.h localStats, Subsumption.cc:50-63)
        long long ts = 0, tt = 0;
        for (size_t i = 0; i < pp->subsumption.localStats.size(); ++i) { ts += pp->subsumption.localStats[i].subsumeSteps; tt += pp->subsumption.localStats[i].strengthSteps; }
        fprintf(fo, "stat par_subsSteps %lld\n", ts);
        fprintf(fo, "stat par_strengthSteps %lld\n", tt);
    }
    fprintf(fo, "stat sub_seconds %.6f\n", pp->subsumption.processTime);
    fprintf(fo, "stat str_seconds %.6f\n", pp->subsumption.strengthTime);
    fprintf(fo, "stat bve_eliminatedVars %d\n", pp->bve.eliminatedVars);
    fprintf(fo, "stat bve_createdClauses %d\n", pp->bve.createdClauses);
    fprintf(fo, "stat bve_removedClauses %d\n", pp->bve.removedClauses);
    fprintf(fo, "stat bve_testedVars %d\n", pp->bve.testedVars);
    fprintf(fo, "stat bve_anticipations %d\n", pp->bve.anticipations);
    fprintf(fo, "stat bve_skippedVars %d\n", pp->bve.skippedVars);
    fprintf(fo, "stat bve_unitsEnqueued %d\n", pp->bve.unitsEnqueued);
    fprintf(fo, "stat bve_removedBC %d\n", pp->bve.removedBC);
    fprintf(fo, "stat bve_totallyAddedClauses %d\n", pp->bve.totallyAddedClauses);
    fprintf(fo, "stat bve_steps %lld\n", (long long)pp->bve.stepper.getCurrentSteps());
    fprintf(fo, "stat bve_seconds %.6f\n", pp->bve.processTime);
    fprintf(fo, "stat bce_steps %d\n", pp->bce.bceSteps);
    fprintf(fo, "stat bce_testedLits %d\n", pp->bce.testedLits);
    fprintf(fo, "stat bce_remBCE %d\n", pp->bce.remBCE);
    fprintf(fo, "stat up_removedClauses %d\n", pp->propagation.removedClauses);
    fprintf(fo, "stat up_removedLiterals %d\n", pp->propagation.removedLiterals);
    if (mode == "time") { fclose(fo); return 0; }

    // same stream as CPnextOutputLit (libcoprocessorc.cc:80-108): trail units, then solver->clauses
    fprintf(fo, "trail %d", S->trail.size());
    for (int i = 0; i < S->trail.size(); ++i) { fprintf(fo, " %d", toDimacs(S->trail[i])); }
    fprintf(fo, "\n");
    fprintf(fo, "clauses %d\n", S->clauses.size());
    for (int i = 0; i < S->clauses.size(); ++i) {
        const Clause& c = S->ca[S->clauses[i]];
        for (int j = 0; j < c.size(); ++j) { fprintf(fo, "%d ", toDimacs(c[j])); }
        fprintf(fo, "0\n");
    }
    // extension (undo) stack, CoprocessorTypes.h:1615-1655: lit_Undef separates entries -> printed as 0
    std::vector<Lit>& undo = pp->data.getUndo();
    fprintf(fo, "undo %d", (int)undo.size());
    for (size_t i = 0; i < undo.size(); ++i) { fprintf(fo, " %d", undo[i] == lit_Undef ? 0 : toDimacs(undo[i])); }
    fprintf(fo, "\n");
    // model extension (Preprocessor::extendModel, Coprocessor.cc:1214: Dense::decompress + undo replay) applied to a
    // fixed pseudo-random assignment of the simplified formula's variables that agrees with the trail
    if (S->okay()) {
        vec<lbool> model;
        for (int v = 0; v < S->nVars(); ++v) { model.push(((((unsigned)(v + 1)) * 2654435761u) >> 13) & 1u ? l_True : l_False); }
        for (int i = 0; i < S->trail.size(); ++i) { model[var(S->trail[i])] = sign(S->trail[i]) ? l_False : l_True; }
        pp->extendModel(model);
        fprintf(fo, "xmodel %d", model.size());
        for (int v = 0; v < model.size(); ++v) { fprintf(fo, " %d", model[v] == l_True ? 1 : 0); }
        fprintf(fo, "\n");
    }
    fclose(fo);
    return 0;
}
