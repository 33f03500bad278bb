// capi.cpp - the libcoprocessorc.h drop-in layer (CP*) on top of the ordered-commit engine.
// Mirrors coprocessor/libcoprocessorc.cc: an opaque handle owns everything; no error codes, UNSAT is
// CPok()==0; unknown options are ignored unless `strict`.
#include <climits>
#include <malloc.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include "engine.h"

namespace {
struct Handle {                         // struct libcp3, libcoprocessorc.cc:12-24
    cp3::Engine eng;
    // CPaddLit calls into a still empty solver are buffered and handed over as ONE stream (device bulk ingest K0)
    // by the first call that looks at the solver; nothing observable depends on when the clauses arrive
    std::vector<int32_t> pend; bool buffering = true;
    std::vector<int32_t> out; size_t outPos = 0; bool outReady = false;
    std::vector<uint8_t> model; size_t modelLit = 0;
    // The reference builds Solver + Preprocessor inside CPinit (libcoprocessorc.cc:35-46); a few options are read by those
    // constructors only: cp3_limited (CoprocessorData::hasLimit, Coprocessor.cc:29 / CoprocessorTypes.h:614), the Stepper
    // limits cp3_sub_limit / cp3_str_limit / cp3_call_inc (Subsumption.cc:22-24) and cp3_bve_limit (BVE.cc:23).  Given to
    // CPinit (preset name or option string) they count, given to CPparseOption* / CPsetPresetConfig afterwards they are
    // accepted and have no effect.  Same here: the values of CPinit time are put back after every later parse.
    struct Ctor { bool limited; int64_t sub_limit, str_limit, bve_limit; int call_inc; } ctor;
    void capture() { const cp3::Options& o = eng.opt; ctor = {o.limited, o.sub_limit, o.str_limit, o.bve_limit, o.call_inc}; }
    void restore() { cp3::Options& o = eng.opt; o.limited = ctor.limited; o.sub_limit = ctor.sub_limit; o.str_limit = ctor.str_limit; o.bve_limit = ctor.bve_limit; o.call_inc = ctor.call_inc; }
};
inline Handle* H(void* p) { return (Handle*)p; }
inline void sync(Handle* h)
{
    h->buffering = false;
    if (h->pend.empty()) { return; }
    std::vector<int32_t> s; s.swap(h->pend);
    h->eng.addStream(s.data(), s.size());
}
inline Handle* HS(void* p) { Handle* h = (Handle*)p; sync(h); return h; }      // handle with the buffered literals delivered
void applyOption(Handle* h, const char* o, bool strict)
{
    int r = h->eng.opt.parse(o);
    h->restore();
    if (r < 0) {
        fprintf(stderr, "cp3b200: option %s selects a reference code path the GPU drop-in does not provide\n", o);
        if (strict) { exit(1); }
    } else if (r == 0 && strict) {
        fprintf(stderr, "ERROR! Unknown flag \"%s\"\n", o);       // riss/utils/Config.h:816-822
        exit(1);
    }
}
void ensureOutput(Handle* h)
{
    if (h->outReady) { return; }
    size_t n = h->eng.output(nullptr, 0);
    h->out.resize(n);
    h->eng.output(h->out.data(), n);
    h->outPos = 0; h->outReady = true;
}
}

extern "C" {

// The engine's work arrays are hundreds of MB that live for one CPpreprocess.  With glibc's defaults every one of them is
// mmap()ed, page-faulted in and unmapped again - 90 ms of a 0.6 s run at 4.6 M clauses.  A host that calls CPpreprocess
// repeatedly can opt in to keeping that working set in the heap (large blocks from the heap instead of mmap, no trimming)
// with CP3_KEEP_HEAP=1.  It is NOT the default: a drop-in library must not change its host's allocator policy.
static void keepHostHeapOnce()
{
    static bool done = false;
    if (done) { return; }
    done = true;
    const char* e = getenv("CP3_KEEP_HEAP");
    if (!e || atoi(e) == 0) { return; }
    mallopt(M_MMAP_MAX, 0);
    mallopt(M_TRIM_THRESHOLD, INT_MAX);
}

void* CPinit(const char* presetConfig)
{
    keepHostHeapOnce();
    Handle* h = new Handle();
    if (presetConfig) { h->eng.opt.preset(presetConfig); }
    h->capture();
    return h;
}
void CPdestroy(void** p) { if (p && *p) { delete H(*p); *p = nullptr; } }
void CPpreprocess(void* p) { HS(p)->eng.preprocess(); H(p)->outReady = false; }

void CPparseOptions(void* p, int* argc, char** argv, int strict)
{
    // like Config::parseOptions (riss/utils/Config.h:779-826): argv[0] is the program name; recognised
    // options are consumed, the rest stays in argv
    int j = 1;
    for (int i = 1; i < *argc; ++i) {
        int r = H(p)->eng.opt.parse(argv[i]);
        H(p)->restore();
        if (r == 0) { if (strict && argv[i][0] == '-') { fprintf(stderr, "ERROR! Unknown flag \"%s\"\n", argv[i]); exit(1); } argv[j++] = argv[i]; }
        else if (r < 0) { applyOption(H(p), argv[i], strict != 0); }
    }
    *argc = j;
}
void CPparseOptionString(void* p, char* s)
{
    std::istringstream is(s ? s : "");
    std::string tok;
    while (is >> tok) { applyOption(H(p), tok.c_str(), false); }
}
void CPsetPresetConfig(void* p, const char* preset) { H(p)->eng.opt.preset(preset); H(p)->restore(); }
void CPaddLit(void* p, int lit)
{
    Handle* h = H(p); h->outReady = false;
    if (h->buffering) { h->pend.push_back(lit); } else { h->eng.addLit(lit); }
}
void CPaddLits(void* p, const int32_t* s, size_t n) { HS(p)->eng.addStream(s, n); H(p)->outReady = false; }
float CPversion(void*) { return 7.1f; }

int CPhasNextOutputLit(void* p) { ensureOutput(HS(p)); return H(p)->outPos < H(p)->out.size() ? 1 : 0; }
int CPnextOutputLit(void* p)
{
    Handle* h = HS(p); ensureOutput(h);
    if (h->outPos >= h->out.size()) { return 0; }
    return h->out[h->outPos++];
}
void CPresetOutput(void* p) { HS(p)->outReady = false; ensureOutput(H(p)); }
size_t CPgetOutput(void* p, int32_t* out, size_t cap) { return HS(p)->eng.output(out, cap); }
size_t CPoutputUnits(void* p) { return HS(p)->eng.nTrail(); }
size_t CPgetUndo(void* p, int32_t* out, size_t cap) { return HS(p)->eng.undoStack(out, cap); }
int64_t CPstat(void* p, const char* name) { return HS(p)->eng.stat(name); }
int CPsetDistributed(void* p, int rank, int world, const void* id) { return H(p)->eng.setDistributed(rank, world, id); }

void CPfreezeVariable(void* p, int variable) { HS(p)->eng.freeze(variable < 0 ? -variable : variable); }
int CPgetReplaceLiteral(void* p, int oldLit) { const int r = HS(p)->eng.importLit(oldLit); return r == 0 ? oldLit : r; }   // libcoprocessorc.cc:198-203

void CPresetModel(void* p) { H(p)->model.clear(); H(p)->modelLit = 0; }
void CPpushModelBool(void* p, int pol) { H(p)->model.push_back(pol == 0 ? cp3::L_FALSE : cp3::L_TRUE); }
// libcoprocessorc.cc:139-147.  DELIBERATE FIX: the reference maps a negative literal -k to variable index k+1 (mkLit(-literal + 1),
// a sign slip - model round trips through negative literals hit the wrong variable); here -k means "variable k is false", which is
// what the header documents.  Positive literals behave identically.  0 is ignored (the reference writes model[-1]).
void CPgiveModelLit(void* p, int literal)
{
    Handle* h = H(p);
    if (literal == 0) { return; }
    size_t v = (size_t)(literal < 0 ? -literal : literal) - 1;
    while (h->model.size() <= v) { h->model.push_back(cp3::L_UNDEF); }
    h->model[v] = literal > 0 ? cp3::L_TRUE : cp3::L_FALSE;
}
void CPpostprocessModel(void* p) { HS(p)->eng.extendModel(H(p)->model); }
int CPgetFinalModelLit(void* p)
{
    Handle* h = H(p);
    if (h->modelLit < h->model.size()) {
        int v = (int)h->modelLit++;
        return h->model[v] == cp3::L_TRUE ? v + 1 : -v - 1;
    }
    return 0;
}
int CPmodelVariables(void* p) { return (int)H(p)->model.size(); }
int CPnVars(void* p) { return HS(p)->eng.nVars(); }
int CPnClss(void* p) { return (int)(HS(p)->eng.nLiveClauses() + H(p)->eng.nTrail()); }
int CPnLits(void* p) { return (int)(HS(p)->eng.nLiveLits() + H(p)->eng.nTrail()); }
int CPok(void* p) { return HS(p)->eng.okay() ? 1 : 0; }
int CPlitFalsified(void* p, int lit) { return HS(p)->eng.litValue(lit) == cp3::L_FALSE ? 1 : 0; }

} // extern "C"
