// engine_options.cpp - CP3Config subset and presets of the path (see engine.h).
#include "engine_util.h"

namespace cp3 {

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

static long long envNum(const char* name, long long dflt) { const char* e = getenv(name); return e ? atoll(e) : dflt; }
static bool envSet(const char* name) { return getenv(name) != nullptr; }
Tuning::Tuning()
    : threads((int)envNum("CP3_THREADS", (long long)std::min(24u, std::max(1u, std::thread::hardware_concurrency())))),
      ingest_min((size_t)envNum("CP3_INGEST_MIN", 262144)), bulk_min((uint32_t)envNum("CP3_BULK_MIN", 4096)),
      parsub_min((size_t)envNum("CP3_PARSUB_MIN", 65536)), static_min((size_t)envNum("CP3_STATIC_MIN", 4096)),
      touched_min((uint64_t)envNum("CP3_TOUCHED_MIN", 262144)), devcommit_maxlist((uint32_t)envNum("CP3_DEVCOMMIT_MAXLIST", 4096)),
      par_min((size_t)envNum("CP3_PAR_MIN", 16384)), bce_maxwords((uint64_t)envNum("CP3_BCE_MAXWORDS", 1ll << 22)),
      no_devcommit(envSet("CP3_NO_DEVCOMMIT")), no_overlap(envSet("CP3_NO_OVERLAP")), no_precompact(envSet("CP3_NO_PRECOMPACT")),
      no_lazy(envSet("CP3_NO_LAZY")), no_plan(envSet("CP3_NO_PLAN")), no_components(envSet("CP3_NO_COMPONENTS")), no_devstatic(envSet("CP3_NO_DEVSTATIC")), no_bve_eval(envSet("CP3_NO_BVE_EVAL")), no_bve_prefetch(envSet("CP3_NO_BVE_PREFETCH")),
      debug_par(envSet("CP3_DEBUG_PAR")), debug_bve(envSet("CP3_DEBUG_BVE")), debug_fine(envSet("CP3_DEBUG_FINE")),
      debug_ondemand(envSet("CP3_DEBUG_ONDEMAND")), debug_pre(envSet("CP3_DEBUG_PRE"))
{
}

} // namespace cp3
