// engine_bce.cpp - stand-alone blocked clause elimination of the host engine (coprocessor/techniques/BCE.cc).
#include "engine_util.h"

namespace cp3 {

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
    const uint64_t row_cap = tune.bce_maxwords;                      // words per variable
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

} // namespace cp3
