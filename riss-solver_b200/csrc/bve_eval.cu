// bve_eval.cu - per-variable BVE evaluation on the device, one warp per candidate variable.
//
// For a candidate variable v the reference's bve_worker (coprocessor/techniques/BVE.cc:388-637) does, on the two occurrence
// lists of v: gate detection with list reordering (findGates, BVE.cc:1143-1267), lazy removal of the deleted entries by
// swap-with-last (BVE.cc:474-489), and the resolvent anticipation over the (pos, neg) pairs the gate limits leave
// (anticipateElimination, BVE.cc:695-880; tryResolve, BVE.h:248-299).  All of it is a pure function of the two lists (in
// their current ORDER), the clauses they name and the deleted flags, so it is evaluated here for every variable that waits in
// the heap, in one launch; the ordered commit on the host consumes a 32-byte record per variable and copies the rewritten
// lists, instead of walking clause headers and literal pools through ~25 dependent cache misses per variable.  A variable
// whose evaluation meets a unit or empty resolvent (the commit must enqueue / fail at that very pair) or whose lists exceed
// the kernel's capacity is flagged and takes the pair-matrix path (K5 k_bve_pairs) on the host.
//
// k_bve_eval      lane 0 replays gate detection and cleaning on a shared-memory copy of the lists; all lanes then count the
//                 resolvents (sizes by sorted merge), per-clause statistics (blocked-clause detection, BVE.cc:1053-1090) and steps
// k_bve_stream    resolvents of the variables the decision rule will most likely eliminate, written in resolveSet's order
//                 (BVE.cc:892-1046) as one literal stream per variable + a size per resolvent
// Latency-bound gathers over short lists; no tensor cores (nothing here is a contraction).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/cp3b200.h"
#include "gpu_db.h"

static constexpr int EV_LCAP = 64;          // list capacity per polarity
static constexpr int EV_WARPS = 4;

struct EvWarp {
    uint32_t pos[EV_LCAP], neg[EV_LCAP];     // clause references, current order
    uint32_t psz[EV_LCAP], nsz[EV_LCAP];     // size | FLAG_DEL, permuted together with the references
    uint32_t pst[EV_LCAP], nst[EV_LCAP];     // per-clause resolvent counts (pos_stats / neg_stats)
    int32_t marks[EV_LCAP]; int nmarks;      // data.ma of findGates: the marked literals
};

// same merge as k_bve_pairs (gpu_db.cu): number of literals of the resolvent on pv, -1 for a tautology
__device__ __forceinline__ int ev_resolvent_size(const int32_t* __restrict__ P, uint32_t pn, const int32_t* __restrict__ N, uint32_t nn, int32_t pv)
{
    const int32_t nv = pv ^ 1;
    uint32_t i = 0, j = 0; int r = 0; int32_t prev = -2;
    while (i < pn || j < nn) {
        int32_t c;
        if (i < pn && P[i] == pv) { ++i; continue; }
        if (j < nn && N[j] == nv) { ++j; continue; }
        if (j >= nn || (i < pn && P[i] < N[j])) { c = P[i++]; } else { c = N[j++]; }
        if (c == prev) { continue; }
        if (c == (prev ^ 1) && prev != -2) { return -1; }
        prev = c; ++r;
    }
    return r;
}

__device__ __forceinline__ bool ev_marked(const EvWarp& s, int32_t x) { for (int i = 0; i < s.nmarks; ++i) { if (s.marks[i] == x) { return true; } } return false; }
__device__ __forceinline__ void ev_mark(EvWarp& s, int32_t x) { if (!ev_marked(s, x) && s.nmarks < EV_LCAP) { s.marks[s.nmarks++] = x; } }
__device__ __forceinline__ void ev_unmark(EvWarp& s, int32_t x) { for (int i = 0; i < s.nmarks; ++i) { if (s.marks[i] == x) { s.marks[i] = s.marks[--s.nmarks]; return; } } }

// findGates (BVE.cc:1143-1267) on the shared copies; lane 0 only.  Returns true and sets the limits when an AND gate is found.
__device__ bool ev_find_gates(EvWarp& s, const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, int32_t v, uint32_t np, uint32_t nn, int& p_limit, int& n_limit)
{
    if (np < 3 && nn < 3) { return false; }
    if (np < 2 || nn < 2) { return false; }
    for (int pn = 0; pn < 2; ++pn) {
        const int32_t pLit = 2 * v + (pn != 0), nLit = 2 * v + (pn == 0);
        uint32_t* pList = pn == 0 ? s.pos : s.neg; uint32_t* pSz = pn == 0 ? s.psz : s.nsz; const uint32_t pN = pn == 0 ? np : nn;
        uint32_t* nList = pn == 0 ? s.neg : s.pos; uint32_t* nSz = pn == 0 ? s.nsz : s.psz; const uint32_t nN = pn == 0 ? nn : np;
        int& pClauses = pn == 0 ? p_limit : n_limit;
        int& nClauses = pn == 0 ? n_limit : p_limit;
        s.nmarks = 0;
        for (uint32_t i = 0; i < pN; ++i) {
            if ((pSz[i] & FLAG_DEL) || (pSz[i] & SIZE_MASK) != 2) { continue; }
            const int32_t* l = lits + hdr[pList[i]].start;
            const int32_t other = l[0] == pLit ? l[1] : l[0];
            ev_mark(s, other ^ 1);
        }
        for (uint32_t i = 0; i < nN; ++i) {
            if (nSz[i] & FLAG_DEL) { continue; }
            const uint32_t sz = nSz[i] & SIZE_MASK;
            const int32_t* l = lits + hdr[nList[i]].start;
            uint32_t j = 0;
            for (; j < sz; ++j) {
                if (l[j] == nLit) { continue; }
                if (!ev_marked(s, l[j])) { break; }
            }
            if (j == sz) {
                pClauses = (int)sz - 1;
                nClauses = 1;
                for (uint32_t k = 0; k < sz; ++k) { ev_unmark(s, l[k]); }
                { const uint32_t t = nList[0]; nList[0] = nList[i]; nList[i] = t; const uint32_t u = nSz[0]; nSz[0] = nSz[i]; nSz[i] = u; }
                uint32_t placed = 0;
                for (uint32_t k = 0; k < pN; ++k) {
                    if ((pSz[k] & FLAG_DEL) || (pSz[k] & SIZE_MASK) != 2) { continue; }
                    const int32_t* l2 = lits + hdr[pList[k]].start;
                    const int32_t other = l2[0] == pLit ? l2[1] : l2[0];
                    if (!ev_marked(s, other ^ 1)) {
                        { const uint32_t t = pList[placed]; pList[placed] = pList[k]; pList[k] = t; const uint32_t u = pSz[placed]; pSz[placed] = pSz[k]; pSz[k] = u; }
                        placed++;
                        ev_mark(s, other ^ 1);
                    }
                }
                return true;
            }
        }
    }
    return false;
}

__global__ void __launch_bounds__(EV_WARPS * 32) k_bve_eval(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t nv,
        const uint32_t* __restrict__ vars, const uint32_t* __restrict__ p_off, uint32_t* __restrict__ p_refs,
        const uint32_t* __restrict__ n_off, uint32_t* __restrict__ n_refs, int use_gates,
        cp3g_bve_rec* __restrict__ recs, uint16_t* __restrict__ p_stats, uint16_t* __restrict__ n_stats)
{
    __shared__ EvWarp sh[EV_WARPS];
    const uint32_t wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * EV_WARPS + wi;
    if (i >= nv) { return; }
    EvWarp& s = sh[wi];
    const uint32_t pb = p_off[i], np0 = p_off[i + 1] - pb, nb = n_off[i], nn0 = n_off[i + 1] - nb;
    const int32_t v = (int32_t)vars[i];
    cp3g_bve_rec r; r.p_limit = (int32_t)np0; r.n_limit = (int32_t)nn0; r.pos_n = np0; r.neg_n = nn0; r.resolvents = 0; r.lit_clauses = 0; r.steps = 0; r.flags = 0;
    if (np0 > (uint32_t)EV_LCAP || nn0 > (uint32_t)EV_LCAP) { if (lane == 0) { r.flags = CP3G_BVE_HOST; recs[i] = r; } return; }
    for (uint32_t k = lane; k < np0; k += 32) { const uint32_t c = p_refs[pb + k]; s.pos[k] = c; s.psz[k] = hdr[c].szfl & (SIZE_MASK | FLAG_DEL); s.pst[k] = 0; }
    for (uint32_t k = lane; k < nn0; k += 32) { const uint32_t c = n_refs[nb + k]; s.neg[k] = c; s.nsz[k] = hdr[c].szfl & (SIZE_MASK | FLAG_DEL); s.nst[k] = 0; }
    __syncwarp();
    int p_limit = (int)np0, n_limit = (int)nn0; uint32_t np = np0, nn = nn0; int found = 0;
    if (lane == 0) {
        if (use_gates) { found = ev_find_gates(s, hdr, lits, v, np0, nn0, p_limit, n_limit) ? 1 : 0; }
        // lazy removal of the deleted entries, BVE.cc:474-489
        for (uint32_t k = 0; k < np; ++k) { if (s.psz[k] & FLAG_DEL) { s.pos[k] = s.pos[np - 1]; s.psz[k] = s.psz[np - 1]; --np; --k; } }
        for (uint32_t k = 0; k < nn; ++k) { if (s.nsz[k] & FLAG_DEL) { s.neg[k] = s.neg[nn - 1]; s.nsz[k] = s.nsz[nn - 1]; --nn; --k; } }
    }
    found = __shfl_sync(0xffffffffu, found, 0); p_limit = __shfl_sync(0xffffffffu, p_limit, 0); n_limit = __shfl_sync(0xffffffffu, n_limit, 0);
    np = __shfl_sync(0xffffffffu, np, 0); nn = __shfl_sync(0xffffffffu, nn, 0);
    __syncwarp();
    // anticipateElimination, BVE.cc:749-790: the pairs the limits leave
    const bool hasDefinition = (p_limit < (int)np || n_limit < (int)nn);
    uint32_t res = 0, lc = 0, steps = 0, special = 0;
    const int32_t pv = 2 * v;
    const uint32_t total = np * nn;
    for (uint32_t t = lane; t < total; t += 32) {
        const uint32_t a = t / nn, b = t - a * nn;
        if ((int)a >= p_limit && (int)b >= n_limit) { continue; }
        if (hasDefinition && (int)a < p_limit && (int)b < n_limit) { continue; }
        ++steps;
        const int sz = ev_resolvent_size(lits + hdr[s.pos[a]].start, s.psz[a] & SIZE_MASK, lits + hdr[s.neg[b]].start, s.nsz[b] & SIZE_MASK, pv);
        if (sz > 1) { ++res; lc += (uint32_t)sz; atomicAdd(&s.pst[a], 1u); atomicAdd(&s.nst[b], 1u); }
        else if (sz == 0 || sz == 1) { special = 1; }
    }
    for (int d = 16; d > 0; d >>= 1) {
        res += __shfl_xor_sync(0xffffffffu, res, d); lc += __shfl_xor_sync(0xffffffffu, lc, d);
        steps += __shfl_xor_sync(0xffffffffu, steps, d); special |= __shfl_xor_sync(0xffffffffu, special, d);
    }
    __syncwarp();
    for (uint32_t k = lane; k < np; k += 32) { p_refs[pb + k] = s.pos[k]; p_stats[pb + k] = (uint16_t)min(s.pst[k], 0xffffu); }
    for (uint32_t k = lane; k < nn; k += 32) { n_refs[nb + k] = s.neg[k]; n_stats[nb + k] = (uint16_t)min(s.nst[k], 0xffffu); }
    if (lane == 0) {
        r.p_limit = p_limit; r.n_limit = n_limit; r.pos_n = np; r.neg_n = nn; r.resolvents = res; r.lit_clauses = lc; r.steps = steps;
        r.flags = (found ? CP3G_BVE_GATE : 0u) | (special ? CP3G_BVE_HOST : 0u);
        recs[i] = r;
    }
}

// resolvents of the selected variables in resolveSet's order (BVE.cc:892-1046): pos-major over the cleaned, reordered lists,
// the pairs the limits leave, tautologies dropped
__global__ void __launch_bounds__(EV_WARPS * 32) k_bve_stream(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t nsel,
        const uint32_t* __restrict__ sel, const uint32_t* __restrict__ vars, const uint32_t* __restrict__ p_off, const uint32_t* __restrict__ p_refs,
        const uint32_t* __restrict__ n_off, const uint32_t* __restrict__ n_refs, const cp3g_bve_rec* __restrict__ recs,
        const unsigned long long* __restrict__ res_off, const unsigned long long* __restrict__ lit_off,
        uint16_t* __restrict__ out_sizes, int32_t* __restrict__ out_lits)
{
    const uint32_t wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t k = blockIdx.x * EV_WARPS + wi;
    if (k >= nsel) { return; }
    const uint32_t i = sel[k];
    const cp3g_bve_rec r = recs[i];
    const uint32_t pb = p_off[i], nb = n_off[i], np = r.pos_n, nn = r.neg_n;
    const bool hasDefinition = (r.p_limit < (int)np || r.n_limit < (int)nn);
    const int32_t pv = 2 * (int32_t)vars[i], nvl = pv ^ 1;
    unsigned long long rbase = res_off[k], lbase = lit_off[k];
    const uint32_t total = np * nn;
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
        const uint32_t t = t0 + lane;
        int sz = 0; uint32_t a = 0, b = 0; ClauseHdr hp, hn;
        if (t < total) {
            a = t / nn; b = t - a * nn;
            const bool skip = ((int)a >= r.p_limit && (int)b >= r.n_limit) || (hasDefinition && (int)a < r.p_limit && (int)b < r.n_limit);
            if (!skip) {
                hp = hdr[p_refs[pb + a]]; hn = hdr[n_refs[nb + b]];
                sz = ev_resolvent_size(lits + hp.start, hp.szfl & SIZE_MASK, lits + hn.start, hn.szfl & SIZE_MASK, pv);
                if (sz < 0) { sz = 0; }
            }
        }
        // exclusive prefix of (has resolvent, literal count) over the 32 pairs of this round
        uint32_t has = sz > 0 ? 1u : 0u, pre_r = has, pre_l = (uint32_t)sz;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, pre_r, d), y = __shfl_up_sync(0xffffffffu, pre_l, d);
            if ((int)lane >= d) { pre_r += x; pre_l += y; }
        }
        const uint32_t tot_r = __shfl_sync(0xffffffffu, pre_r, 31), tot_l = __shfl_sync(0xffffffffu, pre_l, 31);
        if (has) {
            out_sizes[rbase + pre_r - 1] = (uint16_t)sz;
            int32_t* o = out_lits + lbase + (pre_l - (uint32_t)sz);
            const int32_t* P = lits + hp.start; const int32_t* N = lits + hn.start;
            const uint32_t pn = hp.szfl & SIZE_MASK, nq = hn.szfl & SIZE_MASK;
            uint32_t i2 = 0, j2 = 0, w = 0; int32_t prev = -2;
            while (i2 < pn || j2 < nq) {
                int32_t c;
                if (i2 < pn && P[i2] == pv) { ++i2; continue; }
                if (j2 < nq && N[j2] == nvl) { ++j2; continue; }
                if (j2 >= nq || (i2 < pn && P[i2] < N[j2])) { c = P[i2++]; } else { c = N[j2++]; }
                if (c == prev) { continue; }
                prev = c;
                if (w < (uint32_t)sz) { o[w] = c; }
                ++w;
            }
        }
        rbase += tot_r; lbase += tot_l;
    }
}

extern "C" {

int cp3g_bve_eval(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, uint32_t* p_refs, const uint32_t* n_off, uint32_t* n_refs,
                  int use_gates, cp3g_bve_rec* recs, uint16_t* p_stats, uint16_t* n_stats)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    g->ev_nv = 0;
    if (!nv) { return CP3G_OK; }
    cudaSetDevice(g->device);
    const size_t np = p_off[nv], nn = n_off[nv];
    RESERVE(g->bu0, nv, false); RESERVE(g->bu1, (size_t)nv + 1, false); RESERVE(g->bu2, np + 1, false);
    RESERVE(g->bu3, (size_t)nv + 1, false); RESERVE(g->bu4, nn + 1, false);
    RESERVE(g->ev_recs, (size_t)nv * sizeof(cp3g_bve_rec), false); RESERVE(g->ev_pst, np + 1, false); RESERVE(g->ev_nst, nn + 1, false);
    CK(cudaMemcpyAsync(g->bu0.p, vars, (size_t)nv * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu1.p, p_off, ((size_t)nv + 1) * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu2.p, p_refs, np * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu3.p, n_off, ((size_t)nv + 1) * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu4.p, n_refs, nn * 4, cudaMemcpyHostToDevice, g->stream));
    KTimer t(g);
    k_bve_eval<<<grid_for(nv, EV_WARPS), EV_WARPS * 32, 0, g->stream>>>(g->hdr.p, g->lits.p, nv, g->bu0.p, g->bu1.p, g->bu2.p, g->bu3.p, g->bu4.p, use_gates,
                                                                      reinterpret_cast<cp3g_bve_rec*>(g->ev_recs.p), g->ev_pst.p, g->ev_nst.p);
    t.stop();
    CK(cudaMemcpyAsync(recs, g->ev_recs.p, (size_t)nv * sizeof(cp3g_bve_rec), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaMemcpyAsync(p_refs, g->bu2.p, np * 4, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaMemcpyAsync(n_refs, g->bu4.p, nn * 4, cudaMemcpyDeviceToHost, g->stream));
    if (p_stats) { CK(cudaMemcpyAsync(p_stats, g->ev_pst.p, np * 2, cudaMemcpyDeviceToHost, g->stream)); }
    if (n_stats) { CK(cudaMemcpyAsync(n_stats, g->ev_nst.p, nn * 2, cudaMemcpyDeviceToHost, g->stream)); }
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    g->bve_ns += t.collect();
    g->bve_bytes += (int64_t)(np + nn) * (4 + 16 + 16) + (int64_t)nv * 32;      // per listed clause: reference, header, ~4 literals; one record per variable
    g->launches++; g->h2d += (int64_t)(nv * 12 + (np + nn) * 4); g->d2h += (int64_t)((size_t)nv * sizeof(cp3g_bve_rec) + (np + nn) * (4 + (p_stats ? 2 : 0)));
    g->ev_nv = nv;
    return CP3G_OK;
}

int cp3g_bve_stream(cp3g* g, uint32_t nsel, const uint32_t* sel, const uint64_t* res_off, const uint64_t* lit_off, uint16_t* out_sizes, int32_t* out_lits)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!nsel) { return CP3G_OK; }
    if (!g->ev_nv) { g->err = "cp3g_bve_stream without a preceding cp3g_bve_eval"; return CP3G_ERR_ARG; }
    cudaSetDevice(g->device);
    const size_t nres = res_off[nsel], nlit = lit_off[nsel];
    RESERVE(g->scratch_u32, nsel, false); RESERVE(g->bq0, (size_t)nsel + 1, false); RESERVE(g->ev_loff, (size_t)nsel + 1, false);
    RESERVE(g->ev_rsz, nres + 1, false); RESERVE(g->bl0, nlit + 1, false);
    CK(cudaMemcpyAsync(g->scratch_u32.p, sel, (size_t)nsel * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bq0.p, res_off, ((size_t)nsel + 1) * 8, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->ev_loff.p, lit_off, ((size_t)nsel + 1) * 8, cudaMemcpyHostToDevice, g->stream));
    KTimer t(g);
    k_bve_stream<<<grid_for(nsel, EV_WARPS), EV_WARPS * 32, 0, g->stream>>>(g->hdr.p, g->lits.p, nsel, g->scratch_u32.p, g->bu0.p, g->bu1.p, g->bu2.p, g->bu3.p, g->bu4.p,
                                                                        reinterpret_cast<const cp3g_bve_rec*>(g->ev_recs.p), g->bq0.p, g->ev_loff.p, g->ev_rsz.p, g->bl0.p);
    t.stop();
    if (nres) { CK(cudaMemcpyAsync(out_sizes, g->ev_rsz.p, nres * 2, cudaMemcpyDeviceToHost, g->stream)); }
    if (nlit) { CK(cudaMemcpyAsync(out_lits, g->bl0.p, nlit * 4, cudaMemcpyDeviceToHost, g->stream)); }
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    g->bve_ns += t.collect();
    g->bve_bytes += (int64_t)(nres * 2 + nlit * 4);                               // the resolvents written
    g->launches++; g->h2d += (int64_t)nsel * 20; g->d2h += (int64_t)(nres * 2 + nlit * 4);
    return CP3G_OK;
}

} // extern "C"
