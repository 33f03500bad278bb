// bce.cu - K9: the clause-against-clause test of stand-alone blocked clause elimination (coprocessor/techniques/BCE.cc).
//
// BlockedClauseElimination::blockedClauseElimination (BCE.cc:239-601) asks, for a clause c that holds literal l and every live
// clause d that holds ~l, whether the resolvent on l is a tautology (tautologicResolvent, BCE.cc:624-646); c is blocked when all
// of them are.  During BCE proper (no CLE / CLA / BCM) clause contents and occurrence-list orders never change - clauses only get
// deleted - so every answer the run will ever need is a function of the data base at its start: one bit per (clause with +v,
// clause with ~v).  This kernel computes all of them in one launch; the host engine walks the reference's literal heap and
// lists in the reference's order and reads bits (Engine::blockedClauseElimination).
//
// Work unit = one 32-bit mask word = one clause P of occ(+v) against up to 32 consecutive clauses N of occ(~v): a sorted merge
// that stops at the first complementary pair.  Algorithmic bytes per word: 8 + 4*#P once + (4 + 8 + 4*#N) per N read, 4 written.
// HBM/latency bound integer work; the N clauses of a variable are shared by all of its rows (L2 hits).  No tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cp3b200.h"
#include "gpu_db.h"

// 1 iff resolving P (holds pv) with N (holds pv^1) gives no tautology
__device__ __forceinline__ uint32_t non_tautologic(const int32_t* __restrict__ P, uint32_t pn, const int32_t* __restrict__ N, uint32_t nn, int32_t pv)
{
    const int32_t nv = pv ^ 1;
    uint32_t i = 0, j = 0;
    while (i < pn && j < nn) {
        const int32_t a = P[i], b = N[j];
        if (a == pv) { ++i; }
        else if (b == nv) { ++j; }
        else if (a == b) { ++i; ++j; }
        else if (a == (b ^ 1)) { return 0u; }
        else if (a < b) { ++i; }
        else { ++j; }
    }
    return 1u;
}

__global__ void __launch_bounds__(256) k_bce_masks(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t nv,
                                                   const uint32_t* __restrict__ vars, const uint32_t* __restrict__ p_off, const uint32_t* __restrict__ p_refs,
                                                   const uint32_t* __restrict__ n_off, const uint32_t* __restrict__ n_refs,
                                                   const unsigned long long* __restrict__ w_off, unsigned long long total_words, uint32_t* __restrict__ masks)
{
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total_words; w += (unsigned long long)gridDim.x * blockDim.x) {
        // variable that owns word w: last i with w_off[i] <= w
        uint32_t lo = 0, hi = nv;
        while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (w_off[mid] <= w) { lo = mid; } else { hi = mid; } }
        const uint32_t i = lo;
        const uint32_t nb = n_off[i], nn = n_off[i + 1] - nb;
        const uint32_t W = (nn + 31) >> 5;
        const unsigned long long r = w - w_off[i];
        const uint32_t a = (uint32_t)(r / W), wb = (uint32_t)(r - (unsigned long long)a * W);
        const int32_t pv = (int32_t)(vars[i] << 1);
        const ClauseHdr hp = hdr[p_refs[p_off[i] + a]];
        const uint32_t pn = hp.szfl & SIZE_MASK;
        const int32_t* P = lits + hp.start;
        uint32_t word = 0;
        const uint32_t b0 = wb << 5, b1 = min(nn, b0 + 32u);
        for (uint32_t b = b0; b < b1; ++b) {
            const ClauseHdr hn = hdr[n_refs[nb + b]];
            const uint32_t nsz = hn.szfl & SIZE_MASK;
            const int32_t* N = lits + hn.start;
            const uint32_t bit = non_tautologic(P, pn, N, nsz, pv);
            word |= bit << (b - b0);
        }
        masks[w] = word;
    }
}

extern "C" int cp3g_bce_masks(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, const uint32_t* p_refs,
                              const uint32_t* n_off, const uint32_t* n_refs, const uint64_t* word_off, uint32_t* masks)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!nv) { return CP3G_OK; }
    cudaSetDevice(g->device);
    const size_t np = p_off[nv], nn = n_off[nv]; const uint64_t total = word_off[nv];
    if (!total) { return CP3G_OK; }
    RESERVE(g->bu0, nv, false); RESERVE(g->bu1, (size_t)nv + 1, false); RESERVE(g->bu2, np + 1, false);
    RESERVE(g->bu3, (size_t)nv + 1, false); RESERVE(g->bu4, nn + 1, false); RESERVE(g->bq0, (size_t)nv + 1, false);
    RESERVE(g->scratch_u32, total + 1, false);
    CK(cudaMemcpyAsync(g->bu0.p, vars, (size_t)nv * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu1.p, p_off, ((size_t)nv + 1) * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu2.p, p_refs, np * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu3.p, n_off, ((size_t)nv + 1) * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu4.p, n_refs, nn * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bq0.p, word_off, ((size_t)nv + 1) * 8, cudaMemcpyHostToDevice, g->stream));
    // (sharded mode: every rank computes the whole mask set - the host commit that consumes it is replicated anyway and the
    //  masks are 1 bit per pair; no exchange)
    KTimer t(g);
    const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)g->num_sms * 32);
    k_bce_masks<<<blocks, 256, 0, g->stream>>>(g->hdr.p, g->lits.p, nv, g->bu0.p, g->bu1.p, g->bu2.p, g->bu3.p, g->bu4.p, g->bq0.p, total, g->scratch_u32.p);
    t.stop();
    CK(cudaMemcpyAsync(masks, g->scratch_u32.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    g->bce_ns += t.collect();
    g->launches++; g->h2d += (int64_t)(nv * 20 + (np + nn) * 4); g->d2h += (int64_t)total * 4;
    // algorithmic bytes (BVE-style): per mask word the P clause once and up to 32 N clauses, 4 bytes written
    g->bce_bytes += (int64_t)total * 4;
    return CP3G_OK;
}
