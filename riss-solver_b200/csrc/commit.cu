// commit.cu - device-side commit of the first (full-queue) subsumption pass.
//
// Subsumption::subsumption_worker (coprocessor/techniques/Subsumption.cc:244-313) drains the queue from the back; what makes the
// walk look sequential is (1) whether a subsumer is still alive at its turn, (2) lit_occurrence_count - it picks the list a
// subsumer walks and drops at the moment a clause is subsumed (removedClause, CoprocessorTypes.h:1310) - and (3) which entries a
// walk finds deleted and removes by swap-with-last.  All three are functions of WHEN each subsumed clause dies, and in this pass
// clause contents do not change, so every K3 hit stays valid.  For the first pass after initializePreprocessor - every live
// clause queued in clause order, lists pristine (ascending clause index, no deleted entries) - the whole commit runs here:
//
//   k_sp_*      death times: dtime[D] = min over subsumers C of D that are alive at their own turn of T(C), T(C) = ncl-1-C
//               (a later clause index is processed earlier).  Well-founded in T, evaluated by fixpoint iteration over the hit
//               list (2-4 rounds: the depth of "my subsumer was subsumed first" chains, e.g. groups of duplicates).
//   k_sp_dead   the subsumed clauses, and per literal the death times of the clauses that contain it (count, scan, fill)
//   k_sp_choose every queue entry picks its list with the counters as of its own time; walks of lists that never hold a deleted
//               clause just add |list| - 1 steps
//   k_sp_replay one thread per list that holds a deleted clause: replays the walks of that list in time order (the walkers of a
//               list are entries of that list: descending clause index = ascending time), reproducing the lazy swap-with-last
//               cleaning, the size filter in front of it and the step counter exactly
//
// The host receives the subsumed set, the step counter and the lists in their post-pass order (it has not downloaded the
// pristine lists at all in this case) and only does the O(#subsumed) bookkeeping.  Everything here is HBM/latency bound
// integer work on per-clause / per-literal tables; no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/cp3b200.h"
#include "gpu_db.h"

static constexpr int32_t T_INF = 0x7fffffff;

__global__ void k_sp_fill_i32(int32_t* __restrict__ a, int32_t v, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = v; }
}
// one fixpoint round over the hits (req = subsumer C, clause = subsumed D)
__global__ void k_sp_reset(const cp3g_hit* __restrict__ h, uint32_t n, int32_t* __restrict__ dnew)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { dnew[h[i].clause] = T_INF; }
}
__global__ void k_sp_min(const cp3g_hit* __restrict__ h, uint32_t n, const int32_t* __restrict__ dold, int32_t* __restrict__ dnew, uint32_t ncl)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const uint32_t c = h[i].req; const int32_t t = (int32_t)(ncl - 1 - c);
    if (dold[c] >= t) { atomicMin(&dnew[h[i].clause], t); }          // the subsumer is alive at its turn
}
__global__ void k_sp_diff(const cp3g_hit* __restrict__ h, uint32_t n, const int32_t* __restrict__ dold, const int32_t* __restrict__ dnew, uint32_t* __restrict__ changed)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && dold[h[i].clause] != dnew[h[i].clause]) { *changed = 1u; }
}
// subsumed clauses: compact id list, literal total, per-literal death counts
__global__ void k_sp_dead(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, const int32_t* __restrict__ dtime, uint32_t ncl,
                          uint32_t* __restrict__ dead, uint32_t* __restrict__ ndead, unsigned long long* __restrict__ nlits, uint32_t* __restrict__ dl_cnt)
{
    // the two global counters are same-address atomics (~1.2 ns each): one pair per block - ballot + shared-memory atomics for the
    // slots inside the block, shuffle reduction for the literal total (one pair per warp still cost 0.3 ms at C2)
    __shared__ uint32_t s_cnt, s_base; __shared__ unsigned long long s_tot;
    if (threadIdx.x == 0) { s_cnt = 0; s_tot = 0; }
    __syncthreads();
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool is_dead = c < ncl && dtime[c] != T_INF;
    const unsigned m = __ballot_sync(0xffffffffu, is_dead);
    const unsigned lane = threadIdx.x & 31;
    uint32_t sz = 0; ClauseHdr h{};
    if (is_dead) { h = hdr[c]; sz = h.szfl & SIZE_MASK; }
    uint32_t tot = sz;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, d); }
    uint32_t wbase = 0;
    if (lane == 0 && m) { wbase = atomicAdd(&s_cnt, (uint32_t)__popc(m)); atomicAdd(&s_tot, (unsigned long long)tot); }
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) { s_base = atomicAdd(ndead, s_cnt); atomicAdd(nlits, s_tot); }
    __syncthreads();
    if (!is_dead) { return; }
    dead[s_base + wbase + __popc(m & ((1u << lane) - 1u))] = c;
    for (uint32_t k = 0; k < sz; ++k) { atomicAdd(&dl_cnt[lits[h.start + k]], 1u); }
}
__global__ void k_sp_dlfill(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, const int32_t* __restrict__ dtime,
                            const uint32_t* __restrict__ dead, uint32_t ndead, const uint32_t* __restrict__ dl_off, uint32_t* __restrict__ dl_fill,
                            int32_t* __restrict__ dl_time)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ndead) { return; }
    const uint32_t c = dead[i];
    const ClauseHdr h = hdr[c];
    const uint32_t sz = h.szfl & SIZE_MASK; const int32_t t = dtime[c];
    for (uint32_t k = 0; k < sz; ++k) { const int32_t x = lits[h.start + k]; dl_time[dl_off[x] + atomicAdd(&dl_fill[x], 1u)] = t; }
}
// longest list that holds a subsumed clause (the replay of such a list is one thread's sequential work)
__global__ void k_sp_maxlen(const uint32_t* __restrict__ off, const uint32_t* __restrict__ dl_cnt, uint32_t nlit, uint32_t* __restrict__ mx)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < nlit && dl_cnt[x]) { atomicMax(mx, off[x + 1] - off[x]); }
}
// list choice of every queue entry as of its own time (Subsumption.cc:252-258) + steps of the walks over clean lists
__global__ void k_sp_choose(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, const int32_t* __restrict__ dtime, uint32_t ncl,
                            const uint32_t* __restrict__ off, const uint32_t* __restrict__ dl_off, const uint32_t* __restrict__ dl_cnt,
                            const int32_t* __restrict__ dl_time, int32_t* __restrict__ chosen, unsigned long long* __restrict__ steps)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long st = 0;
    if (c < ncl) {
        const ClauseHdr h = hdr[c];
        const int32_t t = (int32_t)(ncl - 1 - c);
        int32_t pick = -1;
        if (!(h.szfl & FLAG_DEL) && dtime[c] >= t) {                 // in the queue and not subsumed before its turn
            const uint32_t sz = h.szfl & SIZE_MASK;
            int32_t best = 0x7fffffff;
            for (uint32_t k = 0; k < sz; ++k) {
                const int32_t x = lits[h.start + k];
                int32_t cnt = (int32_t)(off[x + 1] - off[x]);
                const uint32_t nd = dl_cnt[x];
                if (nd) { const int32_t* q = dl_time + dl_off[x]; for (uint32_t j = 0; j < nd; ++j) { cnt -= q[j] < t ? 1 : 0; } }
                if (cnt < best) { best = cnt; pick = x; }
            }
            if (pick >= 0 && dl_cnt[pick] == 0) { const uint32_t n = off[pick + 1] - off[pick]; st = n ? n - 1 : 0; pick = -1; }   // clean list: every entry but the clause itself
        }
        chosen[c] = pick;
    }
    // block reduction of the step counts
    for (int d = 16; d > 0; d >>= 1) { st += __shfl_down_sync(0xffffffffu, st, d); }
    __shared__ unsigned long long wsum[8];
    if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = st; }
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long s = 0; for (unsigned w = 0; w < (blockDim.x >> 5); ++w) { s += wsum[w]; } if (s) { atomicAdd(steps, s); } }
}
// replay of the walks over one list that holds subsumed clauses (Subsumption.cc:259-273): `orig` is the pristine list (ascending
// clause index), `work` the copy that is cleaned in place
__global__ void k_sp_replay(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ dtime, uint32_t ncl, uint32_t nlit,
                            const uint32_t* __restrict__ off, const uint32_t* __restrict__ dl_cnt, const int32_t* __restrict__ chosen,
                            const uint32_t* __restrict__ orig, uint32_t* __restrict__ work, uint32_t* __restrict__ list_n, uint32_t* __restrict__ list_stale,
                            unsigned long long* __restrict__ steps)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long st = 0;
    if (x < nlit) {
        const uint32_t b = off[x]; uint32_t n = off[x + 1] - b;
        const uint32_t nd = dl_cnt[x];
        uint32_t removed = 0;
        if (nd) {
            const uint32_t n0 = n;
            uint32_t* w = work + b;
            for (uint32_t k = n0; k-- > 0;) {                        // walkers in time order = descending clause index
                const uint32_t cr = orig[b + k];
                if (chosen[cr] != (int32_t)x) { continue; }
                const int32_t t = (int32_t)(ncl - 1 - cr);
                const uint32_t cn = hdr[cr].szfl & SIZE_MASK;
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t d = w[j];
                    if (d == cr) { continue; }
                    ++st;
                    if (dtime[d] < t) {                              // found deleted: size filter first, then lazy removal
                        if ((hdr[d].szfl & SIZE_MASK) < cn) { continue; }
                        w[j] = w[n - 1]; --n; --j; ++removed;
                    }
                }
            }
        }
        list_n[x] = n; list_stale[x] = nd - removed;
    }
    for (int d = 16; d > 0; d >>= 1) { st += __shfl_down_sync(0xffffffffu, st, d); }
    __shared__ unsigned long long wsum[8];
    if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = st; }
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long s = 0; for (unsigned w2 = 0; w2 < (blockDim.x >> 5); ++w2) { s += wsum[w2]; } if (s) { atomicAdd(steps, s); } }
}

// K10: static part of the strengthening pass behind a committed subsumption pass.  One thread per clause; the per-literal tables
// (CSR offsets, deaths, post-pass list lengths, deleted entries still listed) are 4 x 8 MB at C2 and stay L2 resident, the clause
// arena is streamed once: 16 + 4*#C bytes per clause, one 4-byte atomic OR per walker.
__global__ void __launch_bounds__(256) k_str_static(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t ncl,
                                                    const uint32_t* __restrict__ real_bits, const uint32_t* __restrict__ off, const uint32_t* __restrict__ dl_cnt,
                                                    const uint32_t* __restrict__ ln, const uint32_t* __restrict__ ls, uint32_t* __restrict__ walked,
                                                    unsigned long long* __restrict__ out3)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long st = 0; uint32_t ent = 0, small = 0;
    if (c < ncl && !((real_bits[c >> 5] >> (c & 31)) & 1u)) {
        const ClauseHdr h = hdr[c];
        const uint32_t sz = h.szfl & SIZE_MASK;
        if (!(h.szfl & FLAG_DEL)) {
            if (sz < 2) { small = 1; }
            else {
                int32_t mn = 0, best = 0;
                for (uint32_t k = 0; k < sz; ++k) {
                    const int32_t x = lits[h.start + k]; const uint32_t p = (uint32_t)x & ~1u;
                    // lit_occurrence_count after the pass: list length at build time minus twice the deaths (Subsumption.cc:284,305)
                    const int32_t d = (int32_t)(off[p + 2] - off[p]) - 2 * (int32_t)(dl_cnt[p] + dl_cnt[p + 1]);
                    if (k == 0 || best > d) { best = d; mn = x; }
                }
                const uint32_t l1 = ln[mn] - ls[mn], l2 = ln[mn ^ 1] - ls[mn ^ 1];
                st = (l1 ? l1 - 1 : 0) + l2; ent = 1;
                atomicOr(&walked[(uint32_t)mn >> 5], 3u << ((uint32_t)mn & 30u));          // both literals of the variable share a word
            }
        }
    }
    // block totals: steps, entries, small clauses
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { st += __shfl_xor_sync(0xffffffffu, st, d); ent += __shfl_xor_sync(0xffffffffu, ent, d); small += __shfl_xor_sync(0xffffffffu, small, d); }
    __shared__ unsigned long long ws[8]; __shared__ uint32_t we[8], wm[8];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = st; we[threadIdx.x >> 5] = ent; wm[threadIdx.x >> 5] = small; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0, e = 0, m = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) { s += ws[w]; e += we[w]; m += wm[w]; }
        if (s) { atomicAdd(out3, s); } if (e) { atomicAdd(out3 + 1, e); } if (m) { atomicAdd(out3 + 2, m); }
    }
}

extern "C" {

int cp3g_str_static(cp3g* g, const uint32_t* real_bits, int* status, uint64_t* inert_entries, uint64_t* inert_steps, uint64_t* small_clauses,
                    uint32_t* walked_bits)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    if (status) { *status = 1; }
    if (!g->sp_valid || g->sp_build != g->build_count || g->sp_ncl != g->ncl || g->ncl != g->ncl_csr || !g->ovf_host.empty()) { return CP3G_OK; }
    const uint32_t ncl = g->ncl, nlit = 2 * g->nvars;
    const size_t rwords = ((size_t)ncl + 31) / 32, wwords = ((size_t)nlit + 31) / 32;
    RESERVE(g->scratch_u32, rwords + wwords + 8, false);
    uint32_t* d_real = g->scratch_u32.p; uint32_t* d_walked = g->scratch_u32.p + rwords;
    unsigned long long* d_out = (unsigned long long*)(g->counters.p + 8);
    CK(cudaMemcpyAsync(d_real, real_bits, rwords * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemsetAsync(d_walked, 0, wwords * 4, g->stream));
    CK(cudaMemsetAsync(g->counters.p + 8, 0, 8 * sizeof(uint32_t), g->stream));
    KTimer t(g);
    k_str_static<<<grid_for(ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, ncl, d_real, g->occ_off.p, g->sp_dlcnt.p, g->sp_n.p, g->sp_stale.p, d_walked, d_out);
    t.stop();
    unsigned long long out3[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(out3, d_out, sizeof(out3), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaMemcpyAsync(walked_bits, d_walked, wwords * 4, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    g->commit_ns += t.collect(); g->launches++;
    g->commit_bytes += 16ll * (int64_t)ncl + 4ll * (int64_t)g->nlits;
    g->h2d += (int64_t)rwords * 4; g->d2h += (int64_t)wwords * 4 + 24;
    if (inert_steps) { *inert_steps = out3[0]; } if (inert_entries) { *inert_entries = out3[1]; } if (small_clauses) { *small_clauses = out3[2]; }
    if (status) { *status = 0; }
    return CP3G_OK;
}

int cp3g_sub_pass(cp3g* g, uint32_t max_list, int* status, uint32_t* ndead_out, uint64_t* nlits_out, uint64_t* steps_out, uint64_t* checks_out)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    if (status) { *status = 1; }
    g->sp_valid = false;
    if (g->ncl == 0) { return CP3G_OK; }
    // K3 over the full queue (sharded: this rank's slice of the tile table); the hits stay on the device
    uint32_t nh = 0; uint64_t chk = 0;
    uint32_t cap = (uint32_t)std::max<size_t>(g->hits.cap ? g->hits.cap - 1 : 0, std::max<size_t>(4096, g->ncl / 8));
    for (;;) {
        int r = cp3g_scan_all(g, CP3G_SUBSUME, nullptr, cap, &nh, &chk);
        if (r != CP3G_OK) { return r; }
        if (nh <= cap) { break; }
        cap = nh + nh / 8 + 1024;
    }
    if (g->world > 1) {
        // exchange step of the sharded pass: the slices' hit lists are all-gathered over NCCL on the stream that
        // produced them and never leave the device; every rank then runs the commit kernels below on the identical list (the
        // fixpoint over death times is independent of the order of the hits), so the replicas stay bit-identical
        const cp3g_hit* all = nullptr; uint32_t total = 0;
        int r = cp3g_nccl_gather_hits_device(g->nccl, g->stream, g->hits.p, nh, &all, &total, &chk);
        if (r != CP3G_OK) { g->err = "nccl all-gather-v (K3 hits, device) failed"; return CP3G_ERR_CUDA; }
        RESERVE(g->hits, (size_t)total + 1, false);
        if (total) { CK(cudaMemcpyAsync(g->hits.p, all, (size_t)total * sizeof(cp3g_hit), cudaMemcpyDeviceToDevice, g->stream)); }
        nh = total;
    }
    if (!g->ovf_host.empty() || g->ncl_csr != g->ncl || !g->total_occ) { return CP3G_OK; }
    const uint32_t ncl = g->ncl, nlit = 2 * g->nvars;
    KTimer t(g);
    RESERVE(g->sp_da, (size_t)ncl + 1, false); RESERVE(g->sp_db, (size_t)ncl + 1, false); RESERVE(g->sp_chosen, (size_t)ncl + 1, false);
    RESERVE(g->sp_dlcnt, (size_t)nlit + 2, false); RESERVE(g->sp_dloff, (size_t)nlit + 2, false); RESERVE(g->sp_dlfill, (size_t)nlit + 2, false);
    RESERVE(g->sp_n, (size_t)nlit + 1, false); RESERVE(g->sp_stale, (size_t)nlit + 1, false);
    k_sp_fill_i32<<<grid_for(ncl, 256), 256, 0, g->stream>>>(g->sp_da.p, T_INF, ncl);
    k_sp_fill_i32<<<grid_for(ncl, 256), 256, 0, g->stream>>>(g->sp_db.p, T_INF, ncl);
    g->launches += 2;
    int32_t* dold = g->sp_da.p; int32_t* dnew = g->sp_db.p;
    bool converged = nh == 0;
    for (int round = 0; round < 64 && !converged; ++round) {
        CK(cudaMemsetAsync(g->counters.p + 8, 0, sizeof(uint32_t), g->stream));
        k_sp_reset<<<grid_for(nh, 256), 256, 0, g->stream>>>(g->hits.p, nh, dnew);
        k_sp_min<<<grid_for(nh, 256), 256, 0, g->stream>>>(g->hits.p, nh, dold, dnew, ncl);
        k_sp_diff<<<grid_for(nh, 256), 256, 0, g->stream>>>(g->hits.p, nh, dold, dnew, g->counters.p + 8);
        g->launches += 3;
        CK(cudaMemcpyAsync(g->pin_cnt, g->counters.p + 8, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
        CK(cudaStreamSynchronize(g->stream));
        std::swap(dold, dnew);                                       // dold = the newest values
        converged = g->pin_cnt[0] == 0;
    }
    if (!converged) { t.stop(); t.collect(); return CP3G_OK; }
    const int32_t* dtime = dold;
    g->sp_dtime = dold;
    // subsumed set + per-literal death times
    RESERVE(g->sp_dead, (size_t)nh + 1, false);
    CK(cudaMemsetAsync(g->sp_dlcnt.p, 0, ((size_t)nlit + 2) * sizeof(uint32_t), g->stream));
    CK(cudaMemsetAsync(g->sp_dlfill.p, 0, ((size_t)nlit + 2) * sizeof(uint32_t), g->stream));
    CK(cudaMemsetAsync(g->counters.p + 8, 0, 8 * sizeof(uint32_t), g->stream));
    unsigned long long* d_nlits = (unsigned long long*)(g->counters.p + 10);
    unsigned long long* d_steps = (unsigned long long*)(g->counters.p + 12);
    k_sp_dead<<<grid_for(ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, dtime, ncl, g->sp_dead.p, g->counters.p + 8, d_nlits, g->sp_dlcnt.p);
    k_sp_maxlen<<<grid_for(nlit, 256), 256, 0, g->stream>>>(g->occ_off.p, g->sp_dlcnt.p, nlit, g->counters.p + 9);
    g->launches += 2;
    { int r = scan_u32(g, g->sp_dlcnt.p, g->sp_dloff.p, (size_t)nlit + 1); if (r != CP3G_OK) { return r; } }
    uint32_t head[8];
    CK(cudaMemcpyAsync(head, g->counters.p + 8, sizeof(head), cudaMemcpyDeviceToHost, g->stream));
    uint32_t dl_total = 0;
    CK(cudaMemcpyAsync(&dl_total, g->counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    const uint32_t ndead = head[0], maxlen = head[1];
    unsigned long long nl; memcpy(&nl, &head[2], 8);
    if (maxlen > max_list) { t.stop(); t.collect(); return CP3G_OK; }   // a long list with deaths: one thread cannot replay it - host commit
    RESERVE(g->sp_dltime, (size_t)dl_total + 1, false);
    if (ndead) {
        k_sp_dlfill<<<grid_for(ndead, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, dtime, g->sp_dead.p, ndead, g->sp_dloff.p, g->sp_dlfill.p, g->sp_dltime.p);
        g->launches++;
    }
    k_sp_choose<<<grid_for(ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, dtime, ncl, g->occ_off.p, g->sp_dloff.p, g->sp_dlcnt.p, g->sp_dltime.p,
                                                          g->sp_chosen.p, d_steps);
    g->launches++;
    // pristine lists in clause order (refs_tmp) and the working copy that the replay cleans
    { int r = dev_host_order_lists(g); if (r != CP3G_OK) { return r; } }
    RESERVE(g->sp_refs, (size_t)g->total_occ + 1, false);
    CK(cudaMemcpyAsync(g->sp_refs.p, g->refs_tmp.p, (size_t)g->total_occ * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g->stream));
    k_sp_replay<<<grid_for(nlit, 128), 128, 0, g->stream>>>(g->hdr.p, dtime, ncl, nlit, g->occ_off.p, g->sp_dlcnt.p, g->sp_chosen.p, g->refs_tmp.p, g->sp_refs.p,
                                                           g->sp_n.p, g->sp_stale.p, d_steps);
    g->launches++;
    // the device applies the deletions to its own copy (header flag + bitmap)
    { int r = dev_set_deleted(g, g->sp_dead.p, ndead); if (r != CP3G_OK) { return r; } }
    unsigned long long steps = 0;
    CK(cudaMemcpyAsync(&steps, d_steps, sizeof(steps), cudaMemcpyDeviceToHost, g->stream));
    t.stop();
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    g->commit_ns += t.collect();
    // algorithmic bytes: the hits once per fixpoint round (12 B), per clause header + literals + death time + choice, per
    // occurrence the pristine and the working list (read + written)
    g->commit_bytes += 12ll * (int64_t)nh * 3 + (16ll + 8ll) * (int64_t)ncl + 4ll * (int64_t)g->nlits + 12ll * (int64_t)g->total_occ;
    g->sp_valid = true; g->sp_ndead = ndead; g->sp_build = g->build_count; g->sp_ncl = g->ncl;
    if (status) { *status = 0; }
    if (ndead_out) { *ndead_out = ndead; }
    if (nlits_out) { *nlits_out = nl; }
    if (steps_out) { *steps_out = steps; }
    if (checks_out) { *checks_out = chk; }
    return CP3G_OK;
}

int cp3g_sub_pass_fetch(cp3g* g, uint32_t* dead, uint32_t* list_n, uint32_t* list_stale, uint32_t* refs)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!g->sp_valid) { g->err = "cp3g_sub_pass_fetch without a committed pass"; return CP3G_ERR_ARG; }
    cudaSetDevice(g->device);
    const uint32_t nlit = 2 * g->nvars;
    if (dead && g->sp_ndead) { CK(cudaMemcpyAsync(dead, g->sp_dead.p, (size_t)g->sp_ndead * 4, cudaMemcpyDeviceToHost, g->stream)); g->d2h += (int64_t)g->sp_ndead * 4; }
    if (list_n) { CK(cudaMemcpyAsync(list_n, g->sp_n.p, (size_t)nlit * 4, cudaMemcpyDeviceToHost, g->stream)); g->d2h += (int64_t)nlit * 4; }
    if (list_stale) { CK(cudaMemcpyAsync(list_stale, g->sp_stale.p, (size_t)nlit * 4, cudaMemcpyDeviceToHost, g->stream)); g->d2h += (int64_t)nlit * 4; }
    CK(cudaStreamSynchronize(g->stream));
    if (refs && g->total_occ) { int r = d2h_staged(g, refs, g->sp_refs.p, (size_t)g->total_occ * 4); if (r != CP3G_OK) { return r; } g->d2h += (int64_t)g->total_occ * 4; }
    return CP3G_OK;
}

} // extern "C"
