// gpu_db.cu - device-resident clause data base and the sm_100a kernels of the scan service (cp3g_* ABI).
//
// HBM layout (DESIGN.md "data layout"):
//   lits[]     int32 literal pool, clause i = lits[hdr[i].start .. +size), sorted by literal code 2*var+sign
//   hdr[]      16-byte clause header {start:u32, size:24|flags:8, sig:u64}  - one 128-bit load per candidate
//   occ_off[]  u32 CSR offsets per literal code (2*nvars+1), occ_ent[] 16-byte entries {class:2|ref:30, size, sig}.
//              Inside a list the entries are ordered by (class, clause index): the clauses whose full-queue scan
//              walks THIS list ("tagged" for subsumption and/or strengthening) come first, so the list carries its
//              own subsumer set as a prefix.  cp3g_download_occ returns plain ascending clause index
//              (= CoprocessorData::addClause order, CoprocessorTypes.h:812-833)
//   tiles[]    {first entry | aligned flag, first literal} per 128-entry tile of the entry array, rounded down to
//              the start of the variable's list pair when that pair is short (full-queue kernels stage whole lists)
//   ovf[]      clause indices appended after the last CSR build (BVE resolvents); scanned by every request
//
// Kernels (none uses tensor cores - there is no contraction on this path):
//   k_occ_count   K1+K2  64-bit literal signature per clause (new accelerating filter; never changes results)
//                 fused with the occurrence histogram
//   k_occ_*       K2  exclusive scan -> scatter of (class, clause) words -> per-list sort -> tile table -> expand;
//                 large data bases: (literal, class) keys + stable cub radix sort instead of scatter + per-list sort
//   k_full<KIND>  K3/K4 full-queue pass over bulk-copy (cp.async.bulk + mbarrier) staged tiles of the entry array:
//                 flat (subsumer, candidate-chunk) work list per tile; segment / entry forms for cut tiles
//   k_scan<KIND>  K3/K4 request-driven signature filter + sorted-literal merge (Subsumption.cc:244-297,
//                 1276-1521; Clause::ordered_subsumes SolverTypes.h:1003); k_scan_ovf for long lists of young clauses
//   k_bve_pairs   K5  resolvent size per (pos,neg) pair (BVE.h:248-299)
//   k_bve_resolve K6  resolvent generation (BVE.h:208-242)
//   k_dense_*     K7  Dense::compress: mark occurring variables, prefix-sum renumbering, literal rewrite (Dense.cc:26-238)
//   k_compact_*   K8  arena stream compaction: scan of the live sizes + gather (garbageCollect, CoprocessorTypes.h:1396)
#include <cuda_runtime.h>
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <type_traits>
#include <chrono>
#include <mutex>
#include <utility>

#include "../../include/cp3b200.h"
#include "gpu_db.h"

static constexpr int SHORT_LIST = 48;
// occurrence entry: top two bits of `ref` = class of the clause in this list (see k_occ_fill)
static constexpr uint32_t CLS_NOSUB = 1u, CLS_NOSTR = 2u, CLS_NONE = 3u;
// full-queue kernel tiling
#ifndef CP3_TILE
#define CP3_TILE 128
#endif
#ifndef CP3_DCHUNK
#define CP3_DCHUNK 8
#endif
#ifndef CP3_MINBLOCKS
#define CP3_MINBLOCKS 6
#endif
#ifndef CP3_GROUPMAX
#define CP3_GROUPMAX 32
#endif
#ifndef CP3_OFFCAP
#define CP3_OFFCAP CP3_TILE
#endif
#ifndef CP3_WORKCAP
#define CP3_WORKCAP (2 * CP3_TILE)
#endif
#ifndef CP3_WQ
#define CP3_WQ 64
#endif
#ifndef CP3_WH
#define CP3_WH 32
#endif
#ifndef CP3_FQWARPS
#define CP3_FQWARPS 4
#endif
static constexpr uint32_t TILE = CP3_TILE, GROUP_MAX = CP3_GROUPMAX, TCAP = TILE + GROUP_MAX, OFFCAP = CP3_OFFCAP, TILE_ALIGNED = 0x80000000u;
static constexpr uint32_t PIN_HITS = 8192;
static std::mutex g_pin_mu;
static std::vector<std::pair<uint32_t*, cp3g_hit*>> g_pin_free;      // pinned pairs of destroyed handles, reused by the next one


__host__ __device__ __forceinline__ unsigned long long lit_bit(int32_t x)
{
    const uint32_t v = ((uint32_t)x) >> 1;
    const uint32_t h = (v * 0x9E3779B1u) >> 27;            // 5 bits from the variable
    return 1ull << (2 * h + (x & 1));
}
__device__ __forceinline__ unsigned long long var_fold(unsigned long long s)
{
    return (s | (s >> 1)) & 0x5555555555555555ull;
}

// streaming (evict-first) accesses for data that is read or written exactly once per build pass, so that the tables
// the passes gather from at random (offsets, counters, refs, clause headers) stay resident in the 126 MB L2
__device__ __forceinline__ ClauseHdr ld_stream(const ClauseHdr* p)
{
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p));
    ClauseHdr h; h.start = v.x; h.szfl = v.y; h.sig = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
    return h;
}
__device__ __forceinline__ void st_stream(OccEnt* p, const OccEnt& e)
{
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(e.ref, e.szfl, (uint32_t)e.sig, (uint32_t)(e.sig >> 32)));
}

// ------------------------------------------------------------------------------------------ K1 + K2
__global__ void k_occ_count(ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t n,
                            uint32_t* __restrict__ cnt, uint32_t* __restrict__ rank)
{
    // histogram of the literal occurrences; the same pass over the literals refreshes the clause signature (K1).
    // The value the atomic returns is kept as the occurrence's (arbitrary) rank inside its list, stored next to the
    // literal - the scatter pass then needs no second round of atomics.
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const ClauseHdr h = ld_stream(hdr + i);
    const uint32_t sz = h.szfl & SIZE_MASK;
    unsigned long long s = 0;
    const bool live = !(h.szfl & FLAG_DEL);
    if (rank != nullptr) {
        for (uint32_t k = 0; k < sz; ++k) { const int32_t x = __ldcs(lits + h.start + k); s |= lit_bit(x); if (live) { __stcs(rank + h.start + k, atomicAdd(&cnt[x], 1u)); } }
    } else {          // sort path (large data bases): only the histogram
        for (uint32_t k = 0; k < sz; ++k) { const int32_t x = __ldcs(lits + h.start + k); s |= lit_bit(x); if (live) { atomicAdd(&cnt[x], 1u); } }
    }
    if (s != h.sig) { hdr[i].sig = s; }
}

// exclusive scan, three small kernels (tile = 1024 threads x 4 items)
static constexpr int SCAN_T = 1024, SCAN_I = 4, SCAN_TILE = SCAN_T * SCAN_I;
__global__ void k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                             uint32_t* __restrict__ tile_sum)
{
    __shared__ uint32_t warp_tot[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) { v[k] = (base + k < n) ? in[base + k] : 0u; s += v[k]; }
    uint32_t inc = s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) { inc += t; } }
    if (lane == 31) { warp_tot[w] = inc; }
    __syncthreads();
    if (w == 0) {
        uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) { ti += u; } }
        warp_tot[lane] = ti - t;
        if (lane == 31) { tile_sum[blockIdx.x] = ti; }
    }
    __syncthreads();
    uint32_t ex = warp_tot[w] + inc - s;
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) { if (base + k < n) { out[base + k] = ex; } ex += v[k]; }
}
__global__ void k_scan_sums(uint32_t* __restrict__ tile_sum, uint32_t ntiles, uint32_t* __restrict__ total)
{
    // single block; sequential over chunks of blockDim
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) { carry = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sum[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
            uint32_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < ntiles) { tile_sum[i] = carry + sh[threadIdx.x] - v; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { carry += sh[threadIdx.x]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *total = carry; }
}
__global__ void k_scan_add(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ tile_sum)
{
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_I;
    const uint32_t add = tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) { if (base + k < n) { out[base + k] += add; } }
}

// Scatter only the 4-byte clause index: the refs of a list (avg 6.4 x 4 B) share one 32-byte sector and the whole
// array (55 MB at C2) stays L2 resident, whereas scattering the 16-byte entries costs a DRAM read-modify-write per
// partial sector (0.98 ms at C2, profiles/r01b).  The entries are written afterwards by k_occ_expand, coalesced.
//
// The two top bits of the scattered word carry the CLASS of clause i in that list.  A full-queue scan walks, for
// clause i, the shortest list among its literals (subsumption, Subsumption.cc:244-256) resp. the shortest list pair
// occ(x)+occ(~x) (strengthening, Subsumption.cc:1292-1299); i is "tagged" in exactly that list.  class = 0 tagged
// for both, 1 strengthening only, 2 subsumption only, 3 not tagged; the per-list sort on the whole word then puts
// the tagged clauses first, so every list carries the set of clauses that scan it as a prefix (the strengtheners
// exactly, K4 being the more expensive pass; K3 skips the few class-1 entries).
__global__ void k_occ_fill(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t n,
                           const uint32_t* __restrict__ off, const uint32_t* __restrict__ rank, uint32_t* __restrict__ refs)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const ClauseHdr h = ld_stream(hdr + i);
    if (h.szfl & FLAG_DEL) { return; }
    const uint32_t sz = h.szfl & SIZE_MASK;
    uint32_t bs = 0xffffffffu, bt = 0xffffffffu; int32_t xs = -1, xt = -1;
    for (uint32_t k = 0; k < sz; ++k) {
        const int32_t x = lits[h.start + k];
        const uint32_t len = off[x + 1] - off[x];
        const uint32_t both = len + off[(x ^ 1) + 1] - off[x ^ 1];
        if (len < bs) { bs = len; xs = x; }
        if (both < bt) { bt = both; xt = x; }
    }
    for (uint32_t k = 0; k < sz; ++k) {
        const int32_t x = lits[h.start + k];
        const uint32_t cls = (x == xs ? 0u : CLS_NOSUB) | (x == xt ? 0u : CLS_NOSTR);
        refs[off[x] + __ldcs(rank + h.start + k)] = i | (cls << 30);
    }
}
// Large data bases (the refs array does not fit the L2: at 54 M clauses the random 4-byte scatter of k_occ_fill moves
// 163 bytes of DRAM traffic per occurrence): the occurrences are SORTED instead.  key = literal << 2 | class, value =
// clause index, written at the literal's own pool slot (coalesced); a stable LSD radix sort (cub) then yields every
// list in (class, clause index) order - pool slots ascend with the clause index - without atomics or per-list sorts.
__global__ void k_occ_keys(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t n,
                           const uint32_t* __restrict__ off, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const ClauseHdr h = ld_stream(hdr + i);
    if (h.szfl & FLAG_DEL) { return; }
    const uint32_t sz = h.szfl & SIZE_MASK;
    uint32_t bs = 0xffffffffu, bt = 0xffffffffu; int32_t xs = -1, xt = -1;
    for (uint32_t k = 0; k < sz; ++k) {
        const int32_t x = lits[h.start + k];
        const uint32_t len = off[x + 1] - off[x];
        const uint32_t both = len + off[(x ^ 1) + 1] - off[x ^ 1];
        if (len < bs) { bs = len; xs = x; }
        if (both < bt) { bt = both; xt = x; }
    }
    for (uint32_t k = 0; k < sz; ++k) {
        const int32_t x = lits[h.start + k];
        const uint32_t cls = (x == xs ? 0u : CLS_NOSUB) | (x == xt ? 0u : CLS_NOSTR);
        __stcs(keys + h.start + k, ((uint32_t)x << 2) | cls);
        __stcs(vals + h.start + k, i);
    }
}
__global__ void k_refs_from_sorted(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t total, uint32_t* __restrict__ refs)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) { refs[i] = __ldcs(vals + i) | ((__ldcs(keys + i) & 3u) << 30); }
}
// pool slots must ascend with the clause index for the stable sort to give clause order
__global__ void k_check_monotone(const ClauseHdr* __restrict__ hdr, uint32_t n, uint32_t* __restrict__ bad)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < n && hdr[i + 1].start < hdr[i].start + (hdr[i].szfl & SIZE_MASK) && !(hdr[i].szfl & FLAG_DEL) && !(hdr[i + 1].szfl & FLAG_DEL)) { *bad = 1u; }
}
// one thread per entry: gather size + signature of the referenced clause, write the 16-byte entry coalesced.
// The top byte of the size word gets dx = (literal of the entry's list) - (first literal of its tile), saturating
// at 255, so that the full-queue kernels find the list of a staged entry without a search.  The search happens
// here instead, over the handful of offsets of the tile (neighbouring threads walk the same cached words).
__global__ void k_occ_expand(const ClauseHdr* __restrict__ hdr, const uint32_t* __restrict__ refs, OccEnt* __restrict__ ent, uint32_t total,
                             const uint32_t* __restrict__ off, const uint2* __restrict__ tiles)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) { return; }
    const uint32_t r = __ldcs(refs + i);
    uint32_t t = i / TILE;
    uint2 a = tiles[t], b = tiles[t + 1];
    if (i >= (b.x & ~TILE_ALIGNED)) { a = b; b = tiles[t + 2]; }
    uint32_t lo = a.y, hi = (b.x & TILE_ALIGNED) ? b.y : b.y + 1;         // last x with off[x] <= i
    while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (off[mid] <= i) { lo = mid; } else { hi = mid; } }
    const uint32_t dx = min(lo - a.y, 255u);
    const ClauseHdr h = hdr[r & REF_MASK];
    OccEnt e; e.ref = r; e.szfl = (h.szfl & SIZE_MASK) | (dx << 24); e.sig = h.sig;
    st_stream(ent + i, e);
}
// Tile table of the full-queue kernels.  Tile t nominally starts at entry 128*t; the start is moved down to the
// first entry of the variable's list pair (occ(2v), occ(2v+1)) when that pair is short, so that a tile consists
// of whole list pairs (flag TILE_ALIGNED) and fits the shared-memory staging area.  Long lists keep the
// nominal cut and are processed from global memory.
__global__ void k_tiles(const uint32_t* __restrict__ off, uint32_t nlit, uint32_t total, uint32_t ntiles, uint2* __restrict__ tiles)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) { return; }
    const unsigned long long j64 = (unsigned long long)t * TILE;
    if (t == ntiles || j64 >= total) { tiles[t] = make_uint2(total | TILE_ALIGNED, nlit); return; }
    const uint32_t j = (uint32_t)j64;
    uint32_t lo = 0, hi = nlit;                    // last x with off[x] <= j  (off[nlit] = total > j)
    while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (off[mid] <= j) { lo = mid; } else { hi = mid; } }
    const uint32_t x = lo, g = x & ~1u;
    const uint32_t gs = off[g], ge = off[g + 2];
    tiles[t] = (ge - gs <= GROUP_MAX) ? make_uint2(gs | TILE_ALIGNED, g) : make_uint2(j, x);
}
// scatter order is nondeterministic; every list is sorted by (class, clause index) - the scattered word itself.
static constexpr uint32_t SORT_LISTS = 128, SORT_CAP = 4096;
__global__ void __launch_bounds__(SORT_LISTS) k_occ_sort_short(const uint32_t* __restrict__ off, uint32_t nlit, uint32_t* __restrict__ refs,
                                 uint32_t* __restrict__ long_lits, uint32_t* __restrict__ nlong)
{
    // one thread per list, insertion sort; the lists of a block are contiguous in refs[], so the block stages that
    // range in shared memory (coalesced in, coalesced out) unless a long list makes it too large
    __shared__ uint32_t sh[SORT_CAP];
    const uint32_t x0 = blockIdx.x * SORT_LISTS, x = x0 + threadIdx.x, xe = min(x0 + SORT_LISTS, nlit);
    const uint32_t rb = off[x0], rn = off[xe] - rb;
    const bool staged = rn <= SORT_CAP;
    if (staged) { for (uint32_t i = threadIdx.x; i < rn; i += SORT_LISTS) { sh[i] = refs[rb + i]; } }
    __syncthreads();
    if (x < nlit) {
        const uint32_t b = off[x], len = off[x + 1] - b;
        if (len > SHORT_LIST) { long_lits[atomicAdd(nlong, 1u)] = x; }
        else if (len >= 2) {
            uint32_t* a = staged ? sh + (b - rb) : refs + b;
            for (uint32_t i = 1; i < len; ++i) {
                const uint32_t key = a[i];
                uint32_t j = i;
                while (j > 0 && a[j - 1] > key) { a[j] = a[j - 1]; --j; }
                a[j] = key;
            }
        }
    }
    __syncthreads();
    if (staged) { for (uint32_t i = threadIdx.x; i < rn; i += SORT_LISTS) { refs[rb + i] = sh[i]; } }
}
// one block per long list: normalised bitonic sort in global memory.  Every compare-exchange is ascending
// (i < l keeps the smaller ref at i), so the virtual +inf padding above len never has to move and pairs that
// touch it are simply skipped.
__global__ void k_occ_sort_long(const uint32_t* __restrict__ off, const uint32_t* __restrict__ long_lits,
                                uint32_t* __restrict__ refs)
{
    const uint32_t x = long_lits[blockIdx.x];
    uint32_t* a = refs + off[x];
    const uint32_t len = off[x + 1] - off[x];
    uint32_t p2 = 1; while (p2 < len) { p2 <<= 1; }
    for (uint32_t k = 2; k <= p2; k <<= 1) {
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
            const uint32_t l = i ^ (k - 1);
            if (l > i && l < len) { const uint32_t u = a[i], v = a[l]; if (u > v) { a[i] = v; a[l] = u; } }
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
                const uint32_t l = i ^ j;
                if (l > i && l < len) { const uint32_t u = a[i], v = a[l]; if (u > v) { a[i] = v; a[l] = u; } }
            }
            __syncthreads();
        }
    }
}

// Hits are appended to one global list.  A same-address atomicAdd per hit serialises in L2 (~1 ns each: at C2
// the 3*10^5 hits of a full-queue pass cost more than the scan itself), so every block stages its hits in
// shared memory and reserves their slots with ONE global atomic when it finishes.
static constexpr int SINK_CAP = 96;

struct HitSink {
    cp3g_hit* s_hits; uint32_t* s_n;
    cp3g_hit* out; uint32_t cap; uint32_t* nout;
    __device__ __forceinline__ void emit(uint32_t req, uint32_t clause, int32_t lit) const
    {
        const uint32_t p = atomicAdd(s_n, 1u);
        if (p < (uint32_t)SINK_CAP) { s_hits[p].req = req; s_hits[p].clause = clause; s_hits[p].lit = lit; }
        else {      // staging area full (a block with very many hits): fall back to the global counter
            const uint32_t q = atomicAdd(nout, 1u);
            if (q < cap) { out[q].req = req; out[q].clause = clause; out[q].lit = lit; }
        }
    }
    // all threads of the block must call this (contains __syncthreads)
    __device__ __forceinline__ void flush(uint32_t* s_base) const
    {
        __syncthreads();
        const uint32_t n = min(*s_n, (uint32_t)SINK_CAP);
        if (n == 0) { return; }               // uniform: every thread read the same value after the barrier
        if (threadIdx.x == 0) { *s_base = atomicAdd(nout, n); }
        __syncthreads();
        const uint32_t base = *s_base;
        for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) { if (base + k < cap) { out[base + k] = s_hits[k]; } }
        __syncthreads();
        if (threadIdx.x == 0) { *s_n = 0; }
        __syncthreads();
    }
};
#define SINK_DECL __shared__ cp3g_hit sink_hits[SINK_CAP]; __shared__ uint32_t sink_n, sink_base; \
    if (threadIdx.x == 0) { sink_n = 0; } __syncthreads(); \
    const HitSink sink{sink_hits, &sink_n, out, cap, nout}

// ------------------------------------------------------------------------------------------ K3 / K4
// returns true if S (sn literals) is a subset of D (dn literals); both sorted
__device__ __forceinline__ bool merge_subset(const int32_t* __restrict__ S, uint32_t sn,
                                             const int32_t* __restrict__ D, uint32_t dn)
{
    uint32_t i = 0, j = 0;
    while (i < sn && j < dn) {
        const int32_t a = S[i], b = D[j];
        if (a == b) { ++i; ++j; }
        else if (b < a) { ++j; }
        else { return false; }
    }
    return i == sn;
}
// self-subsuming resolution test: exactly one literal of S occurs complemented in D, the rest is in D.
// returns the complemented literal of D, or -1
__device__ __forceinline__ int32_t merge_one_flip(const int32_t* __restrict__ S, uint32_t sn,
                                                  const int32_t* __restrict__ D, uint32_t dn)
{
    uint32_t i = 0, j = 0; int32_t flipped = -1;
    while (i < sn && j < dn) {
        const int32_t a = S[i], b = D[j];
        if (a == b) { ++i; ++j; }
        else if (a == (b ^ 1)) { if (flipped >= 0) { return -1; } flipped = b; ++i; ++j; }
        else if (a < b) { return -1; }
        else { ++j; }
    }
    return (i == sn) ? flipped : -1;
}

template <int KIND>
__device__ __forceinline__ void test_candidate(uint32_t ref, uint32_t szfl, unsigned long long dsig, uint32_t me,
                                               const int32_t* __restrict__ S, uint32_t sn, unsigned long long ssig,
                                               int32_t need_flip /* -1: any */, int32_t forbid_flip /* -1: none */,
                                               const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits,
                                               uint32_t r, const HitSink& sink)
{
    if (ref == me) { return; }
    // the occ entry caches size/sig from build time; the header is authoritative for later edits, so a
    // cheap pre-filter on the entry is only valid for properties that are monotone: clauses only shrink
    // (sig bits only disappear, size only decreases) -> entry sig/size are supersets/upper bounds.
    const uint32_t esz = szfl & SIZE_MASK;
    if (esz < sn) { return; }
    const unsigned long long miss = ssig & ~dsig;
    if (KIND == CP3G_SUBSUME) { if (miss) { return; } }
    else { if (miss & (miss - 1)) { return; } if (var_fold(ssig) & ~var_fold(dsig)) { return; } }
    const ClauseHdr h = hdr[ref];
    if (h.szfl & FLAG_DEL) { return; }
    const uint32_t dn = h.szfl & SIZE_MASK;
    if (dn < sn) { return; }
    const int32_t* D = lits + h.start;
    if (KIND == CP3G_SUBSUME) {
        if ((ssig & ~h.sig) == 0 && merge_subset(S, sn, D, dn)) { sink.emit(r, ref, -1); }
    } else {
        const int32_t f = merge_one_flip(S, sn, D, dn);
        if (f >= 0 && (need_flip < 0 || f == need_flip) && f != forbid_flip) { sink.emit(r, ref, f); }
    }
}

// One warp per request.  Lanes stride over the chosen occurrence list(s); the signature filter rejects
// almost every candidate from the 16-byte entry alone, so clause literals are fetched only for survivors.
template <int KIND>
__global__ void __launch_bounds__(256) k_scan(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits,
        const uint32_t* __restrict__ occ_off, const OccEnt* __restrict__ occ_ent,
        const uint32_t* __restrict__ ovf, uint32_t novf,
        uint32_t req_begin, uint32_t req_end, const uint32_t* __restrict__ self,
        const uint32_t* __restrict__ req_off, const int32_t* __restrict__ req_lits,
        cp3g_hit* __restrict__ out, uint32_t cap, uint32_t* __restrict__ nout,
        unsigned long long* __restrict__ checks)
{
    SINK_DECL;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = req_begin + warp;
    bool live = r < req_end;
    uint32_t me = 0; const int32_t* S = nullptr; uint32_t sn = 0;
    if (live) {
        me = self[r];
        if (req_lits != nullptr) { S = req_lits + req_off[r]; sn = req_off[r + 1] - req_off[r]; }
        else {
            const ClauseHdr h = hdr[me];
            if (h.szfl & FLAG_DEL) { live = false; }
            S = lits + h.start; sn = h.szfl & SIZE_MASK;
        }
        if (sn == 0) { live = false; }
    }
    if (live) {
    // signature of the request and the cheapest list: lanes take literals round robin
    unsigned long long ssig = 0; uint32_t best_len = 0xffffffffu; uint32_t best_idx = 0;
    for (uint32_t k = lane; k < sn; k += 32) {
        const int32_t x = S[k];
        ssig |= lit_bit(x);
        uint32_t len = occ_off[x + 1] - occ_off[x];
        if (KIND == CP3G_STRENGTHEN) { len += occ_off[(x ^ 1) + 1] - occ_off[x ^ 1]; }
        if (len < best_len) { best_len = len; best_idx = k; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ssig |= __shfl_xor_sync(0xffffffffu, ssig, d);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, best_len, d);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, best_idx, d);
        if (ol < best_len || (ol == best_len && oi < best_idx)) { best_len = ol; best_idx = oi; }
    }
    const int32_t x = S[best_idx];
    unsigned long long nchk = 0, nbytes = 0;   // checks and their algorithmic bytes: 4 (entry) + 8 (header) + 4*#D
    {
        const uint32_t b = occ_off[x], e = occ_off[x + 1];
        for (uint32_t i = b + lane; i < e; i += 32) {
            OccEnt en = occ_ent[i]; en.ref &= REF_MASK;
            if (en.ref != me) { nchk += 1; nbytes += 12 + 4 * (en.szfl & SIZE_MASK); }
            // a stale entry could flip on x itself; that hit belongs to the occ(~x) pass below
            test_candidate<KIND>(en.ref, en.szfl, en.sig, me, S, sn, ssig, -1, x ^ 1, hdr, lits, r, sink);
        }
    }
    if (KIND == CP3G_STRENGTHEN) {
        const int32_t nx = x ^ 1;
        const uint32_t b = occ_off[nx], e = occ_off[nx + 1];
        for (uint32_t i = b + lane; i < e; i += 32) {
            OccEnt en = occ_ent[i]; en.ref &= REF_MASK;
            nchk += 1; nbytes += 12 + 4 * (en.szfl & SIZE_MASK);
            test_candidate<KIND>(en.ref, en.szfl, en.sig, me, S, sn, ssig, nx, -1, hdr, lits, r, sink);
        }
    }
    // clauses younger than the CSR (resolvents): brute force.  A short overflow list is walked here; a long one is
    // left to k_scan_ovf, which spreads (request, chunk of the list) pairs over the whole grid
    for (uint32_t i = lane; i < novf; i += 32) {
        const uint32_t ref = ovf[i];
        const ClauseHdr h = hdr[ref];
        if (ref != me) { nchk += 1; nbytes += 12 + 4 * (h.szfl & SIZE_MASK); }
        test_candidate<KIND>(ref, h.szfl, h.sig, me, S, sn, ssig, -1, -1, hdr, lits, r, sink);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { nchk += __shfl_xor_sync(0xffffffffu, nchk, d); nbytes += __shfl_xor_sync(0xffffffffu, nbytes, d); }
    if (lane == 0 && nchk) { atomicAdd(checks, nchk); atomicAdd(checks + 1, nbytes); }
    // (one same-address atomic pair per request: fine for the small incremental batches this kernel serves;
    //  full-queue passes go through k_full, which reduces per block)
    }
    sink.flush(&sink_base);
}

// Overflow part of a request scan for long overflow lists: one warp per (request, 512-entry chunk of the list), so
// the brute force over the young clauses is spread over the grid instead of being one warp's serial loop of
// dependent header gathers (260 us per launch at 21 k young clauses).
static constexpr uint32_t OVF_CHUNK = 512, OVF_INLINE_MAX = 1024;
template <int KIND>
__global__ void __launch_bounds__(256) k_scan_ovf(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits,
        const uint32_t* __restrict__ ovf, uint32_t novf, uint32_t nchunks,
        uint32_t req_begin, uint32_t req_end, const uint32_t* __restrict__ self,
        const uint32_t* __restrict__ req_off, const int32_t* __restrict__ req_lits,
        cp3g_hit* __restrict__ out, uint32_t cap, uint32_t* __restrict__ nout,
        unsigned long long* __restrict__ checks)
{
    SINK_DECL;
    const unsigned long long widx = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = req_begin + (uint32_t)(widx / nchunks), chunk = (uint32_t)(widx % nchunks);
    bool live = r < req_end;
    uint32_t me = 0; const int32_t* S = nullptr; uint32_t sn = 0;
    if (live) {
        me = self[r];
        if (req_lits != nullptr) { S = req_lits + req_off[r]; sn = req_off[r + 1] - req_off[r]; }
        else {
            const ClauseHdr h = hdr[me];
            if (h.szfl & FLAG_DEL) { live = false; }
            S = lits + h.start; sn = h.szfl & SIZE_MASK;
        }
        if (sn == 0) { live = false; }
    }
    if (live) {
        unsigned long long ssig = 0;
        for (uint32_t k = lane; k < sn; k += 32) { ssig |= lit_bit(S[k]); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { ssig |= __shfl_xor_sync(0xffffffffu, ssig, d); }
        unsigned long long nchk = 0, nbytes = 0;
        const uint32_t b = chunk * OVF_CHUNK, e = min(novf, b + OVF_CHUNK);
        for (uint32_t i = b + lane; i < e; i += 32) {
            const uint32_t ref = ovf[i];
            const ClauseHdr h = hdr[ref];
            if (ref != me) { nchk += 1; nbytes += 12 + 4 * (h.szfl & SIZE_MASK); }
            test_candidate<KIND>(ref, h.szfl, h.sig, me, S, sn, ssig, -1, -1, hdr, lits, r, sink);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { nchk += __shfl_xor_sync(0xffffffffu, nchk, d); nbytes += __shfl_xor_sync(0xffffffffu, nbytes, d); }
        if (lane == 0 && nchk) { atomicAdd(checks, nchk); atomicAdd(checks + 1, nbytes); }
    }
    sink.flush(&sink_base);
}

// ------------------------------------------------------------------------------------------ K3 / K4 full queue
// mbarrier + bulk-copy (TMA engine, UBLKCP) helpers: one elected lane moves a contiguous tile global -> shared
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}

// Full-queue pass (every live clause is a subsumer / strengthener).  The clauses whose full-queue scan walks list x are
// exactly the entries of x with the matching class bit clear (k_occ_fill), and they sit at the front of the list - so
// the subsumer set of a candidate is a prefix of its own list and, for a short list, lies in the same tile.
//
// Warps are persistent and fully independent (no block barrier in the loop; tiles are dealt out by static striding -
// a global tile counter was measured and lost).  Each warp owns a double-buffered staging area:
// the tile (<= 159 entries x 16 B, whole list pairs, k_tiles) and the CSR offsets of its lists are moved global ->
// shared by the bulk-copy engine (cp.async.bulk + mbarrier) one tile ahead of the processing.
//   filter   pure tiles: one lane per entry emits (subsumer, <= 8 candidates) units into the warp's work list (slots by
//            one shared-memory atomic per tagged lane), then one lane per unit tests its chunk in straight-line code;
//            tiles inside one long list / cut by one: segment and entry forms.  Size + signature filters only;
//            survivors are queued
//   verify   the lanes verify one queued pair each (merge of the two literal arrays: 4 dependent random loads,
//            which would otherwise stall 31 lanes behind one); the deleted flag is read from the headers loaded
//            here.  The queue is drained at the end of a tile and, in the segment / entry forms, whenever it holds
//            32 pairs (skewed inputs produce thousands of survivors per tile)
//   hits are buffered per warp and reserved in the global list with one atomic per ~16 hits
static constexpr int WQ_CAP = CP3_WQ, WH_CAP = CP3_WH, FQ_WARPS = CP3_FQWARPS;
static constexpr uint32_t WORK_CAP = CP3_WORKCAP, DCHUNK = CP3_DCHUNK;     // flat work list per tile: (subsumer, <= 8 consecutive candidates)
struct TileDesc { uint32_t s, n, x0, xe, xa, noff; bool pure; };
__device__ __forceinline__ TileDesc make_desc(const uint2 a, const uint2 b)
{
    TileDesc d;
    d.s = a.x & ~TILE_ALIGNED; d.n = (b.x & ~TILE_ALIGNED) - d.s; d.x0 = a.y;
    d.xe = (b.x & TILE_ALIGNED) ? b.y : b.y + 1;                       // occ_off[xe] >= end of the tile
    d.xa = d.x0 & ~3u;                                                   // 16-byte aligned source for the bulk copy
    const uint32_t cnt = (d.xe - d.xa + 1 + 3) & ~3u;
    d.noff = cnt <= OFFCAP ? cnt : 0u;                                   // 0: too many (empty) lists, offsets stay global
    d.pure = (a.x & b.x & TILE_ALIGNED) && d.noff;                       // whole short list pairs, offsets staged
    return d;
}
template <int KIND>
__global__ void __launch_bounds__(FQ_WARPS * 32, CP3_MINBLOCKS) k_full(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits,
        const uint32_t* __restrict__ occ_off, const OccEnt* __restrict__ occ_ent, const uint2* __restrict__ tiles,
        uint32_t tile_first, uint32_t tile_last,
        cp3g_hit* __restrict__ out, uint32_t cap, uint32_t* __restrict__ nout, unsigned long long* __restrict__ checks)
{
    __shared__ __align__(16) OccEnt e_s[FQ_WARPS][2][TCAP + DCHUNK];      // + DCHUNK: a unit may read (and ignore) past its chunk
    __shared__ __align__(16) uint32_t o_s[FQ_WARPS][2][OFFCAP];
    __shared__ uint32_t vf_s[FQ_WARPS][TCAP + DCHUNK], work_s[FQ_WARPS][WORK_CAP];    // + DCHUNK: see e_s
    __shared__ __align__(8) unsigned long long bar[FQ_WARPS][2];
    __shared__ unsigned long long sh_chk[FQ_WARPS], sh_byt[FQ_WARPS];
    __shared__ uint32_t q_s[FQ_WARPS][WQ_CAP], q_d[FQ_WARPS][WQ_CAP]; __shared__ int32_t q_l[FQ_WARPS][WQ_CAP];
    __shared__ cp3g_hit h_buf[FQ_WARPS][WH_CAP];
    __shared__ uint32_t q_n[FQ_WARPS], h_n[FQ_WARPS], u_n[FQ_WARPS];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { q_n[w] = 0; h_n[w] = 0; u_n[w] = 0; mbar_init(&bar[w][0], 1); mbar_init(&bar[w][1], 1); }
    __syncwarp();
    unsigned long long nchk = 0, nbytes = 0;
    const uint32_t gwarp = blockIdx.x * FQ_WARPS + w, nwarps = gridDim.x * FQ_WARPS;

    auto verify = [&](uint32_t sref, uint32_t dref, int32_t l, int pass) {
        // pass 0: S holds l, any literal but ~l may be the flipped one; pass 1: S holds ~l, the flip is on l
        const ClauseHdr hs = hdr[sref], hd = hdr[dref];
        if ((hs.szfl | hd.szfl) & FLAG_DEL) { return; }
        const uint32_t sn = hs.szfl & SIZE_MASK, dn = hd.szfl & SIZE_MASK;
        if (dn < sn) { return; }
        int32_t f = -1; bool hit;
        if (KIND == CP3G_SUBSUME) { hit = merge_subset(lits + hs.start, sn, lits + hd.start, dn); }
        else {
            f = merge_one_flip(lits + hs.start, sn, lits + hd.start, dn);
            hit = f >= 0 && (pass == 0 ? f != (l ^ 1) : f == l);
        }
        if (hit) {
            const uint32_t p = atomicAdd(&h_n[w], 1u);
            if (p < (uint32_t)WH_CAP) { h_buf[w][p].req = sref; h_buf[w][p].clause = dref; h_buf[w][p].lit = f; }
            else { const uint32_t g2 = atomicAdd(nout, 1u); if (g2 < cap) { out[g2].req = sref; out[g2].clause = dref; out[g2].lit = f; } }
        }
    };
    auto flush_hits = [&]() {
        __syncwarp();
        const uint32_t n = min(h_n[w], (uint32_t)WH_CAP);
        uint32_t base = 0;
        if (lane == 0 && n) { base = atomicAdd(nout, n); }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < n && base + lane < cap) { out[base + lane] = h_buf[w][lane]; }
        __syncwarp();
        if (lane == 0) { h_n[w] = 0; }
        __syncwarp();
    };
    // Warp-wide verification of the queued pairs, one pair per lane: the dependent random loads of a pair (two headers,
    // two literal arrays) overlap with those of 31 others.  Every lane calls it; nothing happens below `at_least` pairs.
    auto drain = [&](uint32_t at_least) {
        __syncwarp();
        const uint32_t qn = min(q_n[w], (uint32_t)WQ_CAP);
        if (qn < at_least) { return; }
        for (uint32_t k = lane; k < qn; k += 32) { verify(q_s[w][k], q_d[w][k], q_l[w][k] >> 1, q_l[w][k] & 1); }
        __syncwarp();
        if (lane == 0) { q_n[w] = 0; }
        __syncwarp();
        if (h_n[w] >= (uint32_t)(WH_CAP / 2)) { flush_hits(); }
    };
    auto issue = [&](const TileDesc& d, int b) {          // lane 0 only
        const uint32_t bytes_e = d.n * (uint32_t)sizeof(OccEnt), bytes_o = d.noff * 4u;
        mbar_expect_tx(&bar[w][b], bytes_e + bytes_o);
        bulk_g2s(&e_s[w][b][0], occ_ent + d.s, bytes_e, &bar[w][b]);
        if (bytes_o) { bulk_g2s(&o_s[w][b][0], occ_off + d.xa, bytes_o, &bar[w][b]); }
    };

    uint32_t t = tile_first + gwarp, t1 = t + nwarps, t2 = t1 + nwarps;
    if (t < tile_last) {
        TileDesc cur = make_desc(tiles[t], tiles[t + 1]), nxt = cur;
        if (t1 < tile_last) { nxt = make_desc(tiles[t1], tiles[t1 + 1]); }
        if (lane == 0 && cur.n) { issue(cur, 0); }
        uint32_t phase = 0;                                  // bit b = parity the next fill of buffer b completes
        for (uint32_t it = 0; t < tile_last; ++it) {
            const int b = it & 1;
            const bool more = t1 < tile_last;
            if (lane == 0 && more && nxt.n) { issue(nxt, b ^ 1); }
            TileDesc nn = nxt;
            if (t2 < tile_last) { nn = make_desc(tiles[t2], tiles[t2 + 1]); }
            if (cur.n) {
                mbar_wait(&bar[w][b], (phase >> b) & 1u); phase ^= 1u << b;
                const OccEnt* es = e_s[w][b]; const uint32_t* os = o_s[w][b];
                const uint32_t s = cur.s, n = cur.n, xa = cur.xa, noff = cur.noff;
                // A tile is pure iff both of its cuts are aligned to list pairs and its offsets are staged: then every
                // list that meets the tile lies in it together with its complementary list.
                //
                // The pairs to test are (tagged entry S of list x) x (entries D of list x, and of list x^1 for
                // strengthening), D in this tile.  Three forms:
                //  flat     (pure tiles - all of them unless the literal frequencies are skewed) One lane per ENTRY first:
                //           a tagged entry emits work units (S, <= 8 consecutive D) into the warp's list; then one lane
                //           per UNIT keeps S in registers and tests its chunk from shared memory in straight-line code.
                //           Every lane does the same amount of work, whatever the tag counts of the lists are.
                //  segment  (tiles that lie inside ONE long list) 32 tagged entries of that list / its complement at a
                //           time, coalesced from global memory, one per lane; each walks the whole tile, all lanes
                //           reading the same candidate (shared-memory broadcast).
                //  entry    (the rest: tiles cut by a long list, work list overflow) one lane per candidate D walks the
                //           tagged prefix of its list, every access picking shared or global memory.
                const bool pure = cur.pure;
                uint32_t* vf = vf_s[w]; uint32_t* work = work_s[w];
                auto OFF = [&](uint32_t y) -> uint32_t { return (y - xa < noff) ? os[y - xa] : occ_off[y]; };
                auto enqueue = [&](uint32_t sref, uint32_t dref, uint32_t x, uint32_t pass) {
                    const uint32_t q = atomicAdd(&q_n[w], 1u);
                    if (q < (uint32_t)WQ_CAP) { q_s[w][q] = sref; q_d[w][q] = dref; q_l[w][q] = (int32_t)((x << 1) | pass); }
                    else { verify(sref, dref, (int32_t)x, (int)pass); }
                };
                auto entry_form = [&]() {
                    const int npass = KIND == CP3G_STRENGTHEN ? 2 : 1;
                    for (uint32_t base = 0; base < n; base += 32) {          // warp-uniform: the queue is drained in between
                        const uint32_t p = base + lane;
                        const bool act = p < n;
                        OccEnt en; en.ref = 0; en.szfl = 0; en.sig = 0;
                        uint32_t x = cur.x0;
                        if (act) {
                            en = es[p];
                            const uint32_t dx = en.szfl >> 24;
                            x = cur.x0 + dx;
                            if (dx == 255u) {                            // far from the tile's first literal (runs of empty lists)
                                uint32_t lo = x, hi = cur.xe;           // last x with off[x] <= s + p
                                while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (occ_off[mid] <= s + p) { lo = mid; } else { hi = mid; } }
                                x = lo;
                            }
                        }
                        const uint32_t dref = en.ref & REF_MASK, esz = en.szfl & SIZE_MASK;
                        const unsigned long long dboth = var_fold(en.sig) * 3ull;     // both polarity bits of every variable of D
                        uint32_t cnt = 0;
                        for (int pass = 0; pass < npass; ++pass) {
                            const uint32_t xx = pass == 0 ? x : (x ^ 1u);
                            uint32_t k = 0, ke = 0;
                            if (act) { k = OFF(xx); ke = OFF(xx + 1); }
                            while (__any_sync(0xffffffffu, k < ke)) {
                                const uint32_t stop = min(ke, k + 16u);
                                for (; k < stop; ++k) {
                                    const bool in = k - s < n;
                                    const uint2 a = in ? *reinterpret_cast<const uint2*>(&es[k - s]) : *reinterpret_cast<const uint2*>(&occ_ent[k]);
                                    const uint32_t cls = a.x >> 30;
                                    // the tagged classes are a prefix of the list
                                    if (KIND == CP3G_SUBSUME) { if (cls == CLS_NONE) { k = ke; break; } if (cls & CLS_NOSUB) { continue; } }
                                    else { if (cls & CLS_NOSTR) { k = ke; break; } }
                                    const uint32_t sref = a.x & REF_MASK;
                                    if (sref == dref) { continue; }
                                    cnt += 1;
                                    if (esz < (a.y & SIZE_MASK)) { continue; }
                                    const unsigned long long ssig = in ? es[k - s].sig : occ_ent[k].sig;
                                    const unsigned long long miss = ssig & ~en.sig;
                                    if (KIND == CP3G_SUBSUME) { if (miss) { continue; } }
                                    else { if ((ssig & ~dboth) || (miss & (miss - 1))) { continue; } }
                                    enqueue(sref, dref, x, (uint32_t)pass);
                                }
                                drain(32u);
                            }
                        }
                        nchk += cnt; nbytes += (unsigned long long)cnt * (12 + 4 * esz);
                    }
                };
                if (pure) {
                    const uint32_t xb = cur.x0 - xa;
                    uint32_t U = 0; bool fits = true;
                    // one lane per entry: a tagged entry S emits its units, every entry gets its 32-bit variable signature
#pragma unroll 2
                    for (uint32_t base = 0; base < n; base += 32) {
                        const uint32_t p = base + lane;
                        uint32_t cnt = 0, kb = 0, len = 0, kb1 = 0, len1 = 0;
                        if (p < n) {
                            const OccEnt en = es[p];
                            const unsigned long long f = var_fold(en.sig);
                            vf[p] = (uint32_t)f | ((uint32_t)(f >> 32) << 1);          // exact packing of the folded signature
                            if (!((en.ref >> 30) & (KIND == CP3G_SUBSUME ? CLS_NOSUB : CLS_NOSTR))) {
                                const uint32_t xi = xb + (en.szfl >> 24);
                                kb = os[xi] - s; len = os[xi + 1] - os[xi];
                                cnt = (len + DCHUNK - 1) / DCHUNK;
                                if (KIND == CP3G_STRENGTHEN) {
                                    const uint32_t xj = xi ^ 1u;                 // the complementary list (tile starts at an even literal)
                                    kb1 = os[xj] - s; len1 = os[xj + 1] - os[xj];
                                    cnt += (len1 + DCHUNK - 1) / DCHUNK;
                                }
                            }
                        }
                        // the order of the units is irrelevant: every tagged lane reserves its slots with one shared-memory atomic
                        if (cnt) {
                            uint32_t o = atomicAdd(&u_n[w], cnt);
                            if (o + cnt <= WORK_CAP) {
                                for (uint32_t c = 0; c < len; c += DCHUNK) { work[o++] = p | ((kb + c) << 8) | (min(DCHUNK, len - c) << 16); }
                                if (KIND == CP3G_STRENGTHEN) {
                                    for (uint32_t c = 0; c < len1; c += DCHUNK) { work[o++] = p | ((kb1 + c) << 8) | (min(DCHUNK, len1 - c) << 16) | (1u << 20); }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    U = u_n[w]; fits = U <= WORK_CAP;
                    __syncwarp();
                    if (lane == 0) { u_n[w] = 0; }
                    if (fits) {
                        for (uint32_t u = lane; u < U; u += 32) {
                            const uint32_t wk = work[u];
                            const uint32_t sidx = wk & 255u, d0 = (wk >> 8) & 255u, cnt = (wk >> 16) & 15u, pass = wk >> 20;
                            const OccEnt S = es[sidx];
                            const uint32_t sref = S.ref & REF_MASK, sn = S.szfl & SIZE_MASK, svf = vf[sidx];
                            uint32_t c = 0, szsum = 0, cand = 0;
                            // straight-line over the chunk: all loads are issued up front, one rare branch at the end
#pragma unroll
                            for (uint32_t i = 0; i < DCHUNK; ++i) {
                                const uint32_t d = d0 + i;
                                const uint2 dd = *reinterpret_cast<const uint2*>(&es[d]);
                                const uint32_t esz = dd.y & SIZE_MASK;
                                const bool live = i < cnt && (dd.x & REF_MASK) != sref;
                                c += live; szsum += live ? esz : 0u;
                                cand |= (uint32_t)(live && esz >= sn && (svf & ~vf[d]) == 0u) << i;
                            }
                            nchk += c; nbytes += 12ull * c + 4ull * szsum;
                            while (cand) {
                                const uint32_t i = __ffs(cand) - 1; cand &= cand - 1;
                                const uint32_t d = d0 + i;
                                const uint2 dd = *reinterpret_cast<const uint2*>(&es[d]);
                                const unsigned long long miss = S.sig & ~es[d].sig;
                                if (KIND == CP3G_SUBSUME ? miss != 0 : (miss & (miss - 1)) != 0) { continue; }
                                enqueue(sref, dd.x & REF_MASK, cur.x0 + (dd.y >> 24), pass);     // pure tile: dx is exact
                            }
                        }
                    } else {
                        entry_form();
                    }
                } else {
                    // the list of the first entry (uniform over the warp); does it cover the whole tile?
                    const uint32_t dx0 = es[0].szfl >> 24;
                    uint32_t y = cur.x0 + dx0;
                    if (dx0 == 255u) {
                        uint32_t lo = y, hi = cur.xe;
                        while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (occ_off[mid] <= s) { lo = mid; } else { hi = mid; } }
                        y = lo;
                    }
                    const uint32_t lb = OFF(y), le = OFF(y + 1);
                    if (le < s + n) { entry_form(); }
                    else {
                        for (uint32_t p = lane; p < n; p += 32) {
                            const unsigned long long f = var_fold(es[p].sig);
                            vf[p] = (uint32_t)f | ((uint32_t)(f >> 32) << 1);
                        }
                        __syncwarp();
                        const int npass = KIND == CP3G_STRENGTHEN ? 2 : 1;
                        for (int pass = 0; pass < npass; ++pass) {
                            const uint32_t sl = pass == 0 ? y : (y ^ 1u);
                            const uint32_t sb = pass == 0 ? lb : OFF(sl), se = pass == 0 ? le : OFF(sl + 1);
                            for (uint32_t k0 = sb; k0 < se; k0 += 32) {
                                const uint32_t k = k0 + lane;
                                OccEnt S; S.ref = 0xffffffffu; S.szfl = 0; S.sig = 0;
                                if (k < se) { if (k - s < n) { S = es[k - s]; } else { S = occ_ent[k]; } }
                                const uint32_t cls = S.ref >> 30;
                                const bool tagged = !(cls & (KIND == CP3G_SUBSUME ? CLS_NOSUB : CLS_NOSTR));
                                const uint32_t sref = S.ref & REF_MASK, sn = S.szfl & SIZE_MASK;
                                const unsigned long long sf = var_fold(S.sig);
                                const uint32_t svf = (uint32_t)sf | ((uint32_t)(sf >> 32) << 1);
                                uint32_t c = 0; unsigned long long szsum = 0;
                                for (uint32_t db = 0; db < n; db += 16) {             // 16 candidates, then the queue is looked at
                                    if (tagged) {
                                        const uint32_t de = min(db + 16u, n);
                                        for (uint32_t d = db; d < de; ++d) {
                                            const uint2 dd = *reinterpret_cast<const uint2*>(&es[d]);
                                            const uint32_t dref = dd.x & REF_MASK, esz = dd.y & SIZE_MASK;
                                            const bool live = dref != sref;
                                            c += live; szsum += live ? esz : 0u;
                                            if (live && esz >= sn && (svf & ~vf[d]) == 0u) {
                                                const unsigned long long miss = S.sig & ~es[d].sig;
                                                if (KIND == CP3G_SUBSUME ? miss != 0 : (miss & (miss - 1)) != 0) { continue; }
                                                enqueue(sref, dref, y, (uint32_t)pass);
                                            }
                                        }
                                    }
                                    drain(32u);
                                }
                                nchk += c; nbytes += 12ull * c + 4ull * szsum;
                                // the tagged classes are a prefix (strengtheners) resp. everything below class 3 (subsumers)
                                const bool more_s = KIND == CP3G_STRENGTHEN ? tagged : (cls != CLS_NONE);
                                if (!__any_sync(0xffffffffu, more_s)) { break; }
                            }
                        }
                    }
                }
            }
            __syncwarp();                                  // all reads of buffer b are done before it is refilled
            drain(more ? 16u : 1u);                        // verify when the queue is worth a warp pass
            cur = nxt; nxt = nn; t = t1; t1 = t2; t2 += nwarps;
        }
    }
    flush_hits();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { nchk += __shfl_xor_sync(0xffffffffu, nchk, d); nbytes += __shfl_xor_sync(0xffffffffu, nbytes, d); }
    if (lane == 0) { sh_chk[w] = nchk; sh_byt[w] = nbytes; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0, b2 = 0;
        for (int k = 0; k < FQ_WARPS; ++k) { a += sh_chk[k]; b2 += sh_byt[k]; }
        if (a) { atomicAdd(checks, a); atomicAdd(checks + 1, b2); }
    }
}

// ------------------------------------------------------------------------------------------ K5 / K6
// resolvent size of (P containing +v, N containing -v), -1 for a tautology: merged walk that drops the
// pivot, counts distinct literals and stops at the first complementary neighbour pair (BVE.h:248-299)
__device__ __forceinline__ int resolvent_size(const int32_t* __restrict__ P, uint32_t pn,
                                              const int32_t* __restrict__ N, uint32_t nn, int32_t pv)
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
__global__ void k_bve_pairs(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t vfirst, uint32_t nv,
                            const uint32_t* __restrict__ vars, const uint32_t* __restrict__ p_off,
                            const uint32_t* __restrict__ p_refs, const uint32_t* __restrict__ n_off,
                            const uint32_t* __restrict__ n_refs, const unsigned long long* __restrict__ pair_off,
                            int16_t* __restrict__ sizes)
{
    // one warp per variable of this rank's range [vfirst, nv), lanes over the pos x neg pairs
    const uint32_t warp = vfirst + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= nv) { return; }
    const uint32_t pb = p_off[warp], np = p_off[warp + 1] - pb;
    const uint32_t nb = n_off[warp], nn = n_off[warp + 1] - nb;
    const int32_t pv = (int32_t)(vars[warp] << 1);
    const unsigned long long base = pair_off[warp];
    const uint32_t total = np * nn;
    for (uint32_t t = lane; t < total; t += 32) {
        const uint32_t a = t / nn, b = t - a * nn;
        const ClauseHdr hp = hdr[p_refs[pb + a]], hn = hdr[n_refs[nb + b]];
        int s;
        if ((hp.szfl | hn.szfl) & FLAG_DEL) { s = -2; }           // deleted parent: caller skips the pair
        else {
            s = resolvent_size(lits + hp.start, hp.szfl & SIZE_MASK, lits + hn.start, hn.szfl & SIZE_MASK, pv);
        }
        sizes[base + t] = (int16_t)s;
    }
}
__global__ void k_bve_resolve(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t first, uint32_t npairs,
                              const uint32_t* __restrict__ vars, const uint32_t* __restrict__ p_refs,
                              const uint32_t* __restrict__ n_refs, const unsigned long long* __restrict__ out_off,
                              int32_t* __restrict__ out_lits)
{
    const uint32_t t = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npairs) { return; }
    const ClauseHdr hp = hdr[p_refs[t]], hn = hdr[n_refs[t]];
    const int32_t* P = lits + hp.start; const int32_t* N = lits + hn.start;
    const uint32_t pn = hp.szfl & SIZE_MASK, nn = hn.szfl & SIZE_MASK;
    const int32_t pv = (int32_t)(vars[t] << 1), nv = pv ^ 1;
    int32_t* o = out_lits + out_off[t];
    const unsigned long long room = out_off[t + 1] - out_off[t];
    uint32_t i = 0, j = 0; unsigned long long r = 0; int32_t prev = -2;
    while (i < pn || j < nn) {
        int32_t c;
        if (i < pn && P[i] == pv) { ++i; continue; }
        if (j < nn && N[j] == nv) { ++j; continue; }
        if (j >= nn || (i < pn && P[i] < N[j])) { c = P[i++]; } else { c = N[j++]; }
        if (c == prev) { continue; }
        prev = c;
        if (r < room) { o[r] = c; }
        ++r;
    }
}

// ------------------------------------------------------------------------------------------ edits
__global__ void k_set_deleted(ClauseHdr* __restrict__ hdr, const uint32_t* __restrict__ ids, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { hdr[ids[i]].szfl |= FLAG_DEL; }
}
// host clause records are {start:u32, size:24|flags:8, x, y}; flag bit 0 = deleted
__global__ void k_make_hdr(const uint4* __restrict__ rec, ClauseHdr* __restrict__ hdr, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const uint4 r = rec[i];
    ClauseHdr h; h.start = r.x; h.szfl = (r.y & SIZE_MASK) | (((r.y >> 24) & 1u) ? FLAG_DEL : 0u); h.sig = 0;
    hdr[i] = h;
}
__global__ void k_gather_refs(const OccEnt* __restrict__ ent, uint32_t* __restrict__ refs, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { refs[i] = ent[i].ref & REF_MASK; }
}
__global__ void k_shrink(ClauseHdr* __restrict__ hdr, int32_t* __restrict__ lits, const uint32_t* __restrict__ ids,
                         const uint32_t* __restrict__ nstart, const uint32_t* __restrict__ nsize,
                         const int32_t* __restrict__ nlits, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    ClauseHdr h = hdr[ids[i]];
    unsigned long long s = 0;
    for (uint32_t k = 0; k < nsize[i]; ++k) { const int32_t x = nlits[nstart[i] + k]; lits[h.start + k] = x; s |= lit_bit(x); }
    h.szfl = (h.szfl & ~SIZE_MASK) | nsize[i];
    h.sig = s;
    hdr[ids[i]] = h;
}

// ------------------------------------------------------------------------------------------ K0 bulk ingest
static constexpr uint32_t ING_MAX_CLAUSE = 4096, ING_LOCAL = 32;
// clause delimiters of the DIMACS stream + largest variable
__global__ void k_ing_mark(const int32_t* __restrict__ S, uint32_t n, uint32_t* __restrict__ flag, uint32_t* __restrict__ maxvar)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = 0;
    if (i < n) { const int32_t x = S[i]; flag[i] = x == 0 ? 1u : 0u; v = (uint32_t)(x < 0 ? -x : x); }
    else if (i == n) { flag[i] = 0u; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { v = max(v, __shfl_xor_sync(0xffffffffu, v, d)); }
    // same-address atomics serialise (0.4 ms for 575 k warps): only a warp that would raise the maximum issues one
    if ((threadIdx.x & 31) == 0 && v > *reinterpret_cast<volatile uint32_t*>(maxvar)) { atomicMax(maxvar, v); }
}
__global__ void k_ing_ends(const int32_t* __restrict__ S, uint32_t n, const uint32_t* __restrict__ cid, uint32_t* __restrict__ cend)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && S[i] == 0) { cend[cid[i]] = i; }
}
// Solver::addClause_ (Solver.cc:409-499) without assignments: sort, drop duplicates, detect tautologies.  One thread
// per raw clause; the normalised literals are left at the clause's own stream positions in `codes`.
__global__ void k_ing_norm(const int32_t* __restrict__ S, const uint32_t* __restrict__ cend, uint32_t nraw,
                           int32_t* __restrict__ codes, uint32_t* __restrict__ rsize, uint32_t* __restrict__ keep, uint32_t* __restrict__ status)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nraw) { return; }
    if (c == nraw) { rsize[c] = 0; keep[c] = 0; return; }
    const uint32_t b = c ? cend[c - 1] + 1 : 0u, e = cend[c], len = e - b;
    rsize[c] = 0; keep[c] = 0;
    if (len == 0 || len > ING_MAX_CLAUSE) { *status = 1u; return; }
    auto code = [&](uint32_t i) -> int32_t { const int32_t x = S[b + i]; return x < 0 ? 2 * (-x - 1) + 1 : 2 * (x - 1); };
    uint32_t j = 0; bool taut = false;
    if (len <= ING_LOCAL) {
        int32_t a[ING_LOCAL];
        for (uint32_t i = 0; i < len; ++i) {
            const int32_t key = code(i); uint32_t k = i;
            while (k > 0 && a[k - 1] > key) { a[k] = a[k - 1]; --k; }
            a[k] = key;
        }
        int32_t p = -2;
        for (uint32_t i = 0; i < len; ++i) {
            if (a[i] == (p ^ 1) && p >= 0) { taut = true; break; }
            if (a[i] != p) { p = a[i]; codes[b + j++] = p; }
        }
    } else {
        int32_t* a = codes + b;
        for (uint32_t i = 0; i < len; ++i) {
            const int32_t key = code(i); uint32_t k = i;
            while (k > 0 && a[k - 1] > key) { a[k] = a[k - 1]; --k; }
            a[k] = key;
        }
        int32_t p = -2;
        for (uint32_t i = 0; i < len; ++i) {
            const int32_t x = a[i];
            if (x == (p ^ 1) && p >= 0) { taut = true; break; }
            if (x != p) { p = x; a[j++] = p; }
        }
    }
    if (taut) { return; }                      // dropped, like addClause_ returning true early
    if (j == 1) { *status = 1u; return; }
    rsize[c] = j; keep[c] = 1u;
}
__global__ void k_ing_pack(const uint32_t* __restrict__ cend, uint32_t nraw, const int32_t* __restrict__ codes,
                           const uint32_t* __restrict__ rsize, const uint32_t* __restrict__ kid, const uint32_t* __restrict__ nstart,
                           ClauseHdr* __restrict__ hdr, int32_t* __restrict__ lits)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nraw) { return; }
    const uint32_t sz = rsize[c];
    if (!sz) { return; }
    const uint32_t b = c ? cend[c - 1] + 1 : 0u, ns = nstart[c];
    for (uint32_t k = 0; k < sz; ++k) { lits[ns + k] = codes[b + k]; }
    ClauseHdr h; h.start = ns; h.szfl = sz; h.sig = 0;
    hdr[kid[c]] = h;
}
__global__ void k_arena_records(const ClauseHdr* __restrict__ hdr, uint32_t n, uint4* __restrict__ rec)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const ClauseHdr h = hdr[i];
    const uint32_t fl = (h.szfl & FLAG_DEL) ? 1u : 0u;
    rec[i] = make_uint4(h.start, (h.szfl & SIZE_MASK) | ((fl | 2u | 4u) << 24), 0u, 0u);
}

// ------------------------------------------------------------------------------------------ K8 arena compaction
// live size of every clause (0 for deleted) -> exclusive scan -> new pool offsets
__global__ void k_compact_sizes(const ClauseHdr* __restrict__ hdr, uint32_t n, uint32_t* __restrict__ sz)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) { sz[i] = (i < n && !(hdr[i].szfl & FLAG_DEL)) ? (hdr[i].szfl & SIZE_MASK) : 0u; }
}
// gather: one thread per clause moves its literals to the packed pool and rewrites the header; deleted clauses
// become empty (their literals are dropped)
__global__ void k_compact_move(ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ src, int32_t* __restrict__ dst,
                               const uint32_t* __restrict__ nstart, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    ClauseHdr h = hdr[i];
    const uint32_t ns = nstart[i];
    if (h.szfl & FLAG_DEL) { h.szfl = FLAG_DEL; }
    else { const uint32_t sz = h.szfl & SIZE_MASK; for (uint32_t k = 0; k < sz; ++k) { dst[ns + k] = src[h.start + k]; } }
    h.start = ns;
    hdr[i] = h;
}

// ------------------------------------------------------------------------------------------ K7 Dense
// variable v occurs in a live clause (countLiterals, Dense.h:92-121; only "occurs at all" matters)
__global__ void k_dense_mark(const ClauseHdr* __restrict__ hdr, const int32_t* __restrict__ lits, uint32_t n, uint32_t* __restrict__ keep)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const ClauseHdr h = hdr[i];
    if (h.szfl & FLAG_DEL) { return; }
    const uint32_t sz = h.szfl & SIZE_MASK;
    for (uint32_t k = 0; k < sz; ++k) { keep[((uint32_t)lits[h.start + k]) >> 1] = 1u; }      // same value from every writer
}
__global__ void k_dense_keep(const unsigned char* __restrict__ extra, uint32_t nvars, uint32_t* __restrict__ keep)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nvars && extra[v]) { keep[v] = 1u; }
}
// mapping[v] = exclusive prefix sum of keep[] where keep[v], UINT32_MAX otherwise (Dense.cc:49-87: var - diff)
__global__ void k_dense_map(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pre, uint32_t nvars, uint32_t* __restrict__ mapping)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nvars) { mapping[v] = keep[v] ? pre[v] : 0xffffffffu; }
}
// compressClauses, Dense.cc:406-450: literals of live clauses, polarity kept.  The map is monotone, so the
// clauses stay sorted.
__global__ void k_dense_rewrite(const ClauseHdr* __restrict__ hdr, int32_t* __restrict__ lits, uint32_t n, const uint32_t* __restrict__ mapping)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const ClauseHdr h = hdr[i];
    if (h.szfl & FLAG_DEL) { return; }
    const uint32_t sz = h.szfl & SIZE_MASK;
    for (uint32_t k = 0; k < sz; ++k) { const uint32_t x = (uint32_t)lits[h.start + k]; lits[h.start + k] = (int32_t)(2u * mapping[x >> 1] + (x & 1u)); }
}

// ------------------------------------------------------------------------------------------ host side (struct cp3g: gpu_db.h)
// Hits leave the scan kernels in warp-completion order; the host commit consumes them grouped by request and
// ordered by clause index (duplicate ties are decided by index), so they are sorted on the device before the copy.
struct HitLess { __device__ bool operator()(const cp3g_hit& a, const cp3g_hit& b) const { return a.req != b.req ? a.req < b.req : a.clause < b.clause; } };
static int sort_hits(cp3g* g, uint32_t n)
{
    size_t bytes = 0;
    if (cub::DeviceMergeSort::SortKeys(nullptr, bytes, g->hits.p, (int)n, HitLess(), g->stream) != cudaSuccess) { g->fail("cub sort size", cudaGetLastError()); return CP3G_ERR_CUDA; }
    RESERVE(g->sort_tmp, bytes + 16, false);
    cudaError_t e = cub::DeviceMergeSort::SortKeys(g->sort_tmp.p, bytes, g->hits.p, (int)n, HitLess(), g->stream);
    if (e != cudaSuccess) { g->fail("cub sort", e); return CP3G_ERR_CUDA; }
    g->launches += 2;
    return CP3G_OK;
}
static constexpr uint32_t SORT_ON_DEVICE_MIN = 2048;

// refs_tmp := the occurrence lists in the reference's order, plain ascending clause index (CoprocessorData::addClause order,
// CoprocessorTypes.h:812-833); the device lists themselves are ordered (class, clause index)
int dev_host_order_lists(cp3g* g)
{
    const uint32_t nlit = 2 * g->nvars;
    if (!g->total_occ) { return CP3G_OK; }
    RESERVE(g->refs_tmp, g->total_occ, false);
    k_gather_refs<<<grid_for(g->total_occ, 256), 256, 0, g->stream>>>(g->occ_ent.p, g->refs_tmp.p, g->total_occ);
    CK(cudaMemsetAsync(g->counters.p + 1, 0, sizeof(uint32_t), g->stream));
    k_occ_sort_short<<<grid_for(nlit, SORT_LISTS), SORT_LISTS, 0, g->stream>>>(g->occ_off.p, nlit, g->refs_tmp.p, g->long_lits.p, g->counters.p + 1);
    uint32_t nlong = 0;
    CK(cudaMemcpyAsync(&nlong, g->counters.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    if (nlong) { k_occ_sort_long<<<nlong, 1024, 0, g->stream>>>(g->occ_off.p, g->long_lits.p, g->refs_tmp.p); g->launches++; }
    g->launches += 2;
    return CP3G_OK;
}
int dev_set_deleted(cp3g* g, const uint32_t* dev_ids, uint32_t n)
{
    if (!n) { return CP3G_OK; }
    k_set_deleted<<<grid_for(n, 256), 256, 0, g->stream>>>(g->hdr.p, dev_ids, n);
    g->launches++;
    return CP3G_OK;
}

extern "C" {

int cp3g_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

cp3g* cp3g_create(int device)
{
    int n = cp3g_device_count();
    if (device < 0 || device >= n) {
        fprintf(stderr, "cp3b200: no usable CUDA device (%d visible, asked for %d) - there is no CPU fallback\n", n, device);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { return nullptr; }
    cp3g* g = new cp3g();
    g->device = device;
    if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) { delete g; return nullptr; }
    cudaEventCreate(&g->ev0); cudaEventCreate(&g->ev1);
    cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, device);
    {   // keep freed blocks in the default memory pool instead of returning them to the driver at every sync
        cudaMemPool_t pool = nullptr; unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
    }
    if (g->counters.reserve(16, false, g->stream)) { delete g; return nullptr; }
    // pinned counters + hit window: page-locking costs milliseconds, so handles hand their pair on (a client creates one
    // handle per CPinit)
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) { g->pin_cnt = g_pin_free.back().first; g->pin_hits = g_pin_free.back().second; g_pin_free.pop_back(); }
    }
    if (!g->pin_cnt && (cudaHostAlloc((void**)&g->pin_cnt, 64, cudaHostAllocPortable) != cudaSuccess || cudaHostAlloc((void**)&g->pin_hits, PIN_HITS * sizeof(cp3g_hit), cudaHostAllocPortable) != cudaSuccess)) { delete g; return nullptr; }
    return g;
}

void cp3g_destroy(cp3g* g)
{
    if (!g) { return; }
    cudaSetDevice(g->device);
    cudaStreamSynchronize(g->stream);
    g->lits.release(g->stream); g->lits2.release(g->stream); g->hdr.release(g->stream); g->occ_off.release(g->stream); g->occ_tmp.release(g->stream); g->occ_rank.release(g->stream); g->srt_v.release(g->stream); g->srt_k2.release(g->stream); g->srt_v2.release(g->stream); g->tile_sum.release(g->stream);
    g->long_lits.release(g->stream); g->ovf.release(g->stream); g->scratch_u32.release(g->stream); g->occ_ent.release(g->stream); g->tiles.release(g->stream); g->refs_tmp.release(g->stream);
    g->sort_tmp.release(g->stream); g->req_self.release(g->stream); g->req_off.release(g->stream); g->req_lits.release(g->stream); g->hits.release(g->stream); g->counters.release(g->stream);
    g->bu0.release(g->stream); g->bu1.release(g->stream); g->bu2.release(g->stream); g->bu3.release(g->stream); g->bu4.release(g->stream); g->bq0.release(g->stream);
    g->bs0.release(g->stream); g->bl0.release(g->stream);
    g->sp_da.release(g->stream); g->sp_db.release(g->stream); g->sp_chosen.release(g->stream); g->sp_dltime.release(g->stream); g->sp_dlcnt.release(g->stream); g->sp_dloff.release(g->stream);
    g->ev_recs.release(g->stream); g->ev_pst.release(g->stream); g->ev_nst.release(g->stream); g->ev_rsz.release(g->stream); g->ev_loff.release(g->stream);
    g->sp_dlfill.release(g->stream); g->sp_n.release(g->stream); g->sp_stale.release(g->stream); g->sp_dead.release(g->stream); g->sp_refs.release(g->stream);
    cudaStreamSynchronize(g->stream);
    cp3g_nccl_free(g->nccl);
    if (g->pin_cnt && g->pin_hits) {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (g_pin_free.size() < 16) { g_pin_free.push_back({g->pin_cnt, g->pin_hits}); g->pin_cnt = nullptr; g->pin_hits = nullptr; }
    }
    if (g->pin_cnt) { cudaFreeHost(g->pin_cnt); } if (g->pin_hits) { cudaFreeHost(g->pin_hits); }
    cudaEventDestroy(g->ev0); cudaEventDestroy(g->ev1);
    cudaStreamDestroy(g->stream);
    delete g;
}

const char* cp3g_last_error(const cp3g* g) { return g ? g->err.c_str() : "no device"; }

int cp3g_upload(cp3g* g, uint32_t nvars, uint32_t ncl, const int32_t* lits, size_t nlits,
                const uint32_t* start, const uint32_t* size, const uint8_t* flags)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    g->nvars = nvars; g->ncl = ncl; g->nlits = nlits; g->ncl_csr = 0; g->ovf_host.clear();
    RESERVE(g->lits, nlits + nlits / 4 + 1024, false);
    RESERVE(g->hdr, (size_t)ncl + ncl / 4 + 1024, false);
    CK(cudaMemcpyAsync(g->lits.p, lits, nlits * sizeof(int32_t), cudaMemcpyHostToDevice, g->stream));
    std::vector<ClauseHdr> h(ncl);
    for (uint32_t i = 0; i < ncl; ++i) {
        h[i].start = start[i];
        h[i].szfl = (size[i] & SIZE_MASK) | ((flags && (flags[i] & 1)) ? FLAG_DEL : 0u);
        h[i].sig = 0;
    }
    CK(cudaMemcpyAsync(g->hdr.p, h.data(), (size_t)ncl * sizeof(ClauseHdr), cudaMemcpyHostToDevice, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    g->lits.used = nlits; g->hdr.used = ncl;
    g->h2d += (int64_t)(nlits * sizeof(int32_t) + (size_t)ncl * sizeof(ClauseHdr));
    return CP3G_OK;
}

static int finish_upload(cp3g* g, uint32_t ncl, size_t nlits)
{
    CK(cudaStreamSynchronize(g->stream));
    g->lits.used = nlits; g->hdr.used = ncl;
    return CP3G_OK;
}

extern "C" int cp3g_upload_packed(cp3g* g, uint32_t nvars, uint32_t ncl, const int32_t* lits, size_t nlits, const void* records16)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    g->nvars = nvars; g->ncl = ncl; g->nlits = nlits; g->ncl_csr = 0; g->ovf_host.clear(); g->dirty_since_build = true;
    RESERVE(g->lits, nlits + nlits / 4 + 1024, false);
    RESERVE(g->hdr, (size_t)ncl + ncl / 4 + 1024, false);
    RESERVE(g->occ_ent, (size_t)ncl + 1, false);                 // staging for the raw records (rebuilt by cp3g_build anyway)
    { int r = h2d_staged(g, g->lits.p, lits, nlits * sizeof(int32_t)); if (r != CP3G_OK) { return r; } }
    { int r = h2d_staged(g, g->occ_ent.p, records16, (size_t)ncl * 16); if (r != CP3G_OK) { return r; } }
    if (ncl) { k_make_hdr<<<grid_for(ncl, 256), 256, 0, g->stream>>>(reinterpret_cast<const uint4*>(g->occ_ent.p), g->hdr.p, ncl); g->launches++; }
    g->h2d += (int64_t)(nlits * sizeof(int32_t) + (size_t)ncl * 16);
    return finish_upload(g, ncl, nlits);
}

static int build_csr(cp3g* g)
{
    // occurrence entries carry the clause index in 30 bits (two class bits on top): refuse larger data bases loudly
    if (g->ncl > REF_MASK || g->nvars >= (1u << 30)) { g->err = "data base too large for the 30-bit clause references"; fprintf(stderr, "cp3b200: %s\n", g->err.c_str()); return CP3G_ERR_ARG; }
    const uint32_t nlit = 2 * g->nvars;
    RESERVE(g->occ_off, (size_t)nlit + 8, false);          // +8: the staged offset window of a tile is read in 16-byte units
    RESERVE(g->occ_tmp, (size_t)nlit + 2, false);
    const uint32_t ntiles = (nlit + 1 + SCAN_TILE - 1) / SCAN_TILE;
    RESERVE(g->tile_sum, (size_t)ntiles + 1, false);
    RESERVE(g->long_lits, (size_t)nlit + 1, false);
    CK(cudaMemsetAsync(g->occ_tmp.p, 0, ((size_t)nlit + 2) * sizeof(uint32_t), g->stream));
    CK(cudaMemsetAsync(g->counters.p, 0, 16 * sizeof(uint32_t), g->stream));
    // scatter path while the refs array stays L2 resident, sort path beyond (CP3_K2_SORT=0/1 forces one)
    static const int force_sort = getenv("CP3_K2_SORT") ? atoi(getenv("CP3_K2_SORT")) : -1;
    bool use_sort = force_sort >= 0 ? force_sort != 0 : g->nlits * sizeof(uint32_t) > (size_t)(96u << 20);
    if (nlit >= (1u << 30) || g->nlits >= (size_t)0x7fffffff || g->ncl == 0) { use_sort = false; }
    if (use_sort) {
        k_check_monotone<<<grid_for(g->ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->ncl, g->counters.p + 3);
        uint32_t bad = 0;
        CK(cudaMemcpyAsync(&bad, g->counters.p + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
        CK(cudaStreamSynchronize(g->stream));
        g->launches += 1;
        if (bad) { use_sort = false; }
    }
    if (g->ncl) {
        if (!use_sort) { RESERVE(g->occ_rank, g->nlits + 1, false); }
        k_occ_count<<<grid_for(g->ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->ncl, g->occ_tmp.p, use_sort ? nullptr : g->occ_rank.p);
    }
    k_scan_tiles<<<ntiles, SCAN_T, 0, g->stream>>>(g->occ_tmp.p, g->occ_off.p, nlit + 1, g->tile_sum.p);
    k_scan_sums<<<1, 1024, 0, g->stream>>>(g->tile_sum.p, ntiles, g->counters.p + 2);
    k_scan_add<<<ntiles, SCAN_T, 0, g->stream>>>(g->occ_off.p, nlit + 1, g->tile_sum.p);
    g->launches += 4;
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, g->counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    g->total_occ = total;
    RESERVE(g->occ_ent, (size_t)total + 1, false);
    RESERVE(g->refs_tmp, (size_t)total + 1, false);
    bool sorted_by_radix = false;
    if (g->ncl && use_sort) {
        const size_t ns = g->nlits;
        RESERVE(g->occ_rank, ns + 1, false); RESERVE(g->srt_v, ns + 1, false); RESERVE(g->srt_k2, ns + 1, false); RESERVE(g->srt_v2, ns + 1, false);
        CK(cudaMemsetAsync(g->occ_rank.p, 0xff, ns * sizeof(uint32_t), g->stream));      // unused pool slots sort to the end
        k_occ_keys<<<grid_for(g->ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->ncl, g->occ_off.p, g->occ_rank.p, g->srt_v.p);
        int end_bit = 2; while (end_bit < 32 && (1ull << (end_bit - 2)) < (unsigned long long)nlit) { ++end_bit; }
        end_bit = 32;                                  // the 0xffffffff filler keys need every bit
        size_t bytes = 0;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, g->occ_rank.p, g->srt_k2.p, g->srt_v.p, g->srt_v2.p, (int)ns, 0, end_bit, g->stream);
        if (e != cudaSuccess) { g->fail("cub radix sort size", e); return CP3G_ERR_CUDA; }
        RESERVE(g->sort_tmp, bytes + 16, false);
        e = cub::DeviceRadixSort::SortPairs(g->sort_tmp.p, bytes, g->occ_rank.p, g->srt_k2.p, g->srt_v.p, g->srt_v2.p, (int)ns, 0, end_bit, g->stream);
        if (e != cudaSuccess) { g->fail("cub radix sort", e); return CP3G_ERR_CUDA; }
        if (total) { k_refs_from_sorted<<<grid_for(total, 256), 256, 0, g->stream>>>(g->srt_k2.p, g->srt_v2.p, total, g->refs_tmp.p); }
        g->launches += 6;
        sorted_by_radix = true;
    } else if (g->ncl) {
        k_occ_fill<<<grid_for(g->ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->ncl, g->occ_off.p, g->occ_rank.p, g->refs_tmp.p);
        k_occ_sort_short<<<grid_for(nlit, SORT_LISTS), SORT_LISTS, 0, g->stream>>>(g->occ_off.p, nlit, g->refs_tmp.p, g->long_lits.p, g->counters.p + 1);
        g->launches += 2;
    }
    // tile table of the full-queue kernels (independent of the sort)
    g->ntiles = (total + TILE - 1) / TILE;
    RESERVE(g->tiles, (size_t)g->ntiles + 3, false);
    k_tiles<<<grid_for((size_t)g->ntiles + 2, 256), 256, 0, g->stream>>>(g->occ_off.p, nlit, total, g->ntiles + 1, g->tiles.p);
    g->launches += 1;
    uint32_t nlong = 0;
    if (!sorted_by_radix) {
        CK(cudaMemcpyAsync(&nlong, g->counters.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
        CK(cudaStreamSynchronize(g->stream));
    }
    if (nlong) {
        k_occ_sort_long<<<nlong, 1024, 0, g->stream>>>(g->occ_off.p, g->long_lits.p, g->refs_tmp.p);
        g->launches += 1;
    }
    if (total) {
        k_occ_expand<<<grid_for(total, 256), 256, 0, g->stream>>>(g->hdr.p, g->refs_tmp.p, g->occ_ent.p, total, g->occ_off.p, g->tiles.p);
        g->launches += 1;
    }
    CK(cudaGetLastError());
    g->ncl_csr = g->ncl;
    g->ovf_host.clear();
    g->dirty_since_build = false;
    return CP3G_OK;
}

int cp3g_build(cp3g* g)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    KTimer t(g);
    int r = build_csr(g);
    t.stop();
    CK(cudaStreamSynchronize(g->stream));
    g->build_ns += t.collect(); g->build_count++;
    // algorithmic bytes of K1+K2: signature pass 16 + 4*#C per clause, list construction 2*4 per occurrence
    g->build_bytes += 16ll * (int64_t)g->ncl + 4ll * (int64_t)g->nlits + 8ll * (int64_t)g->total_occ;
    return r;
}

int cp3g_download_occ(cp3g* g, uint32_t* off, uint32_t* refs)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    const uint32_t nlit = 2 * g->nvars;
    CK(cudaMemcpyAsync(off, g->occ_off.p, ((size_t)nlit + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    g->d2h += (int64_t)(((size_t)nlit + 1) * sizeof(uint32_t));
    if (refs && g->total_occ) {
        int r = dev_host_order_lists(g);
        if (r != CP3G_OK) { return r; }
        r = d2h_staged(g, refs, g->refs_tmp.p, (size_t)g->total_occ * sizeof(uint32_t));
        if (r != CP3G_OK) { return r; }
        g->d2h += (int64_t)g->total_occ * 4;
    }
    return CP3G_OK;
}

int cp3g_set_deleted(cp3g* g, uint32_t n, const uint32_t* ids)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!n) { return CP3G_OK; }
    cudaSetDevice(g->device);
    RESERVE(g->scratch_u32, n, false);
    CK(cudaMemcpyAsync(g->scratch_u32.p, ids, (size_t)n * 4, cudaMemcpyHostToDevice, g->stream));
    k_set_deleted<<<grid_for(n, 256), 256, 0, g->stream>>>(g->hdr.p, g->scratch_u32.p, n);
    g->launches++; g->h2d += (int64_t)n * 4;        // stream ordered: the next scan on this stream sees it
    return CP3G_OK;
}

int cp3g_shrink(cp3g* g, uint32_t n, const uint32_t* ids, const uint32_t* nstart, const uint32_t* nsize,
                const int32_t* lits, size_t nlits)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!n) { return CP3G_OK; }
    cudaSetDevice(g->device);
    RESERVE(g->bu0, n, false); RESERVE(g->bu1, n, false); RESERVE(g->bu2, n, false); RESERVE(g->bl0, nlits + 1, false);
    CK(cudaMemcpyAsync(g->bu0.p, ids, (size_t)n * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu1.p, nstart, (size_t)n * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu2.p, nsize, (size_t)n * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bl0.p, lits, nlits * 4, cudaMemcpyHostToDevice, g->stream));
    g->dirty_since_build = true;
    k_shrink<<<grid_for(n, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->bu0.p, g->bu1.p, g->bu2.p, g->bl0.p, n);
    g->launches++; g->h2d += (int64_t)n * 12 + (int64_t)nlits * 4;
    return CP3G_OK;
}

int cp3g_append(cp3g* g, uint32_t n, const uint32_t* size, const int32_t* lits, size_t nlits)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!n) { return CP3G_OK; }
    cudaSetDevice(g->device);
    RESERVE(g->lits, g->nlits + nlits, true);
    RESERVE(g->hdr, (size_t)g->ncl + n, true);
    std::vector<ClauseHdr> h(n);
    size_t pos = g->nlits, src = 0;
    for (uint32_t i = 0; i < n; ++i) {
        h[i].start = (uint32_t)pos; h[i].szfl = size[i] & SIZE_MASK;
        unsigned long long s = 0;
        for (uint32_t k = 0; k < size[i]; ++k) { s |= lit_bit(lits[src + k]); }
        h[i].sig = s;
        pos += size[i]; src += size[i];
        g->ovf_host.push_back(g->ncl + i);
    }
    CK(cudaMemcpyAsync(g->lits.p + g->nlits, lits, nlits * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->hdr.p + g->ncl, h.data(), (size_t)n * sizeof(ClauseHdr), cudaMemcpyHostToDevice, g->stream));
    RESERVE(g->ovf, g->ovf_host.size(), false);
    CK(cudaMemcpyAsync(g->ovf.p, g->ovf_host.data(), g->ovf_host.size() * 4, cudaMemcpyHostToDevice, g->stream));
    g->nlits += nlits; g->ncl += n; g->lits.used = g->nlits; g->hdr.used = g->ncl;
    g->h2d += (int64_t)(nlits * 4 + (size_t)n * sizeof(ClauseHdr) + g->ovf_host.size() * 4);
    // many young clauses make the brute-force part of k_scan expensive: fold them into the CSR.  A rebuild costs
    // time proportional to the whole data base (20 ms at 54 M clauses), the brute force time proportional to the
    // overflow list per request - so the list may grow with the data base: ncl/256, within [2048, 65536].
    const size_t fold_at = std::min<size_t>(65536, std::max<size_t>(2048, g->ncl / 256));
    if (g->ovf_host.size() > fold_at) { return build_csr(g); }
    return CP3G_OK;
}

int cp3g_scan(cp3g* g, int kind, uint32_t nreq, const uint32_t* self, const uint32_t* req_off,
              const int32_t* req_lits, cp3g_hit* out, uint32_t cap, uint32_t* nout, uint64_t* checks)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (nout) { *nout = 0; }
    if (!nreq) { return CP3G_OK; }
    cudaSetDevice(g->device);
    RESERVE(g->req_self, nreq, false);
    CK(cudaMemcpyAsync(g->req_self.p, self, (size_t)nreq * 4, cudaMemcpyHostToDevice, g->stream));
    g->h2d += (int64_t)nreq * 4;
    size_t nl = 0;
    if (req_lits) {
        nl = req_off[nreq];
        RESERVE(g->req_off, (size_t)nreq + 1, false);
        RESERVE(g->req_lits, nl + 1, false);
        CK(cudaMemcpyAsync(g->req_off.p, req_off, ((size_t)nreq + 1) * 4, cudaMemcpyHostToDevice, g->stream));
        CK(cudaMemcpyAsync(g->req_lits.p, req_lits, nl * 4, cudaMemcpyHostToDevice, g->stream));
        g->h2d += (int64_t)(((size_t)nreq + 1) * 4 + nl * 4);
    }
    // contiguous request range of this rank (multi-GPU: results are all-gathered below)
    uint32_t rb = 0, re = nreq;
    if (g->world > 1) {
        const uint64_t per = ((uint64_t)nreq + g->world - 1) / g->world;
        rb = (uint32_t)std::min<uint64_t>(nreq, per * g->rank);
        re = (uint32_t)std::min<uint64_t>(nreq, per * (g->rank + 1));
    }
    RESERVE(g->hits, (size_t)cap + 1, false);
    CK(cudaMemsetAsync(g->counters.p, 0, 16 * sizeof(uint32_t), g->stream));
    unsigned long long* dchecks = (unsigned long long*)(g->counters.p + 4);
    KTimer t(g);
    if (re > rb) {
        const unsigned blocks = grid_for((size_t)(re - rb) * 32, 256);
        const int32_t* rl = req_lits ? g->req_lits.p : nullptr;
        const uint32_t novf = (uint32_t)g->ovf_host.size();
        const bool split = novf > OVF_INLINE_MAX;                 // long overflow list: its own grid-wide kernel
        const uint32_t novf_inline = split ? 0u : novf;
        if (kind == CP3G_SUBSUME) {
            k_scan<CP3G_SUBSUME><<<blocks, 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->occ_off.p, g->occ_ent.p, g->ovf.p,
                novf_inline, rb, re, g->req_self.p, g->req_off.p, rl, g->hits.p, cap, g->counters.p, dchecks);
        } else {
            k_scan<CP3G_STRENGTHEN><<<blocks, 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->occ_off.p, g->occ_ent.p, g->ovf.p,
                novf_inline, rb, re, g->req_self.p, g->req_off.p, rl, g->hits.p, cap, g->counters.p, dchecks);
        }
        g->launches++;
        if (split) {
            const uint32_t nchunks = (novf + OVF_CHUNK - 1) / OVF_CHUNK;
            const unsigned long long warps = (unsigned long long)(re - rb) * nchunks;
            const unsigned oblocks = (unsigned)((warps * 32 + 255) / 256);
            if (kind == CP3G_SUBSUME) {
                k_scan_ovf<CP3G_SUBSUME><<<oblocks, 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->ovf.p, novf, nchunks, rb, re,
                    g->req_self.p, g->req_off.p, rl, g->hits.p, cap, g->counters.p, dchecks);
            } else {
                k_scan_ovf<CP3G_STRENGTHEN><<<oblocks, 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->ovf.p, novf, nchunks, rb, re,
                    g->req_self.p, g->req_off.p, rl, g->hits.p, cap, g->counters.p, dchecks);
            }
            g->launches++;
        }
    }
    t.stop();
    uint32_t* cnt = g->pin_cnt;
    const uint32_t spec = (out != nullptr && g->world == 1) ? std::min<uint32_t>(cap, PIN_HITS) : 0u;
    CK(cudaMemcpyAsync(cnt, g->counters.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    if (spec) { CK(cudaMemcpyAsync(g->pin_hits, g->hits.p, (size_t)spec * sizeof(cp3g_hit), cudaMemcpyDeviceToHost, g->stream)); }
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    g->scan_ns += t.collect();
    uint32_t n = cnt[0];
    uint64_t chk, ab; memcpy(&chk, &cnt[4], 8); memcpy(&ab, &cnt[6], 8);
    g->scan_requests += (re - rb); g->scan_checks += (int64_t)chk; g->scan_alg_bytes += (int64_t)ab;
    if (out == nullptr) { if (nout) { *nout = n; } if (checks) { *checks = chk; } return CP3G_OK; }   // count only
    // sharded: the overflow decision is taken on the GATHERED counts, identically on every rank (a rank that returned early on
    // its local count would leave the others alone in the collective); `cap` is the same everywhere - replicated host state
    if (g->world == 1 && n > cap) { if (nout) { *nout = n; } return CP3G_ERR_OVERFLOW; }
    if (g->world > 1) {
        int rr = cp3g_nccl_gather_hits(g->nccl, g->stream, g->hits.p, n, cap, out, &n, &chk);
        if (rr != CP3G_OK) { if (nout) { *nout = n; } return rr; }
    } else if (n >= SORT_ON_DEVICE_MIN) {
        int rr = sort_hits(g, n);
        if (rr != CP3G_OK) { return rr; }
        CK(cudaMemcpyAsync(out, g->hits.p, (size_t)n * sizeof(cp3g_hit), cudaMemcpyDeviceToHost, g->stream));
        CK(cudaStreamSynchronize(g->stream));
    } else if (n) {
        memcpy(out, g->pin_hits, (size_t)std::min(n, spec) * sizeof(cp3g_hit));
        if (n > spec) {
            CK(cudaMemcpyAsync(out + spec, g->hits.p + spec, (size_t)(n - spec) * sizeof(cp3g_hit), cudaMemcpyDeviceToHost, g->stream));
            CK(cudaStreamSynchronize(g->stream));
        }
    }
    g->d2h += (int64_t)n * (int64_t)sizeof(cp3g_hit) + 32;
    if (nout) { *nout = n; }
    if (checks) { *checks = chk; }
    return CP3G_OK;
}

int cp3g_scan_all(cp3g* g, int kind, cp3g_hit* out, uint32_t cap, uint32_t* nout, uint64_t* checks)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (nout) { *nout = 0; }
    cudaSetDevice(g->device);
    static const bool dbg = getenv("CP3_DEBUG_PAR") != nullptr;
    auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double td0 = dbg ? tnow() : 0;
    auto dlap = [&](const char* w) { if (dbg) { double t2 = tnow(); fprintf(stderr, "    scan_all %s %.2f ms\n", w, t2 - td0); td0 = t2; } };
    if (g->dirty_since_build || !g->ovf_host.empty()) { int r = cp3g_build(g); if (r != CP3G_OK) { return r; } }
    dlap("build");
    if (!g->total_occ) { if (checks) { *checks = 0; } return CP3G_OK; }
    RESERVE(g->hits, (size_t)cap + 1, false);
    dlap("reserve hits");
    CK(cudaMemsetAsync(g->counters.p, 0, 16 * sizeof(uint32_t), g->stream));
    unsigned long long* dchecks = (unsigned long long*)(g->counters.p + 4);
    // multi-GPU: contiguous slice of the tile table per rank
    uint32_t tb = 0, te = g->ntiles;
    if (g->world > 1) {
        const uint64_t per = ((uint64_t)g->ntiles + g->world - 1) / g->world;
        tb = (uint32_t)std::min<uint64_t>(g->ntiles, per * g->rank);
        te = (uint32_t)std::min<uint64_t>(g->ntiles, per * (g->rank + 1));
    }
    KTimer t(g);
    if (te > tb) {
        // persistent grid: 5 blocks of 4 independent warps per SM (40 KB of staging + work lists per block), warps stride over tiles
        static const unsigned per_sm = getenv("CP3_FQ_BLOCKS") ? (unsigned)atoi(getenv("CP3_FQ_BLOCKS")) : (unsigned)CP3_MINBLOCKS;
        const unsigned blocks = std::min<unsigned>(grid_for(te - tb, FQ_WARPS), (unsigned)g->num_sms * per_sm);
        if (kind == CP3G_SUBSUME) {
            k_full<CP3G_SUBSUME><<<blocks, FQ_WARPS * 32, 0, g->stream>>>(g->hdr.p, g->lits.p, g->occ_off.p, g->occ_ent.p, g->tiles.p,
                tb, te, g->hits.p, cap, g->counters.p, dchecks);
        } else {
            k_full<CP3G_STRENGTHEN><<<blocks, FQ_WARPS * 32, 0, g->stream>>>(g->hdr.p, g->lits.p, g->occ_off.p, g->occ_ent.p, g->tiles.p,
                tb, te, g->hits.p, cap, g->counters.p, dchecks);
        }
        g->launches++;
    }
    t.stop();
    uint32_t cnt[8];
    CK(cudaMemcpyAsync(cnt, g->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    const int64_t full_ns = t.collect();
    g->scan_ns += full_ns;
    uint32_t n = cnt[0];
    uint64_t chk, ab; memcpy(&chk, &cnt[4], 8); memcpy(&ab, &cnt[6], 8);
    { const int kk = kind == CP3G_SUBSUME ? 0 : 1; g->full_ns[kk] += full_ns; g->full_bytes[kk] += (int64_t)ab; g->full_checks[kk] += (int64_t)chk; g->full_launches[kk]++; }
    g->scan_requests += g->ncl; g->scan_checks += (int64_t)chk; g->scan_alg_bytes += (int64_t)ab;
    dlap("kernel+sync");
    if (out == nullptr) { if (nout) { *nout = n; } if (checks) { *checks = chk; } return CP3G_OK; }
    // sharded: the overflow decision is taken on the GATHERED counts, identically on every rank (a rank that returned early on
    // its local count would leave the others alone in the collective); `cap` is the same everywhere - replicated host state
    if (g->world == 1 && n > cap) { if (nout) { *nout = n; } return CP3G_ERR_OVERFLOW; }
    if (g->world > 1) {
        int rr = cp3g_nccl_gather_hits(g->nccl, g->stream, g->hits.p, n, cap, out, &n, &chk);
        if (rr != CP3G_OK) { if (nout) { *nout = n; } return rr; }
    } else if (n) {
        if (n >= SORT_ON_DEVICE_MIN) { int rr = sort_hits(g, n); if (rr != CP3G_OK) { return rr; } }
        if (dbg) { cudaStreamSynchronize(g->stream); } dlap("sort");
        CK(cudaMemcpyAsync(out, g->hits.p, (size_t)n * sizeof(cp3g_hit), cudaMemcpyDeviceToHost, g->stream));
        CK(cudaStreamSynchronize(g->stream));
        dlap("copy");
    }
    g->d2h += (int64_t)n * (int64_t)sizeof(cp3g_hit) + 32;
    if (nout) { *nout = n; }
    if (checks) { *checks = chk; }
    return CP3G_OK;
}

int cp3g_bve_pairs(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, const uint32_t* p_refs,
                   const uint32_t* n_off, const uint32_t* n_refs, const uint64_t* pair_off, int16_t* sizes)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!nv) { return CP3G_OK; }
    cudaSetDevice(g->device);
    const size_t np = p_off[nv], nn = n_off[nv]; const uint64_t total = pair_off[nv];
    RESERVE(g->bu0, nv, false); RESERVE(g->bu1, (size_t)nv + 1, false); RESERVE(g->bu2, np + 1, false);
    RESERVE(g->bu3, (size_t)nv + 1, false); RESERVE(g->bu4, nn + 1, false); RESERVE(g->bq0, (size_t)nv + 1, false);
    RESERVE(g->bs0, total + 1, false);
    CK(cudaMemcpyAsync(g->bu0.p, vars, (size_t)nv * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu1.p, p_off, ((size_t)nv + 1) * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu2.p, p_refs, np * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu3.p, n_off, ((size_t)nv + 1) * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu4.p, n_refs, nn * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bq0.p, pair_off, ((size_t)nv + 1) * 8, cudaMemcpyHostToDevice, g->stream));
    // multi-GPU: contiguous variable ranges balanced by the number of (pos,neg) pairs; the size slices are
    // all-gathered in place so that every rank commits from the identical matrix
    uint32_t vb = 0, ve = nv; std::vector<size_t> boff;
    if (g->world > 1) {
        boff.assign((size_t)g->world + 1, 0);
        std::vector<uint32_t> cut((size_t)g->world + 1, nv);
        cut[0] = 0;
        for (int r = 1; r < g->world; ++r) {
            const uint64_t want = total * (uint64_t)r / (uint64_t)g->world;
            cut[r] = (uint32_t)(std::lower_bound(pair_off, pair_off + nv + 1, want) - pair_off);
            if (cut[r] > nv) { cut[r] = nv; }
            if (cut[r] < cut[r - 1]) { cut[r] = cut[r - 1]; }
        }
        for (int r = 0; r <= g->world; ++r) { boff[r] = (size_t)pair_off[cut[r]] * sizeof(int16_t); }
        vb = cut[g->rank]; ve = cut[g->rank + 1];
    }
    KTimer t(g);
    if (ve > vb) {
        k_bve_pairs<<<grid_for((size_t)(ve - vb) * 32, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, vb, ve, g->bu0.p, g->bu1.p, g->bu2.p,
                g->bu3.p, g->bu4.p, g->bq0.p, g->bs0.p);
    }
    if (g->world > 1 && cp3g_nccl_allgatherv_inplace(g->nccl, g->stream, g->bs0.p, boff.data()) != CP3G_OK) { g->err = "nccl all-gather-v (K5 sizes) failed"; return CP3G_ERR_CUDA; }
    t.stop();
    CK(cudaMemcpyAsync(sizes, g->bs0.p, total * sizeof(int16_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    t.collect();
    g->launches++; g->h2d += (int64_t)(nv * 20 + (np + nn) * 4); g->d2h += (int64_t)total * 2;
    return CP3G_OK;
}

int cp3g_bve_resolve(cp3g* g, uint32_t npairs, const uint32_t* vars, const uint32_t* p_refs, const uint32_t* n_refs,
                     const uint64_t* out_off, int32_t* out_lits, size_t nlits)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    if (!npairs) { return CP3G_OK; }
    cudaSetDevice(g->device);
    RESERVE(g->bu0, npairs, false); RESERVE(g->bu2, npairs, false); RESERVE(g->bu4, npairs, false);
    RESERVE(g->bq0, (size_t)npairs + 1, false); RESERVE(g->bl0, nlits + 1, false);
    CK(cudaMemcpyAsync(g->bu0.p, vars, (size_t)npairs * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu2.p, p_refs, (size_t)npairs * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bu4.p, n_refs, (size_t)npairs * 4, cudaMemcpyHostToDevice, g->stream));
    CK(cudaMemcpyAsync(g->bq0.p, out_off, ((size_t)npairs + 1) * 8, cudaMemcpyHostToDevice, g->stream));
    // multi-GPU: contiguous pair ranges balanced by resolvent literals, literal streams all-gathered in place
    uint32_t pb = 0, pe = npairs; std::vector<size_t> boff;
    if (g->world > 1) {
        boff.assign((size_t)g->world + 1, 0);
        std::vector<uint32_t> cut((size_t)g->world + 1, npairs);
        cut[0] = 0;
        for (int r = 1; r < g->world; ++r) {
            const uint64_t want = (uint64_t)nlits * (uint64_t)r / (uint64_t)g->world;
            cut[r] = (uint32_t)(std::lower_bound(out_off, out_off + npairs + 1, want) - out_off);
            if (cut[r] > npairs) { cut[r] = npairs; }
            if (cut[r] < cut[r - 1]) { cut[r] = cut[r - 1]; }
        }
        for (int r = 0; r <= g->world; ++r) { boff[r] = (size_t)out_off[cut[r]] * sizeof(int32_t); }
        pb = cut[g->rank]; pe = cut[g->rank + 1];
    }
    KTimer t(g);
    if (pe > pb) {
        k_bve_resolve<<<grid_for(pe - pb, 128), 128, 0, g->stream>>>(g->hdr.p, g->lits.p, pb, pe, g->bu0.p, g->bu2.p, g->bu4.p,
                g->bq0.p, g->bl0.p);
    }
    if (g->world > 1 && cp3g_nccl_allgatherv_inplace(g->nccl, g->stream, g->bl0.p, boff.data()) != CP3G_OK) { g->err = "nccl all-gather-v (K6 literals) failed"; return CP3G_ERR_CUDA; }
    t.stop();
    CK(cudaMemcpyAsync(out_lits, g->bl0.p, nlits * 4, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    t.collect();
    g->launches++; g->h2d += (int64_t)npairs * 20; g->d2h += (int64_t)nlits * 4;
    return CP3G_OK;
}

int cp3g_dense(cp3g* g, const uint8_t* keep, int frag_percent, uint32_t* mapping, uint32_t* nvars_new, int* compressed,
               int32_t* lits_out, size_t nlits)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    const uint32_t nv = g->nvars;
    if (compressed) { *compressed = 0; }
    if (nvars_new) { *nvars_new = nv; }
    if (nv == 0) { return CP3G_OK; }
    if (nlits != g->nlits) { g->err = "cp3g_dense: host and device literal pools differ in size"; return CP3G_ERR_ARG; }
    const uint32_t ntiles = (nv + 1 + SCAN_TILE - 1) / SCAN_TILE;
    RESERVE(g->bu0, (size_t)nv + 2, false); RESERVE(g->bu1, (size_t)nv + 2, false); RESERVE(g->bu2, (size_t)nv + 2, false);
    RESERVE(g->tile_sum, (size_t)ntiles + 1, false);
    RESERVE(g->scratch_u32, ((size_t)nv + 3) / 4 + 1, false);
    KTimer t(g);
    CK(cudaMemsetAsync(g->bu0.p, 0, ((size_t)nv + 2) * sizeof(uint32_t), g->stream));
    CK(cudaMemcpyAsync(g->scratch_u32.p, keep, nv, cudaMemcpyHostToDevice, g->stream));
    if (g->ncl) { k_dense_mark<<<grid_for(g->ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->ncl, g->bu0.p); g->launches++; }
    k_dense_keep<<<grid_for(nv, 256), 256, 0, g->stream>>>(reinterpret_cast<const unsigned char*>(g->scratch_u32.p), nv, g->bu0.p);
    k_scan_tiles<<<ntiles, SCAN_T, 0, g->stream>>>(g->bu0.p, g->bu1.p, nv + 1, g->tile_sum.p);
    k_scan_sums<<<1, 1024, 0, g->stream>>>(g->tile_sum.p, ntiles, g->counters.p + 2);
    k_scan_add<<<ntiles, SCAN_T, 0, g->stream>>>(g->bu1.p, nv + 1, g->tile_sum.p);
    k_dense_map<<<grid_for(nv, 256), 256, 0, g->stream>>>(g->bu0.p, g->bu1.p, nv, g->bu2.p);
    g->launches += 5;
    uint32_t kept = 0;
    CK(cudaMemcpyAsync(&kept, g->counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaMemcpyAsync(mapping, g->bu2.p, (size_t)nv * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    g->h2d += nv; g->d2h += (int64_t)nv * 4 + 4;
    const uint32_t diff = nv - kept;
    // isCompact, Dense.h:124-127 (integer division as there)
    const bool compact = diff == 0 || ((double)frag_percent > 1.0 + (double)((diff * 100u) / nv));
    if (!compact) {
        if (g->ncl) { k_dense_rewrite<<<grid_for(g->ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->ncl, g->bu2.p); g->launches++; }
        CK(cudaMemcpyAsync(lits_out, g->lits.p, nlits * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
        g->d2h += (int64_t)nlits * 4;
        g->nvars = kept; g->dirty_since_build = true;            // signatures and occurrence lists are void now
        if (compressed) { *compressed = 1; }
        if (nvars_new) { *nvars_new = kept; }
    }
    t.stop();
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    t.collect();
    return CP3G_OK;
}

// exclusive scan of n u32 values (in -> out) on the handle stream; the total lands in counters[2]
int scan_u32(cp3g* g, const uint32_t* in, uint32_t* out, size_t n)
{
    const uint32_t nt = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    RESERVE(g->tile_sum, (size_t)nt + 1, false);
    k_scan_tiles<<<nt, SCAN_T, 0, g->stream>>>(in, out, (uint32_t)n, g->tile_sum.p);
    k_scan_sums<<<1, 1024, 0, g->stream>>>(g->tile_sum.p, nt, g->counters.p + 2);
    k_scan_add<<<nt, SCAN_T, 0, g->stream>>>(out, (uint32_t)n, g->tile_sum.p);
    g->launches += 3;
    return CP3G_OK;
}

int cp3g_ingest(cp3g* g, const int32_t* stream, size_t n, uint32_t* nvars, uint32_t* nclauses, size_t* nlits, int* status)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    if (status) { *status = 1; }
    if (n == 0 || n >= (size_t)0xfffffff0u || stream[n - 1] != 0) { return CP3G_OK; }
    const uint32_t N = (uint32_t)n;
    DevBuf<int32_t>& S = g->bl0;                     // the stream; bu0 = flags / sizes, bu1 = scans, bu2 = clause ends
    RESERVE(S, n + 1, false); RESERVE(g->bu0, n + 2, false); RESERVE(g->bu1, n + 2, false);
    { int r = h2d_staged(g, S.p, stream, n * sizeof(int32_t)); if (r != CP3G_OK) { return r; } }
    g->h2d += (int64_t)n * 4;
    CK(cudaMemsetAsync(g->counters.p, 0, 16 * sizeof(uint32_t), g->stream));
    KTimer t(g);
    k_ing_mark<<<grid_for(n + 1, 256), 256, 0, g->stream>>>(S.p, N, g->bu0.p, g->counters.p + 5);
    if (scan_u32(g, g->bu0.p, g->bu1.p, n + 1) != CP3G_OK) { return CP3G_ERR_CUDA; }
    uint32_t head[8];
    CK(cudaMemcpyAsync(head, g->counters.p, sizeof(head), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    const uint32_t nraw = head[2], maxvar = head[5];
    RESERVE(g->bu2, (size_t)nraw + 2, false); RESERVE(g->bu3, (size_t)nraw + 2, false); RESERVE(g->bu4, (size_t)nraw + 2, false);
    RESERVE(g->lits2, n + 1, false);                 // normalised literals at their stream positions
    k_ing_ends<<<grid_for(n, 256), 256, 0, g->stream>>>(S.p, N, g->bu1.p, g->bu2.p);
    // bu0 := rsize, bu3 := keep
    k_ing_norm<<<grid_for((size_t)nraw + 1, 128), 128, 0, g->stream>>>(S.p, g->bu2.p, nraw, g->lits2.p, g->bu0.p, g->bu3.p, g->counters.p + 6);
    g->launches += 3;
    CK(cudaMemcpyAsync(head, g->counters.p, sizeof(head), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    if (head[6]) { t.stop(); CK(cudaStreamSynchronize(g->stream)); t.collect(); return CP3G_OK; }      // units / empty / huge clause: host ingest
    // clause ids (scan of keep -> bu4) and pool offsets (scan of rsize -> bu1)
    if (scan_u32(g, g->bu3.p, g->bu4.p, (size_t)nraw + 1) != CP3G_OK) { return CP3G_ERR_CUDA; }
    uint32_t kept = 0;
    CK(cudaMemcpyAsync(&kept, g->counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    if (scan_u32(g, g->bu0.p, g->bu1.p, (size_t)nraw + 1) != CP3G_OK) { return CP3G_ERR_CUDA; }
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, g->counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    g->nvars = maxvar; g->ncl = kept; g->nlits = total; g->ncl_csr = 0; g->ovf_host.clear(); g->dirty_since_build = true;
    RESERVE(g->lits, (size_t)total + total / 4 + 1024, false);
    RESERVE(g->hdr, (size_t)kept + kept / 4 + 1024, false);
    if (nraw) { k_ing_pack<<<grid_for(nraw, 256), 256, 0, g->stream>>>(g->bu2.p, nraw, g->lits2.p, g->bu0.p, g->bu4.p, g->bu1.p, g->hdr.p, g->lits.p); g->launches++; }
    t.stop();
    int r = finish_upload(g, kept, total);
    g->ingest_ns += t.collect(); g->ingest_bytes += 8ll * (int64_t)n + 16ll * (int64_t)g->ncl;        // 4 read + 4 written per stream literal, 16 per clause header
    if (r != CP3G_OK) { return r; }
    if (nvars) { *nvars = maxvar; } if (nclauses) { *nclauses = kept; } if (nlits) { *nlits = total; }
    if (status) { *status = 0; }
    return CP3G_OK;
}

int cp3g_download_arena(cp3g* g, int32_t* lits, void* records16)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    if (g->ncl) {
        RESERVE(g->refs_tmp, (size_t)g->ncl * 4 + 4, false);     // staging for the 16-byte records (a build-time temporary)
        k_arena_records<<<grid_for(g->ncl, 256), 256, 0, g->stream>>>(g->hdr.p, g->ncl, reinterpret_cast<uint4*>(g->refs_tmp.p));
        g->launches++;
        { int r = d2h_staged(g, records16, g->refs_tmp.p, (size_t)g->ncl * 16); if (r != CP3G_OK) { return r; } }
    }
    if (g->nlits) { int r = d2h_staged(g, lits, g->lits.p, g->nlits * sizeof(int32_t)); if (r != CP3G_OK) { return r; } }
    g->d2h += (int64_t)g->nlits * 4 + (int64_t)g->ncl * 16;
    return CP3G_OK;
}

int cp3g_compact(cp3g* g, uint32_t* new_start, int32_t* lits_out, size_t lits_cap, size_t* nlits_new)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    const uint32_t n = g->ncl;
    if (nlits_new) { *nlits_new = g->nlits; }
    if (n == 0) { return CP3G_OK; }
    const uint32_t ntiles = (n + 1 + SCAN_TILE - 1) / SCAN_TILE;
    RESERVE(g->bu0, (size_t)n + 2, false); RESERVE(g->bu1, (size_t)n + 2, false);
    RESERVE(g->tile_sum, (size_t)ntiles + 1, false);
    RESERVE(g->lits2, g->lits.cap, false);
    KTimer t(g);
    k_compact_sizes<<<grid_for((size_t)n + 1, 256), 256, 0, g->stream>>>(g->hdr.p, n, g->bu0.p);
    k_scan_tiles<<<ntiles, SCAN_T, 0, g->stream>>>(g->bu0.p, g->bu1.p, n + 1, g->tile_sum.p);
    k_scan_sums<<<1, 1024, 0, g->stream>>>(g->tile_sum.p, ntiles, g->counters.p + 2);
    k_scan_add<<<ntiles, SCAN_T, 0, g->stream>>>(g->bu1.p, n + 1, g->tile_sum.p);
    k_compact_move<<<grid_for(n, 256), 256, 0, g->stream>>>(g->hdr.p, g->lits.p, g->lits2.p, g->bu1.p, n);
    g->launches += 5;
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, g->counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    if (new_start) { CK(cudaMemcpyAsync(new_start, g->bu1.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream)); g->d2h += (int64_t)n * 4; }
    t.stop();
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    t.collect();
    std::swap(g->lits.p, g->lits2.p); std::swap(g->lits.cap, g->lits2.cap);
    g->nlits = total; g->lits.used = total;
    // the occurrence CSR stays valid: its entries name clauses, and every literal access goes through hdr[].start
    if (nlits_new) { *nlits_new = total; }
    if (lits_out) {
        if (lits_cap < total) { g->err = "cp3g_compact: output buffer too small"; return CP3G_ERR_ARG; }
        CK(cudaMemcpyAsync(lits_out, g->lits.p, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, g->stream));
        CK(cudaStreamSynchronize(g->stream));
        g->d2h += (int64_t)total * 4;
    }
    return CP3G_OK;
}

int cp3g_nccl_init(cp3g* g, int rank, int world, const void* id128)
{
    if (!g) { return CP3G_ERR_NODEVICE; }
    cudaSetDevice(g->device);
    if (world <= 1) { g->rank = 0; g->world = 1; return CP3G_OK; }
    void* c = cp3g_nccl_create(rank, world, id128);
    if (!c) { g->err = "ncclCommInitRank failed"; return CP3G_ERR_CUDA; }
    g->nccl = c; g->rank = rank; g->world = world;
    return CP3G_OK;
}

int64_t cp3g_counter(const cp3g* g, const char* name)
{
    if (!g) { return -1; }
    if (!strcmp(name, "launches")) { return g->launches; }
    if (!strcmp(name, "kernel_ns")) { return g->kernel_ns; }
    if (!strcmp(name, "h2d_bytes")) { return g->h2d; }
    if (!strcmp(name, "d2h_bytes")) { return g->d2h; }
    if (!strcmp(name, "scan_requests")) { return g->scan_requests; }
    if (!strcmp(name, "scan_checks")) { return g->scan_checks; }
    if (!strcmp(name, "scan_ns")) { return g->scan_ns; }
    if (!strcmp(name, "commit_ns")) { return g->commit_ns; }
    if (!strcmp(name, "commit_bytes")) { return g->commit_bytes; }
    if (!strcmp(name, "build_ns")) { return g->build_ns; }
    if (!strcmp(name, "build_bytes")) { return g->build_bytes; }
    if (!strcmp(name, "build_count")) { return g->build_count; }
    if (!strcmp(name, "ingest_ns")) { return g->ingest_ns; }
    if (!strcmp(name, "ingest_bytes")) { return g->ingest_bytes; }
    if (!strcmp(name, "bve_bytes")) { return g->bve_bytes; }
    if (!strncmp(name, "full", 4) && (name[4] == '0' || name[4] == '1') && name[5] == '_') {
        const int kk = name[4] - '0';
        if (!strcmp(name + 6, "ns")) { return g->full_ns[kk]; }
        if (!strcmp(name + 6, "bytes")) { return g->full_bytes[kk]; }
        if (!strcmp(name + 6, "checks")) { return g->full_checks[kk]; }
        if (!strcmp(name + 6, "launches")) { return g->full_launches[kk]; }
    }
    if (!strcmp(name, "bve_ns")) { return g->bve_ns; }
    if (!strcmp(name, "bce_ns")) { return g->bce_ns; }
    if (!strcmp(name, "bce_bytes")) { return g->bce_bytes; }
    if (!strcmp(name, "scan_alg_bytes")) { return g->scan_alg_bytes; }
    if (!strcmp(name, "nclauses")) { return g->ncl; }
    if (!strcmp(name, "nlits")) { return (int64_t)g->nlits; }
    if (!strcmp(name, "nvars")) { return g->nvars; }
    if (!strcmp(name, "total_occ")) { return g->total_occ; }
    return -1;
}

} // extern "C"
