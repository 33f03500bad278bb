// gpu_db.h - internals shared by the CUDA translation units (gpu_db.cu: data base + scan kernels, commit.cu: device-side
// pass commit) and nccl_comm.cpp.  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/cp3b200.h"

// growable device buffer (cudaMallocAsync on the handle stream; grows by 1.5x, optionally keeping the first `used` elements)
template <class T> struct DevBuf {
    T* p = nullptr; size_t cap = 0; size_t used = 0;
    int reserve(size_t n, bool keep, cudaStream_t st);
    void release(cudaStream_t st);
};

// NCCL plumbing (nccl_comm.cpp; libnccl is dlopen()ed so the library loads on hosts without it)
void* cp3g_nccl_create(int rank, int world, const void* id128);
void  cp3g_nccl_free(void* comm);
// all-gather-v of the ranks' hit lists (device buffer `dhits`, n local hits) into host `out`
int   cp3g_nccl_gather_hits(void* comm, cudaStream_t st, const cp3g_hit* dhits, uint32_t n, uint32_t cap,
                            cp3g_hit* out, uint32_t* nout, uint64_t* checks);
// the same exchange with the gathered list left on the device (*dout: communicator-owned buffer), for the device-side pass commit
int   cp3g_nccl_gather_hits_device(void* comm, cudaStream_t st, const cp3g_hit* dhits, uint32_t n, const cp3g_hit** dout, uint32_t* total, uint64_t* checks);
// in-place all-gather-v of a device buffer whose per-rank byte ranges [off[r], off[r+1]) are known to every rank
int   cp3g_nccl_allgatherv_inplace(void* comm, cudaStream_t st, void* dbuf, const size_t* off);

#ifdef __CUDACC__
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { g->fail(#call, _e); return CP3G_ERR_CUDA; } } while (0)

static constexpr uint32_t FLAG_DEL = 0x80000000u;
static constexpr uint32_t SIZE_MASK = 0x00ffffffu;
// occurrence entry: top two bits of `ref` = class of the clause in this list (see k_occ_fill)
static constexpr uint32_t REF_MASK = 0x3fffffffu;

struct __align__(16) ClauseHdr { uint32_t start; uint32_t szfl; unsigned long long sig; };
struct __align__(16) OccEnt { uint32_t ref; uint32_t szfl; unsigned long long sig; };

struct cp3g {
    int device = 0; int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    uint32_t nvars = 0, ncl = 0, ncl_csr = 0;
    size_t nlits = 0;
    DevBuf<int32_t> lits, lits2; DevBuf<ClauseHdr> hdr;
    DevBuf<uint32_t> occ_off, occ_tmp, occ_rank, srt_v, srt_k2, srt_v2, tile_sum, long_lits, ovf, scratch_u32;
    DevBuf<OccEnt> occ_ent; DevBuf<uint2> tiles; DevBuf<uint32_t> refs_tmp; uint32_t ntiles = 0; bool dirty_since_build = true;
    DevBuf<uint32_t> req_self, req_off; DevBuf<int32_t> req_lits;
    DevBuf<unsigned char> sort_tmp;
    DevBuf<cp3g_hit> hits; DevBuf<uint32_t> counters; // [0]=nout [1]=nlong [2]=total ; checks at +4 (u64)
    DevBuf<uint32_t> bu0, bu1, bu2, bu3, bu4; DevBuf<unsigned long long> bq0; DevBuf<int16_t> bs0; DevBuf<int32_t> bl0;
    // device-side pass commit (commit.cu)
    DevBuf<int32_t> sp_da, sp_db, sp_chosen, sp_dltime; DevBuf<uint32_t> sp_dlcnt, sp_dloff, sp_dlfill, sp_n, sp_stale, sp_dead, sp_refs;
    const int32_t* sp_dtime = nullptr; bool sp_valid = false; uint32_t sp_ndead = 0; int64_t commit_ns = 0; int64_t sp_build = -1; uint32_t sp_ncl = 0;
    // device-side BVE evaluation (bve_eval.cu): the batch stays resident between cp3g_bve_eval and cp3g_bve_stream
    DevBuf<unsigned char> ev_recs; DevBuf<uint16_t> ev_pst, ev_nst, ev_rsz; DevBuf<unsigned long long> ev_loff; uint32_t ev_nv = 0; int64_t bve_ns = 0;
    int64_t bce_ns = 0, bce_bytes = 0;                                 // K9 (bce.cu)
    std::vector<uint32_t> ovf_host;
    uint32_t total_occ = 0;
    int rank = 0, world = 1; void* nccl = nullptr;
    int64_t launches = 0, kernel_ns = 0, h2d = 0, d2h = 0, scan_requests = 0, scan_checks = 0, scan_ns = 0, scan_alg_bytes = 0;
    // per-kernel-family device times and ALGORITHMIC bytes for the roofline report: K1+K2 build, K0 ingest,
    // k_full by kind (0 = K3 subsumption, 1 = K4 strengthening), device pass commit, K5'/K6'
    int64_t build_ns = 0, build_bytes = 0, build_count = 0, ingest_ns = 0, ingest_bytes = 0, commit_bytes = 0, bve_bytes = 0;
    int64_t full_ns[2] = {0, 0}, full_bytes[2] = {0, 0}, full_checks[2] = {0, 0}, full_launches[2] = {0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t* pin_cnt = nullptr; cp3g_hit* pin_hits = nullptr;          // pinned: counters + the first PIN_HITS hits come back with one sync
    void fail(const char* what, cudaError_t e) { err = std::string(what) + ": " + cudaGetErrorString(e); fprintf(stderr, "cp3b200: CUDA error %s\n", err.c_str()); }
};

template <class T> int DevBuf<T>::reserve(size_t n, bool keep, cudaStream_t st)
{
    if (n <= cap) { return 0; }
    size_t ncap = std::max(n, cap + cap / 2 + 64);
    T* np = nullptr;
    // stream-ordered allocation from the device's default pool (its release threshold is lifted in cp3g_create):
    // buffers of a destroyed handle go back to the pool, so the next CPinit/CPpreprocess in the process does not pay
    // for mapping several hundred MB again
    cudaError_t e = cudaMallocAsync((void**)&np, ncap * sizeof(T), st);
    if (e != cudaSuccess) { return (int)e; }
    if (keep && p && used) { cudaMemcpyAsync(np, p, used * sizeof(T), cudaMemcpyDeviceToDevice, st); }
    if (p) { cudaFreeAsync(p, st); }
    p = np; cap = ncap;
    return 0;
}
template <class T> void DevBuf<T>::release(cudaStream_t st) { if (p) { cudaFreeAsync(p, st); } p = nullptr; cap = used = 0; }

#define RESERVE(buf, n, keep) do { int _r = (buf).reserve((n), (keep), g->stream); if (_r) { g->fail("cudaMalloc", (cudaError_t)_r); return CP3G_ERR_CUDA; } } while (0)

static inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

struct KTimer {
    cp3g* g;
    explicit KTimer(cp3g* g_) : g(g_) { cudaEventRecord(g->ev0, g->stream); }
    void stop() { cudaEventRecord(g->ev1, g->stream); }
    int64_t collect() { float ms = 0; if (cudaEventElapsedTime(&ms, g->ev0, g->ev1) == cudaSuccess) { g->kernel_ns += (int64_t)(ms * 1e6); return (int64_t)(ms * 1e6); } return 0; }
};


// helpers defined in gpu_db.cu and used by commit.cu
int d2h_staged(cp3g* g, void* dst, const void* src, size_t bytes);      // staging.cu: large copy into pageable host memory, returns when it is there
int h2d_staged(cp3g* g, void* ddst, const void* src, size_t bytes);     // staging.cu: large copy out of pageable host memory (plain async copy for pinned sources)
extern "C" int scan_u32(cp3g* g, const uint32_t* in, uint32_t* out, size_t n);     // exclusive scan on the handle stream, total in counters[2]
int dev_set_deleted(cp3g* g, const uint32_t* dev_ids, uint32_t n);       // mark clauses deleted (header flag), ids on the device
int dev_host_order_lists(cp3g* g);                                        // refs_tmp := occurrence lists in ascending clause order
#endif
