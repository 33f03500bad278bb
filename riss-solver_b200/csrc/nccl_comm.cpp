// nccl_comm.cpp - multi-GPU exchange for the scan service: every rank scans a contiguous slice of the
// request list on its own GPU; the hit lists are all-gathered (variable length) over NCCL so that all
// ranks commit the identical ordered result.  libnccl.so.2 is dlopen()ed: the
// library loads and its single-GPU path works on hosts without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "gpu_db.h"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclUint32 = 3, ncclUint64 = 5 };
struct Api {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Api* api()
{
    static Api a; static bool tried = false;
    if (tried) { return a.lib ? &a : nullptr; }
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) { break; } }
    if (!a.lib) { fprintf(stderr, "cp3b200: cannot dlopen libnccl.so.2: %s\n", dlerror()); return nullptr; }
#define SYM(f) *(void**)(&a.f) = dlsym(a.lib, "nccl" #f); if (!a.f) { fprintf(stderr, "cp3b200: nccl" #f " missing\n"); a.lib = nullptr; return nullptr; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllGather) SYM(Broadcast) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    return &a;
}
struct Comm {
    ncclComm_t comm = nullptr; int rank = 0, world = 1;
    uint32_t* dcounts = nullptr;      // world x 4 u32: n, checks lo, checks hi, pad
    cp3g_hit* dgather = nullptr; size_t gather_cap = 0;
};
}

extern "C" int cp3g_nccl_unique_id(void* out128)
{
    Api* a = api();
    if (!a) { return CP3G_ERR_NODEVICE; }
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != 0) { return CP3G_ERR_CUDA; }
    memcpy(out128, &id, sizeof(id));
    return CP3G_OK;
}

// One communicator per process and (rank, world): building it and its first collective (NCCL connects the peers lazily) cost a few
// hundred milliseconds - as much as a whole simplification - so a process that creates several handles in a row (one per CPinit)
// keeps the first one.  Every rank of the job takes the same decision, so the cached communicators stay matched; the unique id of
// a later call is simply not used.
static Comm* g_cached = nullptr;

void* cp3g_nccl_create(int rank, int world, const void* id128)
{
    Api* a = api();
    if (!a) { return nullptr; }
    if (g_cached && g_cached->rank == rank && g_cached->world == world) { return g_cached; }
    Comm* c = new Comm();
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    ncclResult_t r = a->CommInitRank(&c->comm, world, id, rank);
    if (r != 0) { fprintf(stderr, "cp3b200: ncclCommInitRank: %s\n", a->GetErrorString(r)); delete c; return nullptr; }
    c->rank = rank; c->world = world;
    if (cudaMalloc((void**)&c->dcounts, sizeof(uint32_t) * 4 * (size_t)(world + 1)) != cudaSuccess) { delete c; return nullptr; }
    if (!g_cached) { g_cached = c; }
    return c;
}

void cp3g_nccl_free(void* comm)
{
    if (!comm || comm == g_cached) { return; }              // the cached communicator lives as long as the process
    Comm* c = (Comm*)comm; Api* a = api();
    if (a && c->comm) { a->CommDestroy(c->comm); }
    if (c->dcounts) { cudaFree(c->dcounts); }
    if (c->dgather) { cudaFree(c->dgather); }
    delete c;
}

int cp3g_nccl_gather_hits(void* comm, cudaStream_t st, const cp3g_hit* dhits, uint32_t n, uint32_t cap,
                          cp3g_hit* out, uint32_t* nout, uint64_t* checks)
{
    Comm* c = (Comm*)comm; Api* a = api();
    if (!c || !a) { return CP3G_ERR_NODEVICE; }
    const int W = c->world;
    uint32_t mine[4] = { n, (uint32_t)(*checks & 0xffffffffu), (uint32_t)(*checks >> 32), 0 };
    uint32_t* dmine = c->dcounts + 4 * (size_t)W;
    if (cudaMemcpyAsync(dmine, mine, sizeof(mine), cudaMemcpyHostToDevice, st) != cudaSuccess) { return CP3G_ERR_CUDA; }
    if (a->AllGather(dmine, c->dcounts, 4, ncclUint32, c->comm, st) != 0) { return CP3G_ERR_CUDA; }
    std::vector<uint32_t> all(4 * (size_t)W);
    if (cudaMemcpyAsync(all.data(), c->dcounts, all.size() * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) { return CP3G_ERR_CUDA; }
    if (cudaStreamSynchronize(st) != cudaSuccess) { return CP3G_ERR_CUDA; }
    uint64_t total = 0, chk = 0;
    std::vector<uint64_t> offs(W + 1, 0);
    for (int r = 0; r < W; ++r) { offs[r] = total; total += all[4 * r]; chk += (uint64_t)all[4 * r + 1] | ((uint64_t)all[4 * r + 2] << 32); }
    offs[W] = total;
    *checks = chk;
    // every rank sees the same counts and the same cap, so every rank takes the same branch here (a local count above
    // cap is included: total >= it).  The caller retries with a cap derived from `total` alone.
    if (total > cap) { *nout = (uint32_t)std::min<uint64_t>(total, 0xffffffffull); return CP3G_ERR_OVERFLOW; }
    if (total > c->gather_cap) {
        if (c->dgather) { cudaFree(c->dgather); }
        c->gather_cap = total + total / 2 + 1024;
        if (cudaMalloc((void**)&c->dgather, c->gather_cap * sizeof(cp3g_hit)) != cudaSuccess) { c->dgather = nullptr; c->gather_cap = 0; return CP3G_ERR_CUDA; }
    }
    if (total) {
        // all-gather-v: one broadcast per rank inside a group (payloads are small; latency-bound)
        a->GroupStart();
        for (int r = 0; r < W; ++r) {
            const size_t bytes = (size_t)all[4 * r] * sizeof(cp3g_hit);
            if (!bytes) { continue; }
            a->Broadcast(dhits, c->dgather + offs[r], bytes, ncclUint8, r, c->comm, st);
        }
        if (a->GroupEnd() != 0) { return CP3G_ERR_CUDA; }
        if (cudaMemcpyAsync(out, c->dgather, total * sizeof(cp3g_hit), cudaMemcpyDeviceToHost, st) != cudaSuccess) { return CP3G_ERR_CUDA; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { return CP3G_ERR_CUDA; }
    }
    *nout = (uint32_t)total;
    return CP3G_OK;
}

// all-gather-v of the ranks' hit lists that STAYS on the device (the device-side pass commit consumes it): counts by ncclAllGather,
// payload as one grouped broadcast per rank, all enqueued on the stream of the kernel that produced the hits.  *dout points into a
// buffer owned by the communicator; *total = number of hits of all ranks, *checks is replaced by the sum over the ranks.
int cp3g_nccl_gather_hits_device(void* comm, cudaStream_t st, const cp3g_hit* dhits, uint32_t n, const cp3g_hit** dout, uint32_t* total_out, uint64_t* checks)
{
    Comm* c = (Comm*)comm; Api* a = api();
    if (!c || !a) { return CP3G_ERR_NODEVICE; }
    const int W = c->world;
    uint32_t mine[4] = { n, (uint32_t)(*checks & 0xffffffffu), (uint32_t)(*checks >> 32), 0 };
    uint32_t* dmine = c->dcounts + 4 * (size_t)W;
    if (cudaMemcpyAsync(dmine, mine, sizeof(mine), cudaMemcpyHostToDevice, st) != cudaSuccess) { return CP3G_ERR_CUDA; }
    if (a->AllGather(dmine, c->dcounts, 4, ncclUint32, c->comm, st) != 0) { return CP3G_ERR_CUDA; }
    std::vector<uint32_t> all(4 * (size_t)W);
    if (cudaMemcpyAsync(all.data(), c->dcounts, all.size() * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) { return CP3G_ERR_CUDA; }
    if (cudaStreamSynchronize(st) != cudaSuccess) { return CP3G_ERR_CUDA; }      // the sizes of the payload: one small read
    uint64_t total = 0, chk = 0;
    std::vector<uint64_t> offs((size_t)W + 1, 0);
    for (int r = 0; r < W; ++r) { offs[r] = total; total += all[4 * r]; chk += (uint64_t)all[4 * r + 1] | ((uint64_t)all[4 * r + 2] << 32); }
    *checks = chk;
    if (total > 0xfffffff0ull) { return CP3G_ERR_OVERFLOW; }
    if (total > c->gather_cap) {
        if (c->dgather) { cudaFree(c->dgather); }
        c->gather_cap = total + total / 2 + 1024;
        if (cudaMalloc((void**)&c->dgather, c->gather_cap * sizeof(cp3g_hit)) != cudaSuccess) { c->dgather = nullptr; c->gather_cap = 0; return CP3G_ERR_CUDA; }
    }
    if (total) {
        a->GroupStart();
        for (int r = 0; r < W; ++r) {
            const size_t bytes = (size_t)all[4 * r] * sizeof(cp3g_hit);
            if (!bytes) { continue; }
            a->Broadcast(dhits, c->dgather + offs[r], bytes, ncclUint8, r, c->comm, st);
        }
        if (a->GroupEnd() != 0) { return CP3G_ERR_CUDA; }
    }
    *dout = c->dgather; *total_out = (uint32_t)total;
    return CP3G_OK;
}

// all-gather-v in place: rank r owns bytes [off[r], off[r+1]) of the device buffer `dbuf`, whose layout every rank
// knows (the partition is a pure function of the request); afterwards all ranks hold the whole buffer.  Used for
// the per-pair resolvent sizes (K5) and the resolvent literal streams (K6) of a variable batch.
int cp3g_nccl_allgatherv_inplace(void* comm, cudaStream_t st, void* dbuf, const size_t* off)
{
    Comm* c = (Comm*)comm; Api* a = api();
    if (!c || !a) { return CP3G_ERR_NODEVICE; }
    a->GroupStart();
    for (int r = 0; r < c->world; ++r) {
        const size_t bytes = off[r + 1] - off[r];
        if (!bytes) { continue; }
        char* ptr = (char*)dbuf + off[r];
        a->Broadcast(ptr, ptr, bytes, ncclUint8, r, c->comm, st);
    }
    if (a->GroupEnd() != 0) { return CP3G_ERR_CUDA; }
    return CP3G_OK;
}
