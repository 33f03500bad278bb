// gpu_db.h - internal helpers shared by gpu_db.cu and nccl_comm.cpp (not part of the C ABI)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
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
// in-place all-gather-v of a device buffer whose per-rank byte ranges [off[r], off[r+1]) are known to every rank
int   cp3g_nccl_allgatherv_inplace(void* comm, cudaStream_t st, void* dbuf, const size_t* off);
