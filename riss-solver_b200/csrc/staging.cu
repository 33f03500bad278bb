// staging.cu - large copies between the device and PAGEABLE host memory (device -> host: the host engine's clause arena and
// occurrence lists; host -> device: a client's DIMACS stream that is not page-locked).
//
// cudaMemcpy into pageable memory is staged by the driver through its own bounce buffer on one thread, and a freshly
// allocated destination takes its first-touch page faults on that thread too: 130 MB of arena at C2 came back at
// ~10 GB/s.  Here the copy runs through process-wide pinned chunks: the DMA fills chunk i+1.. while a few host threads
// copy the chunks that have landed into the destination (and take the page faults in parallel).  Replaces nothing in
// the reference (its data base never leaves the host); it is plumbing of the boundary described in DESIGN.md §2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "gpu_db.h"

namespace {
constexpr size_t CHUNK = 4u << 20;          // bytes per pinned chunk
constexpr int WORKERS = 6, NBUF = 2 * WORKERS;
constexpr size_t STAGED_MIN = 8u << 20;     // below this the plain copy is as good

struct Stage {
    std::mutex mu;                          // one staged copy at a time per process
    uint8_t* buf[NBUF] = {};
    cudaEvent_t ev[NBUF] = {};
    int device = -1; bool ok = false, tried = false;
    bool init(int dev)
    {
        if (tried && device == dev) { return ok; }
        if (tried) { release(); }
        tried = true; device = dev; ok = true;
        for (int b = 0; b < NBUF && ok; ++b) {
            ok = cudaHostAlloc((void**)&buf[b], CHUNK, cudaHostAllocPortable) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) == cudaSuccess;
        }
        if (!ok) { release(); tried = true; device = dev; cudaGetLastError(); }
        return ok;
    }
    void release()
    {
        for (int b = 0; b < NBUF; ++b) { if (buf[b]) { cudaFreeHost(buf[b]); buf[b] = nullptr; } if (ev[b]) { cudaEventDestroy(ev[b]); ev[b] = nullptr; } }
        ok = false; tried = false;
    }
};
Stage& stage() { static Stage* s = new Stage(); return *s; }   // leaked on purpose: no CUDA calls from static destructors
}

// Copies `bytes` from device memory to pageable host memory on the handle's stream and returns when the data is in `dst`.
// The caller's earlier work on the stream is ordered before it as usual.
int d2h_staged(cp3g* g, void* dst, const void* src, size_t bytes)
{
    if (!bytes) { return CP3G_OK; }
    static const bool off = getenv("CP3_NO_STAGED_D2H") != nullptr;
    Stage& st = stage();
    std::unique_lock<std::mutex> lk(st.mu, std::defer_lock);
    if (off || bytes < STAGED_MIN || !lk.try_lock() || !st.init(g->device)) {
        if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g->stream) != cudaSuccess) { g->fail("d2h", cudaGetLastError()); return CP3G_ERR_CUDA; }
        if (cudaStreamSynchronize(g->stream) != cudaSuccess) { g->fail("d2h sync", cudaGetLastError()); return CP3G_ERR_CUDA; }
        return CP3G_OK;
    }
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[nchunks]);
    for (size_t i = 0; i < nchunks; ++i) { done[i].store(0, std::memory_order_relaxed); }
    std::atomic<size_t> issued{0};
    std::atomic<int> failed{0};
    const int nw = (int)std::min<size_t>(WORKERS, nchunks);
    std::vector<std::thread> workers;
    for (int w = 0; w < nw; ++w) {
        workers.emplace_back([&, w]() {
            cudaSetDevice(g->device);
            for (size_t i = (size_t)w; i < nchunks; i += (size_t)nw) {
                while (issued.load(std::memory_order_acquire) <= i) { if (failed.load(std::memory_order_relaxed)) { return; } std::this_thread::yield(); }
                const int b = (int)(i % NBUF);
                if (cudaEventSynchronize(st.ev[b]) != cudaSuccess) { failed.store(1); done[i].store(1, std::memory_order_release); continue; }
                const size_t o = i * CHUNK, len = std::min(CHUNK, bytes - o);
                memcpy((uint8_t*)dst + o, st.buf[b], len);
                done[i].store(1, std::memory_order_release);
            }
        });
    }
    for (size_t i = 0; i < nchunks; ++i) {
        const int b = (int)(i % NBUF);
        if (i >= (size_t)NBUF) { while (!done[i - NBUF].load(std::memory_order_acquire)) { std::this_thread::yield(); } }   // chunk buffer free again
        const size_t o = i * CHUNK, len = std::min(CHUNK, bytes - o);
        if (cudaMemcpyAsync(st.buf[b], (const uint8_t*)src + o, len, cudaMemcpyDeviceToHost, g->stream) != cudaSuccess ||
            cudaEventRecord(st.ev[b], g->stream) != cudaSuccess) { failed.store(1); issued.store(nchunks, std::memory_order_release); break; }
        issued.store(i + 1, std::memory_order_release);
    }
    for (std::thread& t : workers) { t.join(); }
    if (failed.load()) { g->fail("staged d2h", cudaGetLastError()); return CP3G_ERR_CUDA; }
    return CP3G_OK;
}

// Copies `bytes` from host memory to device memory on the handle's stream.  Page-locked sources go straight to the copy engine
// (asynchronously, as before); a large pageable source is copied into the pinned chunks by the worker threads while the chunks
// that are full are on the wire.  Returns when the source has been read (the device copy may still be in flight on the stream
// for the plain path; the staged path waits for its last chunk, since the chunks are shared).
int h2d_staged(cp3g* g, void* ddst, const void* src, size_t bytes)
{
    if (!bytes) { return CP3G_OK; }
    static const bool off = getenv("CP3_NO_STAGED_H2D") != nullptr;
    bool pageable = false;
    if (!off && bytes >= STAGED_MIN) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess) { pageable = at.type == cudaMemoryTypeUnregistered; } else { cudaGetLastError(); }
    }
    Stage& st = stage();
    std::unique_lock<std::mutex> lk(st.mu, std::defer_lock);
    if (!pageable || !lk.try_lock() || !st.init(g->device)) {
        if (cudaMemcpyAsync(ddst, src, bytes, cudaMemcpyHostToDevice, g->stream) != cudaSuccess) { g->fail("h2d", cudaGetLastError()); return CP3G_ERR_CUDA; }
        return CP3G_OK;
    }
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    std::unique_ptr<std::atomic<uint8_t>[]> ready(new std::atomic<uint8_t>[nchunks]);
    for (size_t i = 0; i < nchunks; ++i) { ready[i].store(0, std::memory_order_relaxed); }
    std::atomic<size_t> issued{0};                      // chunks whose device copy + event were put on the stream
    std::atomic<int> failed{0};
    const int nw = (int)std::min<size_t>(WORKERS, nchunks);
    std::vector<std::thread> workers;
    for (int w = 0; w < nw; ++w) {
        workers.emplace_back([&, w]() {
            cudaSetDevice(g->device);
            for (size_t i = (size_t)w; i < nchunks; i += (size_t)nw) {
                const int b = (int)(i % NBUF);
                if (i >= (size_t)NBUF) {                    // the chunk buffer is free once its previous load is on the device
                    while (issued.load(std::memory_order_acquire) <= i - NBUF) { if (failed.load(std::memory_order_relaxed)) { return; } std::this_thread::yield(); }
                    if (cudaEventSynchronize(st.ev[b]) != cudaSuccess) { failed.store(1); ready[i].store(1, std::memory_order_release); continue; }
                }
                const size_t o = i * CHUNK, len = std::min(CHUNK, bytes - o);
                memcpy(st.buf[b], (const uint8_t*)src + o, len);
                ready[i].store(1, std::memory_order_release);
            }
        });
    }
    for (size_t i = 0; i < nchunks; ++i) {
        const int b = (int)(i % NBUF);
        while (!ready[i].load(std::memory_order_acquire)) { if (failed.load(std::memory_order_relaxed)) { break; } std::this_thread::yield(); }
        if (failed.load(std::memory_order_relaxed)) { issued.store(nchunks, std::memory_order_release); break; }
        const size_t o = i * CHUNK, len = std::min(CHUNK, bytes - o);
        if (cudaMemcpyAsync((uint8_t*)ddst + o, st.buf[b], len, cudaMemcpyHostToDevice, g->stream) != cudaSuccess ||
            cudaEventRecord(st.ev[b], g->stream) != cudaSuccess) { failed.store(1); issued.store(nchunks, std::memory_order_release); break; }
        issued.store(i + 1, std::memory_order_release);
    }
    for (std::thread& t : workers) { t.join(); }
    if (failed.load()) { g->fail("staged h2d", cudaGetLastError()); return CP3G_ERR_CUDA; }
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) { g->fail("staged h2d sync", cudaGetLastError()); return CP3G_ERR_CUDA; }
    return CP3G_OK;
}
