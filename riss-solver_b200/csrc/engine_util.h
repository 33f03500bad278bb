// engine_util.h - helpers shared by the engine's translation units (timers, debug laps, the host-core loops).  Internal.
#pragma once
#include "engine.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <functional>

namespace cp3 {

static inline int64_t now_ns()
{
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct PhaseTimer {
    int64_t& acc; int64_t t0;
    explicit PhaseTimer(int64_t& a) : acc(a), t0(now_ns()) {}
    ~PhaseTimer() { acc += now_ns() - t0; }
};

// sort + unique of a clause id list; queues that were filled in clause order (the first round) are already sorted
static inline void sortUnique(std::vector<uint32_t>& v)
{
    bool strictly = true;
    for (size_t i = 1; i < v.size(); ++i) { if (v[i - 1] >= v[i]) { strictly = false; break; } }
    if (strictly) { return; }
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

// large hit lists arrive sorted by (request, clause) from the device (gpu_db.cu sort_hits); small or gathered ones are sorted here
static inline void sortHits(std::vector<cp3g_hit>& h, uint32_t n)
{
    auto less = [](const cp3g_hit& a, const cp3g_hit& b) { return a.req != b.req ? a.req < b.req : a.clause < b.clause; };
    if (!std::is_sorted(h.begin(), h.begin() + n, less)) { std::sort(h.begin(), h.begin() + n, less); }
}

struct DbgLap {
    bool on; int64_t t; const char* who;
    explicit DbgLap(const char* w) : on(getenv("CP3_DEBUG_PAR") != nullptr), t(on ? now_ns() : 0), who(w) {}
    DbgLap(const char* w, bool enabled) : on(enabled), t(on ? now_ns() : 0), who(w) {}      // per-variable callers: no getenv
    void operator()(const char* what) { if (on) { int64_t n2 = now_ns(); fprintf(stderr, "  %s %s %.1f ms\n", who, what, (n2 - t) / 1e6); t = n2; } }
};

static inline bool containsLit(const int32_t* l, uint32_t n, int x)
{
    return std::binary_search(l, l + n, x);
}

template <class V> static inline const Hit* findHit(const V& hits, uint32_t hb, uint32_t he, uint32_t clause)
{
    const Hit* b = hits.data() + hb; const Hit* e = hits.data() + he;
    const Hit* it = std::lower_bound(b, e, clause, [](const Hit& h, uint32_t c) { return h.clause < c; });
    return (it != e && it->clause == clause) ? it : nullptr;
}

template <class A> static inline int indexOf(const A& a, uint32_t x)
{
    for (size_t i = 0; i < a.size(); ++i) { if (a[i] == x) { return (int)i; } }
    return -1;
}

// contiguous chunks of [0,n) on the host cores (the statically inert part of the queue walks)
template <class F> void Engine::parallelFor(size_t n, F f, size_t grain)
{
    const int T = (int)std::min<size_t>((size_t)nthreads, (n + grain - 1) / grain);
    if (T <= 1) { f((size_t)0, n, 0); return; }
    const size_t per = (n + T - 1) / T;
    int tasks = 0;
    for (int t = 0; t < T; ++t) { if (per * t < n) { ++tasks; } }
    poolRun(tasks, [&](int t) { const size_t b = per * (size_t)t, e = std::min(n, per * ((size_t)t + 1)); f(b, e, t); });
}

template <class V, class P> void Engine::parallelFilter(const V& src, std::vector<uint32_t>& out, P pred)
{
    const size_t n = src.size();
    std::vector<size_t> cnt((size_t)nthreads + 2, 0);
    parallelFor(n, [&](size_t b, size_t e, int t) { size_t c = 0; for (size_t i = b; i < e; ++i) { c += pred(src[i]) ? 1 : 0; } cnt[(size_t)t + 1] = c; });
    for (size_t t = 1; t < cnt.size(); ++t) { cnt[t] += cnt[t - 1]; }
    out.resize(cnt.back());
    parallelFor(n, [&](size_t b, size_t e, int t) { size_t o = cnt[(size_t)t]; for (size_t i = b; i < e; ++i) { if (pred(src[i])) { out[o++] = src[i]; } } });
}

} // namespace cp3
