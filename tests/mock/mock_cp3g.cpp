// mock_cp3g.cpp - TEST INFRASTRUCTURE ONLY.  A CPU stand-in for the cp3g_* scan service so that the host
// engine's ordered-commit logic (riss-solver_b200/csrc/engine*.cpp) can be differential-tested against the
// oracle on machines without a GPU (`pytest -m "not gpu"`).  It is compiled into tests/_build/
// libcp3b200_mock.so together with engine*.cpp and capi.cpp; the product library libcp3b200.so never
// contains it and has no CPU path.  The mock answers exactly what the kernels answer (sets of hits).
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/cp3b200.h"

struct cp3g {
    uint32_t nvars = 0;
    std::vector<std::vector<int32_t>> cl; std::vector<uint8_t> del;
    std::vector<size_t> start; size_t nlits = 0;          // pool offsets as uploaded / appended (cp3g_dense writes the pool back)
    std::vector<std::vector<uint32_t>> occ;
    int64_t launches = 0, requests = 0, checks = 0;
    // result of the last cp3g_sub_pass
    // last cp3g_bve_eval batch (cp3g_bve_stream reads it)
    std::vector<uint32_t> ev_vars, ev_poff, ev_prefs, ev_noff, ev_nrefs; std::vector<cp3g_bve_rec> ev_recs;
    bool sp_valid = false; std::vector<uint32_t> sp_dead, sp_n, sp_stale, sp_dlc, sp_len0; std::vector<std::vector<uint32_t>> sp_lists; size_t sp_ncl = 0;
    std::string err;
};

static bool subset(const std::vector<int32_t>& s, const std::vector<int32_t>& d)
{
    size_t i = 0, j = 0;
    while (i < s.size() && j < d.size()) { if (s[i] == d[j]) { ++i; ++j; } else if (d[j] < s[i]) { ++j; } else { return false; } }
    return i == s.size();
}
static int32_t one_flip(const std::vector<int32_t>& s, const std::vector<int32_t>& d)
{
    size_t i = 0, j = 0; int32_t f = -1;
    while (i < s.size() && j < d.size()) {
        if (s[i] == d[j]) { ++i; ++j; }
        else if (s[i] == (d[j] ^ 1)) { if (f >= 0) { return -1; } f = d[j]; ++i; ++j; }
        else if (s[i] < d[j]) { return -1; }
        else { ++j; }
    }
    return i == s.size() ? f : -1;
}
static int res_size(const std::vector<int32_t>& P, const std::vector<int32_t>& N, int32_t pv, std::vector<int32_t>* out)
{
    const int32_t nv = pv ^ 1; size_t i = 0, j = 0; int r = 0; int32_t prev = -2; bool taut = false;
    while (i < P.size() || j < N.size()) {
        int32_t c;
        if (i < P.size() && P[i] == pv) { ++i; continue; }
        if (j < N.size() && N[j] == nv) { ++j; continue; }
        if (j >= N.size() || (i < P.size() && P[i] < N[j])) { c = P[i++]; } else { c = N[j++]; }
        if (c == prev) { continue; }
        if (prev != -2 && c == (prev ^ 1)) { taut = true; if (!out) { return -1; } }
        prev = c; ++r; if (out) { out->push_back(c); }
    }
    return taut ? -1 : r;
}

extern "C" {
int cp3g_device_count(void) { return 1; }
cp3g* cp3g_create(int) { return new cp3g(); }
void cp3g_destroy(cp3g* g) { delete g; }
const char* cp3g_last_error(const cp3g* g) { return g ? g->err.c_str() : ""; }
int cp3g_upload(cp3g* g, uint32_t nvars, uint32_t ncl, const int32_t* lits, size_t nl, const uint32_t* start,
                const uint32_t* size, const uint8_t* flags)
{
    g->nvars = nvars; g->cl.assign(ncl, {}); g->del.assign(ncl, 0); g->start.assign(ncl, 0); g->nlits = nl;
    for (uint32_t i = 0; i < ncl; ++i) { g->cl[i].assign(lits + start[i], lits + start[i] + size[i]); g->del[i] = flags ? (flags[i] & 1) : 0; g->start[i] = start[i]; }
    return CP3G_OK;
}
int cp3g_upload_packed(cp3g* g, uint32_t nvars, uint32_t ncl, const int32_t* lits, size_t nl, const void* rec)
{
    const uint32_t* r = (const uint32_t*)rec;
    g->nvars = nvars; g->cl.assign(ncl, {}); g->del.assign(ncl, 0); g->start.assign(ncl, 0); g->nlits = nl;
    for (uint32_t i = 0; i < ncl; ++i) {
        const uint32_t start = r[4 * i], sz = r[4 * i + 1] & 0xffffffu;
        g->start[i] = start;
        g->cl[i].assign(lits + start, lits + start + sz); g->del[i] = (r[4 * i + 1] >> 24) & 1u;
    }
    return CP3G_OK;
}
int cp3g_build(cp3g* g)
{
    g->occ.assign(2 * (size_t)g->nvars, {});
    for (uint32_t i = 0; i < g->cl.size(); ++i) { if (!g->del[i]) { for (int32_t x : g->cl[i]) { g->occ[x].push_back(i); } } }
    g->launches += 6;
    return CP3G_OK;
}
int cp3g_download_occ(cp3g* g, uint32_t* off, uint32_t* refs)
{
    uint32_t o = 0;
    for (size_t x = 0; x < g->occ.size(); ++x) { off[x] = o; if (refs) { std::copy(g->occ[x].begin(), g->occ[x].end(), refs + o); } o += (uint32_t)g->occ[x].size(); }
    off[g->occ.size()] = o;
    return CP3G_OK;
}
// The device-side commit of the first subsumption pass, mocked by the plain ordered walk it must reproduce
// (Subsumption::subsumption_worker, Subsumption.cc:244-313) over the pristine lists.
int cp3g_sub_pass(cp3g* g, uint32_t max_list, int* status, uint32_t* ndead, uint64_t* nlits, uint64_t* steps, uint64_t* checks)
{
    *status = 1; g->sp_valid = false;
    if (g->cl.empty()) { return CP3G_OK; }
    cp3g_build(g);
    std::vector<std::vector<uint32_t>> L = g->occ;
    std::vector<int> cnt(L.size());
    for (size_t x = 0; x < L.size(); ++x) { cnt[x] = (int)L[x].size(); }
    std::vector<uint8_t> del = g->del; std::vector<uint32_t> dead; uint64_t st = 0, nl = 0;
    for (uint32_t c = (uint32_t)g->cl.size(); c-- > 0;) {
        if (del[c]) { continue; }
        const std::vector<int32_t>& C = g->cl[c];
        int32_t mn = C[0];
        for (size_t k = 1; k < C.size(); ++k) { if (cnt[C[k]] < cnt[mn]) { mn = C[k]; } }
        std::vector<uint32_t>& li = L[mn];
        for (size_t i = 0; i < li.size(); ++i) {
            const uint32_t d = li[i];
            if (d == c) { continue; }
            ++st;
            if (g->cl[d].size() < C.size()) { continue; }
            if (del[d]) { li[i] = li.back(); li.pop_back(); --i; continue; }
            if (subset(C, g->cl[d])) { del[d] = 1; dead.push_back(d); nl += g->cl[d].size(); for (int32_t x : g->cl[d]) { --cnt[x]; } }
        }
    }
    // the kernel refuses lists with subsumed clauses that are longer than max_list
    for (uint32_t d : dead) { for (int32_t x : g->cl[d]) { if (g->occ[x].size() > max_list) { return CP3G_OK; } } }
    g->sp_dead = dead; g->sp_lists = L; g->sp_n.assign(L.size(), 0); g->sp_stale.assign(L.size(), 0);
    for (size_t x = 0; x < L.size(); ++x) { g->sp_n[x] = (uint32_t)L[x].size(); for (uint32_t d : L[x]) { g->sp_stale[x] += del[d] && !g->del[d]; } }
    g->sp_dlc.assign(L.size(), 0); g->sp_len0.assign(L.size(), 0);
    for (size_t x = 0; x < L.size(); ++x) { g->sp_len0[x] = (uint32_t)g->occ[x].size(); }
    for (uint32_t d : dead) { g->del[d] = 1; for (int32_t x : g->cl[d]) { ++g->sp_dlc[x]; } }
    g->sp_ncl = g->cl.size();
    g->sp_valid = true; g->launches += 12;
    *status = 0; *ndead = (uint32_t)dead.size(); *nlits = nl; *steps = st; if (checks) { *checks = st; }
    return CP3G_OK;
}
int cp3g_str_static(cp3g* g, const uint32_t* real_bits, int* status, uint64_t* inert_entries, uint64_t* inert_steps, uint64_t* small_clauses,
                    uint32_t* walked_bits)
{
    *status = 1;
    if (!g->sp_valid || g->sp_ncl != g->cl.size()) { return CP3G_OK; }
    const size_t nlit = g->occ.size();
    for (size_t w = 0; w < (nlit + 31) / 32; ++w) { walked_bits[w] = 0; }
    uint64_t ent = 0, st = 0, small = 0;
    auto cnt = [&](int32_t y) { return (int)g->sp_len0[y] - 2 * (int)g->sp_dlc[y]; };
    for (uint32_t c = 0; c < g->cl.size(); ++c) {
        if (g->del[c] || ((real_bits[c >> 5] >> (c & 31)) & 1u)) { continue; }
        const std::vector<int32_t>& C = g->cl[c];
        if (C.size() < 2) { ++small; continue; }
        int32_t mn = C[0]; int best = cnt(mn & ~1) + cnt(mn | 1);
        for (size_t k = 1; k < C.size(); ++k) { const int d = cnt(C[k] & ~1) + cnt(C[k] | 1); if (best > d) { best = d; mn = C[k]; } }
        const uint32_t l1 = g->sp_n[mn] - g->sp_stale[mn], l2 = g->sp_n[mn ^ 1] - g->sp_stale[mn ^ 1];
        st += (l1 ? l1 - 1 : 0) + l2; ++ent;
        walked_bits[(uint32_t)mn >> 5] |= 3u << ((uint32_t)mn & 30u);
    }
    g->launches++;
    *inert_entries = ent; *inert_steps = st; *small_clauses = small; *status = 0;
    return CP3G_OK;
}
int cp3g_sub_pass_fetch(cp3g* g, uint32_t* dead, uint32_t* list_n, uint32_t* list_stale, uint32_t* refs)
{
    if (!g->sp_valid) { return CP3G_ERR_ARG; }
    if (dead) { std::copy(g->sp_dead.begin(), g->sp_dead.end(), dead); }
    if (list_n) { std::copy(g->sp_n.begin(), g->sp_n.end(), list_n); }
    if (list_stale) { std::copy(g->sp_stale.begin(), g->sp_stale.end(), list_stale); }
    if (refs) { uint32_t o = 0; for (size_t x = 0; x < g->occ.size(); ++x) { std::copy(g->sp_lists[x].begin(), g->sp_lists[x].end(), refs + o); o += (uint32_t)g->occ[x].size(); } }
    return CP3G_OK;
}
int cp3g_set_deleted(cp3g* g, uint32_t n, const uint32_t* ids) { for (uint32_t i = 0; i < n; ++i) { g->del[ids[i]] = 1; } return CP3G_OK; }
int cp3g_shrink(cp3g* g, uint32_t n, const uint32_t* ids, const uint32_t* st, const uint32_t* sz, const int32_t* lits, size_t)
{
    for (uint32_t i = 0; i < n; ++i) { g->cl[ids[i]].assign(lits + st[i], lits + st[i] + sz[i]); }
    return CP3G_OK;
}
int cp3g_append(cp3g* g, uint32_t n, const uint32_t* size, const int32_t* lits, size_t)
{
    size_t p = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t id = (uint32_t)g->cl.size();
        g->cl.emplace_back(lits + p, lits + p + size[i]); g->del.push_back(0);
        g->start.push_back(g->nlits); g->nlits += size[i];
        for (int32_t x : g->cl.back()) { g->occ[x].push_back(id); }
        p += size[i];
    }
    return CP3G_OK;
}
int cp3g_scan(cp3g* g, int kind, uint32_t nreq, const uint32_t* self, const uint32_t* req_off, const int32_t* req_lits,
              cp3g_hit* out, uint32_t cap, uint32_t* nout, uint64_t* checks)
{
    uint32_t n = 0; uint64_t chk = 0;
    g->launches++; g->requests += nreq;
    for (uint32_t r = 0; r < nreq; ++r) {
        std::vector<int32_t> S;
        if (req_lits) { S.assign(req_lits + req_off[r], req_lits + req_off[r + 1]); }
        else { if (g->del[self[r]]) { continue; } S = g->cl[self[r]]; }
        if (S.empty()) { continue; }
        // shortest list like the kernel
        size_t best = 0; size_t bl = (size_t)-1;
        for (size_t k = 0; k < S.size(); ++k) {
            size_t len = g->occ[S[k]].size() + (kind == CP3G_STRENGTHEN ? g->occ[S[k] ^ 1].size() : 0);
            if (len < bl) { bl = len; best = k; }
        }
        const int32_t x = S[best];
        for (int pass = 0; pass < (kind == CP3G_STRENGTHEN ? 2 : 1); ++pass) {
            const int32_t lx = pass == 0 ? x : (x ^ 1);
            for (uint32_t d : g->occ[lx]) {
                if (d == self[r]) { continue; }
                ++chk;
                if (g->del[d] || g->cl[d].size() < S.size()) { continue; }
                if (kind == CP3G_SUBSUME) {
                    if (subset(S, g->cl[d])) { if (n < cap) { out[n] = {r, d, -1}; } ++n; }
                } else {
                    int32_t f = one_flip(S, g->cl[d]);
                    if (f >= 0 && (pass == 0 || f == lx)) {
                        // pass 0 may see a stale entry that flips on x itself; the kernel reports it once
                        if (pass == 0 && f == (x ^ 1)) { continue; }
                        if (n < cap) { out[n] = {r, d, f}; } ++n;
                    }
                }
            }
        }
    }
    g->checks += (int64_t)chk;
    if (nout) { *nout = n; }
    if (checks) { *checks = chk; }
    return n > cap ? CP3G_ERR_OVERFLOW : CP3G_OK;
}
int cp3g_scan_all(cp3g* g, int kind, cp3g_hit* out, uint32_t cap, uint32_t* nout, uint64_t* checks)
{
    // same answer as cp3g_scan over all live clauses, hit.req = clause index
    std::vector<uint32_t> ids;
    for (uint32_t i = 0; i < g->cl.size(); ++i) { if (!g->del[i]) { ids.push_back(i); } }
    uint32_t n = 0;
    int r = cp3g_scan(g, kind, (uint32_t)ids.size(), ids.data(), nullptr, nullptr, out, cap, &n, checks);
    if (nout) { *nout = n; }
    if (r != CP3G_OK) { return r; }
    for (uint32_t k = 0; k < n; ++k) { out[k].req = ids[out[k].req]; }
    return CP3G_OK;
}
int cp3g_bve_pairs(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, const uint32_t* p_refs,
                   const uint32_t* n_off, const uint32_t* n_refs, const uint64_t* pair_off, int16_t* sizes)
{
    g->launches++;
    for (uint32_t i = 0; i < nv; ++i) {
        const uint32_t np = p_off[i + 1] - p_off[i], nn = n_off[i + 1] - n_off[i];
        for (uint32_t a = 0; a < np; ++a) for (uint32_t b = 0; b < nn; ++b) {
            const uint32_t p = p_refs[p_off[i] + a], n = n_refs[n_off[i] + b];
            int s = (g->del[p] || g->del[n]) ? -2 : res_size(g->cl[p], g->cl[n], (int32_t)(vars[i] << 1), nullptr);
            sizes[pair_off[i] + (uint64_t)a * nn + b] = (int16_t)s;
        }
    }
    return CP3G_OK;
}
int cp3g_bce_masks(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, const uint32_t* p_refs,
                   const uint32_t* n_off, const uint32_t* n_refs, const uint64_t* word_off, uint32_t* masks)
{
    g->launches++;
    for (uint32_t i = 0; i < nv; ++i) {
        const uint32_t np = p_off[i + 1] - p_off[i], nn = n_off[i + 1] - n_off[i], W = (nn + 31) / 32;
        for (uint32_t a = 0; a < np; ++a) {
            for (uint32_t w = 0; w < W; ++w) { masks[word_off[i] + (uint64_t)a * W + w] = 0; }
            for (uint32_t b = 0; b < nn; ++b) {
                const uint32_t p = p_refs[p_off[i] + a], n = n_refs[n_off[i] + b];
                if (res_size(g->cl[p], g->cl[n], (int32_t)(vars[i] << 1), nullptr) >= 0) { masks[word_off[i] + (uint64_t)a * W + (b >> 5)] |= 1u << (b & 31); }
            }
        }
    }
    return CP3G_OK;
}
int cp3g_bve_resolve(cp3g* g, uint32_t npairs, const uint32_t* vars, const uint32_t* p_refs, const uint32_t* n_refs,
                     const uint64_t* out_off, int32_t* out_lits, size_t)
{
    g->launches++;
    for (uint32_t t = 0; t < npairs; ++t) {
        std::vector<int32_t> r; res_size(g->cl[p_refs[t]], g->cl[n_refs[t]], (int32_t)(vars[t] << 1), &r);
        for (size_t k = 0; k < r.size() && out_off[t] + k < out_off[t + 1]; ++k) { out_lits[out_off[t] + k] = r[k]; }
    }
    return CP3G_OK;
}
// K5' mocked by the plain per-variable sequence of bve_worker: findGates (BVE.cc:1143-1267) on the lists as given, lazy removal
// of deleted entries (BVE.cc:474-489), anticipateElimination (BVE.cc:695-880)
static bool mock_find_gates(cp3g* g, int v, std::vector<uint32_t>& pos, std::vector<uint32_t>& neg, int& p_limit, int& n_limit)
{
    if (pos.size() < 3 && neg.size() < 3) { return false; }
    if (pos.size() < 2 || neg.size() < 2) { return false; }
    for (int pn = 0; pn < 2; ++pn) {
        const int32_t pLit = 2 * v + (pn != 0), nLit = 2 * v + (pn == 0);
        std::vector<uint32_t>& pList = pn == 0 ? pos : neg; std::vector<uint32_t>& nList = pn == 0 ? neg : pos;
        int& pClauses = pn == 0 ? p_limit : n_limit; int& nClauses = pn == 0 ? n_limit : p_limit;
        std::vector<int32_t> marks;
        auto marked = [&](int32_t x) { return std::find(marks.begin(), marks.end(), x) != marks.end(); };
        for (uint32_t c : pList) {
            if (g->del[c] || g->cl[c].size() != 2) { continue; }
            const int32_t other = g->cl[c][0] == pLit ? g->cl[c][1] : g->cl[c][0];
            if (!marked(other ^ 1)) { marks.push_back(other ^ 1); }
        }
        for (size_t i = 0; i < nList.size(); ++i) {
            const uint32_t c = nList[i];
            if (g->del[c]) { continue; }
            size_t j = 0;
            for (; j < g->cl[c].size(); ++j) { if (g->cl[c][j] == nLit) { continue; } if (!marked(g->cl[c][j])) { break; } }
            if (j == g->cl[c].size()) {
                pClauses = (int)g->cl[c].size() - 1; nClauses = 1;
                for (int32_t x : g->cl[c]) { marks.erase(std::remove(marks.begin(), marks.end(), x), marks.end()); }
                std::swap(nList[0], nList[i]);
                size_t placed = 0;
                for (size_t k = 0; k < pList.size(); ++k) {
                    const uint32_t c2 = pList[k];
                    if (g->del[c2] || g->cl[c2].size() != 2) { continue; }
                    const int32_t other = g->cl[c2][0] == pLit ? g->cl[c2][1] : g->cl[c2][0];
                    if (!marked(other ^ 1)) { std::swap(pList[placed], pList[k]); placed++; marks.push_back(other ^ 1); }
                }
                return true;
            }
        }
    }
    return false;
}
int cp3g_bve_eval(cp3g* g, uint32_t nv, const uint32_t* vars, const uint32_t* p_off, uint32_t* p_refs, const uint32_t* n_off, uint32_t* n_refs,
                  int use_gates, cp3g_bve_rec* recs, uint16_t* p_stats, uint16_t* n_stats)
{
    g->launches++;
    for (uint32_t i = 0; i < nv; ++i) {
        std::vector<uint32_t> pos(p_refs + p_off[i], p_refs + p_off[i + 1]), neg(n_refs + n_off[i], n_refs + n_off[i + 1]);
        cp3g_bve_rec r; r.p_limit = (int)pos.size(); r.n_limit = (int)neg.size(); r.resolvents = r.lit_clauses = r.steps = r.flags = 0;
        r.pos_n = (uint32_t)pos.size(); r.neg_n = (uint32_t)neg.size();
        if (pos.size() > 64 || neg.size() > 64) { r.flags = CP3G_BVE_HOST; recs[i] = r; continue; }      // the kernel's list capacity
        int p_limit = r.p_limit, n_limit = r.n_limit;
        const bool found = use_gates && mock_find_gates(g, (int)vars[i], pos, neg, p_limit, n_limit);
        for (size_t k = 0; k < pos.size(); ++k) { if (g->del[pos[k]]) { pos[k] = pos.back(); pos.pop_back(); --k; } }
        for (size_t k = 0; k < neg.size(); ++k) { if (g->del[neg[k]]) { neg[k] = neg.back(); neg.pop_back(); --k; } }
        const bool hasDef = p_limit < (int)pos.size() || n_limit < (int)neg.size();
        std::vector<uint16_t> ps(pos.size(), 0), ns(neg.size(), 0); bool special = false;
        for (size_t a = 0; a < pos.size(); ++a) for (size_t b = 0; b < neg.size(); ++b) {
            if ((int)a >= p_limit && (int)b >= n_limit) { continue; }
            if (hasDef && (int)a < p_limit && (int)b < n_limit) { continue; }
            ++r.steps;
            const int sz = res_size(g->cl[pos[a]], g->cl[neg[b]], (int32_t)(vars[i] << 1), nullptr);
            if (sz > 1) { ++r.resolvents; r.lit_clauses += (uint32_t)sz; ++ps[a]; ++ns[b]; }
            else if (sz == 0 || sz == 1) { special = true; }
        }
        r.p_limit = p_limit; r.n_limit = n_limit; r.pos_n = (uint32_t)pos.size(); r.neg_n = (uint32_t)neg.size();
        r.flags = (found ? CP3G_BVE_GATE : 0u) | (special ? CP3G_BVE_HOST : 0u);
        recs[i] = r;
        std::copy(pos.begin(), pos.end(), p_refs + p_off[i]); std::copy(neg.begin(), neg.end(), n_refs + n_off[i]);
        if (p_stats) { std::copy(ps.begin(), ps.end(), p_stats + p_off[i]); }
        if (n_stats) { std::copy(ns.begin(), ns.end(), n_stats + n_off[i]); }
    }
    g->ev_vars.assign(vars, vars + nv); g->ev_poff.assign(p_off, p_off + nv + 1); g->ev_noff.assign(n_off, n_off + nv + 1);
    g->ev_prefs.assign(p_refs, p_refs + p_off[nv]); g->ev_nrefs.assign(n_refs, n_refs + n_off[nv]); g->ev_recs.assign(recs, recs + nv);
    return CP3G_OK;
}
int cp3g_bve_stream(cp3g* g, uint32_t nsel, const uint32_t* sel, const uint64_t* res_off, const uint64_t* lit_off, uint16_t* out_sizes, int32_t* out_lits)
{
    g->launches++;
    for (uint32_t k = 0; k < nsel; ++k) {
        const uint32_t i = sel[k]; const cp3g_bve_rec& r = g->ev_recs[i];
        const uint32_t* pos = g->ev_prefs.data() + g->ev_poff[i]; const uint32_t* neg = g->ev_nrefs.data() + g->ev_noff[i];
        const bool hasDef = r.p_limit < (int)r.pos_n || r.n_limit < (int)r.neg_n;
        uint64_t ro = res_off[k], lo = lit_off[k];
        for (uint32_t a = 0; a < r.pos_n; ++a) for (uint32_t b = 0; b < r.neg_n; ++b) {
            if ((int)a >= r.p_limit && (int)b >= r.n_limit) { continue; }
            if (hasDef && (int)a < r.p_limit && (int)b < r.n_limit) { continue; }
            std::vector<int32_t> out;
            const int sz = res_size(g->cl[pos[a]], g->cl[neg[b]], (int32_t)(g->ev_vars[i] << 1), &out);
            if (sz <= 0) { continue; }
            if (ro >= res_off[k + 1] || lo + out.size() > lit_off[k + 1]) { g->err = "stream overflow"; return CP3G_ERR_ARG; }
            out_sizes[ro++] = (uint16_t)sz;
            for (int32_t x : out) { out_lits[lo++] = x; }
        }
    }
    return CP3G_OK;
}
int cp3g_dense(cp3g* g, const uint8_t* keep, int frag, uint32_t* mapping, uint32_t* nvars_new, int* compressed, int32_t* lits_out, size_t nlits)
{
    const uint32_t nv = g->nvars;
    *compressed = 0; *nvars_new = nv;
    if (nv == 0) { return CP3G_OK; }
    if (nlits != g->nlits) { g->err = "pool size"; return CP3G_ERR_ARG; }
    std::vector<uint8_t> used(nv, 0);
    for (size_t i = 0; i < g->cl.size(); ++i) { if (!g->del[i]) { for (int32_t x : g->cl[i]) { used[x >> 1] = 1; } } }
    uint32_t kept = 0;
    for (uint32_t v = 0; v < nv; ++v) { if (used[v] || keep[v]) { mapping[v] = kept++; } else { mapping[v] = 0xffffffffu; } }
    const uint32_t diff = nv - kept;
    if (diff == 0 || ((double)frag > 1.0 + (double)((diff * 100u) / nv))) { return CP3G_OK; }
    for (size_t i = 0; i < g->cl.size(); ++i) {
        if (g->del[i]) { continue; }
        for (size_t k = 0; k < g->cl[i].size(); ++k) { int32_t& x = g->cl[i][k]; x = (int32_t)(2u * mapping[x >> 1] + (x & 1)); lits_out[g->start[i] + k] = x; }
    }
    g->nvars = kept; *compressed = 1; *nvars_new = kept; g->launches += 7;
    return CP3G_OK;
}
int cp3g_ingest(cp3g* g, const int32_t* S, size_t n, uint32_t* nvars, uint32_t* nclauses, size_t* nlits, int* status)
{
    *status = 1;
    if (n == 0 || S[n - 1] != 0) { return CP3G_OK; }
    std::vector<std::vector<int32_t>> cl; std::vector<int32_t> cur; uint32_t maxv = 0;
    for (size_t i = 0; i < n; ++i) {
        const int32_t x = S[i];
        if (x != 0) { const uint32_t v = (uint32_t)(x < 0 ? -x : x); maxv = std::max(maxv, v); cur.push_back(x < 0 ? 2 * (-x - 1) + 1 : 2 * (x - 1)); continue; }
        if (cur.empty() || cur.size() > 4096) { return CP3G_OK; }
        std::sort(cur.begin(), cur.end());
        std::vector<int32_t> out; bool taut = false; int32_t p = -2;
        for (int32_t y : cur) { if (p >= 0 && y == (p ^ 1)) { taut = true; break; } if (y != p) { p = y; out.push_back(y); } }
        cur.clear();
        if (taut) { continue; }
        if (out.size() == 1) { return CP3G_OK; }
        cl.push_back(out);
    }
    g->nvars = maxv; g->cl.swap(cl); g->del.assign(g->cl.size(), 0); g->start.assign(g->cl.size(), 0); g->occ.clear();
    size_t pos = 0;
    for (size_t i = 0; i < g->cl.size(); ++i) { g->start[i] = pos; pos += g->cl[i].size(); }
    g->nlits = pos;
    *nvars = maxv; *nclauses = (uint32_t)g->cl.size(); *nlits = pos; *status = 0;
    g->launches += 8;
    return CP3G_OK;
}
int cp3g_download_arena(cp3g* g, int32_t* lits, void* records16)
{
    uint32_t* r = (uint32_t*)records16;
    for (size_t i = 0; i < g->cl.size(); ++i) {
        std::copy(g->cl[i].begin(), g->cl[i].end(), lits + g->start[i]);
        r[4 * i] = (uint32_t)g->start[i]; r[4 * i + 1] = (uint32_t)g->cl[i].size() | (((g->del[i] ? 1u : 0u) | 6u) << 24); r[4 * i + 2] = r[4 * i + 3] = 0;
    }
    return CP3G_OK;
}
int cp3g_compact(cp3g* g, uint32_t* new_start, int32_t* lits_out, size_t lits_cap, size_t* nlits_new)
{
    size_t pos = 0;
    for (size_t i = 0; i < g->cl.size(); ++i) {
        g->start[i] = pos; if (new_start) { new_start[i] = (uint32_t)pos; }
        if (g->del[i]) { g->cl[i].clear(); continue; }
        if (lits_out) { if (pos + g->cl[i].size() > lits_cap) { return CP3G_ERR_ARG; } std::copy(g->cl[i].begin(), g->cl[i].end(), lits_out + pos); }
        pos += g->cl[i].size();
    }
    g->nlits = pos; if (nlits_new) { *nlits_new = pos; }
    g->launches += 5;
    return CP3G_OK;
}
int cp3g_nccl_unique_id(void* out) { memset(out, 0, 128); return CP3G_OK; }
int cp3g_nccl_init(cp3g*, int, int, const void*) { return CP3G_OK; }
int64_t cp3g_counter(const cp3g* g, const char* name)
{
    if (!strcmp(name, "launches")) { return g->launches; }
    if (!strcmp(name, "scan_requests")) { return g->requests; }
    if (!strcmp(name, "scan_checks")) { return g->checks; }
    if (!strcmp(name, "nclauses")) { return (int64_t)g->cl.size(); }
    if (!strcmp(name, "nlits")) { return (int64_t)g->nlits; }
    if (!strcmp(name, "nvars")) { return (int64_t)g->nvars; }
    return 0;
}
}
