// Host-side construction of the flat GPU index (layout: index.cuh) from BOSS arrays
// W / last / F, as handed over by the reference's DBGSuccinct (boss.hpp:499-525).
// Pure C++; used by the C-ABI (api.cu) and by the host-emulation test harness.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "index.cuh"

namespace mgb {

struct HostIndex {
    std::vector<uint32_t> blocks, blk_rank, sel_last, valid, sfx;
    std::vector<uint2> adj;
    std::vector<uint32_t> sel_W[kSigmaDNA];
    uint64_t n = 0; uint32_t nblk = 0, k = 0, sfx_len = 0, sigma = kSigmaDNA;
    uint64_t F[kMaxSigma], NF[kMaxSigma]; uint32_t total_W[kMaxSigma]; uint64_t num_ones = 0;
    // alphabet-generic layout (index.cuh IndexView::wide)
    bool wide = false;
    std::vector<uint8_t> wW;
    std::vector<uint32_t> wl, wrank, wsel, wadj;
    uint32_t wsel_off[kMaxSigma];

    // view over the host vectors (emulation) — the device view is assembled in api.cu
    IndexView view() const {
        IndexView v;
        std::memset(&v, 0, sizeof(v));
        v.blocks = blocks.data(); v.blk_rank = blk_rank.data(); v.sel_last = sel_last.data();
        for (int c = 0; c < kSigmaDNA; ++c) v.sel_W[c] = sel_W[c].data();
        v.valid = valid.empty() ? nullptr : valid.data();
        v.sfx = sfx.empty() ? nullptr : sfx.data();
        v.adj = adj.empty() ? nullptr : adj.data();
        v.radj = nullptr;
        v.n = n; v.nblk = nblk; v.k = k; v.sfx_len = sfx_len; v.sigma = sigma;
        for (uint32_t c = 0; c < sigma; ++c) { v.F[c] = F[c]; v.NF[c] = NF[c]; v.total_W[c] = total_W[c]; }
        v.num_ones = num_ones;
        v.wide = wide ? 1 : 0;
        v.wW = wW.data(); v.wl = wl.data(); v.wrank = wrank.data(); v.wsel = wsel.data(); v.wadj = wadj.data();
        for (int c = 0; c < kMaxSigma; ++c) v.wsel_off[c] = wsel_off[c];
        return v;
    }
};

// n_plus_1 = number of edges + 1 (position 0 is the placeholder, boss_chunk.cpp:60-62).
// sigma = 5 builds the DNA block layout, unless force_wide; any other sigma <= 32 the generic one.
inline void build_host_index(const uint8_t *W, const uint8_t *last, uint64_t n_plus_1,
                             const uint64_t *F, const uint8_t *valid_bytes, uint32_t k,
                             uint32_t suffix_len, HostIndex *out, uint32_t sigma = kSigmaDNA,
                             bool force_wide = false) {
    HostIndex &h = *out;
    if (n_plus_1 < 2) throw std::invalid_argument("empty BOSS table");
    if (n_plus_1 - 1 >= (1ull << 32) - 64) throw std::invalid_argument("more than 2^32 edges");
    if (k < 2) throw std::invalid_argument("k must be >= 2");
    if (sigma < 2 || sigma > (uint32_t)kMaxSigma) throw std::invalid_argument("alphabet size out of range");
    h.n = n_plus_1 - 1; h.k = k; h.sigma = sigma;
    h.wide = force_wide || sigma != (uint32_t)kSigmaDNA;
    h.nblk = (uint32_t)((h.n >> 6) + 1);
    h.blk_rank.assign(h.nblk + 1, 0);
    std::vector<uint32_t> cnt(sigma, 0);
    std::vector<std::vector<uint32_t>> selW(sigma);
    uint32_t ones = 0;
    h.sel_last.clear();
    h.blocks.clear(); h.wW.clear(); h.wl.clear(); h.wrank.clear(); h.wsel.clear(); h.wadj.clear();
    if (h.wide) {
        h.wW.assign((size_t)h.nblk * 64, 0xFF);
        h.wl.assign((size_t)h.nblk * 4, 0);
        h.wrank.assign((size_t)(h.nblk + 1) * kMaxSigma, 0);
    } else {
        h.blocks.assign((size_t)h.nblk * kBlkWords, 0);
    }
    for (uint32_t b = 0; b < h.nblk; ++b) {
        uint32_t *blk = h.wide ? nullptr : &h.blocks[(size_t)b * kBlkWords];
        h.blk_rank[b] = ones;
        if (h.wide) {
            h.wl[(size_t)b * 4 + 2] = ones;
            for (uint32_t c = 0; c < sigma; ++c) h.wrank[(size_t)b * kMaxSigma + c] = cnt[c];
        } else {
            blk[10] = ones;
            for (uint32_t c = 0; c < sigma; ++c) blk[11 + c] = cnt[c];
        }
        for (int off = 0; off < kBlkEdges; ++off) {
            uint64_t i = ((uint64_t)b << 6) + off;
            uint32_t w = h.wide ? 0xFF : 0xF; uint32_t l = 0;
            if (i >= 1 && i <= h.n) {
                w = W[i];
                if (w >= 2 * sigma) throw std::invalid_argument("W value out of range");
                l = last[i] ? 1 : 0;
            }
            if (h.wide) {
                h.wW[i] = (uint8_t)w;
                h.wl[(size_t)b * 4 + (off >> 5)] |= l << (off & 31);
            } else {
                blk[off >> 3] |= w << (4 * (off & 7));
                blk[8 + (off >> 5)] |= l << (off & 31);
            }
            if (l) {
                if (ones % kSelLastRate == 0) h.sel_last.push_back(b);
                ++ones;
            }
            if (w < sigma) {
                if (cnt[w] % kSelWRate == 0) selW[w].push_back(b);
                ++cnt[w];
            }
        }
    }
    h.blk_rank[h.nblk] = ones;
    h.num_ones = ones;
    h.sel_last.push_back(h.nblk - 1);
    h.sel_last.push_back(h.nblk - 1);
    for (int c = 0; c < kMaxSigma; ++c) { h.wsel_off[c] = 0; h.total_W[c] = 0; h.F[c] = 0; h.NF[c] = 0; }
    for (int c = 0; c < kSigmaDNA; ++c) h.sel_W[c].clear();
    for (uint32_t c = 0; c < sigma; ++c) {
        h.total_W[c] = cnt[c];
        selW[c].push_back(h.nblk - 1);
        selW[c].push_back(h.nblk - 1);
        if (h.wide) {
            h.wrank[(size_t)h.nblk * kMaxSigma + c] = cnt[c];
            h.wsel_off[c] = (uint32_t)h.wsel.size();
            h.wsel.insert(h.wsel.end(), selW[c].begin(), selW[c].end());
        } else {
            h.sel_W[c] = selW[c];
        }
    }
    for (uint32_t c = 0; c < sigma; ++c) h.F[c] = F[c];
    // NF[c] = rank_last(F[c]) (boss.cpp:1095-1101)
    for (uint32_t c = 0; c < sigma; ++c) {
        uint64_t i = F[c], r = 0;
        if (i) {
            r = h.blk_rank[i >> 6];
            for (uint64_t p = (i >> 6) << 6; p <= i; ++p) r += (p >= 1 && last[p]) ? 1 : 0;
        }
        h.NF[c] = r;
    }
    h.valid.clear();
    if (valid_bytes) {
        h.valid.assign((n_plus_1 + 31) / 32 + 1, 0);
        for (uint64_t i = 1; i <= h.n; ++i)
            if (valid_bytes[i]) h.valid[i >> 5] |= 1u << (i & 31);
    }
    // forward adjacency records (index.cuh load_adj): the j-th un-flagged occurrence of label c
    // points to the j-th node whose last character is c (boss.cpp:642-652 fwd); flagged edges share
    // the target of the preceding un-flagged one.
    {
        std::vector<uint32_t> ones_pos;          // select_last(r) = ones_pos[r - 1]
        ones_pos.reserve(ones);
        for (uint64_t i = 1; i <= h.n; ++i) if (last[i]) ones_pos.push_back((uint32_t)i);
        std::vector<uint32_t> mask_all(ones_pos.size() + 1, 0), mask_ok(ones_pos.size() + 1, 0);   // by node rank
        {
            uint64_t r = 1; uint32_t all = 0, ok = 0;
            for (uint64_t i = 1; i <= h.n; ++i) {
                uint32_t c = W[i] % sigma;
                all |= 1u << c;
                if (c && (!valid_bytes || valid_bytes[i])) ok |= 1u << c;
                if (last[i]) { mask_all[r] = all; mask_ok[r] = ok; ++r; all = 0; ok = 0; }
            }
        }
        h.adj.clear();
        if (h.wide) h.wadj.assign(n_plus_1 * 4, 0);
        else h.adj.assign(n_plus_1, uint2{0, 0});
        std::vector<uint64_t> cur(sigma, 0);
        for (uint64_t i = 1; i <= h.n; ++i) {
            uint32_t w = W[i], c = w % sigma;
            if (w < sigma) ++cur[c];
            if (i > 1 && c == 0) continue;                         // sink dummy: no outgoing edges
            uint64_t r = h.NF[c] + cur[c];
            if (r == 0 || r > ones_pos.size()) continue;
            if (h.wide) {
                h.wadj[i * 4] = ones_pos[r - 1]; h.wadj[i * 4 + 1] = mask_all[r]; h.wadj[i * 4 + 2] = mask_ok[r];
            } else {
                h.adj[i] = uint2{ ones_pos[r - 1], mask_all[r] | (mask_ok[r] << 8) };
            }
        }
    }
    // suffix ranges (boss.hpp:516-525, boss_chunk_construct.cpp:260-320): for every string
    // over the sigma-1 real symbols of length s, the [begin, end) edge range of the nodes
    // ending with it, in co-lex index order. Computed by refinement from length s-1 with
    // host-only rank/select over the byte arrays (tighten_range, boss.hpp:682-693).
    h.sfx.clear(); h.sfx_len = 0;
    uint32_t s = suffix_len;
    if (s > k - 1) s = k - 1;
    if (s) {
        // rank_W(i, c) for c in 1..4 sampled every 64 positions; positions of set `last` bits
        const uint64_t np1 = n_plus_1;
        std::vector<uint32_t> wr((np1 / 64 + 1) * sigma, 0);
        std::vector<uint32_t> ones_pos;   // select_last(r) = ones_pos[r - 1]
        ones_pos.reserve(ones);
        {
            std::vector<uint32_t> c2(sigma, 0);
            for (uint64_t i = 0; i < np1; ++i) {
                if (i % 64 == 0) for (uint32_t c = 0; c < sigma; ++c) wr[(i / 64) * sigma + c] = c2[c];
                if (i >= 1) {
                    if (W[i] < sigma) ++c2[W[i]];
                    if (last[i]) ones_pos.push_back((uint32_t)i);
                }
            }
        }
        auto rankW = [&](uint64_t i, uint32_t c) -> uint64_t {     // occurrences in [1..i]
            if (i == 0) return 0;
            uint64_t r = wr[(i / 64) * sigma + c];
            for (uint64_t p = (i / 64) * 64; p <= i; ++p) r += (p >= 1 && W[p] == c);
            return r;
        };
        auto selectLast = [&](uint64_t r) -> uint64_t { return r == 0 ? 0 : ones_pos[r - 1]; };
        auto tighten = [&](uint64_t *rl, uint64_t *ru, uint32_t c) -> bool {
            uint64_t rk_rl = rankW(*rl - 1, c) + 1, rk_ru = rankW(*ru, c);
            if (rk_rl > rk_ru) return false;
            *rl = selectLast(h.NF[c] + rk_rl - 1) + 1;
            *ru = selectLast(h.NF[c] + rk_ru);
            return true;
        };
        std::vector<uint32_t> cur(2 * (sigma - 1)), nxt;
        for (uint32_t c = 1; c < sigma; ++c) {   // length 1: [F[c] + 1, F[c + 1] + 1)
            uint64_t rl = F[c] + 1 < h.n + 1 ? F[c] + 1 : h.n + 1;
            uint64_t ru = c + 1 < sigma ? F[c + 1] : h.n;
            cur[2 * (c - 1)] = (uint32_t)rl;
            cur[2 * (c - 1) + 1] = (uint32_t)(ru + 1);
        }
        uint64_t cur_num = sigma - 1;
        for (uint32_t len = 2; len <= s; ++len) {
            // new index = old_index + (c - 1) * (sigma-1)^(len-1): the appended character is
            // the most significant digit (boss.hpp:651-655)
            nxt.assign(2 * cur_num * (sigma - 1), 1);
            for (uint32_t c = 1; c < sigma; ++c) {
                for (uint64_t idx = 0; idx < cur_num; ++idx) {
                    uint64_t rl = cur[2 * idx], ru = (uint64_t)cur[2 * idx + 1] - 1;
                    uint64_t o = idx + (uint64_t)(c - 1) * cur_num;
                    if (rl <= ru && tighten(&rl, &ru, c)) {
                        nxt[2 * o] = (uint32_t)rl; nxt[2 * o + 1] = (uint32_t)(ru + 1);
                    }
                }
            }
            cur.swap(nxt);
            cur_num *= (sigma - 1);
        }
        for (uint64_t idx = 0; idx < cur_num; ++idx)
            if (cur[2 * idx] >= cur[2 * idx + 1]) { cur[2 * idx] = 1; cur[2 * idx + 1] = 1; }
        h.sfx = std::move(cur);
        h.sfx_len = s;
    }
}

} // namespace mgb
