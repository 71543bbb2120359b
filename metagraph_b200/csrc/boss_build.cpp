// Host-side batch construction of BOSS arrays (W, last, F) from sequences — the index-build
// side of the boundary (not on the timed path). Restates the *result* of
// BOSSConstructor -> construct_boss_chunk (boss_chunk_construct.cpp:341-462: real (k+1)-mers,
// non-redundant dummy sink edges, dummy source edges of every sentinel-prefix length, the
// main dummy edge) and initialize_chunk (boss_chunk.cpp:33-133: last, W with the minus
// flag, F) using one parallel sort of packed (k+1)-mers.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include <parallel/algorithm>
#include <omp.h>

#include "../../include/mgb.h"

namespace {

thread_local std::string g_boss_err;

typedef unsigned __int128 u128;
thread_local std::string g_build_err;

// Key of a (k+1)-mer that does not fit 128 bits (DNA k > 42; the reference moves to 256-bit k-mers there,
// kmer_boss.hpp / boss_construct.cpp): four little-endian words with the handful of operations the construction uses.
struct Wide256 {
    uint64_t q[4] = { 0, 0, 0, 0 };
    Wide256() {}
    Wide256(uint64_t lo) { q[0] = lo; }
    explicit operator uint32_t() const { return (uint32_t)q[0]; }
};
inline Wide256 operator<<(const Wide256 &a, int n) {
    Wide256 r;
    const int wq = n >> 6, bq = n & 63;
    for (int i = 3; i >= 0; --i) {
        uint64_t v = 0;
        if (i - wq >= 0) v = a.q[i - wq] << bq;
        if (bq && i - wq - 1 >= 0) v |= a.q[i - wq - 1] >> (64 - bq);
        r.q[i] = v;
    }
    return r;
}
inline Wide256 operator>>(const Wide256 &a, int n) {
    Wide256 r;
    const int wq = n >> 6, bq = n & 63;
    for (int i = 0; i < 4; ++i) {
        uint64_t v = 0;
        if (i + wq < 4) v = a.q[i + wq] >> bq;
        if (bq && i + wq + 1 < 4) v |= a.q[i + wq + 1] << (64 - bq);
        r.q[i] = v;
    }
    return r;
}
inline Wide256 operator|(Wide256 a, const Wide256 &b) { for (int i = 0; i < 4; ++i) a.q[i] |= b.q[i]; return a; }
inline Wide256 operator&(Wide256 a, const Wide256 &b) { for (int i = 0; i < 4; ++i) a.q[i] &= b.q[i]; return a; }
inline Wide256& operator|=(Wide256 &a, const Wide256 &b) { a = a | b; return a; }
inline Wide256 operator~(Wide256 a) { for (int i = 0; i < 4; ++i) a.q[i] = ~a.q[i]; return a; }
inline Wide256 operator-(const Wide256 &a, const Wide256 &b) {
    Wide256 r; uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        const uint64_t x = a.q[i], y = b.q[i];
        r.q[i] = x - y - borrow;
        borrow = (x < y) || (x == y && borrow);
    }
    return r;
}
inline bool operator==(const Wide256 &a, const Wide256 &b) { return !std::memcmp(a.q, b.q, sizeof(a.q)); }
inline bool operator!=(const Wide256 &a, const Wide256 &b) { return !(a == b); }
inline bool operator<(const Wide256 &a, const Wide256 &b) {
    for (int i = 3; i > 0; --i) if (a.q[i] != b.q[i]) return a.q[i] < b.q[i];
    return a.q[0] < b.q[0];
}

struct Alpha { int sigma, bits; uint8_t code[256]; };

Alpha make_dna() {
    Alpha a; a.sigma = 5; a.bits = 3;
    for (int i = 0; i < 256; ++i) a.code[i] = 5;
    const char *l = "ACGT";
    for (int i = 0; i < 4; ++i) { a.code[(int)l[i]] = i + 1; a.code[(int)l[i] + 32] = i + 1; }
    a.code[(int)'U'] = a.code[(int)'u'] = 4;
    for (int i = 128; i < 256; ++i) a.code[i] = 5;
    return a;
}

Alpha make_protein() {       // kmer/alphabets.hpp:29-38: "$ABCDEFGHIJKLMNOPQRSTUVWYZX", unknown -> 'X'
    Alpha a; a.sigma = 27; a.bits = 5;
    const char *l = "$ABCDEFGHIJKLMNOPQRSTUVWYZX";
    for (int i = 0; i < 256; ++i) a.code[i] = 26;
    for (int i = 1; i < 26; ++i) { a.code[(int)l[i]] = i; a.code[(int)l[i] + 32] = i; }
    return a;
}

template <class It> void psort(It b, It e, int threads) {
    if (threads > 1 && e - b > 100000) __gnu_parallel::sort(b, e);
    else std::sort(b, e);
}

template <class Key>
int build_with_key(const Alpha &al, const char *seqs, const uint64_t *offsets, uint32_t n_seqs, uint32_t K,
                   int alphabet, int force_source_dummies, int num_threads, mgb_boss_t *out) {
    const uint32_t k = K - 1;                       // node length
    const int bits = al.bits;
    const Key cmask = ((Key)1 << bits) - 1;
    const Key node_mask = ((Key)1 << (bits * k)) - 1;

    // key: [a_k ... a_1 | label], a_k most significant (KMerBOSS order, kmer_boss.hpp:58-64)
    std::vector<Key> real;
    std::vector<Key> forced;
    for (uint32_t s = 0; s < n_seqs; ++s) {
        const uint8_t *p = (const uint8_t*)seqs + offsets[s];
        const uint64_t len = offsets[s + 1] - offsets[s];
        uint64_t i = 0;
        while (i < len) {
            uint64_t j = i;
            while (j < len && al.code[p[j]] < al.sigma) ++j;
            if (j - i >= K) {
                const size_t base = real.size();
                real.resize(base + (j - i - K + 1));
                // rolling pack: node part shifts down by one char, new last char enters on top
                Key node = 0;                     // a_k..a_1 of the current window's node
                for (uint32_t t = 0; t < k; ++t) node |= (Key)al.code[p[i + t]] << (bits * t);
                for (uint64_t w = i; w + K <= j; ++w) {
                    Key lbl = al.code[p[w + k]];
                    real[base + (w - i)] = (node << bits) | lbl;
                    node = (node >> bits) | (lbl << (bits * (k - 1)));
                }
                if (force_source_dummies) {
                    for (uint32_t d = 1; d <= k; ++d) {   // d leading sentinels
                        Key nd = 0;
                        for (uint32_t t = d; t < k; ++t) nd |= (Key)al.code[p[i + t - d]] << (bits * t);
                        forced.push_back((nd << bits) | al.code[p[i + k - d]]);
                    }
                }
            }
            i = j + 1;
        }
    }
    psort(real.begin(), real.end(), num_threads);
    real.erase(std::unique(real.begin(), real.end()), real.end());

    // source nodes (sorted, unique) and target nodes of the real edges
    std::vector<Key> src, tgt(real.size());
    src.reserve(real.size());
    for (size_t i = 0; i < real.size(); ++i) {
        Key n = real[i] >> bits;
        if (src.empty() || src.back() != n) src.push_back(n);
    }
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < real.size(); ++i) {
        Key n = real[i] >> bits;
        tgt[i] = (n >> bits) | ((real[i] & cmask) << (bits * (k - 1)));
    }
    psort(tgt.begin(), tgt.end(), num_threads);
    tgt.erase(std::unique(tgt.begin(), tgt.end()), tgt.end());

    std::vector<Key> dummy;
    // sinks: target nodes without a real outgoing edge (boss_chunk_construct.cpp:57-100)
    {
        size_t si = 0;
        for (Key t : tgt) {
            while (si < src.size() && src[si] < t) ++si;
            if (si == src.size() || src[si] != t) dummy.push_back(t << bits);
        }
    }
    // sources with one sentinel: nodes without a real incoming edge (:124-170)
    std::vector<Key> level;
    {
        size_t ti = 0;
        for (Key n : src) {
            while (ti < tgt.size() && tgt[ti] < n) ++ti;
            if (ti == tgt.size() || tgt[ti] != n) {
                Key lbl = (n >> (bits * (k - 1))) & cmask;
                level.push_back((((n << bits) & node_mask) << bits) | lbl);
            }
        }
    }
    std::vector<Key>().swap(src);
    std::vector<Key>().swap(tgt);
    std::sort(level.begin(), level.end());
    level.erase(std::unique(level.begin(), level.end()), level.end());
    for (uint32_t c = 2; c < k + 1; ++c) {          // longer sentinel prefixes (:380-397)
        dummy.insert(dummy.end(), level.begin(), level.end());
        std::vector<Key> next;
        Key prev = ~(Key)0;
        for (Key key : level) {
            Key n = key >> bits;
            if (n == prev) continue;
            prev = n;
            Key lbl = (n >> (bits * (k - 1))) & cmask;
            next.push_back((((n << bits) & node_mask) << bits) | lbl);
        }
        std::sort(next.begin(), next.end());
        next.erase(std::unique(next.begin(), next.end()), next.end());
        level.swap(next);
    }
    dummy.insert(dummy.end(), level.begin(), level.end());
    dummy.insert(dummy.end(), forced.begin(), forced.end());
    dummy.push_back(0);                             // main dummy $...$ -> $ (:404-409)
    std::sort(dummy.begin(), dummy.end());
    dummy.erase(std::unique(dummy.begin(), dummy.end()), dummy.end());

    // merge + W / last / F (boss_chunk.cpp:33-133)
    const uint64_t n_total = real.size() + dummy.size();
    uint8_t *W = (uint8_t*)std::malloc(n_total + 2);
    uint8_t *last = (uint8_t*)std::malloc(n_total + 2);
    if (!W || !last) {
        std::free(W); std::free(last);
        g_boss_err = "out of host memory for W / last (" + std::to_string(2 * (n_total + 2)) + " bytes)";
        return MGB_ERR_NO_MEMORY;
    }
    W[0] = 0; last[0] = 0;
    std::memset(out->F, 0, sizeof(out->F));
    uint64_t curpos = 1;
    uint32_t lastF = 0;
    std::vector<Key> last_kmer(al.sigma, 0);
    std::vector<char> last_set(al.sigma, 0);
    const Key minus1 = ~(((Key)1 << (2 * bits)) - 1);
    size_t ri = 0, di = 0;
    auto peek = [&](bool *ok) -> Key {
        if (ri < real.size() && (di >= dummy.size() || real[ri] < dummy[di])) { *ok = true; return real[ri]; }
        if (di < dummy.size()) { *ok = true; return dummy[di]; }
        *ok = false; return 0;
    };
    auto pop = [&]() {
        if (ri < real.size() && (di >= dummy.size() || real[ri] < dummy[di])) ++ri; else ++di;
    };
    bool ok;
    Key kmer = peek(&ok);
    while (ok) {
        pop();
        bool ok2;
        Key nxt = peek(&ok2);
        uint32_t curW = (uint32_t)(kmer & cmask);
        uint32_t curF = (uint32_t)((kmer >> (bits * k)) & cmask);
        bool same_node = ok2 && (nxt >> bits) == (kmer >> bits);
        if (same_node && curW == 0 && curF > 0) { kmer = nxt; ok = ok2; continue; }  // redundant sink
        last[curpos] = same_node ? 0 : 1;
        if (curW) {
            if (last_set[curW] && last_kmer[curW] != 0
                    && (kmer & minus1) == (last_kmer[curW] & minus1)) {
                curW += al.sigma;
            } else {
                last_kmer[curW] = kmer; last_set[curW] = 1;
            }
        }
        W[curpos] = (uint8_t)curW;
        while (curF > lastF && lastF + 1 < (uint32_t)al.sigma) out->F[++lastF] = curpos - 1;
        ++curpos;
        kmer = nxt; ok = ok2;
    }
    while (++lastF < (uint32_t)al.sigma) out->F[lastF] = curpos - 1;
    out->n_plus_1 = curpos;
    out->W = W; out->last = last; out->k = K; out->alphabet = alphabet;
    return MGB_OK;
}

} // namespace

extern "C" {

const char* mgb_boss_last_error(void) { return g_boss_err.c_str(); }

int mgb_boss_build(const char *seqs, const uint64_t *offsets, uint32_t n_seqs, uint32_t K,
                   int alphabet, int force_source_dummies, int num_threads, mgb_boss_t *out) {
    auto fail = [](int rc, const std::string &why) { g_boss_err = why; return rc; };
    g_boss_err.clear();
    if (!offsets || !out || (n_seqs && !seqs)) return fail(MGB_ERR_INVALID_ARGUMENT, "null argument");
    if (alphabet != MGB_ALPHABET_DNA && alphabet != MGB_ALPHABET_PROTEIN)
        return fail(MGB_ERR_UNSUPPORTED, "unknown alphabet " + std::to_string(alphabet));
    if (K < 2) return fail(MGB_ERR_INVALID_ARGUMENT, "k must be at least 2");
    const Alpha al = alphabet == MGB_ALPHABET_PROTEIN ? make_protein() : make_dna();
    if ((uint64_t)K * al.bits > 256)                                   // DNA: k <= 85, protein: k <= 51
        return fail(MGB_ERR_UNSUPPORTED, "k = " + std::to_string(K) + " does not fit the 256-bit (k+1)-mer keys");
    if (num_threads < 1) num_threads = omp_get_max_threads();
    omp_set_num_threads(num_threads);
    try {
        // MGB_TEST_WIDE_KEYS=1 (tests): every graph through the 256-bit instantiation
        if ((uint64_t)K * al.bits <= 128 && !std::getenv("MGB_TEST_WIDE_KEYS"))
            return build_with_key<u128>(al, seqs, offsets, n_seqs, K, alphabet, force_source_dummies, num_threads, out);
        return build_with_key<Wide256>(al, seqs, offsets, n_seqs, K, alphabet, force_source_dummies, num_threads, out);
    } catch (const std::bad_alloc &) {
        return fail(MGB_ERR_NO_MEMORY, "out of host memory while sorting the (k+1)-mers");
    } catch (const std::exception &e) {
        return fail(MGB_ERR_INVALID_ARGUMENT, e.what());
    }
}

void mgb_boss_free(mgb_boss_t *b) {
    if (!b) return;
    std::free(b->W); std::free(b->last);
    b->W = b->last = nullptr; b->n_plus_1 = 0;
}

int mgb_boss_mask_dummy(const mgb_boss_t *b, uint8_t *valid) {
    // an edge is dummy iff its (k+1)-mer contains '$' (boss.cpp:1736-1775, flipped at
    // dbg_succinct.cpp:903-908). Source dummies are found by walking the sentinel tree.
    if (!b || !valid || !b->W) return MGB_ERR_INVALID_ARGUMENT;
    const uint64_t n = b->n_plus_1 - 1;
    const int sigma = b->alphabet == MGB_ALPHABET_PROTEIN ? 27 : 5;
    const uint32_t k = b->k - 1;
    std::vector<uint64_t> rankW((n / 64 + 2) * sigma, 0), ones;
    {
        uint64_t c2[32] = { 0 };
        for (uint64_t i = 0; i <= n; ++i) {
            if (i % 64 == 0) for (int c = 0; c < sigma; ++c) rankW[(i / 64) * sigma + c] = c2[c];
            if (i >= 1) { if (b->W[i] < sigma) ++c2[b->W[i]]; if (b->last[i]) ones.push_back(i); }
        }
    }
    uint64_t NF[32];
    for (int c = 0; c < sigma; ++c) {
        NF[c] = std::upper_bound(ones.begin(), ones.end(), b->F[c]) - ones.begin();
    }
    auto fwd = [&](uint64_t i, uint32_t c) -> uint64_t {
        uint64_t r = rankW[(i / 64) * sigma + c];
        for (uint64_t p = (i / 64) * 64; p <= i; ++p) r += (p >= 1 && b->W[p] == c);
        uint64_t t = NF[c] + r;
        return t ? ones[t - 1] : 0;
    };
    valid[0] = 0;
    for (uint64_t i = 1; i <= n; ++i) valid[i] = (b->W[i] % sigma) != 0;
    std::vector<std::pair<uint64_t, uint32_t>> stack;
    stack.emplace_back(ones.empty() ? 0 : ones[0], k);
    while (!stack.empty()) {
        auto [lst, depth] = stack.back();
        stack.pop_back();
        uint64_t e = lst;
        if (!e) continue;
        do {
            valid[e] = 0;
            uint32_t w = b->W[e] % sigma;
            if (e > 1 && w && depth > 1) stack.emplace_back(fwd(e, w), depth - 1);
        } while (--e && !b->last[e]);
    }
    return MGB_OK;
}

} // extern "C"
