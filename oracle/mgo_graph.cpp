// ORACLE — TEST INFRASTRUCTURE ONLY (see mgo_graph.hpp header).
#include "mgo_graph.hpp"
#include <cstdlib>

#include <cctype>
#include <stdexcept>
#include <unordered_set>

namespace mgo {

// ---------------------------------------------------------------------------
// Alphabets
// ---------------------------------------------------------------------------

const Alphabet& Alphabet::dna() {
    // kmer/alphabets.hpp:64-79: "$ACGT", invalid -> 5, 'U' == 'T'
    static const Alphabet a = [] {
        Alphabet x;
        x.letters = "$ACGT";
        x.sigma = 5;
        std::fill(x.char_to_code, x.char_to_code + 128, 5);
        const char *acgt = "ACGT";
        for (int i = 0; i < 4; ++i) {
            x.char_to_code[(int)acgt[i]] = i + 1;
            x.char_to_code[(int)tolower(acgt[i])] = i + 1;
        }
        x.char_to_code[(int)'U'] = x.char_to_code[(int)'u'] = 4;
        x.complement_code = { 0, 4, 3, 2, 1, 5 };
        x.bits_per_char = 3;
        return x;
    }();
    return a;
}

const Alphabet& Alphabet::protein() {
    // kmer/alphabets.hpp:29-38: "$ABCDEFGHIJKLMNOPQRSTUVWYZX", everything else -> 'X' (26)
    static const Alphabet a = [] {
        Alphabet x;
        x.letters = "$ABCDEFGHIJKLMNOPQRSTUVWYZX";
        x.sigma = 27;
        std::fill(x.char_to_code, x.char_to_code + 128, 26);
        for (int i = 1; i < 26; ++i) {
            char c = x.letters[i];
            x.char_to_code[(int)c] = i;
            x.char_to_code[(int)tolower(c)] = i;
        }
        x.bits_per_char = 5;
        return x;
    }();
    return a;
}

static const unsigned char* compl_table() {
    // common/seq_tools/reverse_complement.hpp:31-48, generated instead of spelled out:
    // identity except IUPAC pairs (both cases), 'U'->'A', and '`' -> '@'.
    static unsigned char tab[256];
    static bool init = [] {
        for (int i = 0; i < 256; ++i) tab[i] = i;
        const char *from = "ABCDGHKMRTUVY";
        const char *to   = "TVGHCDMKYAABR";
        for (int i = 0; from[i]; ++i) {
            tab[(int)from[i]] = to[i];
            tab[(int)tolower(from[i])] = tolower(to[i]);
        }
        tab[96] = 64;
        return true;
    }();
    (void)init;
    return tab;
}

char complement_char(char c) { return compl_table()[static_cast<unsigned char>(c)]; }

void reverse_complement_inplace(char *begin, char *end) {
    const unsigned char *tab = compl_table();
    int64_t n = end - begin;
    for (int64_t i = 0; i < n / 2; ++i) {
        char c0 = tab[(unsigned char)begin[i]];
        begin[i] = tab[(unsigned char)begin[n - 1 - i]];
        begin[n - 1 - i] = c0;
    }
    if (n & 1)
        begin[n / 2] = tab[(unsigned char)begin[n / 2]];
}

// ---------------------------------------------------------------------------
// SampledSeq
// ---------------------------------------------------------------------------
void SampledSeq::build(const uint8_t *data, uint64_t n, int num_symbols, int num_select_symbols) {
    data_ = data; n_ = n; nsym_ = num_symbols;
    uint64_t nblk = n / kBlock + 1;
    blk_.assign(nblk * nsym_, 0);
    sel_.assign(num_select_symbols, {});
    totals_.assign(nsym_, 0);
    std::vector<uint64_t> cnt(nsym_, 0);
    for (uint64_t i = 0; i < n; ++i) {
        if (i % kBlock == 0)
            std::copy(cnt.begin(), cnt.end(), blk_.begin() + (i / kBlock) * nsym_);
        uint8_t c = data[i];
        if (c < num_select_symbols && cnt[c] % kBlock == 0)
            sel_[c].push_back(i);
        ++cnt[c];
    }
    if (n % kBlock == 0)
        std::copy(cnt.begin(), cnt.end(), blk_.begin() + (n / kBlock) * nsym_);
    totals_ = cnt;
}

uint64_t SampledSeq::rank(uint8_t c, uint64_t i) const {
    assert(i < n_);
    uint64_t b = i / kBlock;
    uint64_t r = blk_[b * nsym_ + c];
    for (uint64_t j = b * kBlock; j <= i; ++j)
        r += data_[j] == c;
    return r;
}

uint64_t SampledSeq::select(uint8_t c, uint64_t r) const {
    assert(r >= 1 && r <= totals_[c]);
    uint64_t s = (r - 1) / kBlock;
    uint64_t pos = sel_[c][s];
    uint64_t have = s * kBlock + 1;   // occurrences up to and including pos
    while (have < r) {
        ++pos;
        have += data_[pos] == c;
    }
    return pos;
}

// ---------------------------------------------------------------------------
// BOSS construction
// ---------------------------------------------------------------------------
namespace {
typedef unsigned __int128 u128;

// 256-bit key for (k+1)-mers beyond 128 bits (the reference switches to KMer<sdsl::uint256_t>, kmer_boss.hpp):
// only the operations BOSS construction needs
struct U256 {
    uint64_t w[4];
    U256() : w{0, 0, 0, 0} {}
    U256(uint64_t x) : w{x, 0, 0, 0} {}
    explicit operator uint64_t() const { return w[0]; }
    friend U256 operator<<(const U256 &a, int n) {
        U256 r;
        if (n >= 256) return r;
        const int ws = n / 64, bs = n % 64;
        for (int i = 3; i >= ws; --i) {
            r.w[i] = a.w[i - ws] << bs;
            if (bs && i - ws - 1 >= 0) r.w[i] |= a.w[i - ws - 1] >> (64 - bs);
        }
        return r;
    }
    friend U256 operator>>(const U256 &a, int n) {
        U256 r;
        if (n >= 256) return r;
        const int ws = n / 64, bs = n % 64;
        for (int i = 0; i + ws < 4; ++i) {
            r.w[i] = a.w[i + ws] >> bs;
            if (bs && i + ws + 1 < 4) r.w[i] |= a.w[i + ws + 1] << (64 - bs);
        }
        return r;
    }
    friend U256 operator|(const U256 &a, const U256 &b) { U256 r; for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] | b.w[i]; return r; }
    friend U256 operator&(const U256 &a, const U256 &b) { U256 r; for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] & b.w[i]; return r; }
    friend U256 operator-(const U256 &a, const U256 &b) {
        U256 r; unsigned __int128 borrow = 0;
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - borrow;
            r.w[i] = (uint64_t)d; borrow = (d >> 64) & 1;
        }
        return r;
    }
    U256 operator~() const { U256 r; for (int i = 0; i < 4; ++i) r.w[i] = ~w[i]; return r; }
    friend bool operator==(const U256 &a, const U256 &b) { return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3]; }
    friend bool operator!=(const U256 &a, const U256 &b) { return !(a == b); }
    friend bool operator<(const U256 &a, const U256 &b) {
        for (int i = 3; i >= 0; --i) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i];
        return false;
    }
};

// Packed (k+1)-mer in KMerBOSS order (kmer/kmer_boss.hpp:58-64, 125-140):
// most significant = last node char a_k ... a_1, least significant = edge label.
template <class Key>
struct Packer {
    int bits; size_t k; // node length
    Key pack(const TAlphabet *x) const { // x[0..k] = a_1..a_k, label
        Key key = 0;
        for (size_t i = k; i-- > 0; )
            key = (key << bits) | x[i];
        return (key << bits) | x[k];
    }
    TAlphabet label(Key key) const { return (TAlphabet)(uint64_t)(key & Key((1u << bits) - 1)); }
    Key node(Key key) const { return key >> bits; } // a_k..a_1 (a_1 least significant)
    TAlphabet node_char(Key key, size_t i /*1-based a_i*/) const {
        return (TAlphabet)(uint64_t)((key >> (bits * i)) & Key((1u << bits) - 1));
    }
};
} // namespace

template <class Key>
static BOSS build_impl(const Alphabet &alph, size_t k, const std::vector<std::string> &seqs,
                       bool force_source_dummies) {
    const size_t K = k + 1;
    const int sigma = alph.sigma;
    Packer<Key> pk { alph.bits_per_char, k };
    const unsigned mask = (1u << pk.bits) - 1;

    // real (k+1)-mers; segments are split at invalid characters
    // (kmer/kmer_extractor.cpp:320-372)
    std::vector<Key> real;
    std::vector<Key> forced;
    for (const std::string &s : seqs) {
        std::vector<TAlphabet> enc = alph.encode(s);
        size_t i = 0;
        while (i < enc.size()) {
            size_t j = i;
            while (j < enc.size() && enc[j] < sigma) ++j;
            if (j - i >= K) {
                for (size_t p = i; p + K <= j; ++p)
                    real.push_back(pk.pack(enc.data() + p));
                if (force_source_dummies) {
                    std::vector<TAlphabet> buf(K, 0);
                    for (size_t d = 1; d <= k; ++d) {   // d leading sentinels
                        std::fill(buf.begin(), buf.begin() + d, 0);
                        std::copy(enc.begin() + i, enc.begin() + i + (K - d), buf.begin() + d);
                        forced.push_back(pk.pack(buf.data()));
                    }
                }
            }
            i = j + 1;
        }
    }
    std::sort(real.begin(), real.end());
    real.erase(std::unique(real.begin(), real.end()), real.end());

    // node sets for redundancy checks
    std::vector<Key> src_nodes(real.size()), tgt_nodes(real.size());
    for (size_t i = 0; i < real.size(); ++i) {
        src_nodes[i] = pk.node(real[i]);
        // target node a_2..a_{k+1}: drop a_1, append the label as the new last char
        Key n = pk.node(real[i]) >> pk.bits;
        tgt_nodes[i] = n | ((Key)pk.label(real[i]) << (pk.bits * (k - 1)));
    }
    // src_nodes is already sorted (node is the major key)
    src_nodes.erase(std::unique(src_nodes.begin(), src_nodes.end()), src_nodes.end());
    std::sort(tgt_nodes.begin(), tgt_nodes.end());
    tgt_nodes.erase(std::unique(tgt_nodes.begin(), tgt_nodes.end()), tgt_nodes.end());

    std::vector<Key> dummy;
    // dummy sink edges a_2..a_{k+1} -> $ (boss_chunk_construct.cpp:57-100)
    for (Key t : tgt_nodes) {
        if (!std::binary_search(src_nodes.begin(), src_nodes.end(), t))
            dummy.push_back(t << pk.bits);
    }
    // dummy source edges with one sentinel: $a_1..a_{k-1} -> a_k (:124-170)
    std::vector<Key> level;
    for (Key n : src_nodes) {
        if (!std::binary_search(tgt_nodes.begin(), tgt_nodes.end(), n)) {
            TAlphabet lbl = (TAlphabet)(uint64_t)((n >> (pk.bits * (k - 1))) & Key(mask)); // a_k
            Key prev_node = (n << pk.bits) & (((Key)1 << (pk.bits * k)) - 1); // $a_1..a_{k-1}
            level.push_back((prev_node << pk.bits) | lbl);
        }
    }
    std::sort(level.begin(), level.end());
    level.erase(std::unique(level.begin(), level.end()), level.end());
    // longer sentinel prefixes (:380-397)
    for (size_t c = 2; c < k + 1; ++c) {
        dummy.insert(dummy.end(), level.begin(), level.end());
        std::vector<Key> next;
        Key prev_n = ~(Key)0;
        for (Key key : level) {
            Key n = pk.node(key);
            if (n == prev_n) continue;
            prev_n = n;
            TAlphabet lbl = (TAlphabet)(uint64_t)((n >> (pk.bits * (k - 1))) & Key(mask));
            Key pn = (n << pk.bits) & (((Key)1 << (pk.bits * k)) - 1);
            next.push_back((pn << pk.bits) | lbl);
        }
        std::sort(next.begin(), next.end());
        next.erase(std::unique(next.begin(), next.end()), next.end());
        level.swap(next);
    }
    dummy.insert(dummy.end(), level.begin(), level.end());
    dummy.insert(dummy.end(), forced.begin(), forced.end());

    std::vector<Key> all;
    all.reserve(real.size() + dummy.size() + 1);
    all.push_back(0); // main dummy source $..$ -> $ (:404-409)
    all.insert(all.end(), real.begin(), real.end());
    all.insert(all.end(), dummy.begin(), dummy.end());
    std::sort(all.begin() + 1, all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());

    // W / last / F (boss_chunk.cpp:33-133)
    BOSS b;
    b.alph = &alph; b.alph_size = sigma; b.k_ = k;
    b.W.push_back(0); b.last.push_back(0);
    b.F.assign(sigma, 0);
    uint64_t curpos = 1;
    TAlphabet lastF = 0;
    std::vector<Key> last_kmer(sigma, 0);
    std::vector<bool> last_kmer_set(sigma, false);
    const Key minus1_mask = ~(((Key)1 << (2 * pk.bits)) - 1); // node chars a_2..a_k
    for (size_t it = 0; it < all.size(); ) {
        Key kmer = all[it];
        TAlphabet curW = pk.label(kmer);
        TAlphabet curF = pk.node_char(kmer, k);
        ++it;
        if (it < all.size() && pk.node(kmer) == pk.node(all[it])) {
            if (curW == 0 && curF > 0)
                continue; // redundant dummy sink
            b.last.push_back(0);
        } else {
            b.last.push_back(1);
        }
        if (curW) {
            // KMER(0).data() == 0 acts as "unset" in the reference
            if (last_kmer_set[curW] && last_kmer[curW] != 0
                    && (kmer & minus1_mask) == (last_kmer[curW] & minus1_mask)) {
                curW += sigma;
            } else {
                last_kmer[curW] = kmer;
                last_kmer_set[curW] = true;
            }
        }
        b.W.push_back(curW);
        while (curF > lastF && lastF + 1 < sigma)
            b.F[++lastF] = curpos - 1;
        ++curpos;
    }
    while (++lastF < sigma)
        b.F[lastF] = curpos - 1;
    b.finalize();
    return b;
}

BOSS BOSS::build(const Alphabet &alph, size_t k, const std::vector<std::string> &seqs,
                 bool force_source_dummies) {
    // MGO_FORCE_U256=1 sends every graph through the wide-key instantiation (cross-check of the two)
    if ((k + 1) * alph.bits_per_char <= 128 && !std::getenv("MGO_FORCE_U256"))
        return build_impl<u128>(alph, k, seqs, force_source_dummies);
    if ((k + 1) * alph.bits_per_char <= 256)
        return build_impl<U256>(alph, k, seqs, force_source_dummies);
    throw std::runtime_error("oracle BOSS::build: k too large for packed k-mers");
}

BOSS BOSS::from_arrays(const Alphabet &alph, size_t k, std::vector<uint8_t> &&W,
                       std::vector<uint8_t> &&last, const std::vector<uint64_t> &F) {
    BOSS b;
    b.alph = &alph; b.alph_size = alph.sigma; b.k_ = k;
    b.W = std::move(W); b.last = std::move(last); b.F = F;
    b.finalize();
    return b;
}

void BOSS::finalize() {
    W_rs_.build(W.data(), W.size(), 2 * alph_size, alph_size);
    last_rs_.build(last.data(), last.size(), 2, 2);
    NF.resize(F.size());
    for (size_t c = 0; c < F.size(); ++c)   // boss.cpp:1095-1101
        NF[c] = rank_last(F[c]);
}

void BOSS::index_suffix_ranges(size_t s) {
    // Semantics of boss_chunk_construct.cpp:260-320: for every string over the
    // sigma-1 real symbols of length s (co-lex rank), the half-open edge range
    // of nodes ending with it. Computed here by running the un-indexed search.
    indexed_suffix_length_ = 0;
    suffix_ranges_.clear();
    if (!s || s > k_) return;
    uint64_t num = 1;
    for (size_t i = 0; i < s; ++i) num *= (alph_size - 1);
    std::vector<uint64_t> ranges(2 * num, 0);
    std::vector<TAlphabet> str(s);
    for (uint64_t idx = 0; idx < num; ++idx) {
        uint64_t x = idx;
        // index = sum (c_j - 1) * (sigma-1)^j  with j = position in the string
        for (size_t j = 0; j < s; ++j) { str[j] = x % (alph_size - 1) + 1; x /= (alph_size - 1); }
        auto [rl, ru, off] = get_initial_range(str.data(), str.data() + 1);
        bool ok = rl <= ru;
        for (size_t j = off; ok && j < s; ++j)
            ok = tighten_range(&rl, &ru, str[j]);
        if (ok) {
            // stored as the edge range of those nodes
            ranges[2 * idx] = pred_last(rl - 1) + 1;
            ranges[2 * idx + 1] = ru + 1;
        }
    }
    // get_initial_range consumes [rl, ru] as ranges of *edges* whose source nodes match
    // (boss.hpp:655-663); empty ranges have begin >= end.
    for (uint64_t idx = 0; idx < num; ++idx) {
        if (ranges[2 * idx] >= ranges[2 * idx + 1]) { ranges[2 * idx] = 1; ranges[2 * idx + 1] = 1; }
    }
    suffix_ranges_ = std::move(ranges);
    indexed_suffix_length_ = s;
}

// ---------------------------------------------------------------------------
// BOSS primitives
// ---------------------------------------------------------------------------
uint64_t BOSS::rank_W(edge_index i, TAlphabet c) const {
    return i == 0 ? 0 : W_rs_.rank(c, i) - (c == 0);
}

edge_index BOSS::succ_W(edge_index i, TAlphabet c) const {
    // wavelet_tree::next: first position >= i holding c, size() if none
    for (; i < W.size(); ++i)
        if (W[i] == c) return i;
    return W.size();
}

std::pair<edge_index, TAlphabet> BOSS::succ_W(edge_index i, TAlphabet c1, TAlphabet c2) const {
    for (; i < W.size(); ++i) {
        if (W[i] == c1) return { i, c1 };
        if (W[i] == c2) return { i, c2 };
    }
    return { W.size(), 0 };
}

uint64_t BOSS::rank_last(edge_index i) const { return i == 0 ? 0 : last_rs_.rank(1, i); }
edge_index BOSS::select_last(uint64_t i) const { return i == 0 ? 0 : last_rs_.select(1, i); }

edge_index BOSS::pred_last(edge_index i) const {
    for (; i > 0; --i)
        if (last[i]) return i;
    return 0;
}

edge_index BOSS::succ_last(edge_index i) const {
    for (; i < last.size(); ++i)
        if (last[i]) return i;
    return last.size();
}

edge_index BOSS::bwd(edge_index i) const {
    uint64_t target_node = rank_last(i - 1) + 1;
    if (target_node == 1)
        return 1;
    TAlphabet c = get_node_last_value(i);
    return W_rs_.select(c, target_node - NF[c]);
}

edge_index BOSS::fwd(edge_index i, TAlphabet c) const {
    return select_last(NF[c] + rank_W(i, c));
}

TAlphabet BOSS::get_node_last_value(edge_index i) const {
    if (i == 0) return 0;
    for (TAlphabet c = 0; c < alph_size; ++c)
        if (F[c] >= i) return c - 1;
    return alph_size - 1;
}

std::pair<TAlphabet, edge_index> BOSS::get_minus_k_value(edge_index i, size_t k) const {
    for (; k > 0; --k)
        i = bwd(i);
    return { get_node_last_value(i), bwd(i) };
}

edge_index BOSS::pick_edge(edge_index edge, TAlphabet c) const {
    do {
        TAlphabet w = get_W(edge);
        if (w == c || w == c + alph_size)
            return edge;
    } while (--edge && !get_last(edge));
    return npos;
}

bool BOSS::is_single_incoming(edge_index i, TAlphabet w) const {
    if (w > alph_size) return false;
    ++i;
    return i == W.size() || succ_W(i, w, w + alph_size).second != w + alph_size;
}

std::tuple<edge_index, edge_index, size_t>
BOSS::get_initial_range(const TAlphabet *begin, const TAlphabet *end) const {
    edge_index rl, ru; size_t offset;
    if (indexed_suffix_length_ && begin + indexed_suffix_length_ <= end
            && !std::count(begin, begin + indexed_suffix_length_, 0)) {
        uint64_t index = 0;
        for (const TAlphabet *it = begin + indexed_suffix_length_ - 1; it != begin - 1; --it)
            index = index * (alph_size - 1) + (*it - 1);
        rl = suffix_ranges_[2 * index];
        ru = suffix_ranges_[2 * index + 1] - 1;
        offset = indexed_suffix_length_;
    } else {
        TAlphabet s = *begin;
        rl = F.at(s) + 1 < W.size() ? F.at(s) + 1 : W.size();
        ru = s + 1 < alph_size ? F[s + 1] : W.size() - 1;
        offset = 1;
    }
    return { rl, ru, offset };
}

bool BOSS::tighten_range(edge_index *rl, edge_index *ru, TAlphabet s) const {
    uint64_t rk_rl = rank_W(*rl - 1, s) + 1;
    uint64_t rk_ru = rank_W(*ru, s);
    if (rk_rl > rk_ru)
        return false;
    *rl = select_last(NF[s] + rk_rl - 1) + 1;
    *ru = select_last(NF[s] + rk_ru);
    return true;
}

edge_index BOSS::index(const TAlphabet *begin, const TAlphabet *end) const {
    if (std::find(begin, end, (TAlphabet)alph_size) != end)
        return npos;
    auto [rl, ru, offset] = get_initial_range(begin, end);
    if (rl > ru)
        return npos;
    for (const TAlphabet *it = begin + offset; it != end; ++it)
        if (!tighten_range(&rl, &ru, *it))
            return npos;
    return ru;
}

std::tuple<edge_index, edge_index, const TAlphabet*>
BOSS::index_range(const TAlphabet *begin, const TAlphabet *end) const {
    if (begin == end)
        return { 1, 1, begin };
    if (std::find(begin, end, (TAlphabet)alph_size) != end)
        return { 0, 0, begin };
    auto [rl, ru, offset] = get_initial_range(begin, end);
    if (rl > ru) {
        TAlphabet s = *begin;
        rl = F.at(s) + 1 < W.size() ? F.at(s) + 1 : W.size();
        ru = s + 1 < alph_size ? F[s + 1] : W.size() - 1;
        if (rl > ru)
            return { 0, 0, begin };
        offset = 1;
    }
    const TAlphabet *it = begin + offset;
    for (; it != end; ++it)
        if (!tighten_range(&rl, &ru, *it))
            return { succ_last(rl), ru, it };
    return { succ_last(rl), ru, it };
}

edge_index BOSS::map_to_edge(const TAlphabet *begin, const TAlphabet *end) const {
    edge_index edge = index(begin, end - 1);
    return edge && *(end - 1) < alph_size ? pick_edge(edge, *(end - 1)) : npos;
}

std::vector<edge_index> BOSS::map_to_edges(const std::vector<TAlphabet> &seq) const {
    std::vector<edge_index> out;
    if (seq.size() <= k_)
        return out;
    out.reserve(seq.size() - k_);
    // common/algorithms.hpp:58-74 drag_and_mark_segments(seq, alph_size, k+1)
    std::vector<uint8_t> invalid(seq.size(), 0);
    {
        size_t last_occ = std::find(seq.begin(), seq.end(), (TAlphabet)alph_size) - seq.begin();
        for (size_t i = last_occ; i < seq.size(); ++i) {
            if (seq[i] == alph_size) last_occ = i;
            if (i - last_occ < k_ + 1) invalid[i] = 1;
        }
    }
    for (size_t i = 0; i + k_ + 1 <= seq.size(); ++i) {
        if (invalid[i + k_]) { out.push_back(npos); continue; }
        edge_index edge = map_to_edge(seq.data() + i, seq.data() + i + k_ + 1);
        out.push_back(edge);
        while (edge && ++i + k_ < seq.size()) {
            if (invalid[i + k_]) { out.push_back(npos); break; }
            edge = fwd(edge, seq[i + k_ - 1]);
            edge = pick_edge(edge, seq[i + k_]);
            out.push_back(edge);
        }
    }
    return out;
}

std::vector<TAlphabet> BOSS::get_node_seq(edge_index x) const {
    std::vector<TAlphabet> ret(k_);
    size_t i = k_;
    ret[--i] = get_node_last_value(x);
    while (i > 0) {
        x = bwd(x);
        ret[--i] = get_node_last_value(x);
    }
    return ret;
}

std::string BOSS::get_node_str(edge_index i) const {
    std::string s;
    for (TAlphabet c : get_node_seq(i)) s += alph->decode(c);
    return s;
}

// ---------------------------------------------------------------------------
// DBGSuccinct
// ---------------------------------------------------------------------------
void DBGSuccinct::mask_dummy_kmers() {
    // boss.cpp:1736-1775 mark_all_dummy_edges, flipped (dbg_succinct.cpp:903-908):
    // an edge is dummy iff its (k+1)-mer contains '$'.
    uint64_t n = boss.num_edges();
    valid_edges.assign(n + 1, 1);
    valid_edges[0] = 0;
    for (uint64_t i = 1; i <= n; ++i)
        if (boss.get_W(i) % boss.alph_size == 0) valid_edges[i] = 0;
    // source dummies: walk the sentinel tree from $$$ (edges of node 1)
    std::vector<std::pair<edge_index, size_t>> stack; // (last edge of node, #leading sentinels)
    stack.emplace_back(boss.succ_last(1), boss.k_);
    while (!stack.empty()) {
        auto [lst, depth] = stack.back();
        stack.pop_back();
        edge_index e = lst;
        do {
            valid_edges[e] = 0;
            TAlphabet w = boss.get_W(e) % boss.alph_size;
            if (e > 1 && w && depth > 1)
                stack.emplace_back(boss.fwd(e, w), depth - 1);
        } while (--e && !boss.get_last(e));
    }
}

uint64_t DBGSuccinct::num_nodes() const {
    if (valid_edges.empty()) return boss.num_edges();
    uint64_t c = 0;
    for (uint8_t v : valid_edges) c += v;
    return c;
}

std::vector<node_index> DBGSuccinct::map_to_nodes_sequentially(std::string_view seq) const {
    if (seq.size() < get_k())
        return {};
    std::vector<edge_index> edges = boss.map_to_edges(boss.alph->encode(seq));
    for (auto &e : edges) e = validate_edge(e);
    return edges;
}

bool DBGSuccinct::has_multiple_outgoing(node_index node) const {
    if (node == 1)
        return boss.succ_last(1) > 2;
    TAlphabet d = boss.get_W(node) % boss.alph_size;
    if (!d) return false;
    return !boss.get_last(boss.fwd(node, d) - 1);
}

bool DBGSuccinct::has_single_incoming(node_index node) const {
    if (node == 1) return false;
    edge_index x = boss.bwd(node);
    TAlphabet w = boss.get_node_last_value(node);
    size_t first_valid = valid_edges.empty() || valid_edges[x];
    if (x + 1 == boss.W.size())
        return first_valid;
    if (first_valid)
        return boss.is_single_incoming(x, w);
    size_t indeg = 0;
    boss.call_incoming_to_target(x, w, [&](edge_index) { ++indeg; });
    return indeg == 2;
}

std::string DBGSuccinct::get_node_sequence(node_index node) const {
    return boss.get_node_str(node) + boss.alph->decode(boss.get_W(node) % boss.alph_size);
}

node_index DBGSuccinct::traverse(node_index node, char c) const {
    TAlphabet s = boss.alph->encode(c);
    if (s == boss.alph_size) return npos;
    TAlphabet w = boss.get_W(node) % boss.alph_size;
    if (node > 1 && !w) return npos;
    edge_index e = boss.fwd(node, w);
    return validate_edge(boss.pick_edge(e, s));
}

void DBGSuccinct::call_nodes_with_suffix_matching_longest_prefix(
        std::string_view str, const std::function<void(node_index, uint64_t)> &cb,
        size_t min_match_length) const {
    if (str.size() < min_match_length)
        return;
    std::vector<TAlphabet> encoded = boss.alph->encode(str);
    if (std::find(encoded.begin(), encoded.end(), (TAlphabet)boss.alph_size) != encoded.end())
        return;
    const TAlphabet *b = encoded.data();
    const TAlphabet *e = std::min(b + get_k() - 1, b + encoded.size());
    auto [first, lst, end] = boss.index_range(b, e);
    size_t match_size = end - b;
    if (str.size() == get_k() && match_size + 1 == get_k()) {
        edge_index edge = boss.pick_edge(lst, encoded.back());
        if (edge && in_graph(edge)) {
            cb(edge, get_k());
            return;
        }
    }
    if (match_size < min_match_length)
        return;
    uint64_t rank_first = boss.rank_last(first);
    uint64_t rank_lst = boss.rank_last(lst);
    for (uint64_t i = rank_first; i <= rank_lst; ++i) {
        edge_index ee = boss.select_last(i);
        boss.call_incoming_to_target(boss.bwd(ee), boss.get_node_last_value(ee),
            [&](edge_index in) {
                if (in_graph(in))
                    cb(in, match_size);
            });
    }
}

// ---------------------------------------------------------------------------
// CanonicalDBG (graph/representation/canonical_dbg.cpp)
// ---------------------------------------------------------------------------
node_index CanonicalDBG::reverse_complement(node_index node) const {   // :521-553
    if (node > offset_)
        return node - offset_;
    if (k_odd_)
        return node + offset_;
    std::string seq = get_node_sequence(node);
    std::string rev_seq = seq;
    reverse_complement_inplace(rev_seq);
    return rev_seq == seq ? node : node + offset_;
}

void CanonicalDBG::reverse_complement(std::string &seq, std::vector<node_index> &path) const {   // :555-565
    reverse_complement_inplace(seq);
    std::vector<node_index> rev_path(path.size());
    std::transform(path.begin(), path.end(), rev_path.rbegin(),
                   [&](node_index i) { return i ? reverse_complement(i) : i; });
    std::swap(path, rev_path);
}

std::string CanonicalDBG::get_node_sequence(node_index index) const {   // :423-432
    node_index node = get_base_node(index);
    std::string seq = g.get_node_sequence(node);
    if (node != index)
        reverse_complement_inplace(seq);
    return seq;
}

// :55-146. The forward strand is mapped up to its first miss; from there on every k-mer is looked up on both
// strands. For odd k (no palindromes) the reference skips the forward lookup of k-mers whose reverse
// complement was found (and of the k-mer the first pass stopped on), for even k the forward hit wins.
std::vector<node_index> CanonicalDBG::map_to_nodes_sequentially(std::string_view sequence) const {
    std::vector<node_index> out;
    if (sequence.size() < get_k())
        return out;
    std::vector<node_index> fwd_all = g.map_to_nodes_sequentially(sequence);
    size_t prefix = 0;
    while (prefix < fwd_all.size() && fwd_all[prefix])
        ++prefix;
    out.assign(fwd_all.begin(), fwd_all.begin() + prefix);
    sequence = sequence.substr(prefix);
    if (sequence.size() < get_k())
        return out;
    std::string rev_seq(sequence);
    reverse_complement_inplace(rev_seq);
    std::vector<node_index> rev_path = g.map_to_nodes_sequentially(rev_seq);
    std::vector<node_index> path(fwd_all.begin() + prefix, fwd_all.end());
    auto it = rev_path.rbegin();
    for (size_t j = 0; j < path.size(); ++j, ++it) {
        if (k_odd_ && (j == 0 || *it))
            path[j] = npos;
        if (path[j] != npos) out.push_back(path[j]);
        else if (*it != npos) out.push_back(*it + offset_);
        else out.push_back(npos);
    }
    return out;
}

// node_first_cache.cpp:122-148 get_prefix_rc / :150-176 get_suffix_rc, uncached
static edge_index rc_range_edge(const BOSS &boss, std::string rev_seq, bool upper) {
    if (rev_seq[0] == '$')
        return 0;
    reverse_complement_inplace(rev_seq);
    std::vector<TAlphabet> encoded = boss.alph->encode(rev_seq);
    auto [e1, e2, end] = boss.index_range(encoded.data(), encoded.data() + encoded.size());
    if (end != encoded.data() + encoded.size())
        return 0;
    return upper ? e2 : e1;
}

void CanonicalDBG::adjacent_incoming_rc_strand(node_index, const std::string &spelling,
                                               const std::function<void(node_index, char)> &cb) const {
    const BOSS &boss = g.boss;
    edge_index rc_edge = rc_range_edge(boss, spelling.substr(0, spelling.size() - 1), true);
    if (!rc_edge)
        return;
    // BOSS::call_outgoing (boss.hpp:779-784): the edges of the node, last one first
    edge_index edge = rc_edge;
    do {
        node_index prev = edge;
        if (g.in_graph(prev)) {
            char c = boss.alph->decode(boss.get_W(edge) % boss.alph_size);
            if (!(spelling.back() == '$' && c == '$'))
                cb(prev, c);
        }
    } while (--edge && !boss.get_last(edge));
}

void CanonicalDBG::adjacent_outgoing_rc_strand(node_index, const std::string &spelling,
                                               const std::function<void(node_index, char)> &cb) const {
    const BOSS &boss = g.boss;
    edge_index rc_edge = rc_range_edge(boss, spelling.substr(1), false);
    if (!rc_edge)
        return;
    // NodeFirstCache::call_incoming_edges + get_first_char (node_first_cache.cpp:9-36)
    boss.call_incoming_to_target(boss.bwd(rc_edge), boss.get_node_last_value(rc_edge),
        [&](edge_index prev_edge) {
            node_index prev = prev_edge;
            if (!g.in_graph(prev))
                return;
            char c = boss.alph->decode(boss.get_minus_k_value(prev_edge, boss.k_ - 1).first);
            if (spelling[0] == '$' && c == '$')
                return;
            cb(prev, c);
        });
}

// alphabet_encoder_ (canonical_dbg.cpp:31-38): position in the graph's alphabet string, '$' included
static inline TAlphabet alphabet_pos(const Alphabet &alph, char c) { return c == '$' ? 0 : alph.encode(c); }

void CanonicalDBG::call_outgoing_kmers(node_index node, const std::function<void(node_index, char)> &cb) const {
    if (node > offset_) {
        call_incoming_kmers(node - offset_, [&](node_index next, char c) {
            cb(reverse_complement(next), complement_char(c));
        });
        return;
    }
    const std::string spelling = get_node_sequence(node);
    const Alphabet &alph = *g.boss.alph;
    std::vector<node_index> children(alph.sigma, npos);
    size_t max_num_edges_left = children.size() - has_sentinel_;
    g.call_outgoing_kmers(node, [&](node_index next, char c) {
        if (c != '$') {
            cb(next, c);
            --max_num_edges_left;
        }
        children[alphabet_pos(alph, c)] = next;
    });
    if (!max_num_edges_left)
        return;
    adjacent_outgoing_rc_strand(node, spelling, [&](node_index next, char c) {
        c = complement_char(c);
        TAlphabet s = alphabet_pos(alph, c);
        if (children[s] != npos && c != '$') {
            if (k_odd_)
                throw std::runtime_error("Forward traversal: Primary graph contains both forward and reverse complement");
            return;                                   // `next` is a palindrome
        }
        next = reverse_complement(next);
        if (c != '$') {
            cb(next, c);
            children[s] = next;
            --max_num_edges_left;
        }
    });
    if (has_sentinel_ && children[alphabet_pos(alph, '$')] && max_num_edges_left + 1 == (size_t)alph.sigma)
        cb(children[alphabet_pos(alph, '$')], '$');
}

void CanonicalDBG::call_incoming_kmers(node_index node, const std::function<void(node_index, char)> &cb) const {
    if (node > offset_) {
        call_outgoing_kmers(node - offset_, [&](node_index prev, char c) {
            cb(reverse_complement(prev), complement_char(c));
        });
        return;
    }
    const std::string spelling = get_node_sequence(node);
    const Alphabet &alph = *g.boss.alph;
    std::vector<node_index> parents(alph.sigma, npos);
    size_t max_num_edges_left = parents.size() - has_sentinel_;
    g.call_incoming_kmers(node, [&](node_index prev, char c) {
        if (c != '$') {
            cb(prev, c);
            --max_num_edges_left;
        }
        parents[alphabet_pos(alph, c)] = prev;
    });
    if (!max_num_edges_left)
        return;
    adjacent_incoming_rc_strand(node, spelling, [&](node_index prev, char c) {
        c = complement_char(c);
        TAlphabet s = alphabet_pos(alph, c);
        if (parents[s] != npos && c != '$') {
            if (k_odd_)
                throw std::runtime_error("Backward traversal: Primary graph contains both forward and reverse complement");
            return;
        }
        prev = reverse_complement(prev);
        if (c != '$') {
            cb(prev, c);
            parents[s] = prev;
            --max_num_edges_left;
        }
    });
    if (has_sentinel_ && parents[alphabet_pos(alph, '$')] && max_num_edges_left + 1 == (size_t)alph.sigma)
        cb(parents[alphabet_pos(alph, '$')], '$');
}

void CanonicalDBG::adjacent_incoming_nodes(node_index node, const std::function<void(node_index)> &cb) const {
    if (node > offset_) {
        adjacent_outgoing_nodes(node - offset_, [&](node_index prev) { cb(reverse_complement(prev)); });
    } else {
        call_incoming_kmers(node, [&](node_index prev, char) { cb(prev); });
    }
}

void CanonicalDBG::adjacent_outgoing_nodes(node_index node, const std::function<void(node_index)> &cb) const {
    if (node > offset_) {
        adjacent_incoming_nodes(node - offset_, [&](node_index next) { cb(reverse_complement(next)); });
    } else {
        call_outgoing_kmers(node, [&](node_index next, char) { cb(next); });
    }
}

bool CanonicalDBG::has_multiple_outgoing(node_index node) const {
    size_t n = 0;
    adjacent_outgoing_nodes(node, [&](node_index) { ++n; });
    return n > 1;
}

bool CanonicalDBG::has_single_incoming(node_index node) const {
    size_t n = 0;
    adjacent_incoming_nodes(node, [&](node_index) { ++n; });
    return n == 1;
}

} // namespace mgo
