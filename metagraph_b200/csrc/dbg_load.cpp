// Loader for the reference's on-disk graph format (`.dbg`, DBGSuccinct::serialize,
// graph/representation/succinct/dbg_succinct.cpp:690-803 -> BOSS::serialize, boss.cpp:262-277, load
// :338-390) into the plain W / last / F arrays that mgb_index_create() flattens for the GPU.
// Host-side, index build time; not part of the timed path.
//
// File layout (numbers written by serialize_number are big-endian u64, common/serialization.cpp:30-57;
// everything inside the sdsl containers is little-endian):
//   F[]            u64 count + values                       (serialize_number_vector_raw)
//   k_             BOSS node length (DBG k = k_ + 1)
//   state          BOSS::State: SMALL = 1, DYN = 2, STAT = 3, FAST = 4   (boss.hpp:325)
//   W              SMALL: wt_huff<rrr_vector<63>>; STAT: wt_huff<bit_vector>; + logsigma (BE)
//                  (wavelet_tree.hpp:226-230, wavelet_tree.cpp:390-419)
//   last           SMALL: bit_vector_small = representation code (BE) + rrr_vector<63> | sd_vector
//                  (bit_vector_adaptive.hpp:48-56, 104-133); STAT: bit_vector + #ones (BE) +
//                  rank_support_v5 + select_support_mcl<1> (bit_vector_sdsl.hpp:240-281, 450-454)
//   mode           DeBruijnGraph::Mode: BASIC = 0, CANONICAL, PRIMARY   (sequence_graph.hpp:160)
//   suffix ranges  length + sd_vector (boss.cpp:392-420) — not read: mgb_index_create builds its own
//
// sdsl-lite (hmusta fork) is not vendored in the reference tree, so the container encodings are restated
// from the library's published serialisation and PINNED on the two graphs the reference ships
// (examples/data/graphs/test_DNA_graph.dbg, test_Protein_graph.dbg, both SMALL state): the decoded
// tables equal a fresh construction from examples/data/test_*_sequences.fa (tests/test_dbg_loader.py).
// That pins: the header, wt_huff's byte_tree, rrr_vector<63> (block classes, offsets in the
// combinatorial number system, blocks with more than half the bits set stored complemented). The STAT
// containers (plain bit_vector, rank_support_v/v5, select_support_mcl) and sd_vector-coded `last`
// follow the same library's layout but have no fixture: parity unpinned; the loader cross-checks F
// against the decoded arrays and refuses files that do not add up. DYN and FAST states are refused.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mgb.h"

namespace {

thread_local std::string g_load_err;

struct Reader {
    std::vector<uint8_t> b;
    size_t p = 0;
    // overflow-safe: a corrupt length field must not wrap around the bound
    void need(uint64_t n) const { if (n > b.size() - p) throw std::runtime_error("unexpected end of file"); }
    // number of 64-bit words of a vector of `bits` bits, checked against what is left of the file
    uint64_t words_of(uint64_t bits) const {
        const uint64_t nw = bits / 64 + (bits % 64 != 0);
        if (nw > (b.size() - p) / 8) throw std::runtime_error("vector longer than the file");
        return nw;
    }
    uint64_t be() { need(8); uint64_t v = 0; for (int i = 0; i < 8; ++i) v = (v << 8) | b[p + i]; p += 8; return v; }
    uint64_t le() { need(8); uint64_t v = 0; for (int i = 7; i >= 0; --i) v = (v << 8) | b[p + i]; p += 8; return v; }
    uint8_t u8() { need(1); return b[p++]; }
    uint16_t le16() { need(2); uint16_t v = (uint16_t)(b[p] | (b[p + 1] << 8)); p += 2; return v; }
    void skip(size_t n) { need(n); p += n; }
};

// sdsl::int_vector<1>: size in bits, then ceil(size / 64) words
struct Bits {
    uint64_t size = 0;
    std::vector<uint64_t> w;
    bool get(uint64_t i) const { return (w[i >> 6] >> (i & 63)) & 1u; }
    uint64_t get_int(uint64_t pos, unsigned len) const {     // len <= 64, bits [pos, pos + len)
        if (!len) return 0;
        if (len > 64 || (pos >> 6) + 1 >= w.size()) throw std::runtime_error("bit vector read out of range");
        uint64_t lo = w[pos >> 6] >> (pos & 63);
        unsigned got = 64 - (unsigned)(pos & 63);
        if (got < len) lo |= w[(pos >> 6) + 1] << got;
        return len == 64 ? lo : (lo & ((1ull << len) - 1));
    }
};
Bits read_bits(Reader &r) {
    Bits v; v.size = r.le();
    const uint64_t nw = r.words_of(v.size);
    v.w.resize(nw + 1, 0);
    for (uint64_t i = 0; i < nw; ++i) v.w[i] = r.le();
    return v;
}
// sdsl::int_vector<0>: size in bits, width byte, data
struct Ints {
    Bits bits; unsigned width = 0;
    uint64_t size() const { return width ? bits.size / width : 0; }
    uint64_t operator[](uint64_t i) const { return bits.get_int(i * width, width); }
};
Ints read_ints(Reader &r) {
    Ints v;
    const uint64_t sz = r.le();
    v.width = r.u8();
    if (v.width < 1 || v.width > 64) throw std::runtime_error("int_vector width " + std::to_string(v.width));
    v.bits.size = sz;
    const uint64_t nw = r.words_of(sz);
    v.bits.w.resize(nw + 1, 0);
    for (uint64_t i = 0; i < nw; ++i) v.bits.w[i] = r.le();
    return v;
}
void skip_ints64(Reader &r) { const uint64_t sz = r.le(); r.skip(r.words_of(sz) * 8); }   // int_vector<64>
// sdsl::select_support_mcl: #args; if any: superblocks, a flag vector, one vector per 4096 args
void skip_select_mcl(Reader &r) {
    const uint64_t cnt = r.le();
    if (!cnt) return;
    read_ints(r);
    const uint64_t sb = cnt / 4096 + (cnt % 4096 != 0);
    Bits mini_or_long = read_bits(r);
    (void)mini_or_long;
    for (uint64_t i = 0; i < sb; ++i) read_ints(r);
}

// --- rrr_vector<63> -------------------------------------------------------------------------------
constexpr int kRrrBlock = 63;
struct Binomial {
    uint64_t c[kRrrBlock + 1][kRrrBlock + 1];
    unsigned space[kRrrBlock + 1];
    Binomial() {
        for (int n = 0; n <= kRrrBlock; ++n)
            for (int k = 0; k <= kRrrBlock; ++k)
                c[n][k] = k == 0 ? 1 : (n == 0 ? 0 : c[n - 1][k - 1] + c[n - 1][k]);
        for (int k = 0; k <= kRrrBlock; ++k) {
            const uint64_t v = c[kRrrBlock][k];
            space[k] = (k == 0 || k == kRrrBlock) ? 0 : 64 - (unsigned)__builtin_clzll(v);
        }
    }
};
const Binomial& binom() { static const Binomial b; return b; }
// offset -> block in the combinatorial number system: scanning positions upwards, the patterns with
// the position clear come first (rrr_helper::bin_to_nr adds C(remaining - 1, k) for a set bit)
uint64_t rrr_pattern(unsigned k, uint64_t nr) {
    const Binomial &B = binom();
    uint64_t bits = 0;
    for (int pos = 0; pos < kRrrBlock && k; ++pos) {
        const uint64_t c = B.c[kRrrBlock - pos - 1][k];
        if (nr >= c) { bits |= 1ull << pos; nr -= c; --k; }
    }
    return bits;
}
Bits read_rrr(Reader &r) {
    const uint64_t size = r.le();
    Ints bt = read_ints(r);
    Bits btnr = read_bits(r);
    read_ints(r);                // sampled pointers into btnr
    read_ints(r);                // sampled ranks
    if (size / kRrrBlock > bt.size()) throw std::runtime_error("rrr_vector: block class array too short");
    Bits out; out.size = size; out.w.assign(size / 64 + 3, 0);
    const Binomial &B = binom();
    uint64_t pos = 0;
    const uint64_t nblocks = size / kRrrBlock + (size % kRrrBlock != 0);
    if (bt.size() < nblocks) throw std::runtime_error("rrr_vector: block class array too short");
    const uint64_t full = (1ull << kRrrBlock) - 1;
    for (uint64_t i = 0; i < nblocks; ++i) {
        const unsigned k = (unsigned)bt[i];
        if (k > (unsigned)kRrrBlock) throw std::runtime_error("rrr_vector: bad block class");
        const unsigned sp = B.space[k];
        if (pos > btnr.size || sp > btnr.size - pos) throw std::runtime_error("rrr_vector: offsets overrun");
        const uint64_t nr = btnr.get_int(pos, sp);
        pos += sp;
        // blocks with more than half of the bits set are stored complemented
        const uint64_t pat = 2 * k > (unsigned)kRrrBlock ? (full ^ rrr_pattern(kRrrBlock - k, nr)) : rrr_pattern(k, nr);
        const uint64_t at = i * kRrrBlock;
        out.w[at >> 6] |= pat << (at & 63);
        if ((at & 63) && (at & 63) + kRrrBlock > 64) out.w[(at >> 6) + 1] |= pat >> (64 - (at & 63));
    }
    for (uint64_t i = size; i < (uint64_t)out.w.size() * 64 && i < nblocks * kRrrBlock; ++i)
        if (out.get(i)) throw std::runtime_error("rrr_vector: bits beyond the end");
    return out;
}

// --- sd_vector --------------------------------------------------------------------------------------
Bits read_sd(Reader &r) {
    const uint64_t size = r.le();
    const unsigned wl = r.u8();
    Ints low = read_ints(r);
    Bits high = read_bits(r);
    skip_select_mcl(r); skip_select_mcl(r);
    if (wl > 63) throw std::runtime_error("sd_vector: bad low width");
    if (size / 64 > (uint64_t)r.b.size() * 64) throw std::runtime_error("sd_vector: length field larger than the file allows");
    Bits out; out.size = size; out.w.assign(size / 64 + 2, 0);
    uint64_t ones = 0;
    for (uint64_t i = 0; i < high.size; ++i) {
        if (!high.get(i)) continue;
        const uint64_t hi = i - ones;                    // number of zeros before this one
        if (ones >= low.size()) throw std::runtime_error("sd_vector: low part too short");
        if (wl && (hi >> (64 - wl))) throw std::runtime_error("sd_vector: position out of range");
        const uint64_t v = (hi << wl) | low[ones];
        if (v >= size) throw std::runtime_error("sd_vector: position out of range");
        out.w[v >> 6] |= 1ull << (v & 63);
        ++ones;
    }
    return out;
}

// --- wt_huff ---------------------------------------------------------------------------------------
struct WtNode { uint64_t bv_pos, bv_pos_rank; uint16_t parent, child[2]; };
void wt_decode(const Bits &bv, const std::vector<WtNode> &nodes, uint16_t v, const std::vector<uint64_t> &idx,
               std::vector<uint8_t> &out, unsigned depth = 0) {
    // a Huffman tree over <= 2 * 27 symbols is at most that deep: anything deeper is a cycle in a corrupt file
    if (depth > 64) throw std::runtime_error("wt_huff: tree too deep (cycle?)");
    const WtNode &nd = nodes.at(v);
    if (nd.child[0] == 0xffff) {                         // leaf: bv_pos_rank holds the symbol
        for (uint64_t i : idx) out[i] = (uint8_t)nd.bv_pos_rank;
        return;
    }
    std::vector<uint64_t> l, rr;
    for (uint64_t j = 0; j < idx.size(); ++j) {
        if (nd.bv_pos >= bv.size || j >= bv.size - nd.bv_pos) throw std::runtime_error("wt_huff: node beyond the bit vector");
        (bv.get(nd.bv_pos + j) ? rr : l).push_back(idx[j]);
    }
    wt_decode(bv, nodes, nd.child[0], l, out, depth + 1);
    wt_decode(bv, nodes, nd.child[1], rr, out, depth + 1);
}
std::vector<uint8_t> read_wt_huff(Reader &r, bool rrr) {
    const uint64_t n = r.le();
    const uint64_t sigma = r.le();
    Bits bv;
    if (rrr) bv = read_rrr(r);                           // rank/select supports of rrr_vector store nothing
    else { bv = read_bits(r); skip_ints64(r); skip_select_mcl(r); skip_select_mcl(r); }
    const uint64_t n_nodes = r.le();
    if (n_nodes > 0xffff || (sigma && n_nodes != 2 * sigma - 1)) throw std::runtime_error("wt_huff: bad tree");
    std::vector<WtNode> nodes(n_nodes);
    for (auto &nd : nodes) {
        nd.bv_pos = r.le(); nd.bv_pos_rank = r.le();
        nd.parent = r.le16(); nd.child[0] = r.le16(); nd.child[1] = r.le16();
    }
    r.skip(256 * 2 + 256 * 8);                           // symbol -> leaf, symbol -> path
    if (n > bv.size + 1 && n_nodes > 1) throw std::runtime_error("wt_huff: more symbols than tree bits");
    if (n > (uint64_t)r.b.size() * 64) throw std::runtime_error("wt_huff: length field larger than the file allows");
    std::vector<uint8_t> out(n, 0);
    if (n) {
        if (nodes.empty()) throw std::runtime_error("wt_huff: empty tree");
        std::vector<uint64_t> idx(n);
        for (uint64_t i = 0; i < n; ++i) idx[i] = i;
        wt_decode(bv, nodes, 0, idx, out);
    }
    return out;
}

// BOSS::serialize (boss.cpp:262-277) followed by the graph mode (dbg_succinct.cpp:780-803); leaves the reader at
// the optional suffix-range index
struct ParsedBoss { uint64_t nf = 0, k_node = 0, state = 0, mode = 0; std::vector<uint64_t> F; std::vector<uint8_t> W; Bits last; };
void parse_boss(Reader &r, ParsedBoss &b) {
    uint64_t &nf = b.nf, &k_node = b.k_node, &state = b.state, &mode = b.mode;
    std::vector<uint64_t> &F = b.F; std::vector<uint8_t> &W = b.W; Bits &last = b.last;
    nf = r.be();
    if (nf != 5 && nf != 27) throw std::runtime_error("unsupported alphabet size " + std::to_string(nf));
    F.assign(nf, 0);
    for (auto &f : F) f = r.be();
    k_node = r.be();
    // node length k_node = k - 1; the device keys hold k <= 85 DNA / k <= 51 protein characters (mgb.h)
    if (k_node < 1 || k_node + 1 > (nf == 27 ? 51u : 85u))
        throw std::runtime_error("k = " + std::to_string(k_node + 1) + " is outside the supported range");
    state = r.be();
    if (state != 1 && state != 3)
        throw std::runtime_error("BOSS state " + std::to_string(state) + " (DYN / FAST) is not supported; "
                                 "convert the graph with `metagraph transform --state small|stat`");
    W = read_wt_huff(r, state == 1);
    const uint64_t logsigma = r.be();
    (void)logsigma;
    if (state == 1) {
        const uint64_t code = r.be();               // bit_vector_adaptive::VectorCode
        if (code == 0) last = read_rrr(r);
        else if (code == 1) last = read_sd(r);
        else throw std::runtime_error("unsupported `last` representation " + std::to_string(code));
    } else {
        last = read_bits(r);
        r.be();                                      // number of set bits
        skip_ints64(r);                              // rank_support_v5
        skip_select_mcl(r);                          // select_support_mcl<1>; select_support_scan<0> stores nothing
    }
    mode = r.be();
}
void read_file(const char *path, Reader &r) {
    std::ifstream in(path, std::ios::binary);
    if (!in.good()) throw std::runtime_error(std::string("cannot open ") + path);
    r.b.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
}

} // namespace

extern "C" {

const char* mgb_dbg_last_error(void) { return g_load_err.c_str(); }

int mgb_dbg_load(const char *path, mgb_boss_t *out, int *mode_out, int *state_out) {
    if (!path || !out) return MGB_ERR_INVALID_ARGUMENT;
    std::memset(out, 0, sizeof(*out));
    ParsedBoss pb;
    try {
        Reader r;
        read_file(path, r);
        parse_boss(r, pb);
        const uint64_t nf = pb.nf, k_node = pb.k_node, state = pb.state, mode = pb.mode;
        const std::vector<uint64_t> &F = pb.F; const std::vector<uint8_t> &W = pb.W; const Bits &last = pb.last;
        if (state_out) *state_out = (int)state;
        if (mode_out) *mode_out = (int)mode;
        // consistency: sizes, W range, F against the decoded arrays (boss_chunk.cpp:105-123)
        const uint64_t n1 = W.size();
        if (n1 < 2 || last.size != n1) throw std::runtime_error("W and last differ in length");
        if (mode > 2) throw std::runtime_error("bad graph mode");
        uint64_t ones = 0;
        for (uint64_t i = 1; i < n1; ++i) {
            if (W[i] >= 2 * nf) throw std::runtime_error("W value out of range");
            ones += last.get(i);
        }
        if (last.get(0) || !last.get(n1 - 1)) throw std::runtime_error("last[] malformed");
        for (uint64_t c = 1; c < nf; ++c)
            if (F[c] < F[c - 1] || F[c] >= n1) throw std::runtime_error("F[] malformed");
        {   // every node with last character c is the target of exactly one un-flagged edge labelled c
            std::vector<uint64_t> cnt(nf, 0);
            for (uint64_t i = 1; i < n1; ++i) if (W[i] < nf) ++cnt[W[i]];
            std::vector<uint64_t> nodes_with(nf, 0);
            uint64_t c = 0;
            for (uint64_t i = 1; i < n1; ++i) {
                while (c + 1 < nf && F[c + 1] < i) ++c;
                if (last.get(i)) ++nodes_with[c];
            }
            for (uint64_t s = 1; s < nf; ++s)
                if (cnt[s] != nodes_with[s]) throw std::runtime_error("F / W / last do not describe a BOSS table");
        }
        out->n_plus_1 = n1;
        out->W = (uint8_t*)std::malloc(n1);
        out->last = (uint8_t*)std::malloc(n1);
        if (!out->W || !out->last) { std::free(out->W); std::free(out->last); out->W = out->last = nullptr; return MGB_ERR_CUDA; }
        for (uint64_t i = 0; i < n1; ++i) { out->W[i] = W[i]; out->last[i] = last.get(i) ? 1 : 0; }
        for (uint64_t c = 0; c < nf; ++c) out->F[c] = F[c];
        out->k = (uint32_t)k_node + 1;
        out->alphabet = nf == 27 ? MGB_ALPHABET_PROTEIN : MGB_ALPHABET_DNA;
        return MGB_OK;
    } catch (const std::exception &e) {
        g_load_err = e.what();
        if (state_out) *state_out = (int)pb.state;       // a refused DYN / FAST file still says what it is
        return MGB_ERR_INVALID_ARGUMENT;
    }
}

int mgb_dbg_load_suffix_ranges(const char *path, uint32_t *suffix_len, uint64_t **ranges, uint64_t *n_ranges) {
    if (!path || !suffix_len || !ranges || !n_ranges) return MGB_ERR_INVALID_ARGUMENT;
    *suffix_len = 0; *ranges = nullptr; *n_ranges = 0;
    try {
        Reader r;
        read_file(path, r);
        ParsedBoss pb;
        parse_boss(r, pb);
        if (r.p == r.b.size()) return MGB_OK;            // written without the index (boss.cpp:402-426 tolerates that)
        // BOSS::serialize_suffix_ranges (boss.cpp:396-400): the length, then an sd_vector whose i-th set bit sits at
        // range value i + i (build_suffix_ranges_sd, boss.cpp:99-118; get_suffix_range(i) = select1(i + 1) - i)
        const uint64_t len = r.be();
        if (len == 0) {
            // `metagraph build` without an index still writes the (empty) vector: both example graphs end like this,
            // which pins the field order of the container (size, low width, low parts, high parts, two select supports)
            const Bits none = read_sd(r);
            if (none.size != 0 || r.p != r.b.size()) throw std::runtime_error("trailing bytes after an empty suffix-range index");
            return MGB_OK;
        }
        if (len > pb.k_node) throw std::runtime_error("bad index of suffix ranges");
        uint64_t expect = 2;
        for (uint64_t i = 0; i < len; ++i) {
            if (expect > (1ull << 40) / (pb.nf - 1)) throw std::runtime_error("bad index of suffix ranges");
            expect *= pb.nf - 1;
        }
        const Bits sd = read_sd(r);
        if (r.p != r.b.size()) throw std::runtime_error("trailing bytes after the suffix-range index");
        if (sd.size != pb.W.size() + expect) throw std::runtime_error("suffix-range index does not fit the graph");
        uint64_t *out = (uint64_t*)std::malloc(expect * sizeof(uint64_t));
        if (!out) { g_load_err = "out of host memory"; return MGB_ERR_NO_MEMORY; }
        uint64_t i = 0;
        for (uint64_t w = 0; w < sd.w.size() && i <= expect; ++w)
            for (uint64_t bits = sd.w[w]; bits; bits &= bits - 1) {
                const uint64_t pos = w * 64 + (uint64_t)__builtin_ctzll(bits);
                if (i < expect) out[i] = pos - i;
                ++i;
            }
        if (i != expect) { std::free(out); throw std::runtime_error("suffix-range index: wrong number of entries"); }
        *suffix_len = (uint32_t)len; *ranges = out; *n_ranges = expect;
        return MGB_OK;
    } catch (const std::exception &e) {
        g_load_err = e.what();
        return MGB_ERR_INVALID_ARGUMENT;
    }
}

void mgb_dbg_free_suffix_ranges(uint64_t *ranges) { std::free(ranges); }

} // extern "C"
