// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the read-only query half of MetaGraph's BOSS / DBGSuccinct
// (the graph side of the `metagraph align` hot path) plus a batch constructor
// used to build test graphs.  Nothing under oracle/ is part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may build, link or call it.
//
// Parity status: the reference cannot be built in this environment (all 24
// submodules, incl. sdsl-lite, are empty), so this restatement is pinned
// against the reference's own golden vectors instead (tests/golden/, see
// tests/test_oracle_golden.py): test_boss.cpp SmallGraphTraversal,
// test_aligner.cpp CIGAR/score triples, integration_tests/test_align.py TSVs.
//
// Each function cites the reference file:line it restates (paths relative to
// /root/reference/metagraph/src).
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <string_view>
#include <tuple>
#include <utility>
#include <vector>

namespace mgo {

typedef uint64_t edge_index;
typedef uint64_t node_index;
typedef uint8_t TAlphabet;
static constexpr node_index npos = 0;

// ---------------------------------------------------------------------------
// Alphabets: kmer/alphabets.hpp:29-38 (protein), :64-79 (DNA);
// kmer/kmer_extractor.cpp:30-44 (encode: negative chars -> table[0]).
// ---------------------------------------------------------------------------
struct Alphabet {
    int sigma = 0;            // alph_size incl. '$'
    std::string letters;      // "$ACGT" / "$ABCDEFGHIJKLMNOPQRSTUVWYZX"
    uint8_t char_to_code[128];
    std::vector<uint8_t> complement_code; // empty for protein
    int bits_per_char = 0;    // bits for one code in a packed k-mer (incl. '$')

    TAlphabet encode(char c) const {
        int8_t s = static_cast<int8_t>(c);
        return s >= 0 ? char_to_code[static_cast<size_t>(s)] : char_to_code[0];
    }
    char decode(TAlphabet c) const { return letters[c]; }
    std::vector<TAlphabet> encode(std::string_view s) const {
        std::vector<TAlphabet> r(s.size());
        for (size_t i = 0; i < s.size(); ++i) r[i] = encode(s[i]);
        return r;
    }
    static const Alphabet& dna();
    static const Alphabet& protein();
};

// common/seq_tools/reverse_complement.hpp:31-78 (seqtk table, generated)
char complement_char(char c);
void reverse_complement_inplace(char *begin, char *end);
inline void reverse_complement_inplace(std::string &s) {
    reverse_complement_inplace(s.data(), s.data() + s.size());
}

// ---------------------------------------------------------------------------
// Plain-array rank/select over a byte sequence (stands in for sdsl wt_huff /
// rank_support_v5 / select_support_mcl; only the mathematical semantics of
// common/vectors/wavelet_tree.cpp:346-372 and bit_vector_sdsl.hpp:130-222
// are observable).
// ---------------------------------------------------------------------------
class SampledSeq {
  public:
    static constexpr uint64_t kBlock = 64;
    void build(const uint8_t *data, uint64_t n, int num_symbols, int num_select_symbols);
    // number of occurrences of c in [0..i] inclusive (i < n)
    uint64_t rank(uint8_t c, uint64_t i) const;
    // position of the r-th (1-based) occurrence of c
    uint64_t select(uint8_t c, uint64_t r) const;
    uint64_t count(uint8_t c) const { return totals_[c]; }
  private:
    const uint8_t *data_ = nullptr;
    uint64_t n_ = 0;
    int nsym_ = 0;
    std::vector<uint64_t> blk_;                  // [block][sym] counts before block
    std::vector<std::vector<uint64_t>> sel_;     // per sym: position of every kBlock-th occurrence
    std::vector<uint64_t> totals_;
};

// ---------------------------------------------------------------------------
// BOSS table (graph/representation/succinct/boss.hpp:28-631).
// k_ = node length = K - 1.  Arrays have a placeholder at position 0
// (boss_chunk.cpp:60-62).
// ---------------------------------------------------------------------------
class BOSS {
  public:
    const Alphabet *alph = nullptr;
    int alph_size = 0;
    size_t k_ = 0;
    std::vector<uint8_t> W;      // values in [0, 2*sigma)
    std::vector<uint8_t> last;   // 0/1
    std::vector<uint64_t> F, NF;
    size_t indexed_suffix_length_ = 0;   // optional suffix-range index (boss.hpp:516-525)
    std::vector<uint64_t> suffix_ranges_; // 2 entries per suffix: [begin, end)

    // Batch construction from sequences (boss_chunk_construct.cpp:341-462 dummy edges,
    // boss_chunk.cpp:33-133 W/last/F).  `k` is the node length.
    // `force_source_dummies` additionally inserts the $-prefixed chain of every sequence start even
    // when it is redundant, which is what the dynamic BOSS::add_sequence leaves behind
    // (boss.cpp:1187-1260; used by the goldens that build graphs with add_sequence).
    static BOSS build(const Alphabet &alph, size_t k, const std::vector<std::string> &seqs,
                      bool force_source_dummies = false);
    // Adopt arrays produced elsewhere (e.g. the product's own constructor).
    static BOSS from_arrays(const Alphabet &alph, size_t k, std::vector<uint8_t> &&W,
                            std::vector<uint8_t> &&last, const std::vector<uint64_t> &F);
    void index_suffix_ranges(size_t suffix_length); // boss_chunk_construct.cpp:260-320

    size_t get_k() const { return k_; }
    uint64_t num_edges() const { return W.size() - 1; }
    uint64_t num_nodes() const { return last_rs_.count(1); }

    TAlphabet get_W(edge_index i) const { return W[i]; }
    bool get_last(edge_index i) const { return last[i]; }

    uint64_t rank_W(edge_index i, TAlphabet c) const;        // boss.cpp:437-441
    edge_index succ_W(edge_index i, TAlphabet c) const;      // boss.cpp:505-509
    std::pair<edge_index, TAlphabet> succ_W(edge_index i, TAlphabet c1, TAlphabet c2) const; // :515-570
    uint64_t rank_last(edge_index i) const;                  // :577-581
    edge_index select_last(uint64_t i) const;                // :588-592
    edge_index pred_last(edge_index i) const;                // :598-607
    edge_index succ_last(edge_index i) const;                // :613-617
    edge_index bwd(edge_index i) const;                      // :623-636
    edge_index fwd(edge_index i, TAlphabet c) const;         // :642-652
    TAlphabet get_node_last_value(edge_index i) const;       // :679-690
    std::pair<TAlphabet, edge_index> get_minus_k_value(edge_index i, size_t k) const; // :696-704
    edge_index pick_edge(edge_index edge, TAlphabet c) const; // :710-722
    template <class CB> void call_incoming_to_target(edge_index edge, TAlphabet d, CB &&cb) const; // :766-786
    bool is_single_incoming(edge_index i, TAlphabet w) const; // :803-816

    std::tuple<edge_index, edge_index, size_t>
    get_initial_range(const TAlphabet *begin, const TAlphabet *end) const;  // boss.hpp:636-680
    bool tighten_range(edge_index *rl, edge_index *ru, TAlphabet s) const;  // boss.hpp:682-693
    edge_index index(const TAlphabet *begin, const TAlphabet *end) const;   // boss.hpp:695-718
    std::tuple<edge_index, edge_index, const TAlphabet*>
    index_range(const TAlphabet *begin, const TAlphabet *end) const;        // boss.hpp:720-764
    edge_index map_to_edge(const TAlphabet *begin, const TAlphabet *end) const; // boss.hpp:766-777
    std::vector<edge_index> map_to_edges(const std::vector<TAlphabet> &seq) const; // boss.cpp:996-1045

    std::vector<TAlphabet> get_node_seq(edge_index i) const;  // boss.cpp:953-973
    std::string get_node_str(edge_index i) const;

    void finalize();   // builds the rank/select structures once W / last / F are set
  private:
    SampledSeq W_rs_, last_rs_;
};

template <class CB>
void BOSS::call_incoming_to_target(edge_index edge, TAlphabet d, CB &&cb) const {
    cb(edge);
    TAlphabet d_next;
    while (++edge < W.size()) {
        std::tie(edge, d_next) = succ_W(edge, d, d + alph_size);
        if (d_next != d + alph_size)
            break;
        cb(edge);
    }
}

// ---------------------------------------------------------------------------
// DBGSuccinct (dbg_succinct.hpp:13-211): DBG node id == BOSS edge index.
// `valid_edges` empty  == mask dropped (cli/align.cpp:335-339);
//              non-empty == dummy k-mers masked (unit-test graphs,
//              tests/graph/all/test_dbg_helpers.cpp:369-385).
// ---------------------------------------------------------------------------
class DBGSuccinct {
  public:
    BOSS boss;
    std::vector<uint8_t> valid_edges;
    // DeBruijnGraph::Mode (sequence_graph.hpp:160): 0 BASIC, 1 CANONICAL (the graph holds the reverse
    // complement of every k-mer), 2 PRIMARY (one k-mer of every reverse-complement pair; aligned to through
    // the CanonicalDBG wrapper below).
    int mode = 0;

    size_t get_k() const { return boss.k_ + 1; }
    uint64_t max_index() const { return boss.num_edges(); }
    bool in_graph(node_index n) const {                       // dbg_succinct.cpp:934-936
        return n > 0 && n <= max_index() && (valid_edges.empty() || valid_edges[n]);
    }
    node_index validate_edge(edge_index e) const { return in_graph(e) ? e : npos; }
    void mask_dummy_kmers();                                  // dbg_succinct.cpp:917-932
    void reset_mask() { valid_edges.clear(); }

    std::vector<node_index> map_to_nodes_sequentially(std::string_view seq) const; // :285-305
    template <class CB> void call_outgoing_kmers(node_index node, CB &&cb) const;  // :110-139
    template <class CB> void call_incoming_kmers(node_index node, CB &&cb) const;  // :141-163
    bool has_multiple_outgoing(node_index node) const;        // :617-630
    bool has_single_incoming(node_index node) const;          // :662-680
    std::string get_node_sequence(node_index node) const;     // :275-281
    node_index traverse(node_index node, char c) const;       // :84-97
    // :307-393 (only the max_num_allowed_matches == SIZE_MAX branch is used by the aligner)
    void call_nodes_with_suffix_matching_longest_prefix(
        std::string_view str, const std::function<void(node_index, uint64_t)> &cb,
        size_t min_match_length) const;
    uint64_t num_nodes() const;
};

template <class CB>
void DBGSuccinct::call_outgoing_kmers(node_index node, CB &&cb) const {
    TAlphabet w = 0;
    if (node > 1 && !(w = boss.get_W(node)))
        return; // sink dummy
    edge_index lst = boss.fwd(node, w % boss.alph_size);
    edge_index first = boss.pred_last(lst - 1) + 1;
    for (edge_index i = std::max<uint64_t>(2, first); i <= lst; ++i) {
        if (in_graph(i))
            cb(i, boss.alph->decode(boss.get_W(i) % boss.alph_size));
    }
}

template <class CB>
void DBGSuccinct::call_incoming_kmers(node_index node, CB &&cb) const {
    boss.call_incoming_to_target(boss.bwd(node), boss.get_node_last_value(node),
        [&](edge_index prev) {
            if (in_graph(prev))
                cb(prev, boss.alph->decode(boss.get_minus_k_value(prev, get_k() - 2).first));
        });
}

// ---------------------------------------------------------------------------
// CanonicalDBG (graph/representation/canonical_dbg.{hpp,cpp}): wraps a PRIMARY-mode DBGSuccinct (one k-mer of
// every reverse-complement pair) and behaves like a CANONICAL-mode graph. Node ids above `offset_` denote the
// reverse complements of base-graph nodes. The caches of the reference (NodeFirstCache, palindrome cache)
// only memoise: the values they would hold are recomputed here (graph_extensions/node_first_cache.cpp).
// ---------------------------------------------------------------------------
class CanonicalDBG {
  public:
    explicit CanonicalDBG(const DBGSuccinct &graph)
        : g(graph), offset_(graph.max_index()), k_odd_(graph.get_k() % 2),
          has_sentinel_(graph.valid_edges.empty()) {}              // canonical_dbg.cpp:23-44
    const DBGSuccinct &g;

    size_t get_k() const { return g.get_k(); }
    uint64_t max_index() const { return offset_ * 2; }
    node_index get_base_node(node_index node) const { return node > offset_ ? node - offset_ : node; }
    node_index reverse_complement(node_index node) const;                              // :521-553
    void reverse_complement(std::string &seq, std::vector<node_index> &path) const;    // :555-565
    std::vector<node_index> map_to_nodes_sequentially(std::string_view seq) const;     // :55-146
    std::string get_node_sequence(node_index node) const;                              // :423-432
    // :158-243 and :245-336; cb(node, char); the spelling hint of the reference is the node's sequence
    void call_outgoing_kmers(node_index node, const std::function<void(node_index, char)> &cb) const;
    void call_incoming_kmers(node_index node, const std::function<void(node_index, char)> &cb) const;
    void adjacent_outgoing_nodes(node_index node, const std::function<void(node_index)> &cb) const; // :352-364
    void adjacent_incoming_nodes(node_index node, const std::function<void(node_index)> &cb) const; // :338-350
    bool has_multiple_outgoing(node_index node) const;                                 // :366-380
    bool has_single_incoming(node_index node) const;                                   // :382-392
  private:
    uint64_t offset_;
    bool k_odd_, has_sentinel_;
    void adjacent_incoming_rc_strand(node_index node, const std::string &spelling,
                                     const std::function<void(node_index, char)> &cb) const;  // :579-625
    void adjacent_outgoing_rc_strand(node_index node, const std::string &spelling,
                                     const std::function<void(node_index, char)> &cb) const;  // :640-680
};

// Graph view used by the aligner: the graph itself, its reverse complement
// (graph/representation/rc_dbg.hpp:17-178), or a PRIMARY graph behind the CanonicalDBG wrapper.
struct GraphView {
    const DBGSuccinct *g = nullptr;
    bool rc = false;
    const CanonicalDBG *canon = nullptr;
    // DBGSuccinct::adjacent_incoming_nodes (dbg_succinct.cpp:176-193), plain view only
    template <class CB> void adjacent_incoming_nodes(node_index node, CB &&cb) const {
        if (canon) { canon->adjacent_incoming_nodes(node, cb); return; }
        g->call_incoming_kmers(node, [&](node_index prev, char) { cb(prev); });
    }
    size_t get_k() const { return g->get_k(); }
    uint64_t max_index() const { return canon ? canon->max_index() : g->max_index(); }
    template <class CB> void call_outgoing_kmers(node_index node, CB &&cb) const {
        if (canon) {
            canon->call_outgoing_kmers(node, cb);
        } else if (!rc) {
            g->call_outgoing_kmers(node, cb);
        } else {   // rc_dbg.hpp:86-97
            g->call_incoming_kmers(node, [&](node_index prev, char c) { cb(prev, complement_char(c)); });
        }
    }
    std::string get_node_sequence(node_index node) const {
        if (canon) return canon->get_node_sequence(node);
        std::string s = g->get_node_sequence(node);
        if (rc) reverse_complement_inplace(s);
        return s;
    }
};

} // namespace mgo
