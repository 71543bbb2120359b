// Exact seeding: BOSS::map_to_edges (boss.cpp:996-1045) for one strand of one read, executed
// by one quad (4 lanes share every 64-byte index block; 8 strands per warp are in flight).
// Also the query preparation (upper-casing, reverse complement, alphabet codes:
// alignment.cpp:1348-1372, kmer_extractor.cpp:30-44, seq_tools/reverse_complement.hpp).
#pragma once
#include "index.cuh"

namespace mgb {

// kBOSSCharToDNA (kmer/alphabets.hpp:67-76) restated as arithmetic: A/a=1 C/c=2 G/g=3
// T/t/U/u=4, everything else (and bytes >= 128) = 5 = sigma (invalid).
MGB_HD uint8_t encode_dna(uint8_t ch) {
    if (ch >= 128) return 5;
    uint8_t u = ch & 0xDF;   // fold case for letters
    if (ch < 64) return 5;
    return u == 'A' ? 1 : u == 'C' ? 2 : u == 'G' ? 3 : (u == 'T' || u == 'U') ? 4 : 5;
}

// COMPL_TAB (seq_tools/reverse_complement.hpp:31-48) restated: IUPAC complement for both
// cases, 'U' -> 'A', '`' -> '@', identity elsewhere.
MGB_HD uint8_t complement_char(uint8_t ch) {
    if (ch == 96) return 64;
    uint8_t u = ch & 0xDF;
    if (ch < 64 || ch >= 128 || u < 'A' || u > 'Z') return ch;
    uint8_t lower = ch & 0x20;
    uint8_t r;
    switch (u) {
        case 'A': r = 'T'; break; case 'B': r = 'V'; break; case 'C': r = 'G'; break;
        case 'D': r = 'H'; break; case 'G': r = 'C'; break; case 'H': r = 'D'; break;
        case 'K': r = 'M'; break; case 'M': r = 'K'; break; case 'R': r = 'Y'; break;
        case 'T': r = 'A'; break; case 'U': r = 'A'; break; case 'V': r = 'B'; break;
        case 'Y': r = 'R'; break; default: r = u;
    }
    return r | lower;
}

// AlignmentResults ctor (alignment.cpp:1357-1358): toupper, bytes >= 128 -> 127
MGB_HD uint8_t sanitize_char(uint8_t ch) {
    if (ch >= 128) return 127;
    return (ch >= 'a' && ch <= 'z') ? ch - 32 : ch;
}

// First pass of the exact seeding, one k-mer per thread: can the k-mer at codes[i..i+K) exist at all?
// 0 = no (an invalid character in it, or the suffix-range table has no node for its first sfx_len
// characters -> BOSS::index returns 0, boss.hpp:695-718), 1 = maybe. On a strand that does not match the
// graph nearly every k-mer ends here, with one table load and no dependent chain; map_to_edges() below
// only resolves the rest. Results are unchanged: a k-mer's node does not depend on how it is looked up.
MGB_HD uint64_t premap_kmer(const IndexView &ix, const uint8_t *codes, int i) {
    const int K = (int)ix.k;
    const int S = (ix.sfx_len && (int)ix.sfx_len <= K - 1) ? (int)ix.sfx_len : 0;
    bool zero = false;
    for (int j = 0; j < K; ++j) {
        const uint32_t c = codes[i + j];
        if (c >= ix.sigma) return 0;
        zero |= c == 0;
    }
    if (!S || zero) return 1;
    const uint64_t base = ix.sigma - 1;
    uint64_t slot = 0;
    for (int j = S - 1; j >= 0; --j) slot = slot * base + (codes[i + j] - 1);
    const uint32_t rl = ldg32(ix.sfx + 2 * slot), ru1 = ldg32(ix.sfx + 2 * slot + 1);
    return rl < ru1 ? 1 : 0;                             // [rl, ru1) non-empty
}

// map_to_edges for codes[0..L). Writes L - K + 1 node ids (validate_edge applied) to `out`
// and, if `flags` != nullptr, per k-mer bit0 = "BOSS fwd() of this edge lands on a node with
// more than one outgoing edge" when the walk continued from it (used for the UniMEM
// terminator, dbg_succinct.cpp:617-630 has_multiple_outgoing).
// All lanes of the group execute this with identical arguments; lane 0 of the group stores.
// hints (optional): bit i = premap_kmer() of k-mer i (first pass); k-mers whose bit is clear are known to
// be absent, `out` is zero-filled beforehand, and the cold search jumps from set bit to set bit.
MGB_HD int next_hint(const uint32_t *hints, int n, int pos) {
    while (pos < n) {
        const uint32_t w = hints[pos >> 5] >> (pos & 31);
        if (w) return pos + ffs32(w) - 1;
        pos = (pos | 31) + 1;
    }
    return n;
}
MGB_HD void map_to_edges(const IndexView &ix, const uint8_t *codes, int L, uint64_t *out,
                         const uint32_t *hints = nullptr) {
    const int K = (int)ix.k;
    if (L < K) return;
    const bool writer = glane() == 0;
    // index of the last invalid character seen in the current window, or -1
    int last_inv = -1, last_zero = -1;         // (code 0, '$', never comes out of encode_dna)
    for (int j = 0; j < K - 1; ++j) {
        if (codes[j] >= ix.sigma) last_inv = j;
        if (codes[j] == 0) last_zero = j;
    }
    LineCache lc;
    // suffix-table slot of codes[w .. w + S): sum of (code - 1) * (sigma - 1)^j, kept as a rolling value
    // so that a cold lookup (one per k-mer on a strand that does not match) costs O(1) instead of O(k)
    const int S = (ix.sfx_len && (int)ix.sfx_len <= K - 1) ? (int)ix.sfx_len : 0;
    const uint64_t base = ix.sigma - 1;
    uint64_t top = 1;                                  // base^(S-1)
    for (int j = 1; j < S; ++j) top *= base;
    uint64_t slot = 0; int slot_at = -1 - S;           // window start the slot belongs to
    const int nk = L - K + 1;
    for (int i = 0; i + K <= L; ++i) {
        if (hints) {
            // jump to the next k-mer the first pass left open: it has no invalid character, so the
            // trackers (only ever compared with the current position) need no update for the gap
            i = next_hint(hints, nk, i);
            if (i >= nk) break;
            last_zero = -1;
            for (int j = 0; j < K; ++j) if (codes[i + j] == 0) last_zero = i + j;
        } else {
            if (codes[i + K - 1] >= ix.sigma) last_inv = i + K - 1;
            if (codes[i + K - 1] == 0) last_zero = i + K - 1;
            if (last_inv >= i) {              // invalid[i + k_]
                if (writer) out[i] = 0;
                continue;
            }
        }
        // map_to_edge (boss.hpp:766-777); all K codes are valid here
        uint64_t edge;
        if (S && last_zero < i) {
            if (i - slot_at >= S || i < slot_at) {     // restart (an invalid code, once clamped, never
                slot = 0;                              // survives S shifts, but a restart is simpler)
                for (int j = S - 1; j >= 0; --j) slot = slot * base + (codes[i + j] - 1);
            } else {
                for (int w = slot_at; w < i; ++w) {
                    const uint32_t c = codes[w + S];
                    const uint64_t digit = (c >= ix.sigma ? ix.sigma - 1 : (c ? c : 1)) - 1;   // clamped: unused windows
                    slot = (base == 4 ? slot >> 2 : slot / base) + digit * top;
                }
            }
            slot_at = i;
            edge = boss_index_slot(ix, codes + i, K - 1, slot);
        } else {
            edge = boss_index(ix, codes + i, K - 1);
        }
        if (edge) edge = pick_edge(ix, lc, edge, codes[i + K - 1]);
        if (writer) out[i] = in_graph(ix, edge) ? edge : 0;
        while (edge && ++i + K - 1 < L) {
            if (codes[i + K - 1] >= ix.sigma) last_inv = i + K - 1;
            if (codes[i + K - 1] == 0) last_zero = i + K - 1;
            if (last_inv >= i) {
                if (writer) out[i] = 0;
                break;
            }
            // fwd(edge, codes[i+K-2]) + pick_edge(.., codes[i+K-1]) through the adjacency record
            const uint32_t c = codes[i + K - 1];
            if (!MGB_WIDE(ix)) {
                const uint2 a = load_adj(ix, edge);
                edge = adj_child(a, c);
                if (writer) out[i] = (edge && ((a.y >> (8 + c)) & 1u)) ? edge : 0;
            } else {
                const Adj a = load_adj_any(ix, edge);
                edge = adj_child(a, c);
                if (writer) out[i] = (edge && ((a.ok >> c) & 1u)) ? edge : 0;
            }
        }
    }
}

} // namespace mgb
