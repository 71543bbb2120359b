// Kernel-side code of the C-ABI library: per-item device functions, the kernels and their launchers.
//
// The product library compiles this file twice (no relocatable device code, so the two sets of device
// functions never meet):
//   api.cu          -DMGB_NARROW_ONLY, MGB_KERNEL_NS = kern_dna : DNA block layout only; the branches
//                   to the alphabet-generic index layout are compiled out of the hot kernels
//   api_generic.cu  -DMGB_WIDE_ONLY,   MGB_KERNEL_NS = kern_any : alphabet-generic layout only (protein)
//   api_canonical.cu -DMGB_NARROW_ONLY -DMGB_CANONICAL_ONLY -DMGB_ALIGN_KERNEL_ONLY, MGB_KERNEL_NS = kern_canon :
//                   k_align for CANONICAL-mode DNA graphs; the other two sets compile that mode's branches out
// The host-emulation build (tests/emu) includes it once with the layout chosen at run time.
#pragma once
#include "align_core.cuh"
#if !defined(MGB_HOST_EMU)
#include <cuda_runtime.h>
#endif

#ifndef MGB_KERNEL_NS
#define MGB_KERNEL_NS kern_dna
#endif

namespace mgb {

struct ReadHdr {              // per read, written by the align kernel
    uint32_t status, n_aln;
    uint64_t heap_off;
    ReadStats stats;
};

// ---------------------------------------------------------------------------------------
// kernels (device) / loops (host emulation)
// ---------------------------------------------------------------------------------------
struct PrepArgs {
    const char *seqs; const uint64_t *offsets; uint32_t n_reads;
    char *qf, *qr; uint8_t *cf, *cr;
    uint8_t code_of[256];     // KmerExtractorBOSS::encode of the index's alphabet
};

MGB_HD void prepare_read(const PrepArgs &a, uint32_t r) {
    const uint64_t b = a.offsets[r];
    const int L = (int)(a.offsets[r + 1] - b);
    for (int i = wlane(); i < L; i += kWarp) {
        uint8_t f = sanitize_char((uint8_t)a.seqs[b + i]);
        uint8_t rc = complement_char(sanitize_char((uint8_t)a.seqs[b + L - 1 - i]));
        a.qf[b + i] = (char)f; a.qr[b + i] = (char)rc;
        a.cf[b + i] = a.code_of[f]; a.cr[b + i] = a.code_of[rc];
    }
}

struct SeedArgs {
    IndexView ix;
    const uint8_t *cf, *cr; const uint64_t *offsets; const uint64_t *koff;
    uint64_t *nodes_f, *nodes_r; uint32_t n_reads; uint32_t n_strands;
    uint32_t hinted;           // 1: k_premap ran, hint_f / hint_r hold its verdicts
    uint32_t *hint_f, *hint_r;
};

MGB_HD void seed_item(const SeedArgs &a, uint64_t item) {
    uint32_t r = (uint32_t)(item / a.n_strands);
    uint32_t s = (uint32_t)(item % a.n_strands);
    const uint64_t b = a.offsets[r];
    const int L = (int)(a.offsets[r + 1] - b);
    const uint32_t *hints = a.hinted ? (s ? a.hint_r : a.hint_f) + (a.koff[r] >> 5) + r : nullptr;
    map_to_edges(a.ix, (s ? a.cr : a.cf) + b, L, (s ? a.nodes_r : a.nodes_f) + a.koff[r], hints);
}
// first pass (premap_kmer) for strand s of read r: one k-mer per lane, one hint word per 32 k-mers
// (the words of read r start at word koff[r] / 32 + r of the strand's hint array)
MGB_HD void premap_item(const SeedArgs &a, uint32_t r, uint32_t s) {
    const uint64_t b = a.offsets[r];
    const int L = (int)(a.offsets[r + 1] - b);
    const int nk = L - (int)a.ix.k + 1;
    if (nk <= 0) return;
    const uint8_t *codes = (s ? a.cr : a.cf) + b;
    uint32_t *hw = (s ? a.hint_r : a.hint_f) + (a.koff[r] >> 5) + r;
    for (int base = 0; base < nk; base += 32) {
        unsigned word = 0;
        for (int o = 0; o < 32; o += kWarp) {                 // one pass on the device
            const int i = base + o + wlane();
            word |= wballot(i < nk && premap_kmer(a.ix, codes, i) != 0) << o;
        }
        if (wlane() == 0) hw[base >> 5] = word;
    }
}

// Sub-k seeding, lookup part (SuffixSeeder::generate_seeds, aligner_seeder_methods.cpp:215-238 ->
// dbg_succinct.cpp:307-330 -> BOSS::index_range, boss.hpp:720-764): the longest prefix of the query at
// every position without a full k-mer hit and its edge range. The lookups are independent of each other,
// so they run GPU-wide, one quad per 16 positions of a strand, before the per-read kernel; k_align
// keeps the order-dependent part (enumeration, per-locus rules, merging with the MEM seeds).
struct SubkArgs {
    IndexView ix;
    const uint8_t *cf, *cr; const uint64_t *offsets, *koff; const uint64_t *nodes_f, *nodes_r;
    uint32_t n_reads, n_strands, min_seed_length, max_len;     // max_len = min(max_seed_length, k - 1)
    uint32_t *first_f, *last_f, *first_r, *last_r; uint8_t *len_f, *len_r;
};
#ifndef MGB_SUBK_CHUNK
#define MGB_SUBK_CHUNK 1
#endif
static constexpr int kSubkChunk = MGB_SUBK_CHUNK;
MGB_HD void subk_item(const SubkArgs &a, uint32_t r, uint32_t s, uint32_t chunk) {
    const uint64_t b = a.offsets[r];
    const int L = (int)(a.offsets[r + 1] - b);
    const int K = (int)a.ix.k;
    if (L < (int)a.min_seed_length) return;
    const int n_pos = L - (int)a.min_seed_length + 1;
    const uint8_t *codes = (s ? a.cr : a.cf) + b;
    const uint64_t *nodes = L >= K ? (s ? a.nodes_r : a.nodes_f) + a.koff[r] : nullptr;
    uint32_t *of = (s ? a.first_r : a.first_f) + b, *ol = (s ? a.last_r : a.last_f) + b;
    uint8_t *on = (s ? a.len_r : a.len_f) + b;
    const int nk = L >= K ? L - K + 1 : 0;
    for (int i = (int)chunk * kSubkChunk; i < n_pos && i < (int)(chunk + 1) * kSubkChunk; ++i) {
        uint64_t first = 0, lst = 0; int matched = 0;
        uint8_t code = 0xFF;                              // not computed: a full k-mer matches here
        if (!(i < nk && nodes && nodes[i] != 0)) {
            const int len = (int)a.max_len < L - i ? (int)a.max_len : L - i;
            bool ok = len >= (int)a.min_seed_length;
            for (int t = 0; t < len && ok; ++t) ok = codes[i + t] < a.ix.sigma;
            if (ok) boss_index_range(a.ix, codes + i, len < K - 1 ? len : K - 1, &first, &lst, &matched, (int)a.min_seed_length);
            code = (uint8_t)matched;
        }
        if (glane() == 0) { of[i] = (uint32_t)first; ol[i] = (uint32_t)lst; on[i] = code; }
    }
}

struct AlignArgs {
    IndexView ix; DevConfig cfg; Caps caps;
    WarpLayout lay;             // arena regions of a lane group (offsets)
    SmemLayout slay;            // on-chip working set of a lane group (offsets)
    int use_fast;               // 0 disables the register fast path (test knob)
    unsigned long long *phase_out;   // MGB_PHASE_TIMERS builds: cycles per phase (setup, seeds, fwd, backtrack, align total)
    const char *qf, *qr; const uint8_t *cf, *cr; const uint64_t *offsets, *koff;
    const uint64_t *nodes_f, *nodes_r;
    const uint32_t *sub_first_f, *sub_last_f, *sub_first_r, *sub_last_r;    // k_subk results or nullptr
    const uint8_t *sub_len_f, *sub_len_r;
    const uint32_t *read_list; uint32_t n_list;
    char *arena; size_t arena_stride;
    ReadHdr *hdr; char *heap; uint64_t heap_cap; unsigned long long *heap_used;
    unsigned int *next;
};

// `act`: this lane group has a read; a group without one still walks through the lock-step loops of the
// aligner (the lane groups of a warp run them together)
MGB_HD void align_read(const AlignArgs &a, uint32_t r, const WarpMem &mem, const WarpSmem &sm, bool act) {
    ReadAligner al(a.ix, a.cfg, a.caps, mem, sm);
    al.use_fast = a.use_fast != 0;
    const uint64_t b = act ? a.offsets[r] : 0;
    const int L = act ? (int)(a.offsets[r + 1] - b) : 0;
    int order[kMaxAlt];
    const bool has_k = L >= (int)a.ix.k;
    if (a.sub_len_f) {
        al.subk_first[0] = a.sub_first_f + b; al.subk_last[0] = a.sub_last_f + b; al.subk_len[0] = a.sub_len_f + b;
        if (a.sub_len_r) { al.subk_first[1] = a.sub_first_r + b; al.subk_last[1] = a.sub_last_r + b; al.subk_len[1] = a.sub_len_r + b; }
    }
    const uint64_t ko = (act && has_k) ? a.koff[r] : 0;
    int n = al.run(act, L, a.qf + b, a.qr + b, a.cf + b, a.cr + b,
                   has_k ? a.nodes_f + ko : nullptr,
                   has_k && a.cfg.forward_and_reverse_complement ? a.nodes_r + ko : nullptr, order);
    if (!act) return;
#if defined(MGB_PHASE_TIMERS) && MGB_DEVICE_CODE
    if (wlane() == 0 && a.phase_out) {
        for (int p = 0; p < 8; ++p) atomicAdd((unsigned long long*)a.phase_out + p, (unsigned long long)al.phase_cycles[p]);
    }
#endif
    ReadHdr h;
    h.status = al.overflow ? MGB_READ_OVERFLOW : MGB_READ_OK;
    h.n_aln = 0; h.heap_off = 0; h.stats = al.stats;
    if (!al.overflow && n) {
        // bytes: per alignment OutAln + nodes*8 + cigar*4 + seq (padded to 8)
        uint64_t bytes = 0;
        for (int i = 0; i < n; ++i) {
            const AlnHdr ah = *mem.slot(SLOT_AGG + order[i]).h;
            bytes += sizeof(OutAln) + (a.cfg.result_nodes ? 0ull : 8ull * ah.n_nodes) + ((4ull * ah.n_cigar + 7) & ~7ull)
                   + (((uint64_t)ah.seq_len + 7) & ~7ull);
        }
        unsigned long long off = 0;
#if MGB_DEVICE_CODE
        if (wlane() == 0) off = atomicAdd(a.heap_used, (unsigned long long)bytes);
        off = wbcast64(off, 0);
#else
        off = *a.heap_used; *a.heap_used += bytes;
#endif
        if (off + bytes > a.heap_cap) {
            h.status = MGB_READ_OVERFLOW;
        } else {
            h.n_aln = n; h.heap_off = off;
            char *p = a.heap + off;
            for (int i = 0; i < n; ++i) {
                const AlnSlot sl = mem.slot(SLOT_AGG + order[i]);
                const AlnHdr ah = *sl.h;
                OutAln o;
                o.orientation = ah.orientation; o.score = ah.score; o.offset = ah.offset;
                o.query_begin = al.aln_clipping(sl); o.query_len = ah.q_len;
                o.n_nodes = ah.n_nodes; o.seq_len = ah.seq_len; o.n_cigar = ah.n_cigar;
                if (wlane() == 0) *(OutAln*)p = o;
                p += sizeof(OutAln);
                if (!a.cfg.result_nodes) {
                    uint64_t *pn = (uint64_t*)p;
                    for (int t = wlane(); t < ah.n_nodes; t += kWarp) pn[t] = sl.nodes[t];
                    p += 8ull * ah.n_nodes;
                }
                uint32_t *pc = (uint32_t*)p;
                for (int t = wlane(); t < ah.n_cigar; t += kWarp) pc[t] = sl.cigar[t];
                p += (4ull * ah.n_cigar + 7) & ~7ull;
                for (int t = wlane(); t < ah.seq_len; t += kWarp) p[t] = sl.seq[t];
                p += ((uint64_t)ah.seq_len + 7) & ~7ull;
            }
        }
    }
    if (wlane() == 0) {
        a.hdr[r] = h;
        mem.epoch_store()[0] = sm.ctx()[0].conv_epoch; mem.epoch_store()[1] = sm.ctx()[1].conv_epoch;
    }
    wsync();
}

// once per arena: convergence-table slots start with epoch 0 (never equal to a live epoch)
MGB_HD void init_arena(const AlignArgs &a, char *arena) {
    const WarpMem mem { arena, &a.lay };
    for (int e = 0; e < 2; ++e)
        for (uint32_t i = wlane(); i < a.caps.hash_size; i += kWarp) mem.conv_slots(e)[i].epoch = 0;
    if (wlane() == 0) { mem.epoch_store()[0] = 0; mem.epoch_store()[1] = 0; }
    wsync();
}

// One refinement level of the suffix-range table (boss.hpp:651-655 index order: the appended
// character is the most significant digit): entry o = idx + (c-1)*cur_num of the new table is
// tighten_range(cur[idx], c).
struct SfxArgs { IndexView ix; const uint32_t *cur; uint32_t *nxt; uint64_t cur_num; };

MGB_HD void sfx_extend_item(const SfxArgs &a, uint64_t o) {
    const uint64_t idx = o % a.cur_num;
    const uint32_t c = (uint32_t)(o / a.cur_num) + 1;
    uint64_t rl = a.cur[2 * idx], ru = (uint64_t)a.cur[2 * idx + 1] - 1;
    uint32_t b = 1, e = 1;
    if (rl <= ru && tighten_range(a.ix, &rl, &ru, c)) { b = (uint32_t)rl; e = (uint32_t)(ru + 1); }
    if (glane() == 0) { a.nxt[2 * o] = b; a.nxt[2 * o + 1] = e; }
}

// Reverse adjacency construction (index.cuh load_radj): per edge bwd(e) + "source node has several
// incoming edges", then k-2 gather rounds c_{j+1}[e] = c_j[bwd(e)] that move the last node character
// (boss.cpp:679-690) to the first position of the k-mer.
struct RadjArgs { IndexView ix; uint32_t *bwd_arr; uint8_t *c_cur; uint8_t *c_nxt; uint8_t *multi; uint2 *radj; uint64_t n;
                  // k-mer hash index build (nullptr: not built): the packed k-mer of every edge, accumulated over the same
                  // gather rounds that move the node's last character to the first position
                  unsigned long long *kmer; };
static constexpr unsigned long long kKmerBad = ~0ull;      // the k-mer holds a '$'
inline uint32_t radj_multi_shift(const IndexView &ix) { return ix.wide ? 7u : 3u; }   // index.cuh radj_multi

MGB_HD void radj_bwd_item(const RadjArgs &a, uint64_t e) {
    LineCache lc;
    uint64_t x = bwd(a.ix, lc, e);
    uint32_t d = node_last_value(a.ix, e);
    uint32_t multi = 0;
    if (x + 1 <= a.ix.n) {
        uint32_t w;
        succ_W2(a.ix, lc, x + 1, d, &w);
        multi = w == d + a.ix.sigma;
    }
    if (a.kmer) {
        // last two characters of the k-mer: the edge label and the node's last character
        const uint32_t lab = lc.get_W(a.ix, e) % a.ix.sigma;
        const uint32_t K = a.ix.k;
        if (glane() == 0)
            a.kmer[e] = (lab == 0 || d == 0) ? kKmerBad
                      : ((unsigned long long)(lab - 1) << (2 * (K - 1))) | ((unsigned long long)(d - 1) << (2 * (K - 2)));
    }
    if (glane() == 0) { a.bwd_arr[e] = (uint32_t)x; a.c_cur[e] = (uint8_t)d; a.multi[e] = (uint8_t)multi; }
}

// PRIMARY graphs: the rc-strand jump tables of IndexView (rcs / rcp / palin) for edge e. The k-mer is spelled
// through the reverse adjacency records (BOSS::get_node_seq, boss.cpp:953-973), its two (k-1)-mers are reverse
// complemented and looked up (NodeFirstCache::get_suffix_rc / get_prefix_rc, node_first_cache.cpp:122-176).
constexpr int kMaxPrimaryK = 128;
struct RcArgs { IndexView ix; uint32_t *rcs; uint32_t *rcp; uint32_t *palin; uint64_t n; };
MGB_HD void rc_tables_item(const RcArgs &a, uint64_t e) {
    const IndexView &ix = a.ix;
    const int K = (int)ix.k;
    uint8_t km[kMaxPrimaryK];
    {
        LineCache lc;
        km[K - 1] = (uint8_t)(lc.get_W(ix, e) % ix.sigma);
        uint64_t cur = e;
        for (int i = K - 2; i >= 0; --i) {
            km[i] = (uint8_t)node_last_value(ix, cur);
            cur = load_radj(ix, cur).x;
        }
    }
    uint8_t rc[kMaxPrimaryK];
    uint32_t r_s = 0, r_p = 0;
    bool clean = true;                                   // no '$' in x[1..k)
    for (int i = 1; i < K; ++i) clean = clean && km[i] != 0;
    if (clean) {
        for (int t = 0; t < K - 1; ++t) rc[t] = (uint8_t)(ix.sigma - km[K - 1 - t]);
        uint64_t first, lst; int matched;
        boss_index_range(ix, rc, K - 1, &first, &lst, &matched);
        if (matched == K - 1) r_s = (uint32_t)lst;
    }
    if (km[0] != 0) {                                    // then x[0..k-1) holds no '$' at all
        for (int t = 0; t < K - 1; ++t) rc[t] = (uint8_t)(ix.sigma - km[K - 2 - t]);
        uint64_t first, lst; int matched;
        boss_index_range(ix, rc, K - 1, &first, &lst, &matched);
        if (matched == K - 1) r_p = (uint32_t)lst;
    }
    bool pal = a.palin != nullptr && clean && km[0] != 0;
    for (int t = 0; t < K && pal; ++t) pal = km[t] == ix.sigma - km[K - 1 - t];
    if (glane() == 0) {
        a.rcs[e] = r_s; a.rcp[e] = r_p;
        if (pal) {
#if MGB_DEVICE_CODE
            atomicOr(a.palin + (e >> 5), 1u << (e & 31));
#else
            a.palin[e >> 5] |= 1u << (e & 31);
#endif
        }
    }
}

} // namespace mgb

#if !defined(MGB_HOST_EMU)
namespace MGB_KERNEL_NS {
using namespace mgb;

#if !defined(MGB_ALIGN_KERNEL_ONLY)
__global__ void __launch_bounds__(128) k_radj_bwd(RadjArgs a) {
    uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    uint64_t nquads = ((uint64_t)gridDim.x * blockDim.x) >> 2;
    for (uint64_t e = 1 + quad; e <= a.n; e += nquads) radj_bwd_item(a, e);
}

__global__ void __launch_bounds__(128) k_sfx_extend(SfxArgs a) {
    uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    uint64_t nquads = ((uint64_t)gridDim.x * blockDim.x) >> 2;
    const uint64_t total = a.cur_num * (a.ix.sigma - 1);
    for (uint64_t o = quad; o < total; o += nquads) sfx_extend_item(a, o);
}

// Sub-k lookups (subk_item) for 32 query positions of one strand per warp. Most positions end with one load: the
// first sfx_len characters do not occur in the graph, and a miss in a suffix-range table no longer than
// min_seed_length is final (boss_index_range). So every LANE first settles its own position against the table;
// only the survivors (a matched prefix of at least sfx_len characters: near an error on the matching strand) go
// through the quad-cooperative range search, 8 at a time. (One quad per position kept 4 lanes on every one-load
// miss.) Same results as subk_item(), which the host emulation still runs.
__global__ void __launch_bounds__(128) k_subk(SubkArgs a, uint32_t chunks_per_strand) {
    const IndexView &ix = a.ix;
    const int K = (int)ix.k;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t blocks32 = (chunks_per_strand * (uint32_t)kSubkChunk + 31) / 32;    // 32-position blocks per strand
    const uint64_t per_read = (uint64_t)a.n_strands * blocks32;
    const uint64_t items = (uint64_t)a.n_reads * per_read;
    const int S = (int)ix.sfx_len;
    const uint64_t base = ix.sigma - 1;
    for (uint64_t it = warp; it < items; it += nwarps) {
        const uint32_t r = (uint32_t)(it / per_read), rem = (uint32_t)(it % per_read);
        const uint32_t s = rem / blocks32, blk = rem % blocks32;
        const uint64_t b = a.offsets[r];
        const int L = (int)(a.offsets[r + 1] - b);
        if (L < (int)a.min_seed_length) continue;
        const int n_pos = L - (int)a.min_seed_length + 1;
        const uint8_t *codes = (s ? a.cr : a.cf) + b;
        const uint64_t *nodes = L >= K ? (s ? a.nodes_r : a.nodes_f) + a.koff[r] : nullptr;
        uint32_t *of = (s ? a.first_r : a.first_f) + b, *ol = (s ? a.last_r : a.last_f) + b;
        uint8_t *on = (s ? a.len_r : a.len_f) + b;
        const int nk = L >= K ? L - K + 1 : 0;
        const int i = (int)blk * 32 + lane;
        // ---- per lane: positions that need no search
        bool survivor = false;
        int len = 0;
        if (i < n_pos) {
            if (i < nk && nodes && nodes[i] != 0) {
                of[i] = 0; ol[i] = 0; on[i] = 0xFF;                 // a full k-mer matches here: not computed
            } else {
                len = (int)a.max_len < L - i ? (int)a.max_len : L - i;
                bool ok = len >= (int)a.min_seed_length;
                bool zero = false;
                for (int t = 0; t < len && ok; ++t) { const uint32_t c = codes[i + t]; ok = c < ix.sigma; zero = zero || c == 0; }
                const int qlen = len < K - 1 ? len : K - 1;         // what boss_index_range is asked for
                survivor = ok;
                if (ok && S && S <= qlen && S <= (int)a.min_seed_length && !zero && S > 1) {
                    uint64_t slot = 0;
                    for (int j = S - 1; j >= 0; --j) slot = slot * base + (codes[i + j] - 1);
                    const uint32_t rl = ldg32(ix.sfx + 2 * slot), ru1 = ldg32(ix.sfx + 2 * slot + 1);
                    if (rl >= ru1) survivor = false;                // empty range: matched = 0, final
                }
                if (!survivor) { of[i] = 0; ol[i] = 0; on[i] = 0; }
            }
        }
        // ---- the survivors, 8 per round, one quad each
        uint32_t todo = __ballot_sync(0xffffffffu, survivor);
        while (todo) {
            // the (lane >> 2)-th set bit of `todo` is this quad's position, if there are that many
            const int q = lane >> 2;
            const int src = popc32(todo) > q ? nth_set32(todo, q + 1) : -1;
            if (src >= 0) {
                const int pi = (int)blk * 32 + src;
                const int plen = (int)a.max_len < L - pi ? (int)a.max_len : L - pi;
                uint64_t first = 0, lst = 0; int matched = 0;
                boss_index_range(ix, codes + pi, plen < K - 1 ? plen : K - 1, &first, &lst, &matched, (int)a.min_seed_length);
                if (glane() == 0) { of[pi] = (uint32_t)first; ol[pi] = (uint32_t)lst; on[pi] = (uint8_t)matched; }
            }
            // drop the (up to) 8 lowest set bits
            for (int t = 0; t < 8 && todo; ++t) todo &= todo - 1;
        }
    }
}

// one lane group per read strand
__global__ void __launch_bounds__(256) k_premap(SeedArgs a) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) / kWarp;
    const uint64_t items = (uint64_t)a.n_reads * a.n_strands;
    for (uint64_t it = warp; it < items; it += nwarps) premap_item(a, (uint32_t)(it / a.n_strands), (uint32_t)(it % a.n_strands));
}

// BOSS::map_to_edges (boss.cpp:996-1045) for 32 read strands per warp, one per lane. The two regimes of the walk
// have opposite shapes, so the warp alternates between them instead of letting every strand run its own loop:
//   cold  a k-mer is looked up from scratch (BOSS::index: suffix-range table + tighten_range steps over 64-byte
//         blocks): quad-cooperative, 8 strands at a time, the strand's state fetched from / returned to its lane;
//   warm  the next k-mer is the child of the current edge: ONE 8-byte adjacency load per step and lane, 32
//         independent pointer chases per warp, all lanes in the same three-instruction-deep loop.
// (One quad per strand running both regimes kept 8 of 32 lanes busy on average: 8 quads in 8 different states.)
// Same results as map_to_edges() in seed_core.cuh, which the host emulation still runs.
__global__ void __launch_bounds__(128) k_seed(SeedArgs a) {
    const IndexView &ix = a.ix;
    const int K = (int)ix.k;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t items = (uint64_t)a.n_reads * a.n_strands;
    if (ix.kh_slots) {
        // ---- with the k-mer hash index: one strand per lane from start to end. Every k-mer costs one load: the
        // adjacency record while the walk is warm, a hash probe (+ the value on a hit) when it is not; the packed
        // k-mer is a rolling value. Every lane of a warp makes the same number of steps for reads of one length.
        const uint32_t sigma = ix.sigma;
        for (uint64_t item = warp * 32 + lane; item < items; item += nwarps * 32) {
            const uint32_t r = (uint32_t)(item / a.n_strands), s = (uint32_t)(item % a.n_strands);
            const uint64_t b = a.offsets[r];
            const int L = (int)(a.offsets[r + 1] - b);
            if (L < K) continue;
            const int nk = L - K + 1;
            const uint8_t *codes = (s ? a.cr : a.cf) + b;
            uint64_t *out = (s ? a.nodes_r : a.nodes_f) + a.koff[r];
            unsigned long long key = 0;
            int last_bad = -1;
            for (int j = 0; j < K - 1; ++j) {
                const uint32_t c = codes[j];
                if (c < 1 || c >= sigma) last_bad = j;
                key = (key >> 2) | ((unsigned long long)((c - 1) & 3u) << (2 * (K - 1)));
            }
            uint64_t edge = 0;
            const unsigned long long *keys = ix.kh_keys;
            const uint64_t slots = ix.kh_slots;
            // one step per k-mer, written with selects rather than branches: the lanes of a warp are in different
            // regimes (warm / cold / inside an invalid window) at the same step and should still issue together
            for (int pos = 0; pos < nk; ++pos) {
                const uint32_t c = codes[pos + K - 1];
                if (c < 1 || c >= sigma) last_bad = pos + K - 1;
                key = (key >> 2) | ((unsigned long long)((c - 1) & 3u) << (2 * (K - 1)));
                const bool ok = last_bad < pos;                        // no invalid character in the window
                const bool warm = ok && edge != 0, cold = ok && edge == 0;
                const unsigned long long want = key | (1ull << 63);
                uint64_t slot = kh_slot_of(want, slots);
                uint2 ar = make_uint2(0u, 0u);
                ulonglong2 ka = make_ulonglong2(0ull, 0ull), kb = ka;
                if (warm) ar = load_adj(ix, edge);                     // boss.cpp:1024-1043
                if (cold) {                                            // map_to_edge (boss.hpp:766-777): the k-mer's bucket
                    ka = __ldg(reinterpret_cast<const ulonglong2*>(keys + slot));
                    kb = __ldg(reinterpret_cast<const ulonglong2*>(keys + slot + 2));
                }
                int h = ka.x == want ? 0 : ka.y == want ? 1 : kb.x == want ? 2 : kb.y == want ? 3 : -1;
                // a full bucket without the k-mer: its run goes on in the next buckets (rare at this load factor)
                bool more = cold && h < 0 && ka.x != 0 && ka.y != 0 && kb.x != 0 && kb.y != 0;
                while (more) {
                    slot += 4; if (slot >= slots) slot = 0;
                    ka = __ldg(reinterpret_cast<const ulonglong2*>(keys + slot));
                    kb = __ldg(reinterpret_cast<const ulonglong2*>(keys + slot + 2));
                    h = ka.x == want ? 0 : ka.y == want ? 1 : kb.x == want ? 2 : kb.y == want ? 3 : -1;
                    more = h < 0 && ka.x != 0 && ka.y != 0 && kb.x != 0 && kb.y != 0;
                }
                uint64_t found = 0;
                if (cold && h >= 0) found = ix.kh_vals[slot + h];
                const uint64_t child = adj_child(ar, c);              // 0 for lanes that are not warm (ar = 0)
                const bool child_ok = (ar.y >> (8 + c)) & 1u;
                edge = warm ? child : found;                           // (0 when the window is invalid)
                const bool store = warm ? (child != 0 && child_ok) : (found != 0 && in_graph(ix, found));
                if (store) out[pos] = edge;
            }
        }
        return;
    }
    const int S = (ix.sfx_len && (int)ix.sfx_len <= K - 1) ? (int)ix.sfx_len : 0;
    const uint64_t base = ix.sigma - 1;
    for (uint64_t chunk = warp * 32; chunk < items; chunk += nwarps * 32) {
        const uint64_t item = chunk + lane;
        // this lane's strand
        const uint8_t *codes = nullptr; uint64_t *out = nullptr; const uint32_t *hints = nullptr;
        int nk = 0;
        if (item < items) {
            const uint32_t r = (uint32_t)(item / a.n_strands), s = (uint32_t)(item % a.n_strands);
            const uint64_t b = a.offsets[r];
            const int L = (int)(a.offsets[r + 1] - b);
            nk = L >= K ? L - K + 1 : 0;
            codes = (s ? a.cr : a.cf) + b;
            out = (s ? a.nodes_r : a.nodes_f) + a.koff[r];
            hints = a.hinted ? (s ? a.hint_r : a.hint_f) + (a.koff[r] >> 5) + r : nullptr;
        }
        int pos = 0;                 // first k-mer not resolved yet
        uint64_t edge = 0;           // warm: the edge of k-mer pos - 1
        while (__any_sync(0xffffffffu, pos < nk)) {
            if (pos >= nk) edge = 0;
            // ---- cold: strands without an edge look for their next k-mer that exists. The quads take the waiting
            // strands in lane order, 8 per round (quad q serves the q-th waiting lane)
            uint32_t waiting = __ballot_sync(0xffffffffu, edge == 0 && pos < nk);
            while (waiting) {
                const int q = lane >> 2;
                const bool need = popc32(waiting) > q;
                const int src = need ? nth_set32(waiting, q + 1) : 0;
                const uint64_t codes_s = __shfl_sync(0xffffffffu, (unsigned long long)codes, src);
                const uint64_t hints_s = __shfl_sync(0xffffffffu, (unsigned long long)hints, src);
                const uint64_t out_s = __shfl_sync(0xffffffffu, (unsigned long long)out, src);
                const int nk_s = __shfl_sync(0xffffffffu, nk, src);
                int i = __shfl_sync(0xffffffffu, pos, src);
                uint64_t found = 0;
                if (need) {                                            // uniform within the quad
                    const uint8_t *cd = (const uint8_t*)codes_s;
                    const uint32_t *hn = (const uint32_t*)hints_s;
                    LineCache lc;
                    for (; i < nk_s; ++i) {
                        if (hn) { i = next_hint(hn, nk_s, i); if (i >= nk_s) break; }
                        // map_to_edge (boss.hpp:766-777)
                        bool valid = true, zero = false;
                        for (int j = 0; j < K; ++j) { const uint32_t c = cd[i + j]; valid = valid && c < ix.sigma; zero = zero || c == 0; }
                        uint64_t e = 0;
                        if (valid) {
                            if (S && !zero) {
                                uint64_t slot = 0;
                                for (int j = S - 1; j >= 0; --j) slot = slot * base + (cd[i + j] - 1);
                                e = boss_index_slot(ix, cd + i, K - 1, slot);
                            } else {
                                e = boss_index(ix, cd + i, K - 1);
                            }
                            if (e) e = pick_edge(ix, lc, e, cd[i + K - 1]);
                        }
                        if (e) { found = e; break; }
                    }
                    if (found && glane() == 0) ((uint64_t*)out_s)[i] = in_graph(ix, found) ? found : 0;
                }
                // back to the strand's lane: the r-th waiting lane was served by quad r (if r < 8)
                const int my_rank = popc32(waiting & ((1u << lane) - 1u));
                const bool served = ((waiting >> lane) & 1u) && my_rank < 8;
                const uint64_t f_back = __shfl_sync(0xffffffffu, (unsigned long long)found, 4 * (my_rank & 7));
                const int i_back = __shfl_sync(0xffffffffu, i, 4 * (my_rank & 7));
                if (served) {
                    edge = f_back;
                    pos = f_back ? i_back + 1 : nk;                    // nothing left to find: the strand is done
                }
                for (int t = 0; t < 8 && waiting; ++t) waiting &= waiting - 1;
            }
            // ---- warm: every lane with an edge follows it (boss.cpp:1024-1043) until its next k-mer is missing; the
            // warp goes back to cold lookups once 8 strands wait for one (a full round of quads), or nobody is warm
            while (true) {
                const bool warm = edge != 0 && pos < nk;
                // one warp reduction counts both kinds of lanes: warm ones in the low byte, waiting ones above
                const uint32_t cnt = __reduce_add_sync(0xffffffffu, warm ? 1u : ((edge == 0 && pos < nk) ? 256u : 0u));
                if (!(cnt & 255u) || (cnt >> 8) >= 8u) break;
                if (warm) {
                    const uint32_t c = codes[pos + K - 1];
                    if (c >= ix.sigma) {                               // invalid character: this k-mer and the walk end
                        out[pos] = 0; edge = 0; ++pos;
                    } else if (!MGB_WIDE(ix)) {
                        const uint2 ar = load_adj(ix, edge);
                        edge = adj_child(ar, c);
                        out[pos] = (edge && ((ar.y >> (8 + c)) & 1u)) ? edge : 0;
                        ++pos;
                    } else {
                        const Adj ar = load_adj_any(ix, edge);
                        edge = adj_child(ar, c);
                        out[pos] = (edge && ((ar.ok >> c) & 1u)) ? edge : 0;
                        ++pos;
                    }
                }
            }
            if (pos >= nk) edge = 0;
        }
    }
}

#endif // MGB_ALIGN_KERNEL_ONLY

#ifndef MGB_ALIGN_MIN_BLOCKS
#define MGB_ALIGN_MIN_BLOCKS 4
#endif
#ifndef MGB_ALIGN_THREADS
#define MGB_ALIGN_THREADS 128
#endif
static constexpr int kAlignThreads = MGB_ALIGN_THREADS;
static constexpr int kGroupsPerBlock = kAlignThreads / kWarp;   // reads in flight per block of k_align
// one lane group (kWarp lanes, common.cuh) per read; kAlignThreads / kWarp groups per block
__global__ void __launch_bounds__(kAlignThreads, MGB_ALIGN_MIN_BLOCKS) k_align(const __grid_constant__ AlignArgs a) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    char *arena = a.arena + (size_t)warp * a.arena_stride;
    extern __shared__ __align__(16) char smem_raw[];
    char *smem = smem_raw + (threadIdx.x / kWarp) * a.slay.total;
    init_arena(a, arena);
    // the group's arena and on-chip working set: base + offsets from the kernel parameters
    const WarpMem mem { arena, &a.lay };
    const WarpSmem sm { smem, &a.slay };
    while (true) {
        unsigned int t = 0;
        if (wlane() == 0) t = atomicAdd(a.next, 1u);
        t = wbcast(t, 0);
        const bool have = t < a.n_list;
        if (!wany_full(have)) break;
        align_read(a, have ? a.read_list[t] : 0u, mem, sm, have);
    }
}

// launchers (the only entry points the host code uses)
#if !defined(MGB_ALIGN_KERNEL_ONLY)
cudaError_t launch_radj_bwd(unsigned grid, const RadjArgs &a) { k_radj_bwd<<<grid, 128>>>(a); return cudaGetLastError(); }
cudaError_t launch_sfx_extend(unsigned grid, const SfxArgs &a) { k_sfx_extend<<<grid, 128>>>(a); return cudaGetLastError(); }
cudaError_t launch_subk(unsigned grid, cudaStream_t s, const SubkArgs &a, uint32_t chunks_per_strand) {
    k_subk<<<grid, 128, 0, s>>>(a, chunks_per_strand);
    return cudaGetLastError();
}
cudaError_t launch_premap(unsigned grid, cudaStream_t s, const SeedArgs &a) { k_premap<<<grid, 256, 0, s>>>(a); return cudaGetLastError(); }
cudaError_t launch_seed(unsigned grid, cudaStream_t s, const SeedArgs &a) { k_seed<<<grid, 128, 0, s>>>(a); return cudaGetLastError(); }
#endif
cudaError_t launch_align(unsigned grid, size_t smem_block, cudaStream_t s, const AlignArgs &a) {
    k_align<<<grid, kAlignThreads, smem_block, s>>>(a);
    return cudaGetLastError();
}
// raises the kernel's dynamic shared memory limit to `smem_limit` and reports the resident blocks per SM
// for blocks of `smem_block` bytes
cudaError_t align_occupancy(size_t smem_limit, size_t smem_block, int *blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(k_align, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_align, kAlignThreads, smem_block);
}

#if !defined(MGB_WIDE_ONLY) && !defined(MGB_ALIGN_KERNEL_ONLY)
// PRIMARY graphs are DNA graphs in the block layout: first translation unit only
__global__ void __launch_bounds__(128) k_rc_tables(RcArgs a) {
    uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    uint64_t nquads = ((uint64_t)gridDim.x * blockDim.x) >> 2;
    for (uint64_t e = 1 + quad; e <= a.n; e += nquads) rc_tables_item(a, e);
}
// alphabet-independent kernels live in the first translation unit only
__global__ void __launch_bounds__(256) k_prepare(PrepArgs a) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    uint32_t nwarps = (gridDim.x * blockDim.x) / kWarp;
    for (uint32_t r = warp; r < a.n_reads; r += nwarps) prepare_read(a, r);
}
// round r (0-based) brings character k - 3 - r of every edge's k-mer to c_nxt[e]; kmer_shift = 2 * (k - 3 - r)
__global__ void __launch_bounds__(256) k_radj_gather(RadjArgs a, int kmer_shift) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nt = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = 1 + t; e <= a.n; e += nt) {
        const uint8_t c = a.c_cur[a.bwd_arr[e]];
        a.c_nxt[e] = c;
        if (a.kmer) {
            const unsigned long long km = a.kmer[e];
            if (km != kKmerBad) a.kmer[e] = c == 0 ? kKmerBad : (km | ((unsigned long long)(c - 1) << kmer_shift));
        }
    }
}
// k-mer hash index: one slot per k-mer without '$' (every such edge spells a distinct k-mer)
__global__ void __launch_bounds__(256) k_kmer_insert(const unsigned long long *kmer, uint64_t n, unsigned long long *keys,
                                                     uint32_t *vals, uint64_t slots) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nt = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = 1 + t; e <= n; e += nt) {
        const unsigned long long km = kmer[e];
        if (km == kKmerBad) continue;
        const unsigned long long key = km | (1ull << 63);
        uint64_t i = kh_slot_of(key, slots);
        while (atomicCAS(keys + i, 0ull, key) != 0ull) { if (++i == slots) i = 0; }
        vals[i] = (uint32_t)e;
    }
}
__global__ void __launch_bounds__(256) k_radj_pack(RadjArgs a, uint32_t multi_shift) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nt = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = 1 + t; e <= a.n; e += nt)
        a.radj[e] = make_uint2(a.bwd_arr[e], (uint32_t)a.c_cur[e] | ((uint32_t)a.multi[e] << multi_shift));
}
#endif

} // namespace MGB_KERNEL_NS
#endif
