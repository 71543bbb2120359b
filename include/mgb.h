/* mgb.h — C ABI of the B200-native `metagraph align` hot path.
 *
 * This is the drop-in boundary: a maintainer of ratschlab/metagraph binds these entry
 * points from a thin `class B200Aligner : public IDBGAligner` (see INTEGRATION.md and
 * metagraph_b200/csrc/b200_aligner.hpp).  Plain pointers and sizes only; no exceptions,
 * STL or torch types cross this boundary.  Every function returns MGB_OK (0) or a
 * negative error code; mgb_last_error() returns a thread-local message.
 *
 * Reference interfaces replaced (paths relative to metagraph/src):
 *   mgb_index_create      <- DBGSuccinct / boss::BOSS query state: W (wavelet_tree), last
 *                            (bit_vector), F (graph/representation/succinct/boss.hpp:499-525),
 *                            optional dummy mask valid_edges_ (dbg_succinct.hpp:195)
 *   mgb_map_to_nodes      <- map_to_nodes_sequentially() (graph/representation/base/
 *                            sequence_graph.cpp:541-551 -> dbg_succinct.cpp:285-305 ->
 *                            boss.cpp:996-1045 BOSS::map_to_edges)
 *   mgb_config_check      <- DBGAligner::DBGAligner (graph/alignment/dbg_aligner.cpp:26-60)
 *   mgb_align_batch       <- IDBGAligner::align_batch (graph/alignment/dbg_aligner.hpp:32-33,
 *                            dbg_aligner.cpp:251-355) with Seeder = SuffixSeeder<UniMEMSeeder>,
 *                            Extender = DefaultColumnExtender
 *   mgb_config_t          <- DBGAlignerConfig (graph/alignment/aligner_config.hpp:18-94)
 *   mgb_alignment_t       <- Alignment (graph/alignment/alignment.hpp:132-331)
 *   mgb_boss_build        <- BOSSConstructor / BOSS::Chunk (boss_chunk_construct.cpp:341-462,
 *                            boss_chunk.cpp:33-133); host-side, not part of the timed path
 */
#ifndef MGB_H_
#define MGB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_OK 0
#define MGB_ERR_INVALID_ARGUMENT (-1)
#define MGB_ERR_CUDA (-2)
#define MGB_ERR_BAD_CONFIG (-3)   /* reference: std::runtime_error in DBGAligner ctor (dbg_aligner.cpp:55-56) */
#define MGB_ERR_UNSUPPORTED (-4)  /* a configuration or graph the kernels do not serve (chaining, labels, ...) */
#define MGB_ERR_OVERFLOW (-5)     /* a read exceeded the largest per-read work arena */
#define MGB_ERR_NO_DEVICE (-6)
#define MGB_ERR_NO_MEMORY (-7)    /* host allocation failed (index construction) */

#define MGB_ALPHABET_DNA 0      /* "$ACGT", sigma = 5 (kmer/alphabets.hpp:64-79) */
#define MGB_ALPHABET_PROTEIN 1  /* "$ABCDEFGHIJKLMNOPQRSTUVWYZX", sigma = 27 (kmer/alphabets.hpp:29-38) */

/* Cigar::Operator (graph/alignment/aligner_cigar.hpp:18-25); printed as "SX=DIG" */
#define MGB_OP_CLIPPED 0
#define MGB_OP_MISMATCH 1
#define MGB_OP_MATCH 2
#define MGB_OP_DELETION 3
#define MGB_OP_INSERTION 4
#define MGB_OP_NODE_INSERTION 5

typedef struct mgb_index mgb_index_t;
typedef struct mgb_results mgb_results_t;

/* DBGAlignerConfig (aligner_config.hpp:23-61). Penalties are negative as in the struct
 * (cli/align.cpp:33-69 negates the positive CLI values). */
typedef struct mgb_config {
    uint64_t num_alternative_paths;
    uint64_t min_seed_length;
    uint64_t max_seed_length;
    uint64_t max_num_seeds_per_locus;
    int32_t min_cell_score;
    int32_t min_path_score;
    int32_t xdrop;
    int32_t reserved0;
    double min_exact_match;
    double max_nodes_per_seq_char;
    double max_ram_per_alignment;
    double rel_score_cutoff;
    int8_t gap_opening_penalty;
    int8_t gap_extension_penalty;
    int8_t left_end_bonus;
    int8_t right_end_bonus;
    uint8_t forward_and_reverse_complement;
    uint8_t global_xdrop;
    uint8_t allow_left_trim;
    uint8_t no_backtrack;
    uint8_t seed_complexity_filter; /* sdust(T=20, W=64) on seed windows, restated from its definition (library not vendored) */
    /* What mgb_alignment_t::nodes carries (not a DBGAlignerConfig field): MGB_NODES_U64 = the node path (default),
     * MGB_NODES_NONE = nothing (nodes == NULL, num_nodes still set) for consumers that only print alignments
     * (the TSV branch of cli/align.cpp:254-307 never reads Alignment::get_nodes()): the node ids are 4/5 of the
     * bytes a 150 bp alignment brings back from the device. */
    uint8_t result_nodes;
    /* 1 = never take the exact-path shortcut: when every k-mer of a strand is in the graph, the seed at query
     * position 0 extends, provably, to the whole read matched along those nodes (DESIGN.md "Exact-path shortcut":
     * conditions on the scores, proof, what the reference does instead); the kernel then writes that alignment
     * without running the extension. Results are identical either way; the switch exists for measurements. */
    uint8_t no_exact_path_shortcut;
    uint8_t reserved1[5];
    int8_t score_matrix[128][128];
} mgb_config_t;
enum { MGB_NODES_U64 = 0, MGB_NODES_NONE = 1 };

/* One alignment of one read (Alignment, alignment.hpp:323-331). Pointers reference memory
 * owned by the enclosing mgb_results_t. */
typedef struct mgb_alignment {
    uint32_t read_index;
    uint8_t orientation;       /* 1: the reverse complement of the read was aligned */
    uint8_t pad[3];
    int32_t score;
    uint32_t offset;           /* chars discarded from the first node's prefix */
    uint32_t query_begin;      /* == leading clipping; query_view = strand[query_begin, +query_len) */
    uint32_t query_len;
    uint32_t num_nodes;
    uint32_t sequence_len;
    uint32_t num_cigar_ops;
    uint32_t reserved;
    const uint64_t *nodes;     /* DBG node ids == BOSS edge indexes */
    const char *sequence;      /* spelling of the path (not NUL-terminated) */
    const uint32_t *cigar;     /* (len << 3) | op, clipping included */
} mgb_alignment_t;

/* Per-batch counters (dbg_aligner.cpp:341-351 trace line; dp_cells is the GCUPS numerator,
 * SURVEY 8d). */
typedef struct mgb_stats {
    uint64_t num_seeds;
    uint64_t num_extensions;
    uint64_t num_explored_nodes;
    uint64_t dp_cells;
    uint64_t dp_columns;
    uint64_t num_reads_retried;   /* reads re-run with a larger work arena */
    double seed_kernel_ms;        /* device time of the seeding kernel(s) */
    double align_kernel_ms;       /* device time of the extension kernel(s) */
    double h2d_ms, d2h_ms;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t kernel_launches;
} mgb_stats_t;

const char* mgb_last_error(void);
int mgb_device_count(void);

/* ---- index -------------------------------------------------------------------------- */

/* Uploads a BOSS table into HBM in the flat block layout described in DESIGN.md.
 *   W[0..n]     edge labels incl. the position-0 placeholder, values in [0, 2*sigma)
 *   last[0..n]  0/1 bytes
 *   F[sigma]    boss F array
 *   valid       optional 0/1 bytes (dummy-k-mer mask); NULL = mask dropped, which is what
 *               `metagraph align` does (cli/align.cpp:335-339)
 *   k           DBG k-mer length (BOSS node length + 1)
 *   suffix_len  length of the indexed suffix ranges (boss.hpp:516-525); 0 = choose */
int mgb_index_create(const uint8_t *W, const uint8_t *last, uint64_t n_plus_1,
                     const uint64_t *F, const uint8_t *valid, uint32_t k, int alphabet,
                     uint32_t suffix_len, int device, mgb_index_t **out);
/* DeBruijnGraph::get_mode() of the graph (sequence_graph.hpp:160): 0 = BASIC (default), 1 = CANONICAL (built with
 * --mode canonical: the graph holds the reverse complement of every k-mer; dbg_aligner.cpp:224-226, 646-722).
 * 2 = PRIMARY (built with --mode primary: one k-mer of every reverse-complement pair): the index answers with the
 * semantics of the CanonicalDBG wrapper `metagraph align` puts around such graphs (canonical_dbg.cpp): node ids
 * above mgb_index_num_edges() denote reverse complements. The first call builds two 4-byte-per-edge tables on the
 * device; mgb_map_to_nodes then follows CanonicalDBG::map_to_nodes_sequentially (both strands, canonical_dbg.cpp:55-146).
 * Set the mode once, right after mgb_index_create and before the index is shared between threads. */
int mgb_index_set_mode(mgb_index_t *index, int mode);
void mgb_index_destroy(mgb_index_t *index);
uint64_t mgb_index_num_edges(const mgb_index_t *index);
uint64_t mgb_index_device_bytes(const mgb_index_t *index);
uint32_t mgb_index_k(const mgb_index_t *index);

/* ---- hot path ------------------------------------------------------------------------ */

/* Default configs. mgb_config_init = DBGAlignerConfig{} + dna_scoring_matrix(2,-1,-2)
 * (what the reference unit tests use); mgb_config_init_cli = `metagraph align` defaults
 * (cli/config/config.hpp:114-145) for a graph with k-mer length k: DNA match 2 / mismatch -3, protein
 * BLOSUM62 and forward strand only. seed_complexity_filter is left 0 (the CLI turns it on; its sdust
 * dependency is restated here) — set it to 1 for flag-free `metagraph align` behaviour. */
void mgb_config_init(mgb_config_t *config);
void mgb_config_init_cli(mgb_config_t *config, uint32_t k, int alphabet);

/* What the DBGAligner constructor checks (dbg_aligner.cpp:37-60: seed-length normalisation, check_config_scores ->
 * std::runtime_error("Error: sum of min_cell_score and lowest penalty too low.")) plus what this build does not serve:
 * MGB_OK, MGB_ERR_BAD_CONFIG or MGB_ERR_UNSUPPORTED with the text in mgb_last_error(). mgb_align_batch makes the same
 * check; this entry point lets a binding fail at construction time, as the reference does. No device work. */
int mgb_config_check(const mgb_index_t *index, const mgb_config_t *config);

/* Exact seeding only: node ids of all k-mers of each sequence (0 = not in graph), forward
 * strand. seqs: concatenated characters, offsets[n_seqs + 1]. out_nodes must hold
 * sum(max(0, len_i - k + 1)) entries, laid out consecutively per sequence. */
int mgb_map_to_nodes(const mgb_index_t *index, const char *seqs, const uint64_t *offsets,
                     uint32_t n_seqs, uint64_t *out_nodes);

/* Full seed-and-extend of a batch of reads; results in input order. */
int mgb_align_batch(const mgb_index_t *index, const mgb_config_t *config, const char *seqs,
                    const uint64_t *offsets, uint32_t n_reads, mgb_results_t **out);

/* A batch of >= 128k reads is split into up to 8 contiguous pieces that run on two streams, so the
 * download and unpacking of one piece overlap the kernels of the next (reads are independent,
 * dbg_aligner.cpp:251-355). max_pieces caps the split; 1 = one piece (device timers in mgb_stats_t
 * then do not overlap), 0 = automatic (default). Process-wide. */
void mgb_set_pipeline_pieces(uint32_t max_pieces);
/* Host threads for the unpacking of results (mgb_align_batch, mgb_results_export / import): launchers such as
 * torchrun start every rank with OMP_NUM_THREADS=1; a rank that owns cores/ranks of them says so here. Process-wide.
 * The library never uses more than 16 threads per call for this work whatever the setting (the loops are short). */
void mgb_set_host_threads(int num_threads);

uint32_t mgb_results_num_reads(const mgb_results_t *results);
/* alignments of read i are [first, first + count) in mgb_results_alignments() */
void mgb_results_read_range(const mgb_results_t *results, uint32_t read, uint64_t *first,
                            uint32_t *count);
uint64_t mgb_results_num_alignments(const mgb_results_t *results);
const mgb_alignment_t* mgb_results_alignments(const mgb_results_t *results);
const mgb_stats_t* mgb_results_stats(const mgb_results_t *results);
void mgb_results_free(mgb_results_t *results);

/* Multi-GPU: reads shard across ranks (one process per GPU; dbg_aligner.cpp:263-354 keeps no state
 * between reads) and rank 0 collects the per-rank results. mgb_results_export writes a result set as one
 * relocatable byte block (the records as the kernel packed them, no pointers) into caller memory -- e.g. a pinned
 * buffer handed to ncclSend; mgb_results_import rebuilds a result set from such a block on the receiving rank,
 * with read_index_base added to the read indexes (the shard's first read). */
uint64_t mgb_results_export_bytes(const mgb_results_t *results);
int mgb_results_export(const mgb_results_t *results, void *dst, uint64_t capacity);
int mgb_results_import(const void *block, uint64_t bytes, uint32_t read_index_base, mgb_results_t **out);

/* ---- host-side construction (index build, untimed) ------------------------------------ */

/* Builds BOSS arrays from sequences (batch construction incl. dummy edges). Caller frees
 * with mgb_boss_free. force_source_dummies mimics BOSS::add_sequence leftovers. (k+1)-mers are packed into 128-bit
 * keys, 256-bit ones beyond (DNA: k <= 85, protein: k <= 51; larger k: MGB_ERR_UNSUPPORTED). */
typedef struct mgb_boss {
    uint64_t n_plus_1;
    uint8_t *W;
    uint8_t *last;
    uint64_t F[32];
    uint32_t k;        /* DBG k */
    int32_t alphabet;
} mgb_boss_t;
int mgb_boss_build(const char *seqs, const uint64_t *offsets, uint32_t n_seqs, uint32_t k,
                   int alphabet, int force_source_dummies, int num_threads, mgb_boss_t *out);
/* Text for the last failed mgb_boss_build() on this thread ("" after a success). */
const char* mgb_boss_last_error(void);
void mgb_boss_free(mgb_boss_t *boss);
/* Dummy-k-mer mask as DBGSuccinct::mask_dummy_kmers computes it (dbg_succinct.cpp:917-932);
 * valid must hold n_plus_1 bytes. */
int mgb_boss_mask_dummy(const mgb_boss_t *boss, uint8_t *valid);

/* Reads a graph file written by the reference (`.dbg`: DBGSuccinct::serialize, dbg_succinct.cpp:690-803 ->
 * BOSS::serialize, boss.cpp:262-277) into plain BOSS arrays for mgb_index_create(). States SMALL and STAT
 * (boss.hpp:325); DYN / FAST files are refused. *mode = DeBruijnGraph::Mode (0 BASIC, 1 CANONICAL,
 * 2 PRIMARY; pass it to mgb_index_set_mode), *state = BOSS::State. The suffix
 * range index and the optional .edgemask file are not read. Free with mgb_boss_free(). On failure returns
 * MGB_ERR_INVALID_ARGUMENT and mgb_dbg_last_error() describes why. */
int mgb_dbg_load(const char *path, mgb_boss_t *out, int *mode, int *state);
/* The file's own index of suffix ranges (BOSS::serialize_suffix_ranges / load_suffix_ranges, boss.cpp:396-426; what
 * BOSS::get_initial_range consults, boss.hpp:638-664), decoded: *suffix_len = indexed suffix length s (0: the file has
 * none), ranges[2 i], ranges[2 i + 1] = first edge and one past the last edge of the nodes whose last s characters
 * are the i-th s-mer (index = sum over positions p of (code_p - 1) (sigma - 1)^p; an empty range has both equal).
 * mgb_index_create builds its own, deeper table on the device and does not need this one: the entry point exists for
 * callers that keep using the file's table and for checking one against the other. Free with
 * mgb_dbg_free_suffix_ranges(). */
int mgb_dbg_load_suffix_ranges(const char *path, uint32_t *suffix_len, uint64_t **ranges, uint64_t *n_ranges);
void mgb_dbg_free_suffix_ranges(uint64_t *ranges);
const char* mgb_dbg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MGB_H_ */
