// Stand-alone check of the C++ IDBGAligner-shaped shim (metagraph_b200/csrc/b200_aligner.hpp)
// against a reference golden (tests/graph/test_aligner.cpp:264 align_straight_with_N).
// Linked by tests/test_shim_cpp.py against a C-ABI library (host emulation on CPU-only machines).
#include <cstdio>
#include <string>
#include <vector>

#include "../../metagraph_b200/csrc/b200_aligner.hpp"

int main(int argc, char **argv) {
    const std::string ref = "AGCTTCGAGGCCAA";
    uint64_t offsets[2] = { 0, ref.size() };
    mgb_boss_t boss;
    if (mgb_boss_build(ref.data(), offsets, 1, 4, MGB_ALPHABET_DNA, 0, 1, &boss) != MGB_OK) return 2;
    std::vector<uint8_t> valid(boss.n_plus_1);
    if (mgb_boss_mask_dummy(&boss, valid.data()) != MGB_OK) return 3;
    mgb_shim::B200Graph graph(boss.W, boss.last, boss.n_plus_1, boss.F, valid.data(), 4);
    mgb_boss_free(&boss);
    mgb_config_t cfg;
    mgb_config_init(&cfg);
    mgb_shim::B200Aligner aligner(graph, cfg);
    int n = 0, bad = 0;
    aligner.align_batch({ { "q1", "AGCTNCGAGGCCAA" }, { "q2", "AGCTTCGAGGCCAA" } },
        [&](const std::string &header, mgb_shim::AlignmentResults &&r) {
            std::printf("%s\t%s\t%s\t%d\n", header.c_str(), r.query.c_str(),
                        r.alignments.empty() ? "*" : r.alignments[0].cigar_string().c_str(),
                        r.alignments.empty() ? 0 : r.alignments[0].score);
            const char *exp = n == 0 ? "4=1X9=" : "14=";
            int exp_score = n == 0 ? 24 : 28;
            if (r.alignments.size() != 1 || r.alignments[0].cigar_string() != exp
                    || r.alignments[0].score != exp_score || r.alignments[0].nodes.size() != 11)
                ++bad;
            ++n;
        });
    // bad config must surface as an exception, like the reference's ctor (dbg_aligner.cpp:55-56)
    cfg.min_cell_score = INT32_MIN;
    bool thrown = false;
    try { mgb_shim::B200Aligner(graph, cfg).align("AGCTTCGAGG"); } catch (const std::runtime_error&) { thrown = true; }
    // a graph file written by the reference (examples/data/graphs/test_DNA_graph.dbg, k = 20): every
    // query of examples/data/test_DNA_query.fa is a path of the graph
    int dbg_bad = 0;
    if (argc > 1) {
        const std::string dbg_path = argv[1];
        mgb_shim::B200Graph g2(dbg_path);
        mgb_config_t c2;
        mgb_config_init_cli(&c2, g2.get_k(), MGB_ALPHABET_DNA);
        auto res = mgb_shim::B200Aligner(g2, c2).align("ACGTACGTACGTACGTACGTACGTACGTACGTACGT");
        if (res.alignments.size() != 1 || res.alignments[0].cigar_string() != "36=") ++dbg_bad;
        std::printf("dbg\t%s\n", res.alignments.empty() ? "*" : res.alignments[0].cigar_string().c_str());
    }
    return (n == 2 && bad == 0 && thrown && dbg_bad == 0) ? 0 : 1;
}
