#!/usr/bin/env python3
"""Hand transcription of the (graph, query, config) -> (CIGAR, sequence, clipping) triples that
metagraph/tests/graph/test_aligner.cpp asserts for DBGSuccinct graphs in BASIC mode.

Every entry cites the line of the TYPED_TEST in /root/reference/metagraph/tests/graph/test_aligner.cpp.
Config keys not listed keep the DBGAlignerConfig struct defaults (aligner_config.hpp:23-54).
`mask`: true  = build_graph_batch<DBGSuccinct> (dummy k-mers masked, test_dbg_helpers.cpp:369-385)
        false = std::make_shared<DBGSuccinct>(k) + add_sequence (no mask)
`extend`: the test also runs check_extend (max_seed_length = inf must give identical paths).
Running this script rewrites test_aligner_goldens.json next to it.
"""
import json, os

D = dict
G = []
def add(name, line, k, refs, query, expect, cfg=None, mask=True, extend=True, rc_query=False):
    G.append(D(name=name, line=line, k=k, refs=refs, query=query, rc_query=rc_query,
               cfg=cfg or {}, mask=mask, extend=extend, expect=expect))

M212 = D(matrix=[2, -1, -2])
def one(cigar, seq=None, clip=0, eclip=0, nm=None, size=None, offset=0):
    e = D(n_paths=1, cigar=cigar if isinstance(cigar, list) else [cigar], clipping=clip,
          end_clipping=eclip, offset=offset)
    if seq is not None: e["sequence"] = seq if isinstance(seq, list) else [seq]
    if nm is not None: e["num_matches"] = nm
    if size is not None: e["path_size"] = size
    return e

add("align_empty", 99, 4, ["CATTT"], "", D(n_paths=0), M212, extend=False)
add("align_sequence_much_too_short", 112, 4, ["CATTT"], "CA", D(n_paths=0), M212, extend=False)
add("align_sequence_too_short", 125, 4, ["CATTT"], "CAT", D(n_paths=0), D(matrix=[2,-1,-2], min_seed_length=4), extend=False)
add("align_big_self_loop", 137, 3, ["AAAA"], "AAAAAAAAA", one("9=", "AAAAAAAAA", nm=9, size=7), M212)
add("align_single_node", 166, 3, ["CAT"], "CAT", one("3=", "CAT", nm=3, size=1), M212)
add("align_straight", 195, 4, ["AGCTTCGAGGCCAA"], "AGCTTCGAGGCCAA", one("14=", "AGCTTCGAGGCCAA", nm=14, size=11), M212)
add("align_straight_min_path_score", 248, 4, ["AGCTTCGAGGCCAA"], "AGCTTCGAGGCCAA", D(n_paths=0),
    D(matrix=[2,-1,-2], min_path_score=100))
add("align_straight_with_N", 264, 4, ["AGCTTCGAGGCCAA"], "AGCTNCGAGGCCAA", one("4=1X9=", "AGCTTCGAGGCCAA", nm=13, size=11), M212)
add("align_straight_forward_and_reverse_complement", 296, 4, ["AGCTTCGAGGCCAA"], "AGCTTCGAGGCCAA",
    one("14=", "AGCTTCGAGGCCAA", nm=14, size=11), M212, rc_query=True)
add("align_ending_branch", 385, 4, ["AGCTTCGAA", "AGCTTCGAC"], "AGCTTCGAC", one("9=", "AGCTTCGAC", nm=9, size=6), M212)
add("align_branch", 417, 6, ["AGCTTCGAATATTTGTT", "AGCTTCGACGATTTGTT"], "AGCTTCGACGATTTGTT",
    one("17=", "AGCTTCGACGATTTGTT", nm=17, size=12), M212)
add("align_branch_with_cycle", 449, 4, ["AGCTTCGAATATTTGTT", "AGCTTCGACGATTTGTT"], "AGCTTCGACGATTTGTT",
    one("17=", "AGCTTCGACGATTTGTT", nm=17, size=14), M212)
add("repetitive_sequence_alignment", 481, 3, ["AGGGGGGGGGAAAAGGGGGGG"], "AGGGGG", one("6=", "AGGGGG", nm=6, size=4), M212)
add("variation", 510, 4, ["AGCAACTCGAAA"], "AGCAATTCGAAA", one("5=1X6=", "AGCAACTCGAAA", nm=11, size=9), M212)
add("variation_in_branching_point", 540, 4, ["TTAAGCAACTCGAAA", "TTAAGCAAGTCGAAA"], "TTAAGCAATGGGAAA",
    one("8=3X4=", ["TTAAGCAACTCGAAA", "TTAAGCAAGTCGAAA"], nm=12, size=12),
    D(matrix=[2,-1,-2], gap_open=-3, gap_ext=-1))
add("multiple_variations", 579, 4, ["ACGCAACTCTCTGAACTTGT"], "ACGCAATTCTCTGTATTTGT",
    one("6=1X6=1X1=1X4=", "ACGCAACTCTCTGAACTTGT", nm=17, size=17), M212)
add("align_noise_in_branching_point", 609, 4, ["AAAACTTTTTT", "AAAATTGGGGG"], "AAAATTTTTTT",
    D(n_paths=1, alternatives=[D(orientation=0, cigar="4=1D7=", sequence="AAAACTTTTTTT"),
                               D(orientation=1, cigar="7=1D4=", sequence="AAAAAAACTTTT")],
      clipping=0, end_clipping=0, offset=0, num_matches=11, path_size=9),
    D(matrix=[2,-3,-3], gap_open=-3, gap_ext=-1))
add("alternative_path_basic", 650, 4, ["ACAATTTTTTTT", "ACAATTTTTGTT", "ACAAGTTTTTTT", "ACAAGTTTTGTT"], "ACAACTTTTCTT",
    D(n_paths=2, cigar=["4=1X4=1X2="], clipping=0, end_clipping=0, offset=0, num_matches=10),
    D(matrix=[2,-1,-2], gap_open=-3, gap_ext=-1, num_alternative_paths=2))
add("align_multiple_misalignment", 684, 4, ["AAAGCGGACCCTTTCCGTTAT"], "AAAGGGGACCCTTTTCGTTAT",
    one("4=1X9=1X6=", "AAAGCGGACCCTTTCCGTTAT", nm=19, size=18), M212)
add("align_insert_non_existent", 714, 4, ["TTTCCTTGTT"], "TTTCCATTGTT", one("5=1I5=", "TTTCCTTGTT", nm=10, size=7),
    D(matrix=[2,-1,-2], gap_open=-3, gap_ext=-3))
add("align_insert_multi", 746, 4, ["TTTCCTTGTT"], "TTTCCAATTGTT", one("5=2I5=", "TTTCCTTGTT", nm=10, size=7),
    D(matrix=[2,-1,-2], gap_open=-3, gap_ext=-3))
add("align_insert_long", 779, 4, ["TTTCCTTGTT"], "TTTCCAAAAAAAAATTGTT", one("5=9I5=", "TTTCCTTGTT", nm=10, size=7),
    D(matrix=[2,-1,-1], gap_open=-1, gap_ext=-1))
add("align_insert_long_offset", 812, 5, ["TTTCCGGTTGTTA"], "TTTCCGCAAAAAAAAATTGTTA",
    one(["6=1X9I6=", "6=9I1X6="], "TTTCCGGTTGTTA", nm=12, size=9), D(matrix=[2,-1,-1], gap_open=-1, gap_ext=-1))
add("align_delete", 847, 4, ["TTCGATTGGCCT"], "TTCGATGGCCT", one(["6=1D5=", "5=1D6="], "TTCGATTGGCCT", size=9),
    D(matrix=[2,-1,-2], gap_open=-3, gap_ext=-3), extend=False)
add("align_gap", 885, 4, ["TTTCTGTATACCTTGGCGCTCTC"], "TTTCTGTATAGGCGCTCTC", one("10=4D9=", "TTTCTGTATACCTTGGCGCTCTC", nm=19, size=20),
    D(matrix=[2,-1,-2], gap_open=-3, gap_ext=-3))
add("align_gap_after_seed", 918, 4, ["TTTCCCTTGGCGCTCTC"], "TTTCGGCGCTCTC", one("4=4D9=", "TTTCCCTTGGCGCTCTC", nm=13, size=14),
    D(matrix=[2,-1,-2], gap_open=-3, gap_ext=-1))
add("align_loop_deletion", 951, 4, ["AAAATTTTCGAGGCCAA"], "AAAACGAGGCCAA", one("4=3D9=", "AAAATTTCGAGGCCAA", nm=13, size=13),
    D(unit=1, gap_open=-1, gap_ext=-1))
add("align_straight_long_xdrop", 986, 4,
    ["AGCTTCGAGGCCAAGCCTGACTGATCGATGCATGCTAGCTAGTCAGTCAGCGTGAGCTAGCAT",
     "AGCTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"],
    "AGCTTCGAGGCCAAGCCTGACTGATCGATGCATGCTAGCTAGTCAGTCAGCGTGAGCTAGCAT",
    one("63=", "AGCTTCGAGGCCAAGCCTGACTGATCGATGCATGCTAGCTAGTCAGTCAGCGTGAGCTAGCAT", nm=63, size=60),
    D(matrix=[2,-3,-3], xdrop=30, rel_score_cutoff=0.8))
add("align_drop_seed", 1018, 4, ["TTTCCCTGGCGCTCTC"], "TTTCCGGGGCGCTCTC", one("7S9=", "GGCGCTCTC", clip=7, nm=9, size=6),
    D(matrix=[2,-6,-6], gap_open=-10, gap_ext=-4, xdrop=6))
add("align_long_gap_after_seed", 1052, 4, ["TTTCCCTTAAGGCGCTCTC"], "TTTCGGCGCTCTC", one("4S9=", "GGCGCTCTC", clip=4, nm=9, size=6),
    D(matrix=[2,-1,-2], gap_open=-5, gap_ext=-1))
add("align_repeat_sequence_no_delete_after_insert", 1084, 27,
    ["TTTGTGGCTAGAGCTCGAGATCGCGCGGCCACAATTGACAAATGAGATCTAATTAAACTAAAGAGCTTCTGCACAGCAAAAGAAACTGTCATC"],
    "TTTGTGGCTAGAGCTCGAGATCGCGCGGCCACAATTGACAAATGACAAATGTGATCTAATGAAACTAAAGAGCTTCTGCACAGCAAAAGAAACTGTCATC",
    one(["45=7I8=1X39=", "45=5I1=2I7=1X39=", "44=2I1=5I8=1X39=", "44=3I1=4I8=1X39=", "44=4I1=3I8=1X39="],
        "TTTGTGGCTAGAGCTCGAGATCGCGCGGCCACAATTGACAAATGAGATCTAATTAAACTAAAGAGCTTCTGCACAGCAAAAGAAACTGTCATC", nm=92, size=67),
    D(matrix=[2,-3,-3], gap_open=-3, gap_ext=-3), extend=False)
add("align_clipping1", 1147, 4, ["GGCCTGTTTG"], "ACCCTGTTTG", one("2S8=", "CCTGTTTG", clip=2, nm=8, size=5), M212)
add("align_clipping2", 1178, 4, ["AAAAGCTTCGAGGCCAA"], "TTAGCTTCGAGGCCAA", one("2S14=", "AGCTTCGAGGCCAA", clip=2, nm=14, size=11), M212)
add("align_long_clipping", 1208, 4, ["TTTTTTTAAAAGCTTCGAGGCCAA"], "CCCCCCCAAAAGCTTCGAGGCCAA",
    one("7S17=", "AAAAGCTTCGAGGCCAA", clip=7, nm=17, size=14), M212)
add("align_end_clipping", 1239, 4, ["AAAAGCTTCGAGGCCAATTTTTTT"], "AAAAGCTTCGAGGCCAACCCCCCC",
    one("17=7S", "AAAAGCTTCGAGGCCAA", eclip=7, nm=17, size=14), M212)
add("align_clipping_min_cell_score", 1269, 7, ["AAAAGCTTTCGAGGCCAA"], "ACCTTTCGAGGCCAA",
    one("2S13=", "CTTTCGAGGCCAA", clip=2, nm=13, size=7),
    D(matrix=[2,-1,-2], min_cell_score=-2147483548, min_path_score=-2147483548))
add("align_low_similarity", 1301, 27, ["CTAGAACTTAAAGTATAATAATACTAATAATAAAATAAAATACA"],
    "CTAGAACTTAAAGTATAATAATACTAATAAAAGTACAATACA", D(n_paths=1), D(matrix=[2,-3,-3]), extend=False)
add("align_low_similarity2", 1325, 27, ["GCCACAATTGACAAATGAGATCTAATTAAACTAAAGAGCTTCTGCACAGCAAAAGAAACTGTCATC"],
    "GCCACAATTGACAAATGACAAATGTGATCTAATGAAACTAAAGAGCTTCTGCACAGCAAAAGAAACTGTCATC", D(n_paths=1), D(matrix=[2,-3,-3]), extend=False)
add("align_low_similarity5", 1432, 31, ["GTCGTCAGATCGGAAGAGCGTCGTGTAGGGAAAGGTCTTCGCCTGTGTAGATCTCGGTGGTCG"],
    "GTCAGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTTCCTGGTGGTGTAGATC", D(n_paths=1), D(matrix=[2,-3,-3]), mask=False)
add("align_suffix_seed_snp_min_seed_length", 1447, 7, ["AAAAGCTTTCGAGGCCAA"], "ACCTTTCGAGGCCAA",
    one("2S13=", "CTTTCGAGGCCAA", clip=2, nm=13, size=7),
    D(matrix=[2,-1,-2], min_seed_length=2, min_cell_score=-2147483548, min_path_score=-2147483548), mask=False)
add("align_both_directions2", 1579, 11, ["GTAGTGCTAGCTGTAGTCGTGCTGATGC"], "GTAGTGCTACCTGTAGTCGTGGTGATGC",
    D(n_paths=1, sequence=["GTAGTGCTAGCTGTAGTCGTGCTGATGC"], path_size=18, offset=0), M212)
add("align_nodummy_fwd_only", 1634, 7, ["AAAAGCTTTCGAGGCCAA"], "AAAAGTTTTCGAGGCCAA",
    one("6S12=", "TTTCGAGGCCAA", clip=6, nm=12, size=6), D(matrix=[2,-1,-2], forward_and_reverse_complement=False))
add("align_nodummy_both", 1634, 7, ["AAAAGCTTTCGAGGCCAA"], "AAAAGTTTTCGAGGCCAA",
    one("5=1X12=", "AAAAGCTTTCGAGGCCAA", nm=17, size=12), D(matrix=[2,-1,-2], forward_and_reverse_complement=True))
add("align_seed_to_end", 1679, 5, ["ATCCCTTTTAAAA"], "ATCCCGGGGGGGGGGGGGGGGGTTTTAAAA", D(n_paths=1), M212)
add("align_bfs_vs_dfs_xdrop", 1696, 31,
    ["TCGGGGCAAGAAACACACAGCCTTCTCATCCAAGGGCCTCAGTGATGAAGAGTACGATGAGTACAAGAGGATCAGAGAAGAAAGGAATGGCAAATACTCCATAGAAGAGTACCTTCAGGACAGGGACAGATACTATGAGGAGGTGGCCAT",
     "TCGGGGCAAGAAACACACAGCCTTCTCATCCAAGGGCCTCAGTGATGAAGAGTACGATGAGTACAAGAGAATCAGAGAGGAGAGGAATGGCAAATACTCAATAGAGGAATACCTCCAAGATAGGGACAGATACTATGAAGAGCTTGCCAT"],
    "TCGGGGCAAGAAACACACAGCCTTCTCATCCAAGGGCCTCAGTGATGATGAGTACGATGAGTACAAGAGCATCAGAGAGGAGAGGAATGGCAAATACTCAATAGAGGAATACCTCCAAGATAGGGACAGATACTATGAAGAGCTTGCCAT",
    D(n_paths=1, cigar=["48=1X20=1X80="]), D(matrix=[2,-3,-3], xdrop=27, min_seed_length=0, max_seed_length=0, rel_score_cutoff=0.8),
    extend=False)
add("align_dummy", 1720, 7, ["AAAAGCTTTCGAGGCCAA"], "AAAAGTTTTCGAGGCCAA", one("5=1X12=", "AAAAGCTTTCGAGGCCAA", nm=17, size=12),
    D(matrix=[2,-1,-2], min_seed_length=5), mask=False)
add("align_extended_insert_after_match", 1751, 27,
    ["CGTGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAAGCC",
     "CGTGGCCCAGGCCCAGGCCCAGCCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAGGCCCAAGCC"],
    "CGTGGCCCAGGCCCAGGCCCAGTGGGCGTTGGCCCAGGCGGCCACGGTGGCTGCGCAGGCCCGCCTGGCACAAGCCACGCTG",
    D(n_paths=1, score=52), D(matrix=[2,-3,-3], min_seed_length=15), mask=False)

here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "test_aligner_goldens.json"), "w") as f:
    json.dump(G, f, indent=1)
print(len(G), "goldens written")
