"""Pins the CPU restatement (oracle/) against the reference's own golden vectors:
  * tests/graph/succinct/test_boss.cpp:137-161 (W / last / F of test_construct.fa, k=3)
  * tests/graph/succinct/test_boss.cpp:166-221 (SmallGraphTraversal outgoing-edge table)
  * tests/graph/succinct/test_boss.cpp:2317-2345 (map_to_edges closed form)
  * tests/graph/test_aligner.cpp (47 (graph, query, config) -> CIGAR/sequence triples)
  * integration_tests/test_align.py:22-57, 176-206 (CLI-default TSV lines on genome.MT.fa)
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from metagraph_b200.config import (DBGAlignerConfig, SIZE_MAX, cli_defaults, dna_scoring_matrix,
                                   struct_defaults, unit_scoring_matrix)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def read_fasta(path):
    seqs, cur = [], []
    names = []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur)); cur = []
            names.append(line[1:])
        elif line:
            cur.append(line)
    if cur:
        seqs.append("".join(cur))
    return names, seqs


def read_fastq(path):
    lines = [l.rstrip("\n") for l in open(path)]
    return [l[1:] for l in lines[0::4]], lines[1::4]


# ---------------------------------------------------------------- BOSS ----
def test_boss_construction_small():
    _, seqs = read_fasta(os.path.join(GOLD, "test_construct.fa"))
    g = O.OracleGraph(4, seqs)       # BOSS k = 3  <=>  DBG k = 4
    W, last, F, _ = g.arrays()
    assert "".join(str(x) for x in last) == "00011101101111111111111"
    # golden is quoted for the DNA5 alphabet (sigma = 6: flagged A == 7); DNA4 has sigma = 5
    gold_W = [0, 0, 1, 3, 1, 1, 2, 4, 4, 3, 4, 0, 1, 0, 1, 4, 1, 7, 2, 0, 4, 3, 3]
    gold_W = [w - 1 if w >= 6 else w for w in gold_W]
    assert list(W) == gold_W
    assert list(F) == [0, 3, 11, 13, 17]


def test_boss_small_graph_traversal():
    _, seqs = read_fasta(os.path.join(GOLD, "test_construct.fa"))
    g = O.OracleGraph(4, seqs)
    L = O.lib()
    outgoing_edges = [0, 3, 4, 14, 5, 7, 12, 18, 19, 15, 20, 0, 8, 0, 10, 21, 11, 11, 13, 0, 22, 16, 17]
    assert len(outgoing_edges) == g.num_edges + 1
    W, last, F, _ = g.arrays()
    dummy_edge = L.mgo_boss_select_last(g.h, 1)
    assert L.mgo_boss_pick_edge(g.h, dummy_edge, 0) == 1
    assert L.mgo_boss_fwd(g.h, 1) == dummy_edge
    for i in range(1, g.num_edges + 1):
        if W[i] != 0:
            e = L.mgo_boss_pick_edge(g.h, L.mgo_boss_succ_last(g.h, i), int(W[i]) % 5)
            assert outgoing_edges[i] == L.mgo_boss_fwd(g.h, e), i
            f = L.mgo_boss_fwd(g.h, i)
            assert last[f] == 1


@pytest.mark.parametrize("k", range(1, 10))
def test_boss_map_to_edges_closed_form(k):
    # BOSS(k) <=> DBG k+1; built by add_sequence in the reference, identical edge set here
    g = O.OracleGraph(k + 1, ["A" * 100 + "C" * 100], dynamic=True)
    expected = [0, 0, k + 2, k + 2, k + 2] + [k + 2 + i for i in range(1, k + 1)] + [k + 2 + k + 1] * k
    seq = "T" * 2 + "A" * (k + 3) + "C" * (2 * k)
    assert list(g.map_to_nodes(seq)) == expected


def test_mt_graph_node_counts():
    _, seqs = read_fasta(os.path.join(GOLD, "genome.MT.fa"))
    g = O.OracleGraph(11, seqs, mask=True)
    assert g.num_nodes == 16438          # integration_tests/test_align.py:37


# ------------------------------------------------------------- aligner ----
def make_cfg(c):
    cfg = struct_defaults()
    if "matrix" in c:
        cfg.score_matrix = dna_scoring_matrix(*c["matrix"])
    if "unit" in c:
        cfg.score_matrix = unit_scoring_matrix(c["unit"])
    ren = {"gap_open": "gap_opening_penalty", "gap_ext": "gap_extension_penalty"}
    for k, v in c.items():
        if k in ("matrix", "unit"):
            continue
        setattr(cfg, ren.get(k, k), v)
    return cfg


def score_cigar(cfg, ref, query, cigar_str):
    """aligner_config.cpp:68-126 restated independently in Python (query = full read)."""
    import re
    ops = [(int(n), op) for n, op in re.findall(r"(\d+)([SX=DIG])", cigar_str)]
    score = 0
    if ops[0][1] != "S":
        score += cfg.left_end_bonus
    if ops[-1][1] != "S":
        score += cfg.right_end_bonus
    qi = ri = 0
    for n, op in ops:
        if op == "S":
            qi += n
        elif op in "=X":
            for _ in range(n):
                score += cfg.score_matrix[ord(ref[ri])][ord(query[qi])]
                ri += 1; qi += 1
        elif op == "I":
            score += cfg.gap_opening_penalty + (n - 1) * cfg.gap_extension_penalty
            qi += n
        elif op == "D":
            score += cfg.gap_opening_penalty + (n - 1) * cfg.gap_extension_penalty
            ri += n
    assert ri == len(ref) and qi == len(query)
    return score


GOLDENS = json.load(open(os.path.join(GOLD, "test_aligner_goldens.json")))


@pytest.mark.parametrize("gold", GOLDENS, ids=[g["name"] for g in GOLDENS])
def test_aligner_golden(gold):
    g = O.OracleGraph(gold["k"], gold["refs"], mask=gold["mask"], dynamic=not gold["mask"])
    cfg = make_cfg(gold["cfg"])
    query = revcomp(gold["query"]) if gold["rc_query"] else gold["query"]
    variants = [cfg]
    if gold["extend"]:
        uni = make_cfg(gold["cfg"]); uni.max_seed_length = SIZE_MAX   # check_extend
        variants.append(uni)
    results = []
    for c in variants:
        line = g.align_tsv(c, [query], with_nodes=True)[0]
        _, q, alns = O.parse_tsv_line(line, with_nodes=True)
        results.append(alns)
        e = gold["expect"]
        assert len(alns) == e["n_paths"], line
        if not alns:
            continue
        a = alns[0]
        full_q = {"+": q, "-": revcomp(q)}[a["strand"]]
        # is_valid: score == score_cigar
        assert a["score"] == score_cigar(c, a["seq"], full_q, a["cigar"]), line
        if "alternatives" in e:
            alt = [x for x in e["alternatives"] if x["orientation"] == (a["strand"] == "-")][0]
            assert a["cigar"] == alt["cigar"] and a["seq"] == alt["sequence"], line
        if "cigar" in e:
            assert a["cigar"] in e["cigar"], line
        if "sequence" in e:
            assert a["seq"] in e["sequence"], line
        if "num_matches" in e:
            assert a["nm"] == e["num_matches"], line
        if "path_size" in e:
            assert len(a["nodes"]) == e["path_size"], line
        if "offset" in e:
            assert a["offset"] == e["offset"], line
        if "score" in e:
            assert a["score"] == e["score"], line
        import re
        if "clipping" in e:
            m = re.match(r"^(\d+)S", a["cigar"])
            assert (int(m.group(1)) if m else 0) == e["clipping"], line
        if "end_clipping" in e:
            m = re.search(r"(\d+)S$", a["cigar"])
            assert (int(m.group(1)) if m else 0) == e["end_clipping"], line
    if gold["extend"]:
        assert results[0] == results[1], (results[0], results[1])


MT_FWD = [  # integration_tests/test_align.py:49-57 (--align-only-forwards --align-min-exact-match 0.0)
    "MT-10/1\tAACAGAGAATAGTTTAAATTAGAATCTTAGCTTTGGGTGCTAATGGTGGAGTTAAAGACTTTTTCTCTGATTTGTCCTTGGAAAAAGGTTTTCATCTCCGGTTTACAAGACTGGTGTATTAGTTTATACTACAAGGACAGGCCCATTTGA\t+\tTAGAATCTTAG\t22\t11\t19S11=120S\t0",
    "MT-8/1\tAAAACTAACCCCCTAATAAAATTAATTAACCACTCATTCATCGACCTCCCCACCCCATCCAACATCTCCGCATGATGAAACTTCGGCTCACTCCTTGGCGCCTGCCTGATCCTCCAAATCACCACAGGACTATTCCTAGCCATGCACTAC\t+\tAAAACTAACCCCCTAATAAAATTAATTAACCACTCATTCATCGACCTCCCCACCCCATCCAACATCTCCGCATGATGAAACTTCGGCTCACTCCTTGGCGCCTGCCTGATCCTCCAAATCACCACAGGACTATTCCTAGCCATGCACTAC\t310\t150\t150=\t0",
    "MT-6/1\tATATGACTAGCTTACACAATAGCTTTTATAGTAAAGATACCTCTTTACGGACTCCACTTATGACTCCCTAAAGCCCATGTCGAAGCCCCCATCGCTGGGTCAATAGTACTTGCCGCAGTACTCTTAAAACTAGGCGGCTATGGTATAATA\t+\tATATGACTAGCTTACACAATAGCTTTTATAGTAAAGATACCTCTTTACGGACTCCACTTATGACTCCCTAAAGCCCATGTCGAAGCCCCCATCGCTGGGTCAATAGTACTTGCCGCAGTACTCTTAAAACTAGGCGGCTATGGTATAATA\t310\t150\t150=\t0",
    "MT-4/1\tAGTATAGTAGTTCGCTTTGACTGGTGAAGTCTTAGCATGTACTGCTCGGAGGTTCGGTTCTGCTCCGAGGTCGCCCCAACCGAAATTTTTAATGCAGGTTTGGTAGTTTAGGACCTGTGGGTTTGTTAGGTACTGTTTGCATTAATAAAT\t*\t*\t0\t*\t*\t*",
    "MT-2/1\tTGTGTTAATTAATTAATGCTTGTAGGACATAATAATAACAATTGAATGTCTGCACAGCCACTTTCCACACAGACATCATAACAAAAAATTTCCACCAAACCCCCCCTCCCCCGCTTCTGGCCACAGCACTTAAACACATCTCTGCCAAAC\t+\tTGTGTTAATTAATTAATGCTTGTAGGACATAATAATAACAATTGAATGTCTGCACAGCCACTTTCCACACAGACATCATAACAAAAAATTTCCACCAAACCCCCCCTCCCCCGCTTCTGGCCACAGCACTTAAACACATCTCTGCCAAAC\t310\t150\t150=\t0",
]
MT_BOTH = [  # integration_tests/test_align.py:198-206 (--align-min-exact-match 0.0)
    "MT-10/1\tAACAGAGAATAGTTTAAATTAGAATCTTAGCTTTGGGTGCTAATGGTGGAGTTAAAGACTTTTTCTCTGATTTGTCCTTGGAAAAAGGTTTTCATCTCCGGTTTACAAGACTGGTGTATTAGTTTATACTACAAGGACAGGCCCATTTGA\t-\tTCAAATGGGCCTGTCCTTGTAGTATAAACTAATACACCAGTCTTGTAAACCGGAGATGAAAACCTTTTTCCAAGGACAAATCAGAGAAAAAGTCTTTAACTCCACCATTAGCACCCAAAGCTAAGATTCTAATTTAAACTATTCTCTGTT\t310\t150\t150=\t0",
    MT_FWD[1], MT_FWD[2],
    "MT-4/1\tAGTATAGTAGTTCGCTTTGACTGGTGAAGTCTTAGCATGTACTGCTCGGAGGTTCGGTTCTGCTCCGAGGTCGCCCCAACCGAAATTTTTAATGCAGGTTTGGTAGTTTAGGACCTGTGGGTTTGTTAGGTACTGTTTGCATTAATAAAT\t-\tATTTATTAATGCAAACAGTACCTAACAAACCCACAGGTCCTAAACTACCAAACCTGCATTAAAAATTTCGGTTGGGGCGACCTCGGAGCAGAACCCAACCTCCGAGCAGTACATGCTAAGACTTCACCAGTCAAAGCGAACTACTATACT\t305\t149\t95=1X54=\t0",
    MT_FWD[4],
]


@pytest.mark.parametrize("both", [False, True])
def test_integration_mt_cli_defaults(both):
    _, seqs = read_fasta(os.path.join(GOLD, "genome.MT.fa"))
    names, reads = read_fastq(os.path.join(GOLD, "genome_MT1.fq"))
    assert len(reads) == 7
    g = O.OracleGraph(11, seqs, mask=False)      # `align` drops the dummy mask
    cfg = cli_defaults(11, min_exact_match=0.0, forward_and_reverse_complement=both)
    lines = g.align_tsv(cfg, reads, headers=names)
    gold = MT_BOTH if both else MT_FWD
    for got, exp in zip(lines[:5], gold):
        assert got == exp
    f = lines[5].split("\t")
    assert f[0] == "MT-11/1" and f[4] == "22"


MAP_COUNTS = ["1/140/1", "140/140/140", "140/140/140", "0/140/0", "140/140/140", "1/140/1", "1/140/1"]


def _map_counts(nodes):
    nodes = [int(x) for x in nodes]
    return "%d/%d/%d" % (sum(1 for x in nodes if x), len(nodes), len(set(x for x in nodes if x)))


def test_map_count_kmers_golden():
    """`metagraph align --map --count-kmers` on the masked MT graph (integration_tests/test_align.py:58-87,
    cli/align.cpp:108-165): discovered / total / unique k-mers per read pin map_to_nodes_sequentially."""
    _, seqs = read_fasta(os.path.join(GOLD, "genome.MT.fa"))
    _, reads = read_fastq(os.path.join(GOLD, "genome_MT1.fq"))
    g = O.OracleGraph(11, seqs, mask=True)
    assert [_map_counts(g.map_to_nodes(r)) for r in reads] == MAP_COUNTS


MAP_COUNTS_SUBK = ["3/141/3", "141/141/141", "141/141/141", "1/141/1", "141/141/141", "4/141/4", "3/141/3"]


def test_map_count_kmers_subk_golden():
    """`metagraph align --map --count-kmers --align-length 10` on the k = 11 MT graph (integration_tests/test_align.py:
    90-122, cli/align.cpp:113-131): per 10-mer the first node of call_nodes_with_suffix_matching_longest_prefix —
    pins BOSS::index_range + the enumeration of the nodes that end with the match."""
    _, seqs = read_fasta(os.path.join(GOLD, "genome.MT.fa"))
    _, reads = read_fastq(os.path.join(GOLD, "genome_MT1.fq"))
    for mask in (False, True):           # the CLI drops the mask before mapping (cli/align.cpp:336-339); same counts
        g = O.OracleGraph(11, seqs, mask=mask)
        got = [_map_counts([g.suffix_match_first(r[i:i + 10], 10) for i in range(len(r) - 9)]) for r in reads]
        assert got == MAP_COUNTS_SUBK


def test_c1_transcripts_1000_oracle():
    """BASELINE configs[0] fixture (tests/data/transcripts_1000.fa: 1000 records, 1 490 627 bp, 59 .. 11 666 bp): k = 12
    graph of the transcripts, CLI defaults. The reference stores no expected output for it; SURVEY 8c states what the
    algorithm must give -- every read aligns to itself, CIGAR {L}=, score 2L + 10 -- and the restatement does."""
    from metagraph_b200.config import cli_defaults
    names, seqs = read_fasta(os.path.join(GOLD, "transcripts_1000.fa"))
    assert len(seqs) == 1000 and sum(len(s) for s in seqs) == 1490627
    g = O.OracleGraph(12, seqs)
    out = g.align_tsv(cli_defaults(12), seqs, threads=8)
    for s_, l in zip(seqs, out):
        f = l.split("\t")
        assert f[2] == "+" and f[3] == s_ and int(f[4]) == 2 * len(s_) + 10 and f[6] == "%d=" % len(s_) and f[7] == "0"
