"""PRIMARY-mode graphs (`metagraph build --mode primary`), SURVEY 8f-2: the ORACLE restates the CanonicalDBG wrapper
(graph/representation/canonical_dbg.cpp, graph_extensions/node_first_cache.cpp:122-176), the sub-k seeding of the
reverse complement (aligner_seeder_methods.cpp:95-139, 251-314) and the CanonicalDBG branches of
Alignment::reverse_complement (alignment.cpp:583-640) and is pinned here on the reference's integration goldens
(integration_tests/test_align.py:270-330). The kernels are checked against it in tests/test_emu_parity.py,
tests/test_fuzz_parity.py and tests/test_primary_gpu.py."""
import os

import oracle_lib as O
from metagraph_b200.config import cli_defaults
from parity_common import primary_contigs
from test_oracle_canonical import CANONICAL, CANONICAL_SUBK
from test_oracle_golden import GOLD, read_fasta, read_fastq


def _graph():
    _, seqs = read_fasta(os.path.join(GOLD, "genome.MT.fa"))
    contigs = primary_contigs(seqs, 11)
    g = O.OracleGraph(11, contigs, mask=True)
    assert g.num_nodes == 16391                       # test_align.py:281
    g = O.OracleGraph(11, contigs)                    # `align` drops the dummy mask
    g.set_mode(2)
    return g


def test_primary_mode_cli_defaults():
    # test_align.py:270-300: the first five rows are those of the canonical-mode test
    names, reads = read_fastq(os.path.join(GOLD, "genome_MT1.fq"))
    got = _graph().align_tsv(cli_defaults(11, min_exact_match=0.0), reads, headers=names)
    assert len(got) == 7
    for i, exp in CANONICAL:
        assert got[int(i)] == exp.encode().decode("unicode_escape")
    assert got[6].split("\t")[4] == "310"
    f = got[5].split("\t")
    assert f[0] == "MT-11/1" and f[1] == reads[5] and f[4] == "22"


def test_primary_mode_subk():
    # test_align.py:302-330
    names, reads = read_fastq(os.path.join(GOLD, "genome_MT1.fq"))
    got = _graph().align_tsv(cli_defaults(11, min_exact_match=0.0, min_seed_length=10), reads, headers=names)
    assert len(got) == 7
    for i, exp in CANONICAL_SUBK:
        assert got[int(i)] == exp.encode().decode("unicode_escape")


def test_mode_known_answers():
    """tests/graph/test_aligner.cpp: align_suffix_seed_snp_canonical (:1483-1539, PRIMARY behind CanonicalDBG and
    CANONICAL), align_both_directions (:1541-1577), align_low_similarity4_rep_primary (:1600-1631),
    align_suffix_seed_no_full_seeds (:1773-1800), and the BASIC-mode tests that need the transcripts fixture or the
    seed complexity filter: align_straight_forward_and_reverse_complement_batch (:338-381), align_low_similarity3
    (:1345-1363: with the filter on the read must stay unaligned, the one known answer of the reference that depends
    on sdust), align_low_similarity4 (:1365-1424). The oracle satisfies the reference's expectations."""
    from parity_common import check_mode_kats
    check_mode_kats(oracle_only=True)
