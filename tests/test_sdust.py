"""Seed complexity filter (`is_low_complexity`, aligner_seeder_methods.cpp:21-29 -> sdust(seq, T=20, W=64)).
sdust is not vendored in the reference tree: PARITY UNPINNED against the library. What is pinned here:
the oracle's C++ restatement equals an independent brute-force statement of the symmetric-DUST definition,
and the reference's integration goldens (produced with the filter ON, cli default) are reproduced with it on."""
import os
import random

import oracle_lib as O
import parity_common as P
from metagraph_b200.config import cli_defaults
from test_oracle_golden import GOLD, read_fasta, read_fastq


def brute(seq, T=20, W=64):
    runs, cur = [], []
    for ch in seq.upper().replace("U", "T"):
        if ch in "ACGT":
            cur.append(ch)
        else:
            runs.append(cur); cur = []
    runs.append(cur)
    for run in runs:
        words = ["".join(run[i:i + 3]) for i in range(len(run) - 2)]
        for a in range(len(words)):
            for b in range(a + 1, min(len(words), a + W - 2)):
                w = words[a:b + 1]
                r = sum(w.count(x) * (w.count(x) - 1) // 2 for x in set(w))
                if r * 10 > T * (b - a):
                    return True
    return False


def test_definition_vs_restatement():
    rng = random.Random(5)
    assert O.is_low_complexity("A" * 31) and O.is_low_complexity("CA" * 15) and O.is_low_complexity("acg" * 12)
    assert not O.is_low_complexity("") and not O.is_low_complexity("ACGT") and not O.is_low_complexity("AAAAA")
    n_low = 0
    for t in range(1500):
        L = rng.choice([5, 12, 19, 25, 31, 40, 70, 100])
        alpha = rng.choice(["ACGT", "ACGT", "AC", "ACGTN", "AAAC", "acgtACGT"])
        s = "".join(rng.choice(alpha) for _ in range(L))
        if rng.random() < 0.3:
            p = rng.randrange(0, max(1, L - 8)); s = s[:p] + rng.choice("ACGT") * rng.randrange(4, 12) + s[p:]
        exp = brute(s)
        assert O.is_low_complexity(s) == exp, s
        n_low += exp
    assert 100 < n_low < 1400


def test_mt_goldens_with_filter_on():
    """test_align.py:49-57 / :198-206 were produced by the CLI with the filter on."""
    from test_oracle_golden import MT_FWD, MT_BOTH   # expected TSV lines
    _, seqs = read_fasta(os.path.join(GOLD, "genome.MT.fa"))
    names, reads = read_fastq(os.path.join(GOLD, "genome_MT1.fq"))
    g = O.OracleGraph(11, seqs)
    for both, exp in ((False, MT_FWD), (True, MT_BOTH)):
        cfg = cli_defaults(11, min_exact_match=0.0, forward_and_reverse_complement=both, seed_complexity_filter=True)
        got = g.align_tsv(cfg, reads, headers=names)
        for l, e in zip(got[:5], exp):
            assert l == e
        f = got[5].split("\t")
        assert f[0] == "MT-11/1" and f[4] == "22"
