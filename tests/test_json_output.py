"""`metagraph align --json` (cli/align.cpp:289-304, Alignment::to_json, alignment.cpp:704-963) through the
host mirror in metagraph_b200/aligner.py, against the reference's own goldens
(tests/data/genome_MT1.align.json, genome_MT1.align.edit.json; integration_tests/test_align.py:333-382).
The records carry the NODE IDS of every path, so this also pins the graph's node numbering (BOSS edge
indices of a fresh construction) against the graphs the reference built."""
import os
import subprocess

import pytest

from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex, format_alignment_json
from metagraph_b200.config import cli_defaults
from test_oracle_golden import GOLD, read_fasta, read_fastq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "build", "libmgb_emu.so")


def check(lib, gold_file, edit):
    _, seqs = read_fasta(os.path.join(GOLD, "genome.MT.fa"))
    names, reads = read_fastq(os.path.join(GOLD, "genome_MT1.fq"))
    idx = DBGSuccinctIndex(BOSSTable.from_sequences(11, seqs, lib=lib), lib=lib)
    cfg = cli_defaults(11, min_exact_match=0.0, edit_distance=edit)
    res = B200Aligner(idx, cfg).align_batch(list(zip(names, reads)))
    gold = [l.rstrip() for l in open(os.path.join(GOLD, gold_file)) if l.strip()]
    assert len(gold) == 5                       # the reference keeps the first five records
    for n, r, g in zip(names, res, gold):
        assert format_alignment_json(n, r, 11) == g
    idx.close()


@pytest.mark.parametrize("gold_file,edit", [("genome_MT1.align.json", False), ("genome_MT1.align.edit.json", True)])
def test_json_goldens_emu(gold_file, edit):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    check(EMU, gold_file, edit)


@pytest.mark.gpu
@pytest.mark.parametrize("gold_file,edit", [("genome_MT1.align.json", False), ("genome_MT1.align.edit.json", True)])
def test_json_goldens_gpu(gold_file, edit):
    check(None, gold_file, edit)
