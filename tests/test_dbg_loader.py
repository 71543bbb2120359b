"""`.dbg` loader against the two graphs the reference ships (examples/data/graphs/*.dbg, SMALL
state, copied to tests/golden/example_graphs/ with the FASTA files they were built from).

Two things are pinned at once: (1) mgb_dbg_load decodes the reference's on-disk containers, (2) the oracle's
and the product's BOSS construction reproduce, edge for edge, what `metagraph build` wrote — for DNA (k=20)
and for the protein alphabet (k=20)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import parity_common as P
from metagraph_b200 import _lib
from metagraph_b200.aligner import BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "build", "libmgb_emu.so")
EX = os.path.join(ROOT, "tests", "golden", "example_graphs")
CASES = [("test_DNA_graph.dbg", "test_DNA_sequences.fa", "test_DNA_query.fa", "dna", 0),
         ("test_Protein_graph.dbg", "test_Protein_sequences.fa", "test_Protein_query.fa", "protein", 1)]


@pytest.fixture(scope="module", autouse=True)
def _emu_built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)


def fasta(path):
    seqs = []
    for line in open(path):
        if line.startswith(">"):
            seqs.append("")
        else:
            seqs[-1] += line.strip()
    return seqs


@pytest.mark.parametrize("case", CASES, ids=[c[3] for c in CASES])
def test_loader_matches_fresh_construction(case):
    dbg, fa, _, alpha, code = case
    t = BOSSTable.from_dbg(os.path.join(EX, dbg), lib=EMU)
    assert (t.k, t.alphabet, t.mode, t.state) == (20, code, 0, 1)          # BASIC graph, SMALL state
    seqs = fasta(os.path.join(EX, fa))
    g = O.OracleGraph(t.k, seqs, alphabet=alpha)
    W, last, F, _ = g.arrays()
    assert (t.W == W).all() and (t.last[1:] == last[1:]).all() and (t.F == F).all()
    b = BOSSTable.from_sequences(t.k, seqs, lib=EMU, alphabet=code)
    assert (b.W == t.W).all() and (b.last[1:] == t.last[1:]).all() and (b.F == t.F).all()


@pytest.mark.parametrize("case", CASES, ids=[c[3] for c in CASES])
def test_align_against_loaded_graph(case):
    dbg, fa, qfa, alpha, code = case
    t = BOSSTable.from_dbg(os.path.join(EX, dbg), lib=EMU)
    idx = DBGSuccinctIndex(t, lib=EMU)
    reads = fasta(os.path.join(EX, qfa)) + [s[3:50] for s in fasta(os.path.join(EX, fa))]
    cfg = cli_defaults(t.k, alphabet=alpha)
    g = O.OracleGraph(t.k, arrays=(t.W, t.last, t.F), alphabet=alpha)
    exp = g.align_tsv(cfg, reads, with_nodes=True)
    got, _ = P.run_lines(idx, cfg, reads)
    assert got == exp
    assert all(l.split("\t")[6] == "%d=" % len(r) for l, r in zip(got, reads))     # every query is a path
    idx.close()


def test_loader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.dbg"
    p.write_bytes(open(os.path.join(EX, "test_DNA_graph.dbg"), "rb").read()[:500])
    with pytest.raises(_lib.MgbError):
        BOSSTable.from_dbg(str(p), lib=EMU)
    p.write_bytes(b"\x00" * 64)
    with pytest.raises(_lib.MgbError):
        BOSSTable.from_dbg(str(p), lib=EMU)
