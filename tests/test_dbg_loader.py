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


MUTATE = r'''
import os, sys
sys.path.insert(0, ROOT)
import numpy as np
from metagraph_b200 import _lib
from metagraph_b200.aligner import BOSSTable
src = open(SRC, "rb").read()
rng = np.random.default_rng(11)
ok = bad = 0
def attempt(buf):
    global ok, bad
    open(TMP, "wb").write(bytes(buf))
    try:
        BOSSTable.from_dbg(TMP, lib=EMU); ok += 1
    except _lib.MgbError:
        bad += 1
# every 8-byte field in turn replaced by all-ones / a huge length / a small wrong value
for off in range(0, len(src) - 8, 8 if len(src) < 3500 else 16):
    for patt in (b"\xff" * 8, b"\x00" * 7 + b"\x80", b"\x01" + b"\x00" * 7):
        b = bytearray(src); b[off:off + 8] = patt; attempt(b)
# random single-byte damage (also hits the width bytes and the 16-bit tree links)
for _ in range(600):
    b = bytearray(src); b[int(rng.integers(0, len(src)))] = int(rng.integers(0, 256)); attempt(b)
# truncation at every length up to the first KB, then sparsely
for n in list(range(0, 1024, 7)) + list(range(1024, len(src), 131)):
    attempt(src[:n])
print("MUTATIONS_DONE", ok, bad)
'''


@pytest.mark.parametrize("case", CASES, ids=[c[3] for c in CASES])
def test_loader_survives_damaged_files(case, tmp_path):
    """a malformed file must end in MgbError (or load, if the damage hit a part the loader skips) —
    never in a crash, a hang or an out-of-memory kill. Run in a child so that a crash fails this test only."""
    import sys
    script = tmp_path / "mutate.py"
    script.write_text("ROOT = %r\nEMU = %r\nSRC = %r\nTMP = %r\n" % (ROOT, EMU, os.path.join(EX, case[0]),
                                                                   str(tmp_path / "m.dbg")) + MUTATE)
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-3000:]
    assert "MUTATIONS_DONE" in out.stdout
    ok, bad = map(int, out.stdout.split("MUTATIONS_DONE")[1].split()[:2])
    assert bad > 300          # most damage is detected; the rest hit rank/select payloads the loader skips


def test_boss_build_reports_why():
    with pytest.raises(_lib.MgbError, match="256-bit"):
        BOSSTable.from_sequences(90, ["ACGT" * 50], lib=EMU)
    with pytest.raises(_lib.MgbError, match="at least 2"):
        BOSSTable.from_sequences(1, ["ACGT" * 50], lib=EMU)
