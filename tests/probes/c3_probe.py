import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_genome, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 20_000_000)); N = int(os.environ.get("N", 50_000)); RATE = float(os.environ.get("RATE", 0.05))
lib = sys.argv[1] if len(sys.argv) > 1 else None
genome = make_genome(G)
rng = np.random.default_rng(42)
starts = rng.integers(0, G - 150, N)
reads = genome[starts[:, None] + np.arange(150)[None, :]]
comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
rc = rng.random(N) < 0.5
reads[rc] = comp[reads[rc]][:, ::-1]
m = rng.random((N, 150)) < RATE
reads[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
buf = np.ascontiguousarray(reads.reshape(-1)); off = np.arange(N + 1, dtype=np.uint64) * 150
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)), lib=lib)
index = DBGSuccinctIndex(boss, lib=lib)
al = B200Aligner(index, cli_defaults(K, min_exact_match=0.0))
al.set_pipeline_pieces(1)
for i in range(2):
    res = al.align_batch_raw(buf, off); st = al.stats_of(res); al.free_raw(res)
    print("c3 probe: seed_ms %.2f align_ms %.2f seeds %d ext %d cols %d" % (st["seed_kernel_ms"], st["align_kernel_ms"], st["num_seeds"], st["num_extensions"], st["dp_columns"]), flush=True)
# CPU restatement on a bounded sample of the same reads (test infrastructure, timing only)
if os.environ.get("C3_CPU", "1") == "1":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    g = O.OracleGraph(K, arrays=(boss.W, boss.last, boss.F))
    ns = int(os.environ.get("C3_CPU_READS", 4000))
    sample = ["".join(map(chr, reads[i])) for i in range(ns)]
    cfg = cli_defaults(K, min_exact_match=0.0)
    for th in (32, 64):
        t0 = time.time(); exp = g.align_tsv(cfg, sample, threads=th); dt = time.time() - t0
        print("c3 cpu restatement: %d reads, %d threads, %.2f s -> %.0f reads/s" % (ns, th, dt, ns / dt), flush=True)
    got = al.align_batch([("", r) for r in sample])
    from metagraph_b200.aligner import format_alignment
    same = sum(format_alignment("", r, 0) == e for r, e in zip(got, exp))
    print("c3 parity on the sample: %d / %d lines identical" % (same, ns), flush=True)
