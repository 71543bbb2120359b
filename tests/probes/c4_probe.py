"""BASELINE configs[3] shape: protein alphabet (sigma = 27), k = 10, synthetic 100 aa reads with 5 % substitutions
vs a graph of a uniform random protein sequence; alphabet-generic kernels. Prints device times and the CPU
restatement's rate on a bounded sample (parity of the sample is checked too)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex, format_alignment
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 50_000_000)); N = int(os.environ.get("N", 200_000)); K = 10; L = 100
AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
rng = np.random.default_rng(32)
genome = AA[rng.integers(0, 20, G, dtype=np.uint8)]
t0 = time.time()
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)), alphabet=1)
index = DBGSuccinctIndex(boss)
print("graph: %d edges, index %.2f GB, build %.1f s" % (boss.num_edges, index.device_bytes / 1e9, time.time() - t0), flush=True)
rng = np.random.default_rng(42)
starts = rng.integers(0, G - L, N)
reads = genome[starts[:, None] + np.arange(L)[None, :]]
m = rng.random((N, L)) < 0.05
reads[m] = AA[rng.integers(0, 20, int(m.sum()))]
buf = np.ascontiguousarray(reads.reshape(-1)); off = np.arange(N + 1, dtype=np.uint64) * L
cfg = cli_defaults(K, alphabet="protein", min_exact_match=0.0)
al = B200Aligner(index, cfg)
al.set_pipeline_pieces(1)
for i in range(2):
    res = al.align_batch_raw(buf, off); st = al.stats_of(res); al.free_raw(res)
    dev = st["seed_kernel_ms"] + st["align_kernel_ms"]
    print("c4 probe: seed_ms %.2f align_ms %.2f -> %.0f reads/s (device); seeds %d ext %d cols %d" % (
        st["seed_kernel_ms"], st["align_kernel_ms"], N / dev * 1e3, st["num_seeds"], st["num_extensions"], st["dp_columns"]), flush=True)
if os.environ.get("C4_CPU", "1") == "1":
    import oracle_lib as O
    g = O.OracleGraph(K, arrays=(boss.W, boss.last, boss.F), alphabet="protein")
    ns = int(os.environ.get("C4_CPU_READS", 4000))
    sample = [bytes(reads[i]).decode() for i in range(ns)]
    t0 = time.time(); exp = g.align_tsv(cfg, sample, threads=32); dt = time.time() - t0
    print("c4 cpu restatement: %d reads, 32 threads, %.2f s -> %.0f reads/s" % (ns, dt, ns / dt), flush=True)
    got = al.align_batch([("", r) for r in sample])
    same = sum(format_alignment("", r, 0) == e for r, e in zip(got, exp))
    print("c4 parity on the sample: %d / %d lines identical" % (same, ns), flush=True)
