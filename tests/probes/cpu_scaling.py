import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import BOSSTable
from metagraph_b200.config import cli_defaults
import oracle_lib as O
G = int(os.environ.get("G", 100_000_000))
genome = make_genome(G)
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
g = O.OracleGraph(K, arrays=(boss.W, boss.last, boss.F))
buf, off = make_reads(genome, 200000, 42)
reads = [bytes(buf[int(off[i]):int(off[i+1])]) for i in range(200000)]
cfg = cli_defaults(K, min_seed_length=K, max_seed_length=K)
for th in (1, 8, 32, 64, 128):
    n = min(len(reads), max(2000, 1500 * th))
    t = time.time(); g.align_tsv(cfg, reads[:n], threads=th); dt = time.time() - t
    print("threads %3d: %7d reads in %.2fs = %.0f reads/s (%.0f per thread)" % (th, n, dt, n / dt, n / dt / th), flush=True)
