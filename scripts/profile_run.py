"""Small fixed workload for ncu captures (same code path as bench.py, smaller sizes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 20_000_000)); N = int(os.environ.get("N", 100_000)); STEPS = int(os.environ.get("STEPS", 2))
genome = make_genome(G)
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
index = DBGSuccinctIndex(boss)
buf, off = make_reads(genome, N, 42)
al = B200Aligner(index, cli_defaults(K, min_seed_length=K, max_seed_length=K))
al.set_pipeline_pieces(1)
for i in range(STEPS):
    res = al.align_batch_raw(buf, off); st = al.stats_of(res); al.free_raw(res)
    print("step", i, "seed_ms %.2f align_ms %.2f" % (st["seed_kernel_ms"], st["align_kernel_ms"]), flush=True)
