"""Wall-clock split of one mgb_align_batch call at bench scale (MGB_DEBUG=1 prints the stages)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 100_000_000)); N = int(os.environ.get("N", 1_000_000))
genome = make_genome(G)
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
index = DBGSuccinctIndex(boss)
buf, off = make_reads(genome, N, 42)
al = B200Aligner(index, cli_defaults(K, min_seed_length=K, max_seed_length=K))
for i in range(4):
    if i == 3: os.environ["MGB_DEBUG"] = "1"
    t0 = time.perf_counter()
    res = al.align_batch_raw(buf, off)
    t1 = time.perf_counter()
    st = al.stats_of(res); al.free_raw(res)
    t2 = time.perf_counter()
    print("call %d: align_batch %.1f ms, free %.1f ms; seed %.1f align %.1f h2d %.1f d2h %.1f" % (
        i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), st["seed_kernel_ms"], st["align_kernel_ms"], st["h2d_ms"], st["d2h_ms"]), flush=True)
