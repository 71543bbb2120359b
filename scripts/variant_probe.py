import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 20_000_000)); N = int(os.environ.get("N", 100_000))
genome = make_genome(G)
buf, off = make_reads(genome, N, 42)
for lib in sys.argv[1:]:
    boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)), lib=lib)
    index = DBGSuccinctIndex(boss, lib=lib)
    al = B200Aligner(index, cli_defaults(K, min_seed_length=K, max_seed_length=K))
    best = None
    for i in range(3):
        res = al.align_batch_raw(buf, off); st = al.stats_of(res); al.free_raw(res)
        best = st if best is None or st["align_kernel_ms"] < best["align_kernel_ms"] else best
    print(os.path.basename(lib), "seed_ms %.2f align_ms %.2f" % (best["seed_kernel_ms"], best["align_kernel_ms"]), flush=True)
    index.close()
