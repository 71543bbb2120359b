"""k_seed time vs suffix-range table length (cold k-mer lookups on the non-matching strand)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 20_000_000)); N = int(os.environ.get("N", 200_000))
genome = make_genome(G)
buf, off = make_reads(genome, N, 42)
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
for s in [int(x) for x in sys.argv[1:]]:
    index = DBGSuccinctIndex(boss, suffix_len=s)
    al = B200Aligner(index, cli_defaults(K, min_seed_length=K, max_seed_length=K))
    al.set_pipeline_pieces(1)
    best = None
    for i in range(3):
        res = al.align_batch_raw(buf, off); st = al.stats_of(res); al.free_raw(res)
        best = st if best is None or st["seed_kernel_ms"] < best["seed_kernel_ms"] else best
    print("sfx_len %d: seed_ms %.2f align_ms %.2f index %.2f GB" % (s, best["seed_kernel_ms"], best["align_kernel_ms"], index.device_bytes / 1e9), flush=True)
    index.close()
