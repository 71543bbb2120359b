"""Wall clock of mgb_align_batch at bench scale for different piece counts (MGB_TEST_PIECES)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 100_000_000)); N = int(os.environ.get("N", 1_000_000))
genome = make_genome(G)
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
index = DBGSuccinctIndex(boss)
b, o = make_reads(genome, N, 42)
bp = torch.empty(len(b), dtype=torch.uint8, pin_memory=True); bp.numpy()[:] = b
op = torch.empty(len(o), dtype=torch.int64, pin_memory=True); op.numpy()[:] = o.astype(np.int64)
buf, off = bp.numpy(), op.numpy().view(np.uint64)
al = B200Aligner(index, cli_defaults(K, min_seed_length=K, max_seed_length=K))
for pieces in [int(x) for x in sys.argv[1:]]:
    os.environ["MGB_TEST_PIECES"] = str(pieces)
    ts = []
    for i in range(5):
        t0 = time.perf_counter(); res = al.align_batch_raw(buf, off); t1 = time.perf_counter()
        al.free_raw(res); ts.append(1e3 * (t1 - t0))
    print("pieces %d: %s ms" % (pieces, " ".join("%.1f" % t for t in ts)), flush=True)
