"""Ad-hoc GPU probe: synthetic graph + reads, parity on a sample vs the oracle, timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from metagraph_b200.aligner import *
from metagraph_b200.config import *

G = int(os.environ.get("G", 10_000_000)); N = int(os.environ.get("N", 100_000)); RATE = float(os.environ.get("RATE", 0))
MODE = os.environ.get("MODE", "c2")
k = 31
rng = np.random.default_rng(32)
codes = rng.integers(0, 4, G, dtype=np.uint8)
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
t = time.time(); boss = BOSSTable.from_sequences(k, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
print("boss build %.1fs edges %d" % (time.time() - t, boss.num_edges), flush=True)
t = time.time(); index = DBGSuccinctIndex(boss); print("index %.1fs bytes %d" % (time.time() - t, index.device_bytes), flush=True)
rng = np.random.default_rng(42)
starts = rng.integers(0, G - 150, N)
comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
reads = np.empty((N, 150), np.uint8)
for i in range(N):
    r = genome[starts[i]:starts[i] + 150]
    reads[i] = comp[r][::-1] if (i & 1) else r
if RATE > 0:
    m = rng.random((N, 150)) < RATE
    reads[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
buf = np.ascontiguousarray(reads.reshape(-1)); offsets = (np.arange(N + 1, dtype=np.uint64) * 150)
cfg = cli_defaults(k, min_seed_length=31, max_seed_length=31) if MODE == "c2" else cli_defaults(k, min_exact_match=0.0)
al = B200Aligner(index, cfg)
for it in range(3):
    t = time.time(); res = al.align_batch_raw(buf, offsets); dt = time.time() - t
    st = al.stats_of(res); al.free_raw(res)
    print("iter %d: e2e %.3fs = %.0f reads/s | seed %.2f ms align %.2f ms h2d %.2f d2h %.2f | retried %d cols %d cells %d"
          % (it, dt, N / dt, st["seed_kernel_ms"], st["align_kernel_ms"], st["h2d_ms"], st["d2h_ms"],
             st["num_reads_retried"], st["dp_columns"], st["dp_cells"]), flush=True)
# parity on a sample
import oracle_lib as O
S = min(N, 2000)
o = O.OracleGraph(k, arrays=(boss.W, boss.last, boss.F))
rs = [bytes(reads[i]).decode() for i in range(S)]
exp = o.align_tsv(cfg, rs, with_nodes=True, threads=8)
got = [format_alignment("", r, cfg.min_path_score, with_nodes=True) for r in al.align_batch([("", r) for r in rs])]
bad = [i for i in range(S) if exp[i] != got[i]]
print("parity sample: %d/%d differ" % (len(bad), S))
for i in bad[:3]:
    print(" exp", exp[i][:300]); print(" got", got[i][:300])
