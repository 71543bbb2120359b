"""Throughput of the graph modes on the bench shape (not run yet on a GPU: the modes were finished after the round's GPU
budget was spent). BASIC: the bench graph. PRIMARY: the same BOSS table with mode = 2 — in a random genome no k-mer
(k = 31) occurs together with its reverse complement, so the genome is its own primary contig set; the reads taken
from the reverse strand then map through reverse-complement node ids. CANONICAL: genome + reverse complement.
Prints per mode: index build time of the mode (k_rc_tables for PRIMARY), kernel times, reads/s, and parity of a
sample against the CPU restatement."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex, format_alignment
from metagraph_b200.config import cli_defaults
G = int(os.environ.get("G", 20_000_000)); N = int(os.environ.get("N", 200_000)); SAMPLE = int(os.environ.get("SAMPLE", 2000))
lib = sys.argv[1] if len(sys.argv) > 1 else None
genome = make_genome(G)
buf, off = make_reads(genome, N, 7)
comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
both = np.concatenate([genome, comp[genome][::-1]])
import oracle_lib as O
for name, mode in (("basic", 0), ("primary", 2), ("canonical", 1)):
    if mode == 1:
        boss = BOSSTable.from_sequences(K, None, packed=(both, np.array([0, G, 2 * G], dtype=np.uint64)), lib=lib)
    else:
        boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)), lib=lib)
    t0 = time.time(); index = DBGSuccinctIndex(boss, lib=lib, mode=mode); t_index = time.time() - t0
    for kw in (dict(min_seed_length=K, max_seed_length=K), dict()):       # exact seeder (bench config) / CLI defaults
        cfg = cli_defaults(K, **kw)
        al = B200Aligner(index, cfg)
        for it in range(2):
            t0 = time.time(); res = al.align_batch_raw(buf, off); dt = time.time() - t0
            st = al.stats_of(res); al.free_raw(res)
        print("%s %s: index %.2f s, seed_ms %.2f align_ms %.2f, %.0f reads/s end to end" % (
            name, "exact" if kw else "cli-default", t_index, st["seed_kernel_ms"], st["align_kernel_ms"], N / dt), flush=True)
        g = O.OracleGraph(K, arrays=(boss.W, boss.last, boss.F)); g.set_mode(mode)
        sample = [bytes(buf[off[i]:off[i + 1]]).decode() for i in range(SAMPLE)]
        exp = g.align_tsv(cfg, sample, threads=32)
        got = al.align_batch([("", r) for r in sample])
        same = sum(format_alignment("", r, cfg.min_path_score) == e for r, e in zip(got, exp))
        print("%s parity on the sample: %d / %d lines identical" % (name, same, SAMPLE), flush=True)
    index.close()
