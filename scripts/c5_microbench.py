"""BASELINE config[4]: DP-extension microbenchmark. GCUPS (DP cells written / k_align device time) on a
linear synthetic graph, sweeping the band width (set through xdrop: the band of a column holds the
cells within xdrop of the best score) and the number of seeds (= reads, one seed extension each)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults

G = int(os.environ.get("G", 20_000_000))
genome = make_genome(G)
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
index = DBGSuccinctIndex(boss)
rows = []
counts = [int(x) for x in os.environ.get("COUNTS", "1000,10000,100000,1000000").split(",")]
for band_target in (8, 16, 32, 64):
    xdrop = band_target            # ge = -2: a column keeps ~xdrop/2 cells on each side of the diagonal
    # error-free reads: the exact-path shortcut would skip the DP this benchmark measures
    cfg = cli_defaults(K, min_seed_length=K, max_seed_length=K, xdrop=xdrop, no_exact_path_shortcut=True)
    al = B200Aligner(index, cfg)
    for n in counts:
        buf, off = make_reads(genome, n, 42)
        best = None
        for it in range(3):
            res = al.align_batch_raw(buf, off); st = al.stats_of(res); al.free_raw(res)
            if best is None or st["align_kernel_ms"] < best["align_kernel_ms"]:
                best = st
        band = best["dp_cells"] / max(best["dp_columns"], 1)
        gcups = best["dp_cells"] / (best["align_kernel_ms"] * 1e-3) / 1e9
        rows.append({"xdrop": xdrop, "mean_band": round(band, 1), "seeds": n, "align_ms": round(best["align_kernel_ms"], 3),
                     "gcups": round(gcups, 3), "reads_per_s": round(n / (best["align_kernel_ms"] * 1e-3))})
        print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "c5_microbench.json"), "w"), indent=1)
