"""Where the C2 end-to-end step spends its wall clock: MGB_DEBUG split of one default (pipelined) call, then the
wall time per step for a few piece / lane counts. Same workload and call sequence as bench.py's e2e region."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import make_genome, make_reads, K
from metagraph_b200.aligner import B200Aligner, BOSSTable, DBGSuccinctIndex
from metagraph_b200.config import cli_defaults
from metagraph_b200._lib import mgb_alignment_t
G = int(os.environ.get("G", 100_000_000)); N = int(os.environ.get("N", 1_000_000))
genome = make_genome(G)
boss = BOSSTable.from_sequences(K, None, packed=(genome, np.array([0, G], dtype=np.uint64)))
index = DBGSuccinctIndex(boss)
buf_np, off_np = make_reads(genome, N, 42)
buf_pin = torch.empty(len(buf_np), dtype=torch.uint8, pin_memory=True); buf_pin.numpy()[:] = buf_np
off_pin = torch.empty(len(off_np), dtype=torch.int64, pin_memory=True); off_pin.numpy()[:] = off_np.astype(np.int64)
buf = buf_pin.numpy(); off = off_pin.numpy().view(np.uint64)
al = B200Aligner(index, cli_defaults(K, min_seed_length=K, max_seed_length=K, result_nodes=1))
L = al._L
L.mgb_set_host_threads(os.cpu_count())
aln_dtype = np.dtype({"names": ["score"], "formats": ["<i4"], "offsets": [mgb_alignment_t.score.offset],
                      "itemsize": ctypes.sizeof(mgb_alignment_t)})
def step():
    t0 = time.time()
    res = al.align_batch_raw(buf, off)
    t1 = time.time()
    n = int(L.mgb_results_num_alignments(res)); a = L.mgb_results_alignments(res)
    raw = (ctypes.c_char * (n * ctypes.sizeof(mgb_alignment_t))).from_address(ctypes.addressof(a.contents))
    chk = int(np.add.reduce(np.frombuffer(raw, dtype=aln_dtype, count=n)["score"], dtype=np.int64))
    t2 = time.time()
    al.free_raw(res)
    t3 = time.time()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, chk
def run(tag, pieces=None, lanes=None, threads=None, reps=4):
    for k, v in (("MGB_TEST_PIECES", pieces), ("MGB_TEST_LANES", lanes)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    if threads: L.mgb_set_host_threads(threads)
    for _ in range(2): step()
    t = [step() for _ in range(reps)]
    print("%-34s call %.1f  score read %.1f  free %.1f  -> %.1f ms/step" % (
        tag, np.mean([x[0] for x in t]), np.mean([x[1] for x in t]), np.mean([x[2] for x in t]),
        np.mean([sum(x[:3]) for x in t])), flush=True)
al.set_pipeline_pieces(0)
L.mgb_set_host_threads(32)
run("default (8 pieces, 2 lanes)")
for pieces, lanes in ((8, 3), (8, 4), (6, 3), (12, 3), (12, 4), (16, 4), (4, 2), (4, 4), (6, 2), (8, 2)):
    run("%d pieces, %d lanes" % (pieces, lanes), pieces, lanes, reps=6)
os.environ["MGB_DEBUG"] = "1"; sys.stderr.write("== split of one default call (8 pieces, 2 lanes)\n"); run("split", 8, 2, reps=1); os.environ.pop("MGB_DEBUG")
