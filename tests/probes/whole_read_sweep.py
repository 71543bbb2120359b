"""One-off randomized sweep around the whole-read form of the exact-path shortcut (parity_common.whole_read_case with
random k, read length, graph size and configuration): kernels == oracle, shortcut on == off.
usage: python tests/probes/whole_read_sweep.py FIRST_SEED LAST_SEED [LIB]   (default LIB: the host-emulation build)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_common as P
from metagraph_b200 import _lib
from metagraph_b200.config import cli_defaults, dna_scoring_matrix
lo, hi = int(sys.argv[1]), int(sys.argv[2])
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "tests", "emu", "build", "libmgb_emu.so")
rng = np.random.default_rng(lo)
hits = n = 0
for seed in range(lo, hi):
    k = int(rng.integers(5, 32))
    kw = dict(min_seed_length=k, max_seed_length=k)
    r = rng.random()
    if r < 0.2: kw.update(rel_score_cutoff=float(rng.choice([0.0, 0.5, 1.0])))
    elif r < 0.35: kw.update(min_path_score=int(rng.integers(0, 200)))
    elif r < 0.5: kw.update(left_end_bonus=int(rng.integers(0, 8)), right_end_bonus=int(rng.integers(0, 8)))
    elif r < 0.6: kw.update(min_exact_match=float(rng.choice([0.0, 0.5, 1.0])))
    elif r < 0.7: kw.update(xdrop=int(rng.integers(5, 40)))
    elif r < 0.8: kw.update(num_alternative_paths=int(rng.integers(1, 4)))
    elif r < 0.9: kw.update(score_matrix=dna_scoring_matrix(int(rng.integers(1, 4)), -int(rng.integers(1, 5)), -int(rng.integers(1, 5))))
    L = int(rng.integers(max(k + 6, 36), 120))
    try:
        h = P.whole_read_case(lib, seed, k, cli_defaults(k, **kw), G=int(rng.integers(2300, 3500)), L=L)
    except _lib.MgbError as e:            # a configuration the aligner's constructor rejects
        print("configuration rejected:", seed, e)
        continue
    hits += h or 0; n += 1
print("whole-read sweep: %d cases from seed %d, no mismatch; %d reads took the whole-read exit" % (n, lo, hits))
