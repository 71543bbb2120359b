"""Long randomized parity sweep on the GPU (seeds given on the command line), see parity_common.fuzz_case."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as P
lo, hi = int(sys.argv[1]), int(sys.argv[2])
lib = sys.argv[3] if len(sys.argv) > 3 else None
t0 = time.time(); nbad = 0; n = 0
for seed in range(lo, hi):
    if time.time() - t0 > float(os.environ.get("FUZZ_SECONDS", 150)):
        break
    bad, info = P.fuzz_case(lib, seed)
    n += 1
    if bad:
        nbad += 1
        print("MISMATCH seed", seed, {a: b for a, b in info[1].items() if a != "score_matrix"}, "k", info[0], flush=True)
        print(" read", info[3]); print(" exp", info[4][:300]); print(" got", info[5][:300], flush=True)
print("fuzz sweep: %d cases from seed %d, %d mismatching" % (n, lo, nbad))
