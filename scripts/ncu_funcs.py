"""Per-function view of an `ncu --page source --csv --print-source cuda,sass` export of k_align:
instruction share, stall-reason share and opcode mix, using the function boundaries of
metagraph_b200/csrc/align_core.cuh as it is in the working tree (profile the tree you analyse).

usage: python scripts/ncu_funcs.py export.csv [lo hi]   # lo..hi: also list lines of align_core.cuh in that range
"""
import bisect, collections, csv, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, -1)
src = open(os.path.join(ROOT, "metagraph_b200/csrc/align_core.cuh")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s+MGB_HD\s+[\w:<>\*& ]+?\s+\**&?(\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
STALLS = {"long_sb": 37, "short_sb": 45, "no_inst": 42, "wait": 48, "branch": 33, "math": 38,
          "not_sel": 43, "sel": 44, "lg": 36, "mio": 40}
cur = None
curline = None
byf = collections.defaultdict(lambda: collections.Counter())
ops = collections.Counter()
lines = {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0].isdigit():
        curline = (cur, int(r[0]), r[1].strip()[:100])
        continue
    if len(r) > 48 and r[2].startswith("0x") and curline:
        f, ln, text = curline
        if f == "align_core.cuh":
            k = bisect.bisect_right(starts, ln) - 1
            name = funcs[k][1] if k >= 0 else "(types)"
        else:
            name = f
        def g(c):
            try:
                return int(r[c])
            except ValueError:
                return 0
        ins = r[3].strip().split()
        if ins and ins[0].startswith("@"):
            ins = ins[1:]
        n = g(7)
        a = byf[name]
        a["ins"] += n
        for k2, c in STALLS.items():
            a[k2] += g(c)
        a["smp"] += g(4)
        if ins:
            op = ins[0]
            if op in ("IMAD.MOV.U32", "IMAD.MOV", "MOV"):
                a["mov"] += n
            ops[op.split(".")[0] if not op.startswith("IMAD.MOV") else "IMAD.MOV"] += n
        la = lines.setdefault((f, ln), [text, 0, 0, collections.Counter()])
        la[1] += n
        la[2] += g(4)
        for k2, c in STALLS.items():
            la[3][k2] += g(c)
tot_i = sum(a["ins"] for a in byf.values())
tot_s = sum(a["smp"] for a in byf.values())
print("warp instructions %d, samples %d" % (tot_i, tot_s))
tot_st = collections.Counter()
for a in byf.values():
    for k2 in STALLS:
        tot_st[k2] += a[k2]
print("stall mix: " + " ".join("%s=%.1f%%" % (k2, 100.0 * v / tot_s) for k2, v in tot_st.most_common()))
print("%-26s %6s %6s %6s | %s" % ("function", "ins%", "smp%", "mov%", " ".join("%8s" % k2 for k2 in STALLS)))
for n, a in sorted(byf.items(), key=lambda x: -x[1]["smp"])[:22]:
    print("%-26s %6.1f %6.1f %6.1f | %s" % (n, 100.0 * a["ins"] / tot_i, 100.0 * a["smp"] / tot_s, 100.0 * a["mov"] / tot_i,
                                             " ".join("%8.1f" % (100.0 * a[k2] / tot_s) for k2 in STALLS)))
print("opcode mix: " + " ".join("%s=%.1f%%" % (k2, 100.0 * v / tot_i) for k2, v in ops.most_common(18)))
if hi >= lo:
    print("--- align_core.cuh lines %d..%d with >= 0.15%% of instructions or samples" % (lo, hi))
    for (f, ln), (text, n, s, st) in sorted(lines.items()):
        if f == "align_core.cuh" and lo <= ln <= hi and (n >= 0.0015 * tot_i or s >= 0.0015 * tot_s):
            top = ",".join("%s %.1f" % (k2, 100.0 * v / tot_s) for k2, v in st.most_common(2) if v)
            print("%5.2f%% ins %5.2f%% smp  %d  %-90s [%s]" % (100.0 * n / tot_i, 100.0 * s / tot_s, ln, text, top))
