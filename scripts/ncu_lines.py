"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line:
stall samples and executed instructions; prints the top lines."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = csv.reader(open(path))
cur_file = None; hdr = None
agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] and r[0].isdigit():
        try:
            samples = int(r[4]) if r[4] not in ("-", "") else 0
            inst = int(r[7]) if r[7] not in ("-", "") else 0
        except ValueError:
            continue
        agg.append((samples, inst, cur_file, int(r[0]), r[1].strip()[:110]))
tot_s = sum(a[0] for a in agg); tot_i = sum(a[1] for a in agg)
print("total samples", tot_s, "total warp instr", tot_i)
print("--- by stall samples")
for s, i, f, ln, src in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), f, ln, src))
print("--- by instructions")
for s, i, f, ln, src in sorted(agg, key=lambda a: -a[1])[:top // 2]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), f, ln, src))
