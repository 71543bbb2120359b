"""SASS instructions attributed to given source lines of align_core.cuh in an `ncu --page source --csv
--print-source cuda,sass` export, with executed warp-instruction counts.
usage: python scripts/ncu_sass_of_line.py export.csv LINE [LINE ...]"""
import csv, sys
path = sys.argv[1]; want = set(int(x) for x in sys.argv[2:])
cur = None; curline = None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0].isdigit(): curline = (cur, int(r[0])); continue
    if len(r) > 8 and r[2].startswith("0x") and curline and curline[0] == "align_core.cuh" and curline[1] in want:
        print(curline[1], r[2], r[7].rjust(12), r[3].strip())
