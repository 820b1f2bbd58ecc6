"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by source line.
usage: python tools/ncu_lines.py file.csv [function-substring] [top]"""
import csv, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
fn = fp = None; hdr = None
per = collections.defaultdict(lambda: collections.Counter())   # fn -> (file,line,src) -> inst
stall = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if not r: continue
    if r[0] == "File Path": fp = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; ii = r.index("Instructions Executed"); si = r.index("# Samples"); continue
    if hdr is None or len(r) <= ii: continue
    if r[2] == "-" and r[0].isdigit():          # a source line summary row
        try: per[fn][(fp, int(r[0]), r[1].strip()[:110])] += int(r[ii]); stall[fn][(fp, int(r[0]), r[1].strip()[:110])] += int(r[si])
        except ValueError: pass
for f, c in per.items():
    if want not in f: continue
    tot = sum(c.values()); st = sum(stall[f].values())
    print(f"== {f}: {tot} warp instructions, {st} samples")
    for (file, line, src), n in c.most_common(top):
        print(f"{100*n/tot:5.1f}% inst {100*stall[f][(file,line,src)]/max(st,1):5.1f}% smp  {file}:{line}  {src}")
