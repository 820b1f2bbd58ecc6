"""Attribute executed warp instructions of k_me<N> to code regions (by SASS address; inlined helpers take the region of the
nearest preceding instruction that maps to hb_kernels_me.cu).  usage: python tools/ncu_regions.py cuda_sass.csv "k_me<(int)8>" """
import csv, sys, collections, bisect
path, want = sys.argv[1], sys.argv[2]
REG = [(0, 85, "setup/load cur"), (86, 143, "half-pel strips"), (144, 161, "walk replay/control"), (162, 212, "setup/load cur"), (213, 233, "mv_cost"), (234, 261, "exchange"), (262, 302, "round4 (SAD)"),
       (303, 420, "walk replay/control"), (421, 457, "patch staging"), (458, 488, "H planes"), (489, 500, "cur->smem"), (501, 536, "quarter subpel4"),
       (537, 585, "half-pel strips"), (586, 607, "subpel decide"), (608, 633, "pred write"), (634, 700, "result")]
def region(line):
    for lo, hi, n in REG:
        if lo <= line <= hi: return n
    return "?"
rows = list(csv.reader(open(path)))
fn = fp = None; hdr = None; cur_line = None
ins = []   # (addr, file, line, count, sass, samples)
for r in rows:
    if not r: continue
    if r[0] == "File Path": fp = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; ii = r.index("Instructions Executed"); si = r.index("# Samples"); continue
    if hdr is None or want not in (fn or "") or len(r) <= ii: continue
    if r[2] == "-" and r[0].isdigit(): cur_line = int(r[0]); continue
    if r[2].startswith("0x"):
        try: ins.append((int(r[2], 16), fp, cur_line, int(r[ii]), r[3].strip(), int(r[si])))
        except ValueError: pass
ins.sort()
tot = collections.Counter(); smp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
last = "setup/load cur"
for addr, f, line, n, sass, s in ins:
    if f == "hb_kernels_me.cu": last = region(line)
    tot[last] += n; smp[last] += s
    ops[last][sass.split()[0] if not sass.startswith("@") else sass.split()[1]] += n
T = sum(tot.values()); S = sum(smp.values())
print(want, T, "warp instructions")
for k, v in tot.most_common():
    print(f"{100*v/T:5.1f}% inst {100*smp[k]/max(S,1):5.1f}% smp  {k:22s} " + " ".join(f"{o}:{100*c/v:.0f}" for o, c in ops[k].most_common(8)))
