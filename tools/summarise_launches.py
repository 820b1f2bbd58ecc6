"""ncu launch list (csv written by `ncu --metrics ... --csv --log-file X`) -> per-kernel summary of ONE pre-pass frame.
usage: python tools/summarise_launches.py launches.csv out.json [first_kernel_substring]
The frame taken is the LAST complete run of consecutive pre-pass kernels (k_me<64> ... k_tq<4>) in the list."""
import csv, json, sys, collections, re, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_sha
src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if not r[ix['ID']].isdigit(): continue
    k = int(r[ix['ID']])
    e = agg.setdefault(k, {"kernel": re.sub(r"<unnamed>::|void |\(.*\)$", "", r[ix['Kernel Name']]), "grid": r[ix['Grid Size']]})
    e[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
ks = list(agg.values())
# last index of k_me<64>, then everything up to the next k_me<64> / non pre-pass kernel
# a frame starts with the plane kernel (k_subpel_planes) when the plan builds the planes per picture, else with k_me<64
first = "k_subpel_planes" if any(k["kernel"].startswith("k_subpel_planes") for k in ks) else "k_me<64"
starts = [i for i, k in enumerate(ks) if k["kernel"].startswith(first)]
pre = ("k_me<", "k_me_ctu", "k_mc", "k_tq", "k_subpel")
frames = []
for s in starts:
    e = s + 1
    while e < len(ks) and ks[e]["kernel"].startswith(pre) and not ks[e]["kernel"].startswith(first): e += 1
    frames.append((s, e))
s, e = max(frames, key=lambda f: (f[1] - f[0], f[0]))
out = []
for k in ks[s:e]:
    out.append({"kernel": k["kernel"], "grid": k["grid"], "time_us": k["gpu__time_duration.sum"] / 1e3, "warp_inst": k["smsp__inst_executed.sum"],
                "issue_active_pct": k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps_active_pct": k.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "dram_read_bytes": k.get("dram__bytes_read.sum"), "dram_write_bytes": k.get("dram__bytes_write.sum")})
tot_t = sum(k["time_us"] for k in out)
for k in out: k["share_of_frame_time"] = round(k["time_us"] / tot_t, 4)
fam = collections.Counter()
for k in out: fam[k["kernel"].split("<")[0]] += k["time_us"] / tot_t
json.dump({"kernel_source_sha": kernel_source_sha(), "workload": "1920x1080, one frame of the pre-pass (%d kernels), per-launch ncu metrics, --clock-control none (cold cache, serialised)" % len(out),
           "source": src.split("/")[-1], "total_warp_inst": sum(k["warp_inst"] for k in out), "total_time_us": tot_t,
           "family_share_of_time": {k: round(v, 4) for k, v in fam.items()}, "kernels": out}, open(dst, "w"), indent=1)
print(len(out), "kernels", round(tot_t, 1), "us", round(sum(k["warp_inst"] for k in out) / 1e6, 1), "M warp inst", dict(fam))
