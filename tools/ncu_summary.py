"""Selected raw-page metrics per launch out of an .ncu-rep (ncu --set full): the text files kept under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep "header line" > profiles/ncu_XXX.txt"""
import csv, io, subprocess, sys
rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print(header)
print(f"report: {rep} (scratch, not committed); numbers under ncu are cold-cache, serialised launches -- not bench values\n")
for r in rows[2:]:
    if len(r) < len(hdr): continue
    print(f"== {r[ix['Kernel Name']]} grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}")
    for m in WANT:
        if m in ix: print(f"   {m:84s} {r[ix[m]]:>16s} {units[ix[m]]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or (h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")):
            try: st.append((float(r[ix[h]].replace(",", "")), h.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", "")))
            except ValueError: pass
    st.sort(reverse=True)
    if st: print("   stalled warps per issue: " + " ".join(f"{n}:{v:.2f}" for v, n in st[:8]))
    print()
