"""Condense an ncu report (.ncu-rep) into the text summary committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"kernel: {name}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:75s} {r[i]:>18s} {units[i]}")
    stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
    for v, h in sorted(stalls, reverse=True)[:5]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:.2f} warps/issue-cycle")
    if "dram__bytes_read.sum" in hdr:
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        print(f"  traffic (dram read+write)                                                    {rd + wr:18.3f} {units[hdr.index('dram__bytes_read.sum')]}")
    print()
