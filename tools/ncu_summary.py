"""Reduce an .ncu-rep (ncu --set full) to the per-launch summary CSVs kept under profiles/.
usage: python tools/ncu_summary.py in.ncu-rep out.csv [kernel-substring]"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
stall = [c for c in h if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
cols = [c for c in WANT if c in h] + stall
idx = [h.index(c) for c in cols]
kn = h.index("Kernel Name")
flt = sys.argv[3] if len(sys.argv) > 3 else ""
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + cols)
    w.writerow([""] + [units[i] for i in idx])
    for r in rows[2:]:
        if len(r) == len(h) and flt in r[kn]:
            w.writerow([r[kn]] + [r[i] for i in idx])
            print(r[kn][:70], {c.split("__")[-1][:40]: r[i] for c, i in zip(cols[:13], idx[:13])})
