"""Aggregate an ncu launch-list CSV (gpu__time_duration + dram bytes per launch) per kernel.
usage: python tools/launch_list_summary.py launches.csv shares.json conv_traffic.json [first_id last_id [window.csv]]
With an ID window only the launches first_id..last_id (one step of the bench) are aggregated, and the rows of the window
are optionally written to window.csv (the file kept under profiles/)."""
import csv, json, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ci = {c: i for i, c in enumerate(h)}
agg = collections.defaultdict(lambda: {"launches": set(), "time_ns": 0.0, "rd": 0.0, "wr": 0.0})
TU = {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}; BU = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 1 << 60)
kept = [h]
for r in rows[hdr + 1:]:
    if len(r) < len(h): continue
    if not (lo <= int(r[ci['ID']]) <= hi): continue
    kept.append(r)
    name = re.sub(r'\(.*', '', r[ci['Kernel Name']]).replace('void ', '').replace('cai::', '')
    name = re.sub(r'<\((int|bool)\)(\d)>', r'<\2>', name)
    m = r[ci['Metric Name']]; v = float(r[ci['Metric Value']].replace(',', '')); u = r[ci['Metric Unit']]
    a = agg[name]; a["launches"].add(r[ci['ID']])
    if m == 'gpu__time_duration.sum': a["time_ns"] += v * TU.get(u, 1)
    elif m == 'dram__bytes_read.sum': a["rd"] += v * BU.get(u, 1)
    elif m == 'dram__bytes_write.sum': a["wr"] += v * BU.get(u, 1)
if len(sys.argv) > 6:
    csv.writer(open(sys.argv[6], 'w', newline='')).writerows(kept)
tot = sum(a["time_ns"] for a in agg.values())
out = [{"kernel": k, "launches": len(a["launches"]), "time_ms": a["time_ns"] / 1e6, "share": a["time_ns"] / tot,
        "avg_ms": a["time_ns"] / 1e6 / len(a["launches"]), "dram_read_bytes": a["rd"], "dram_write_bytes": a["wr"]}
       for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_ns"])]
for o in out[:14]:
    print(f'{o["kernel"][:44]:44s} n={o["launches"]:4d} {o["time_ms"]:9.2f} ms {100*o["share"]:5.1f}% avg {o["avg_ms"]:.3f} ms  rd {o["dram_read_bytes"]/1e9:.2f} GB wr {o["dram_write_bytes"]/1e9:.2f} GB')
json.dump(out, open(sys.argv[2], 'w'), indent=1)
conv = [o for o in out if o["kernel"].startswith(("conv_gemm_kernel", "conv_tma_kernel"))]
n = sum(o["launches"] for o in conv); b = sum(o["dram_read_bytes"] + o["dram_write_bytes"] for o in conv)
json.dump({"dram_bytes_per_launch": b / n, "launches": n, "conv_time_ms": sum(o["time_ms"] for o in conv),
           "source": (sys.argv[6] if len(sys.argv) > 6 else sys.argv[1]) + " (ncu dram__bytes_read.sum + dram__bytes_write.sum over all conv_gemm_kernel / conv_tma_kernel launches of one step)"},
          open(sys.argv[3], 'w'), indent=1)
print("conv launches", n, "bytes/launch", b / n, "conv total ms", sum(o["time_ms"] for o in conv))
