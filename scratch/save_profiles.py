"""Condenses gpurun_out ncu outputs into the tracked profiles/ directory."""
import csv, io, json, subprocess, sys
from collections import defaultdict
tag = sys.argv[1]          # e.g. r1_final
launches = sys.argv[2]     # launches csv
rep = sys.argv[3]          # ncu-rep
L = sys.argv[4] if len(sys.argv) > 4 else "8"
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    try: d[r[ki]].append(float(r[vi].replace(",", "")))
    except Exception: pass
summary = {k[:90]: {"launches": len(v), "mean_us": sum(v) / len(v) / 1000.0} for k, v in d.items()}
tot = sum(x["launches"] * x["mean_us"] for k, x in summary.items() if "mgb::" in k and "flush" not in k)
for k, x in summary.items():
    if "mgb::" in k and "flush" not in k:
        x["share_of_step"] = x["launches"] * x["mean_us"] / tot
json.dump({"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 120 python bench.py --steps 5 --warmup 3 --cpu-reps 0",
           "note": "per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes", "kernels": summary},
          open(f"profiles/{tag}_launches.json", "w"), indent=1)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); h = rr[0]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
out, traffic, seen = [], {}, set()
for r in rr[2:]:
    if len(r) < len(h): continue
    name = r[h.index("Kernel Name")]
    rec = {"kernel": name[:90]}
    for k in keep:
        if k in h: rec[k] = r[h.index(k)] + " " + rr[1][h.index(k)]
    out.append(rec)
    short = "element_kernel" if "element_kernel" in name else ("gather_kernel" if "gather_kernel" in name else ("patch_kernel" if "patch_kernel" in name else name[:30]))
    if short not in seen:
        seen.add(short)
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd = float(r[h.index("dram__bytes_read.sum")]) * scale[rr[1][h.index("dram__bytes_read.sum")]]
        wr = float(r[h.index("dram__bytes_write.sum")]) * scale[rr[1][h.index("dram__bytes_write.sum")]]
        traffic[short] = int(rd + wr)
json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:element_kernel|gather_kernel -s 6 -c 2 python bench.py --steps 5 --warmup 3 --cpu-reps 0",
           "launches": out}, open(f"profiles/{tag}_ncu_full_summary.json", "w"), indent=1)
try: tj = json.load(open("profiles/r1_traffic.json"))
except Exception: tj = {}
tj[f"L{L}"] = traffic
tj["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture named in *_ncu_full_summary.json"
json.dump(tj, open("profiles/r1_traffic.json", "w"), indent=1)
print(json.dumps(summary, indent=1)[:1500]); print(traffic)
