"""Summarise an `ncu --set full` report: per kernel name the mean of dram__bytes_read.sum + dram__bytes_write.sum per launch
and a few utilisation figures -> profiles/ncu_traffic.json (read by bench.py for roofline.traffic) and a markdown table.

    python tools/ncu_traffic.py gpurun_out/prof_step.ncu-rep profiles/r02_ncu_full.md"""
import csv
import io
import json
import os
import subprocess
import sys

rep, md = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sectors.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


per = {}
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("flatq::", "")
    d = per.setdefault(name, {"n": 0})
    d["n"] += 1
    for m in want:
        if m in ix and r[ix[m]] not in ("", "n/a"):
            v = to_bytes(r[ix[m]], units[ix[m]]) if "bytes" in m else float(r[ix[m]].replace(",", ""))
            d[m] = d.get(m, 0.0) + v
out = {}
lines = ["| kernel | launches | time us | DRAM read MB | DRAM write MB | L2 traffic MB (sectors x 32 B) | DRAM % | tensor % | issue % | warps % | regs |", "|---|---|---|---|---|---|---|---|---|---|---|"]
for name, d in sorted(per.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
    n = d["n"]
    g = lambda m: d.get(m, 0.0) / n  # noqa: E731
    t_us = g("gpu__time_duration.sum") * ({"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1))
    traffic = g("dram__bytes_read.sum") + g("dram__bytes_write.sum")
    base = name.split("<")[0]
    out.setdefault(base, {"bytes_per_launch": 0.0, "launches": 0, "source": os.path.basename(rep)})
    o = out[base]
    o["bytes_per_launch"] = (o["bytes_per_launch"] * o["launches"] + traffic * n) / (o["launches"] + n)
    o["launches"] += n
    lines.append(f"| `{name[:70]}` | {n} | {t_us:.1f} | {g('dram__bytes_read.sum') / 1e6:.2f} | {g('dram__bytes_write.sum') / 1e6:.2f} | "
                 f"{g('lts__t_sectors.sum') * 32 / 1e6:.1f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g('launch__registers_per_thread'):.0f} |")
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json"), "w"), indent=1)
open(md, "a").write("\n".join(lines) + "\n")
print("\n".join(lines))
