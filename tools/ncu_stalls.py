"""Summarise the SASS page of an ncu report exported with
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv
(one section per captured launch): stall reasons over the whole kernel, then the instructions with the most samples.
    python tools/ncu_stalls.py src.csv [section index, default last] [top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(starts) - 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
s0 = starts[which]
s1 = starts[which + 1] if which + 1 < len(starts) else len(rows)
hdr = rows[s0 + 1]
ix = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[s0 + 2:s1] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(f"section {which} of {len(starts)}:", rows[s0][1][:110], "| instructions", len(data), "| samples", tot)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {s}: {v} ({100 * v / max(tot, 1):.1f}%)")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    st = {s: int(r[ix[s]] or 0) for s in stalls}
    main = max(st, key=st.get)
    print(f"{int(r[ix['# Samples']]):6d}  #{i:5d}  {r[ix['Source']].strip()[:100]:100s} {main}")
