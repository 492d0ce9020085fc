"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, mean us, share of the total.
    python tools/ncu_nodes_summary.py [--last-step] profiles/r02_decode_nodes_b1.csv [more.csv ...]   -> markdown on stdout
--last-step: only the launches of the last decode step in the list (after the previous lm_head GEMM up to the last one),
in launch order -- the list of a whole process also holds the weight initialisation and the prefill."""
import collections
import csv
import re
import sys


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    out = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        if x["Metric Unit"] in ("ns", "nsecond"):
            v /= 1000.0
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", x["Kernel Name"]))
        out.append((re.sub(r"^flatq::", "", name)[:90], v, x.get("Grid Size", "")))
    return out


def last_step(path):
    ls = launches(path)
    heads = [i for i, (n, _, _) in enumerate(ls) if n.startswith("nvjet")]
    step = ls[heads[-2] + 1:heads[-1] + 1]
    total = sum(v for _, v, _ in step)
    out = [f"### `{path}` — last decode step: {len(step)} launches, {total:.0f} us summed (cold-cache, serialised per launch)", "",
           "| # | kernel | grid | us |", "|---|---|---|---|"]
    out += [f"| {i} | `{n}` | {g} | {v:.2f} |" for i, (n, v, g) in enumerate(step)]
    return "\n".join(out)


def table(path):
    agg = collections.OrderedDict()
    for name, v, grid in launches(path):
        x = {"Grid Size": grid}
        a = agg.setdefault(name, [0, 0.0, x.get("Grid Size", "")])
        a[0] += 1
        a[1] += v
    total = sum(t for _, t, _ in agg.values())
    out = [f"### `{path}` — {sum(c for c, _, _ in agg.values())} launches, {total:.0f} us (cold-cache, serialised: compare shares)", "",
           "| kernel | grid | launches | mean us | share |", "|---|---|---|---|---|"]
    for name, (c, t, grid) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {grid} | {c} | {t / c:.2f} | {100 * t / total:.1f} % |")
    return "\n".join(out)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    fn = last_step if "--last-step" in sys.argv else table
    print("\n\n".join(fn(p) for p in args))
