#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per kernel, the Blackwell SASS mnemonics the profiling recipe
asks for (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, HMMA = legacy mma.sync, IDP4A = dp4a)
and registers / shared memory from `cuobjdump --dump-resource-usage`.

    python tools/sass_report.py > profiles/r01_sass_static.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flatquant_b200", "lib", "libflatq_b200.so")
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA",
         "IMMA", "IDP.4A", "LDSM", "LDGSTS", "ACQBULK", "UCGABAR"]


def run(*cmd):
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def demangle(names):
    out = run("cu++filt", *names).splitlines() if names else []
    return dict(zip(names, out))


def main():
    sass = run("cuobjdump", "-sass", LIB)
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[cur][w] += 1
    res = {}
    usage = run("cuobjdump", "--dump-resource-usage", LIB)
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", usage):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
    names = demangle(order)

    def short(n):
        d = names.get(n, n)
        d = re.sub(r"^void ", "", d)
        if d.endswith(")"):  # drop the trailing argument list (template arguments may contain parentheses themselves)
            depth = 0
            for i in range(len(d) - 1, -1, -1):
                depth += (d[i] == ")") - (d[i] == "(")
                if depth == 0:
                    d = d[:i]
                    break
        return d.replace("flatq::", "").replace("(anonymous namespace)::", "")

    print("# r01 — static SASS evidence of `libflatq_b200.so` (sm_100a), from `tools/sass_report.py`\n")
    print("`cuobjdump -sass` mnemonic counts per kernel and `cuobjdump --dump-resource-usage` (registers per thread, static")
    print("shared memory; dynamic shared memory is set at launch).  UTC?MMA = `tcgen05.mma` (I = kind::i8, H = kind::f16),")
    print("LDTM/STTM = `tcgen05.ld/st`, UTMALDG = TMA tensor load, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier ops,")
    print("HMMA = legacy `mma.sync` (only the fallback transform kernel for odd factor shapes), IDP4A = `dp4a` (decode GEMV).\n")
    cols = [w for w in WATCH if any(counts[k][w] for k in order)]
    print("| kernel | instr | regs | " + " | ".join(cols) + " |")
    print("|---|---:|---:|" + "---:|" * len(cols))
    for k in sorted(order, key=short):
        c = counts[k]
        r = res.get(k, ("?", 0, 0))
        print(f"| `{short(k)}` | {c['_total']} | {r[0]} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    tc = [short(k) for k in order if counts[k]["UTCIMMA"] or counts[k]["UTCHMMA"] or counts[k]["UTCQMMA"]]
    print(f"\n{len(order)} kernels; {len(tc)} issue tcgen05 MMAs; "
          f"{sum(1 for k in order if counts[k]['UTMALDG'])} use TMA tensor loads; "
          f"{sum(1 for k in order if counts[k]['HMMA'])} use legacy HMMA; no HGMMA/IGMMA (wgmma) anywhere: "
          f"{'confirmed' if not re.search(r'[HIQ]GMMA', sass) else 'VIOLATED'}.")


if __name__ == "__main__":
    sys.exit(main())
