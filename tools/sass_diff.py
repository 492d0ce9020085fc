"""Static regression check without a GPU: compare the SASS instruction streams of two builds of libflatq_b200.so kernel
by kernel (addresses and encodings stripped, operands -- including constant-bank offsets of kernel parameters -- kept).
Used when a change must not touch the code of kernels that have already been verified on hardware, e.g. a new template
parameter with a default:

    cp flatquant_b200/lib/libflatq_b200.so /tmp/before.so;  <edit>;  python __graft_entry__.py --force
    python tools/sass_diff.py /tmp/before.so flatquant_b200/lib/libflatq_b200.so [--map 'REGEX=>REPL']

--map rewrites mangled names of the OLD build before matching (a new defaulted template argument changes the mangling)."""
import re
import subprocess
import sys


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, name, buf = {}, None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = buf
            name, buf = m.group(1), []
        elif name:
            mm = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?)\s*/\*", line)
            if mm:
                buf.append(mm.group(1))
    if name:
        out[name] = buf
    return out


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    maps = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--map=")]
    if "--map" in sys.argv:
        maps.append(sys.argv[sys.argv.index("--map") + 1])
        args = [a for a in args if a not in maps]
    old, new = functions(args[0]), functions(args[1])
    same = changed = removed = 0
    matched = set()
    for k in sorted(old):
        k2 = k
        for m in maps:
            pat, repl = m.split("=>")
            k2 = re.sub(pat, repl, k2)
        if k2 not in new:
            removed += 1
            print("REMOVED  ", demangle(k)[:120])
            continue
        matched.add(k2)
        if old[k] == new[k2]:
            same += 1
        else:
            changed += 1
            nd = sum(1 for x, y in zip(old[k], new[k2]) if x != y) + abs(len(old[k]) - len(new[k2]))
            print(f"CHANGED   {demangle(k)[:100]}: {len(old[k])} -> {len(new[k2])} instructions, {nd} differ")
    fresh = [k for k in new if k not in matched]
    for k in fresh:
        print(f"NEW       {demangle(k)[:110]} ({len(new[k])} instructions)")
    print(f"identical {same}, changed {changed}, removed {removed}, new {len(fresh)}")
    return 1 if changed or removed else 0


if __name__ == "__main__":
    sys.exit(main())
