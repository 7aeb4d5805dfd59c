"""Static evidence of the built library, no GPU needed:
    python tools/sass_resources.py [path/to/libkaldicnn_b200.so] > profiles/rNN_sass_resources.md
Registers / shared / stack / local memory of every kernel (cuobjdump --dump-resource-usage) and the
Blackwell instruction mnemonics in its SASS (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor
load, UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier), per kernel family."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    res = []
    for n in out:
        n = n.replace("(anonymous namespace)::", "")
        depth, cut = 0, len(n)
        for i, ch in enumerate(n):               # drop the argument list: the first '(' outside <...>
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        res.append(n[:cut])
    return res


def family(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    return name


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "kaldi-cnn_b200", "lib", "libkaldicnn_b200.so")
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
    rows, fn = [], None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            rows.append([fn] + [int(x) for x in m.groups()])
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, cur = collections.defaultdict(collections.Counter), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and m.group(1) in MNEMONICS:
            counts[cur][m.group(1)] += 1
    names = demangle([r[0] for r in rows])
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("# Static resources and Blackwell instructions of libkaldicnn_b200.so (sm_100a)\n")
    print("`python tools/sass_resources.py` (cuobjdump --dump-resource-usage, cuobjdump -sass; no GPU involved).\n")
    print("%d kernels; kernels with a stack frame or local memory (spills): **%d**.\n"
          % (len(rows), sum(1 for r in rows if r[2] > 0 or r[4] > 0)))
    print("Whole library: " + ", ".join("%s x %d" % (k, total[k]) for k in MNEMONICS if total[k]) + "\n")
    print("| kernel | regs | static smem B | stack B | local B | tensor / TMA / barrier instructions |")
    print("|---|---|---|---|---|---|")
    for n, r in sorted(zip(names, rows), key=lambda t: t[0]):
        c = counts.get(r[0], {})
        ins = ", ".join("%s %d" % (k, c[k]) for k in MNEMONICS if c.get(k) and k not in ("FFMA",))
        print("| `%s` | %d | %d | %d | %d | %s |" % (family(n), r[1], r[3], r[2], r[4], ins))


if __name__ == "__main__":
    main()
