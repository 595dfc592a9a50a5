"""Static evidence from the built library (no GPU needed): per kernel the SASS instruction count, the mnemonics
that show which units it uses (tcgen05 = UTC*MMA / UTCBAR / LDTM, TMA / bulk copy = UTMALDG / UBLKCP, FP64 tensor
= DMMA, FP64 pipe = DFMA / DADD / DMUL, SFU = MUFU) and registers / stack / shared memory from cuobjdump.
Usage: python scripts/sass_summary.py > profiles/rNN_static_sass.md"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sitator_b200", "lib", "libsitator_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "DFMA",
        "DADD", "DMUL", "MUFU", "FFMA", "ATOMS", "ATOMG", "RED", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
            total[cur] += 1
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = tuple(int(x) for x in m.groups())
    names = demangle(list(total))
    print("# Static SASS summary of libsitator_b200.so (sm_100a, `cuobjdump -sass` / `--dump-resource-usage`)\n")
    print("| kernel | SASS instr. | regs | stack B | static smem B | unit mnemonics (static counts) |")
    print("|---|---|---|---|---|---|")
    for f in sorted(total, key=lambda k: -total[k]):
        nm = re.sub(r"\(.*", "", names[f]).replace("void ", "").replace("sitb::", "")
        r = usage.get(f, ("?", "?", "?"))
        mn = ", ".join("%s %d" % (k, counts[f][k]) for k in KEYS if counts[f][k])
        print("| `%s` | %d | %s | %s | %s | %s |" % (nm, total[f], r[0], r[1], r[2], mn))


if __name__ == "__main__":
    main()
