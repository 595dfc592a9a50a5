"""Join an ncu SASS source page with nvdisasm line info: per-source-line instruction counts and stall samples.

usage: ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> [top_n]
(developer tool; reads the in-tree libsitator_b200.so)
"""
import csv, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "sitator_b200", "lib", "libsitator_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = []   # per SASS instruction: source line
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname not in out:
        continue
    cur_fn, cur_line, in_k = None, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            in_k = kname in m.group(1)
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_fn, cur_line = os.path.basename(m.group(1)), int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((cur_fn, cur_line, m.group(2)))
    if lines:
        break
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
his = [i for i, r in enumerate(rows) if len(r) > 1 and r[0] == "Address"]
which = int(os.environ.get("NCU_LAUNCH", "0"))          # which captured launch of the report to read
hi = his[which]
rows = rows[:his[which + 1]] if which + 1 < len(his) else rows
hdr = rows[hi]
ix_i, ix_s, ix_t = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
sass = [(r[1], int(r[ix_i]), int(r[ix_s]), int(r[ix_t])) for r in rows[hi + 1:] if len(r) > ix_t and r[ix_i].isdigit()]
print("sass instrs: ncu %d, nvdisasm %d" % (len(sass), len(lines)))
n = min(len(sass), len(lines))
agg = {}
for i in range(n):
    key = (lines[i][0], lines[i][1])
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += sass[i][1]; a[1] += sass[i][2]; a[2] += sass[i][3]
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
src_cache = {}
def src(fn, ln):
    if fn not in src_cache:
        p = os.path.join(root, "sitator_b200", "csrc", fn)
        src_cache[fn] = open(p).read().splitlines() if os.path.isfile(p) else []
    s = src_cache[fn]
    return s[ln - 1].strip() if ln and 0 < ln <= len(s) else ""
print("total warp-instr %d, samples %d" % (ti, ts))
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samp lanes %4.1f  %s:%s  %s" % (100.0 * a[0] / ti, 100.0 * a[1] / max(ts, 1), a[2] / max(a[0], 1), fn, ln, src(fn, ln)[:100]))

# per-phase totals for sitb_fill.cu (line ranges located by the step comments)
if os.environ.get("PHASES"):
    fill = open(os.path.join(root, "sitator_b200", "csrc", "sitb_fill.cu")).read().splitlines()
    marks = []
    for i, l in enumerate(fill, 1):
        for tag in ("// ---- stage", "// ---- 1.", "// ---- 2.", "// ---- 3.", "// 3a.", "// 3b.", "// 3c.", "// 3d.", "// 3e.", "if (MODE == MODE_ASSIGN) {", "// ---- flush"):
            if tag in l:
                marks.append((i, tag))
    marks.sort()
    def phase(fn, ln):
        if fn != "sitb_fill.cu" or ln is None:
            return "inlined helpers (%s)" % fn
        if ln < marks[0][0]:
            return "helpers (cutoff_factor, nth_root, ...)"
        cur = marks[0][1]
        for i, t in marks:
            if ln >= i:
                cur = t
        return cur
    ph = {}
    for (fn, ln), a in agg.items():
        k = phase(fn, ln)
        b = ph.setdefault(k, [0, 0, 0])
        b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
    print("--- phases ---")
    for k, a in sorted(ph.items(), key=lambda kv: -kv[1][0]):
        print("%5.1f%% inst %5.1f%% samp lanes %4.1f  %s" % (100.0 * a[0] / ti, 100.0 * a[1] / max(ts, 1), a[2] / max(a[0], 1), k))
