import csv, os, re, subprocess, sys, tempfile
rep, kname = sys.argv[1], sys.argv[2]
root = "/root/repo"
lib = os.path.join(root, "sitator_b200", "lib", "libsitator_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname not in out: continue
    cur_fn = cur_line = None; in_k = False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: in_k = kname in m.group(1); continue
        if not in_k: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur_fn, cur_line = os.path.basename(m.group(1)), int(m.group(2)); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m: lines.append((cur_fn, cur_line, m.group(2)))
    if lines: break
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = [i for i, r in enumerate(rows) if len(r) > 1 and r[0] == "Address"][0]
hdr = rows[hi]
ix_i, ix_s, ix_t = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
sass = [(r[1], int(r[ix_i]), int(r[ix_s]), int(r[ix_t])) for r in rows[hi + 1:] if len(r) > ix_t]
marks = [("// ---- 1.", "1 wrap"), ("// ---- 2.", "2 lattice"), ("// ---- 3.", "3 claim"), ("// 3a.", "3a"), ("// 3b'.", "3b grid"),
         ("// 3b.", "3b full"), ("// 3c.", "3c"), ("// 3d.", "3d items"), ("// component values", "3d root"), ("// 3e.", "3e"), ("// ---- flush", "flush")]
bounds = [(0, 'setup')]
for i, l in enumerate(open(os.path.join(root, "sitator_b200", "csrc", "sitb_fill.cu")).read().splitlines(), 1):
    for tag, name in marks:
        if tag in l and name not in [b[1] for b in bounds]:
            bounds.append((i, name))
def ph(l):
    c='setup'
    for b,n in bounds:
        if l>=b: c=n
    return c
cur='setup'; agg={}
tot=0
for i in range(min(len(sass),len(lines))):
    fn,l,_=lines[i]
    if fn=='sitb_fill.cu' and l and l>=bounds[1][0]-40: cur=ph(l)
    a=agg.setdefault(cur,[0,0,0]); a[0]+=sass[i][1]; a[1]+=sass[i][2]; a[2]+=sass[i][3]; tot+=sass[i][1]
ntask=float(sys.argv[3]) if len(sys.argv)>3 else 1
for k,a in agg.items(): print("%-10s %5.1f%% inst  %7.1f inst/task  samp %5.1f%% lanes %.1f"%(k,100*a[0]/tot,a[0]/ntask,100*a[1]/sum(x[1] for x in agg.values()),a[2]/max(a[0],1)))
