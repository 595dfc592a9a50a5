"""Developer tool: turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown table kept
under profiles/ (share, total ms, launches, ms per launch, kernel).  usage: launch_list_md.py launches.csv "title" > out.md"""
import csv, sys
rows = []
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if len(r) > 10 and r[0].isdigit():
        rows.append(r)
hdr = None
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if r and r[0] == "ID":
        hdr = r
        break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else (v if u in ("ms", "msecond") else v * 1e3))
    a = agg.setdefault(r[ik], [0.0, 0])
    a[0] += ms; a[1] += 1
tot = sum(a[0] for a in agg.values())
print("# %s\n" % (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"))
print("`ncu --metrics gpu__time_duration.sum --clock-control none` on a B200; times are cold-cache and serialised: compare SHARES.\n")
print("| share | total ms | launches | ms/launch | kernel |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
    print("| %.2f%% | %.3f | %d | %.3f | `%s` |" % (100 * a[0] / tot, a[0], a[1], a[0] / a[1], k[:80]))
print("\ntotal %.1f ms over %d launches" % (tot, sum(a[1] for a in agg.values())))
