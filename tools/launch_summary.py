"""Per-kernel totals of an ncu launch list (csv, --metrics gpu__time_duration.sum[,dram__bytes_*]).
Usage: python tools/launch_summary.py launches.csv [--all]   (--all also prints every launch in order)"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
launches = collections.OrderedDict()
for r in rows:
    d = launches.setdefault(int(r[0]), {"name": re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", ""), "grid": r[8]})
    d[r[12]] = float(r[14]) * ({"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(r[13], 1.0))
agg = collections.OrderedDict()
for d in launches.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"{len(launches)} launches, {tot:.1f} us (serialised, cold cache: compare shares)")
print(f"{'kernel':58s} {'n':>3s} {'us':>8s} {'share':>6s} {'dram MB':>8s}")
for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:58]:58s} {c:3d} {t:8.1f} {100 * t / tot:5.1f}% {b:8.1f}")
if "--all" in sys.argv:
    for i, d in launches.items():
        print(i, d["name"][:50], d["grid"], round(d.get("gpu__time_duration.sum", 0), 1), round(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0), 1))
