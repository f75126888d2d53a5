"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST n launches (one step).
usage: python tools/launch_summary.py file.csv [launches_per_step]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, recs = None, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if "gpu__time_duration" in d.get("Metric Name", ""):
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            us = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
            recs.append((d["Kernel Name"], us, d["Grid Size"], d["Block Size"]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(recs)
recs = recs[-n:]
agg = collections.OrderedDict()
for k, us, g, b in recs:
    a = agg.setdefault(k[:110], [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("total %.1f us in %d launches" % (tot, len(recs)))
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%9.1f us %5.1f %% %4d  %s" % (us, 100 * us / tot, c, k))
