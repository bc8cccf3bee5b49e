"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/summarize_launches.py in.csv "title" """
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(d["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for k, v in agg.items() if "fp32_peak" not in k)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = "" if "fp32_peak" in k else f"share={v[1] / tot * 100:5.1f}%"
    print(f"{k[:56]:56s} n={v[0]:4d} total={v[1]:9.3f} ms avg={v[1] / v[0]:8.4f} ms {share}")
