#!/usr/bin/env python
"""Sum an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of the convolution launches
of one pass into the JSON bench.py reads for `roofline.traffic`.  Usage: python tools/conv_traffic.py <csv> <out.json> "<source note>" """
import collections, csv, json, re, sys

path, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.defaultdict(lambda: {"launches": 0, "time_ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
ids = collections.defaultdict(set)
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").strip()
    unit, metric = row["Metric Unit"], row["Metric Name"]
    ids[name].add(row["ID"])
    if metric.startswith("gpu__time_duration"):
        per[name]["time_ms"] += v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        per[name]["dram_read_bytes" if "read" in metric else "dram_write_bytes"] += v * scale
for k in per:
    per[k]["launches"] = len(ids[k])
res = {"source": note,
       "dram_read_bytes": sum(p["dram_read_bytes"] for p in per.values()),
       "dram_write_bytes": sum(p["dram_write_bytes"] for p in per.values()),
       "launches": sum(p["launches"] for p in per.values()), "per_kernel": per}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "per_kernel"}))
