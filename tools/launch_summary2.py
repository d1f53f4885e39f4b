"""Summarise an ncu launch list CSV (gpu__time_duration.sum) by kernel name: python tools/launch_summary2.py file.csv [top]"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
kn, kv, ku = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in data:
    s = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r[ku]]
    name = re.sub(r"\(.*", "", r[kn].replace("void ", ""))
    tot[name] += float(r[kv].replace(",", "")) * s; cnt[name] += 1
T = sum(tot.values())
print("total %.2f ms over %d launches" % (T, sum(cnt.values())))
ours = sum(v for k, v in tot.items() if k.startswith("gim::"))
print("own kernels %.2f ms (%d launches), torch kernels %.2f ms (%d launches)" % (ours, sum(c for k, c in cnt.items() if k.startswith("gim::")), T - ours, sum(c for k, c in cnt.items() if not k.startswith("gim::"))))
for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 70]:
    print("%9.3f ms %5.1f%% %5d %s" % (v, 100 * v / T, cnt[n], n[:110]))
