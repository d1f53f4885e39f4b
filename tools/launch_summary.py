"""Summarise an ncu launch list (CSV of `--metrics gpu__time_duration.sum`) into profiles/launches_r01_summary.txt:

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
      python tools/ncu_step.py --batch 128
  python tools/launch_summary.py gpurun_out/launches.csv
"""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TC_FWD = ("conv_fwd_tc2_kernel", "conv_fwd_tc_kernel")


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    k_name, k_val, k_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in data:
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r[k_unit]]
        name = re.sub(r"\(.*", "", r[k_name])
        if name.startswith("void "):
            name = name[5:]
        if "conv_" in r[k_name] or "first_block" in r[k_name] or "attn_" in r[k_name]:          # keep template arguments of our own kernels
            name = re.sub(r"\(.*", "", r[k_name].replace("void ", ""))
        tot[name] += float(r[k_val].replace(",", "")) * scale
        cnt[name] += 1
    total = sum(tot.values())
    tc = sum(v for k, v in tot.items() if any(t in k for t in TC_FWD))
    out = ["ncu launch list (gpu__time_duration.sum, --clock-control none, cold-cache serialised): one training iteration, O config (1x32x32, m=n=k=5), B=128, bf16 path",
           "command: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python tools/ncu_step.py --batch 128   (round 2, final kernels)",
           "total %.1f ms over %d launches (serialised under ncu; the CUDA-graph step with two-stream overlap is shorter, see bench.py)" % (total, sum(cnt.values())),
           "tcgen05 forward/dgrad kernels (conv_fwd_tc2_kernel<0>, <1>, conv_fwd_tc_kernel): %.1f ms = %.1f %% of the launch list" % (tc, 100 * tc / total), ""]
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append("%9.3f ms %5.1f%% %6d  %s" % (v, 100 * v / total, cnt[name], name[:110]))
    text = "\n".join(out) + "\n"
    out_name = sys.argv[2] if len(sys.argv) > 2 else "launches_r02_summary.txt"
    open(os.path.join(ROOT, "profiles", out_name), "w").write(text)
    print("\n".join(out[:45]))


if __name__ == "__main__":
    main()
