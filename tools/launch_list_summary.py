"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total time, share).

    python tools/launch_list_summary.py profiles/r2_launch_list.csv "<command the list came from>" > profiles/r2_launch_list_summary.json

Per-launch times under ncu are cold-cache and serialised: the SHARES are what is compared with bench.py's live
CUDA-event shares (`roofline.share_of_step`, `roofline_attn.share_of_step`), not the absolute times.
"""
import csv
import json
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    source = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = [l for l in open(path) if l.startswith('"')]
    agg = OrderedDict()
    for r in csv.DictReader(rows):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r["Kernel Name"]).split("(")[0]
        ns = float(r["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    kernels = [{"kernel": k, "launches": v[0], "total_us": round(v[1] / 1e3, 1), "share": round(v[1] / total, 4)}
               for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    print(json.dumps({"source": source, "total_ms": round(total / 1e6, 1), "launches": sum(v[0] for v in agg.values()),
                      "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
