"""profiles/ncu_traffic.json from an `ncu --set full` capture taken INSIDE a bench.py step (tools/gpu_call*.sh):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel, averaged over the captured
launches, with the per-launch rows kept.  bench.py reads the file and puts the value into `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/r2_bench_gemm.ncu-rep gemm_bf16 "what was captured" [-o profiles/ncu_traffic.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import raw_page  # noqa: E402


def main():
    rep, kind, what = sys.argv[1], sys.argv[2], sys.argv[3]
    out = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    rows = [r for r in raw_page(rep) if "dram_bytes" in r]
    per = [{"grid": r.get("grid"), "duration_ms": r.get("duration_ms"), "dram_bytes": r["dram_bytes"], "tensor_pipe_pct": r.get("tensor_pipe_pct"),
            "l2_hit_pct": r.get("l2_hit_pct")} for r in rows]
    try:
        cur = json.load(open(out))
    except Exception:  # noqa: BLE001
        cur = {}
    cur[kind] = {"dram_bytes_per_launch": sum(p["dram_bytes"] for p in per) / len(per), "launches_captured": len(per),
                 "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, captured inside a bench.py step): {what}",
                 "per_launch": per}
    json.dump(cur, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in cur[kind].items() if k != "per_launch"}))


if __name__ == "__main__":
    main()
