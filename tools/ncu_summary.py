"""Summarise an `ncu --set full` report into the JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/some.ncu-rep [-o profiles/rN_xxx_ncu_summary.json] [--roles]

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed) and keeps, per captured launch: duration,
DRAM bytes (read + write) and GB/s, tensor-pipe / XU (MUFU) / FMA / ALU / issue utilisation, L2 hit rate,
registers, shared memory, grid.  With --roles it also reads the SASS-level sampling page
(`--page source --print-source sass`) and splits the samples of a warp-specialised kernel by instruction class:
how often the warps were sampled on tensor-core issue (UTCHMMA / UTCBAR), TMA, MUFU, mbarrier waits - the view that
showed the attention kernel's issuer warp busy issuing 64 % of the time in round 1.
"""
import argparse
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.avg": "sm_cycles",
}


def _ncu(args):
    r = subprocess.run(["ncu", *args], capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(f"ncu failed: {r.stderr[:400]}")
    return r.stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(_ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "?"), "id": d.get("ID")}
        for k, name in KEEP.items():
            if k in d and d[k] != "":
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                rec[name] = v
                if name in ("duration", "dram_read_bytes", "dram_write_bytes"):
                    rec[name + "_unit"] = u.get(k, "")
        dur_s = _seconds(rec.get("duration"), rec.get("duration_unit"))
        tot = _bytes(rec.get("dram_read_bytes"), rec.get("dram_read_bytes_unit")) + _bytes(rec.get("dram_write_bytes"), rec.get("dram_write_bytes_unit"))
        if dur_s:
            rec["duration_ms"] = dur_s * 1e3
            rec["dram_bytes"] = tot
            rec["dram_gbs"] = tot / dur_s / 1e9
        out.append(rec)
    return out


def _seconds(v, unit):
    if v is None:
        return None
    return v * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}.get(unit, 1e-9)


def _bytes(v, unit):
    if v is None:
        return 0.0
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


CLASSES = [("tensor issue (UTCHMMA/UTCBAR)", ("UTCHMMA", "UTCBAR")), ("TMA (UTMALDG/UTMASTG)", ("UTMALDG", "UTMASTG")),
           ("MUFU", ("MUFU",)), ("mbarrier wait (SYNCS...TRYWAIT + its branch)", ("SYNCS.PHASECHK", "TRYWAIT")),
           ("TMEM ld/st (LDTM/STTM)", ("LDTM", "STTM", "UTCLD", "UTCST")), ("R2UR/ELECT", ("R2UR", "ELECT")),
           ("barrier (BAR/WARPSYNC)", ("BAR.", "WARPSYNC"))]


def roles(rep):
    rows = list(csv.reader(io.StringIO(_ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    # the page holds one table per kernel: a "Kernel Name" row, a header row, then instruction rows
    out, hdr, cur = [], None, None
    for row in rows:
        if row and row[0] == "Kernel Name":
            cur = {"kernel": row[1], "samples": 0, "not_issued": 0, "by_class": {c: 0 for c, _ in CLASSES}, "other": 0}
            out.append(cur)
            hdr = None
            continue
        if row and row[0] == "Address":
            hdr = row
            continue
        if cur is None or hdr is None or len(row) != len(hdr):
            continue
        d = dict(zip(hdr, row))
        try:
            s = int(d.get("Warp Stall Sampling (All Samples)") or 0)
            n = int(d.get("Warp Stall Sampling (Not-issued Samples)") or 0)
        except ValueError:
            continue
        cur["samples"] += s
        cur["not_issued"] += n
        src = d.get("Source", "")
        for cname, keys in CLASSES:
            if any(k in src for k in keys):
                cur["by_class"][cname] += s
                break
        else:
            cur["other"] += s
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("-o", "--out")
    ap.add_argument("--roles", action="store_true")
    a = ap.parse_args()
    res = {"source": a.report, "launches": raw_page(a.report)}
    if a.roles:
        res["sampling"] = roles(a.report)
    text = json.dumps(res, indent=1)
    if a.out:
        open(a.out, "w").write(text + "\n")
    else:
        print(text)


if __name__ == "__main__":
    main()
