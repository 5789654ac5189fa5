#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`:
writes the JSON that bench.py reads for roofline.traffic (DRAM bytes read + written per launch).
usage: extract_traffic.py raw.csv out.json pairs_per_gpu read_len"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def to_ms(v, unit):
    return float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)


out = {"workload": {"pairs_per_gpu": int(sys.argv[3]), "read_len": int(sys.argv[4])}, "kernels": {}}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = ("scan_r1" if "k_scan<0" in name or "k_scan<(int)0" in name else "scan_r2" if "k_scan<1" in name or "k_scan<(int)1" in name
           else "emit" if "k_emit" in name else name.split("(")[0])
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    out["kernels"][key] = {
        "kernel": name.split("(")[0], "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic": rd + wr,
        "duration_ms_under_ncu": to_ms(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]),
        "warp_instructions": float(r[col["smsp__inst_executed.sum"]].replace(",", "")),
        "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
        "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
        "registers_per_thread": float(r[col["launch__registers_per_thread"]]),
        "dram_throughput_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
    }
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
