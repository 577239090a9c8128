#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (.ncu-rep) per launch: duration, DRAM traffic, achieved DRAM GB/s, SM throughput,
achieved occupancy, registers, warp instructions and the two largest warp-stall reasons.

    python tools/ncu_full_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full_summary.txt
"""
import csv
import io
import subprocess
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return float("nan")


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    stall = [h for h in head if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]

    def to_bytes(v, u):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def to_us(v, u):
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u.replace("second", "s").replace("nsecond", "ns"), 1)

    print("%-44s %8s %9s %9s %8s %6s %6s %5s %10s  %s" % ("kernel", "us", "dramR_MB", "dramW_MB", "GB/s", "sm%", "occ%", "regs",
                                                          "warp_inst", "top stalls (warps per issue)"))
    for r in body:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")[:44]
        us = to_us(num(r[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]])
        rd = to_bytes(num(r[col["dram__bytes_read.sum"]]), units[col["dram__bytes_read.sum"]])
        wr = to_bytes(num(r[col["dram__bytes_write.sum"]]), units[col["dram__bytes_write.sum"]])
        sm = num(r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]])
        occ = num(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])
        regs = num(r[col["launch__registers_per_thread"]])
        inst = num(r[col["smsp__inst_executed.sum"]])
        st = sorted(((num(r[col[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall),
                    reverse=True)[:2]
        print("%-44s %8.1f %9.1f %9.1f %8.0f %6.1f %6.1f %5d %10.3g  %s" % (
            name, us, rd / 1e6, wr / 1e6, (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0, sm, occ, regs, inst,
            ", ".join("%s %.1f" % (n, v) for v, n in st)))


if __name__ == "__main__":
    main(sys.argv[1])
