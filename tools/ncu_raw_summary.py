#!/usr/bin/env python
"""Summarise `ncu -i x.ncu-rep --page raw --csv` exports (made on the GPU box; the .ncu-rep files of a
--set full capture exceed what gpurun brings back) into a markdown table.

    python tools/ncu_raw_summary.py profiles/out.md gpurun_out/r2g_*.raw.csv
"""
import csv
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "MB rd", 1.0),
    ("dram__bytes_write.sum", "MB wr", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1.0),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tc smem %", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__grid_size", "grid", 1.0),
]


def rows_of(path):
    rd = list(csv.reader(open(path)))
    hdr, units = rd[0], rd[1]
    out = []
    for r in rd[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"name": re.sub(r"\(.*", "", d.get("Kernel Name", "?")).replace("void ", "").replace("acnn::", "")}
        for k, _, _ in KEYS:
            v = d.get(k, "")
            try:
                val = float(v.replace(",", ""))
            except ValueError:
                val = float("nan")
            unit = u.get(k, "")
            if k.startswith("dram__bytes"):
                val *= {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(unit, 1.0)
            if k == "gpu__time_duration.sum":
                val *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
            rec[k] = val
        out.append(rec)
    return out


def main():
    dst, srcs = sys.argv[1], sys.argv[2:]
    with open(dst, "w") as fh:
        fh.write("# ncu --set full --clock-control none captures (round 2), raw-page summary\n\n")
        fh.write("Captured under gpurun on one B200; `ncu -i <rep> --page raw --csv` was run on the GPU box and "
                 "only the CSV exports were brought back (gpurun_out/*.raw.csv; tools/ncu_raw_summary.py). "
                 "Times are single cold-cache replays. HBM peak for the fractions: 6572.9 GB/s measured copy "
                 "bandwidth (MEASURED_PEAKS.json). `tc smem %` = l1tex__data_pipe_tc_wavefronts_mem_shared: the "
                 "tensor core's shared-memory operand reads (what paces the N <= 64 tiles, DESIGN section 4).\n\n")
        fh.write("| capture | kernel | us | DRAM MB (rd+wr) | GB/s | of 6573 | DRAM % | L2 % | tensor % | tc smem % | issue % | occ % | regs | grid |\n")
        fh.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for src in srcs:
            cap = re.sub(r".*r2g_|.*r02f_|\.raw\.csv", "", src)
            for r in rows_of(src):
                us = r["gpu__time_duration.sum"]
                mb = r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]
                gbs = mb / 1e3 / (us * 1e-6) if us > 0 else float("nan")
                fh.write("| %s | `%s` | %.1f | %.1f | %.0f | %.2f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %d | %d |\n" % (
                    cap, r["name"][:70], us, mb, gbs, gbs / 6572.9,
                    r["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
                    r["lts__throughput.avg.pct_of_peak_sustained_elapsed"],
                    r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"],
                    r["l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"],
                    r["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                    r["sm__warps_active.avg.pct_of_peak_sustained_active"],
                    r["launch__registers_per_thread"], r["launch__grid_size"]))
    print("wrote", dst)


if __name__ == "__main__":
    main()
