#!/usr/bin/env python
"""Turn ncu outputs (brought back in gpurun_out/) into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches.csv profiles/rNN_launch_summary.md
    python tools/summarize_profiles.py full gpurun_out/prof_x.ncu-rep profiles/rNN_ncu_x.md
    python tools/summarize_profiles.py dram gpurun_out/launches_dram.csv profiles/rNN_launch_dram.md \
        profiles/conv_dram_traffic.json
"""
import json
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        name = re.sub(r"[<(].*", "", row["Kernel Name"]).replace("void ", "").replace("acnn::", "")
        agg[name][0] += 1
        agg[name][1] += float(row["Metric Value"].replace(",", "")) / 1e6
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list of one training step (Assemble-ResNet-50, B=256, 224x224)\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off "
                 "python tools/profile_step.py --ncu` (one eager step after warm-up; cold-cache, "
                 "serialised: compare shares).\n\n")
        fh.write("%d launches, sum of kernel durations %.2f ms\n\n" % (n, tot))
        fh.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.1f%% | %.1f |\n" % (k, v[0], v[1], 100 * v[1] / tot,
                                                                1e3 * v[1] / v[0]))
    print("wrote", dst)


def dram(src, dst, dst_json=None):
    """Launch list taken with gpu__time_duration + dram__bytes_{read,write} + lts__t_bytes: per-kernel
    time, DRAM traffic and L2 traffic; the tcgen05 GEMM totals go to a json that bench.py reports
    as roofline.traffic."""
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.defaultdict(dict)
    scale_t = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        per[i]["k"] = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("acnn::", "")
        v = float(row["Metric Value"].replace(",", ""))
        n, u = row["Metric Name"], row["Metric Unit"]
        if n == "gpu__time_duration.sum":
            v *= scale_t.get(u, 1.0)
        elif "bytes" in n:
            v *= scale_b.get(u, 1.0)
        per[i][n] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[d["k"]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("lts__t_bytes.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list with DRAM / L2 traffic, one training step (Assemble-ResNet-50, B=256)\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                 "sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none "
                 "--profile-from-start off python tools/profile_step.py --ncu` (eager step, every kernel "
                 "replayed in isolation: cold-cache times, and DRAM bytes of a replay can be BELOW the "
                 "algorithmic traffic when the operands of the previous pass are still in the 126 MB L2).\n\n")
        fh.write("%d launches, sum of kernel durations %.2f ms\n\n" % (len(per), tot))
        fh.write("| kernel | launches | total ms | DRAM GB | DRAM GB/s | L2 GB |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.2f | %.0f | %.2f |\n"
                     % (k, a[0], a[1], a[2] / 1e9, a[2] / 1e6 / a[1] if a[1] else 0, a[3] / 1e9))
    print("wrote", dst)
    if dst_json:
        g = [d for d in per.values() if "gemm_kernel" in d["k"] or "conv_halo_kernel" in d["k"]]
        out = {"source": dst, "launches": len(g),
               "dram_bytes_per_step": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in g),
               "l2_bytes_per_step": sum(d.get("lts__t_bytes.sum", 0) for d in g),
               "kernel_ms_sum_under_ncu": sum(d.get("gpu__time_duration.sum", 0) for d in g)}
        json.dump(out, open(dst_json, "w"), indent=1)
        print("wrote", dst_json, out)


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write("# ncu --set full --clock-control none: %s\n\n" % src.split("/")[-1])
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            fh.write("## %s\n\n| metric | unit | value |\n|---|---|---:|\n" % d.get("Kernel Name", "?")[:160])
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS:
                    fh.write("| %s | %s | %s |\n" % (h, u, v))
            fh.write("\n")
        src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source",
                                  "sass"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src_csv.splitlines()))
        if len(srows) > 2:
            h = srows[1]
            try:
                ia, isamp = h.index("Source"), h.index("# Samples")
                data = [(int(r[isamp]), r[ia].strip()) for r in srows[2:]
                        if len(r) > max(ia, isamp) and r[isamp].isdigit()]
                tot = sum(x[0] for x in data) or 1
                fh.write("Top warp-stall sample sites (SASS, all captured kernels of this report):\n\n| samples | share | instruction |\n|---:|---:|---|\n")
                for sm, ins in sorted(data, reverse=True)[:8]:
                    fh.write("| %d | %.1f%% | `%s` |\n" % (sm, 100 * sm / tot, ins[:90]))
            except ValueError:
                pass
    print("wrote", dst)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dram":
    dram(*sys.argv[2:])
    sys.exit(0)
if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
