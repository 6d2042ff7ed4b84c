#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into the tracked text files of profiles/.
usage: python profiles/summarize.py launches <launches.csv> <out.txt> | raw <rep> <out.txt> [regex]"""
import collections, csv, subprocess, sys


def launches(path, out):
    rows = list(csv.reader(open(path)))
    i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[i0 + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("stab::", "").replace("void ", "")
        t = float(r[ix["Metric Value"]].replace(",", "")) / 1e6
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised launches: compare SHARES)\n")
        f.write(f"# source: {path}; total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:62s} n={v[0]:6d} total={v[1]:10.2f} ms  {100 * v[1] / tot:5.1f}%\n")


KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source report: {rep} (scratch, not tracked)\n")
        for r in rows[2:]:
            f.write("---\n")
            for k in KEYS:
                if k in ix:
                    f.write(f"{k} = {r[ix[k]]} {units[ix[k]]}\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        raw(sys.argv[2], sys.argv[3])
