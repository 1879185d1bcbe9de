"""Summarise ncu outputs into profiles/: launch list (csv from --metrics gpu__time_duration.sum)
and raw pages of full captures.  usage: ncu_summary.py launches <csv> | raw <ncu-rep>"""
import collections
import csv
import subprocess
import sys

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| %s | %d | %.1f | %.1f | %.3f |" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, sum(v) / tot))
else:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print("### %s" % name.split("(")[0])
        for k in keys:
            if k in h:
                print("- %s = %s %s" % (k, r[h.index(k)], rows[1][h.index(k)]))
