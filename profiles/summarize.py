"""Turns gpurun_out/{launches_*.csv, prof_*.ncu-rep} into the small tracked summaries under profiles/.

usage: python profiles/summarize.py <tag> <launches.csv> [<report.ncu-rep> <kernel regex>]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "mio_throttle", "barrier", "wait", "math_pipe_throttle", "lg_throttle",
               "not_selected", "dispatch_stall", "branch_resolving", "no_instruction", "membar", "tex_throttle", "imc_miss"]


def to_num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        a = agg.setdefault(name, dict(kernel=name, launches=0, total_us=0.0, grid=r[gi], block=r[bi], last_us=0.0))
        a["launches"] += 1
        a["total_us"] += v
        a["last_us"] = v
    out = list(agg.values())
    for a in out:
        a["mean_us"] = a["total_us"] / a["launches"]
    return out


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, units, data = r, rows[i + 1], rows[i + 2:]
            break
    res = []
    for r in data:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT + [STALLS % s for s in STALL_NAMES]:
            if w in hdr:
                d[w] = to_num(r[hdr.index(w)])
                d[w + ".unit"] = units[hdr.index(w)]
        res.append(d)
    return res


if __name__ == "__main__":
    tag = sys.argv[1]
    out = {"tag": tag, "launch_list": launches(sys.argv[2])}
    steady = {a["kernel"]: a["last_us"] for a in out["launch_list"]}
    out["note"] = "per-launch times are from ncu's serialised replay (cold caches): compare shares, not absolutes"
    if len(sys.argv) > 3:
        out["ncu_full"] = ncu_raw(sys.argv[3])
    with open(os.path.join(HERE, f"{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["launch_list"], indent=1))
