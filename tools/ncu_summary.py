#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (.ncu-rep) of k_advect_step into markdown + JSON.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_advect_tuned  [kernel-regex]

Writes <out>.md (human-readable), <out>.json (key metrics per captured launch) and, when
the report has source-level data, <out>_opmix.csv (executed warp instructions per opcode).
Needs only the ncu CLI (no GPU).
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    n_items = float(sys.argv[3]) if len(sys.argv) > 3 else None      # work items (buoys) per launch, for persistent kernels
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}
    launches = []
    for r in data:
        d = {"kernel": r[ix["Kernel Name"]][:100]}
        for k in KEYS:
            if k in ix:
                v, u = r[ix[k]], units[ix[k]]
                try:
                    d[k] = to_bytes(v, u) if "byte" in u else float(v.replace(",", ""))
                except ValueError:
                    d[k] = v
                if k == "gpu__time_duration.sum":
                    d[k] = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)   # -> us
        d["dram_bytes_per_launch"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        launches.append(d)
    json.dump(launches, open(out + ".json", "w"), indent=1)

    md = ["# ncu --set full summary: %s" % rep.split("/")[-1], ""]
    for n, d in enumerate(launches):
        md.append("## launch %d: `%s`" % (n, d["kernel"]))
        md.append("")
        md.append("| metric | value |")
        md.append("|---|---|")
        for k in KEYS + ["dram_bytes_per_launch"]:
            if k in d:
                v = d[k]
                md.append("| %s | %s |" % (k, ("%.4g" % v) if isinstance(v, float) else v))
        md.append("")

    src = ncu(["-i", rep, "--page", "source", "--csv"])
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    mix = collections.Counter()
    for r in rows:
        if r and r[0] == "Address":
            if h is not None:
                break                                   # first kernel instance only
            h = {k: i for i, k in enumerate(r)}
            continue
        if h is None or len(r) < len(h):
            continue
        try:
            ie = int(r[h["Instructions Executed"]])
        except ValueError:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[h["Source"]])
        mix[m.group(1) if m else "?"] += ie
    if mix:
        tot = sum(mix.values())
        grid = launches[0].get("launch__grid_size", 0) * launches[0].get("launch__block_size", 0) / 32 or 1
        if n_items:
            grid = n_items / 32.0                       # "per warp" then means per 32 work items
        with open(out + "_opmix.csv", "w") as f:
            f.write("opcode,warp_instructions,percent,per_warp\n")
            for k, v in mix.most_common():
                f.write("%s,%d,%.2f,%.1f\n" % (k, v, 100 * v / tot, v / grid))
        dp = sum(v for k, v in mix.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
        md.append("## executed instruction mix (launch 0)")
        md.append("")
        md.append("total warp instructions %d = %.0f per warp; FP64-pipe (DFMA+DADD+DMUL+DSETP) %.0f per warp"
                  % (tot, tot / grid, dp / grid))
        md.append("")
        md.append("| opcode | per warp | % |")
        md.append("|---|---|---|")
        for k, v in mix.most_common(16):
            md.append("| %s | %.1f | %.1f |" % (k, v / grid, 100 * v / tot))
    open(out + ".md", "w").write("\n".join(md) + "\n")
    print("\n".join(md[:40]))


if __name__ == "__main__":
    main()
