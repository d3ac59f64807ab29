#!/usr/bin/env python3
"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into the small text summaries committed under profiles/.
usage: python tools/summarize_profiles.py r01a [report.ncu-rep ...]   (tag = round + letter; without reports: every .ncu-rep in gpurun_out/)"""
import collections
import csv
import io
import os
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

LAUNCHES = os.environ.get("LAUNCHES", "launches.csv")           # launch list inside gpurun_out/
CMD = os.environ.get("PROF_CMD", "python bench.py --steps 2 --warmup 3")
if os.path.exists(os.path.join(G, LAUNCHES)) and not os.environ.get("NO_LAUNCHES"):
    lines = [l for l in open(os.path.join(G, LAUNCHES)).read().splitlines() if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
        order.append((name, row["Grid Size"], row["Block Size"], float(row["Metric Value"].replace(",", ""))))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` of `{CMD}`\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  Includes one-time setup (k_build_table,\nk_pow_table, k_ntt_setup) and the L2-flush fill kernel of the command.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:80]}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |\n")
        f.write("\n## one MSM step (last complete MSM in the list), launch by launch\n\n| kernel | grid | block | us |\n|---|---|---|---:|\n")
        idx = [i for i, o in enumerate(order) if "k_recode" in o[0]]
        if idx:
            i0 = idx[-1]
            i1 = next((i for i in range(i0 + 1, len(order)) if "k_gridsum_final" in order[i][0] or "k_bitsum_final" in order[i][0] or "k_finish" in order[i][0]), len(order) - 1)
            for o in order[i0:i1 + 1]:
                f.write(f"| `{o[0][:70]}` | {o[1]} | {o[2]} | {o[3] / 1e3:.1f} |\n")
    print("wrote", f"{tag}_launches.md")

only = [os.path.basename(a) for a in sys.argv[2:]]
for rep in sorted(os.listdir(G)):
    if not rep.endswith(".ncu-rep") or (only and rep not in only):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    name = rep.replace(".ncu-rep", "")
    if name.startswith(tag): name = name[len(tag):].lstrip("_")
    with open(os.path.join(P, f"{tag}_{name}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({tag}, {rep})\n\nCaptured under `python tools/prof_cmd.py` (bench-sized inputs, L2 flushed between iterations) on B200; one column per captured launch.\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |\n|---|---|" + "---:|" * (len(rows) - 2) + "\n")
        for w in ["Kernel Name", "Grid Size", "Block Size"] + WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"| {w} | {units[i]} | " + " | ".join(r[i][:60] for r in rows[2:]) + " |\n")
    print("wrote", f"{tag}_{name}.md")
