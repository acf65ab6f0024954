#!/usr/bin/env python
"""Turn an ncu report into the small, committed summaries under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_pose2pose2      # --set full capture
    python tools/summarize_ncu.py --launches gpurun_out/launches.csv profiles/r01_launches.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out + "_metrics.csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["launch", "kernel"] + [k for k in KEYS if k in idx])
        w.writerow(["", ""] + [units[idx[k]] for k in KEYS if k in idx])
        for n, r in enumerate(rows[2:]):
            w.writerow([n, r[idx["Kernel Name"]]] + [r[idx[k]] for k in KEYS if k in idx])
    # dynamic instruction mix of the first launch of each distinct kernel
    seen = set()
    with open(out + "_instmix.md", "w") as fh:
        for n, r in enumerate(rows[2:]):
            name = r[idx["Kernel Name"]]
            if name in seen:
                continue
            seen.add(name)
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(n),
                                  "--launch-count", "1"], capture_output=True, text=True).stdout
            srows = list(csv.reader(src.splitlines()))
            h = srows[1]
            ia, ie, iad = h.index("Source"), h.index("Instructions Executed"), h.index("Address")
            byop, tot, first = collections.Counter(), 0, set()
            for s in srows[2:]:
                try:
                    c = int(s[ie])
                except (ValueError, IndexError):
                    continue
                if s[iad] in first:
                    break
                first.add(s[iad])
                t = s[ia].strip().split()
                op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                byop[op] += c
                tot += c
            fh.write(f"## launch {n}: `{name}`\n\nwarp-level instructions executed: {tot}\n\n| opcode | executed | share |\n|---|---:|---:|\n")
            for op, c in byop.most_common(24):
                fh.write(f"| {op} | {c} | {100 * c / tot:.1f}% |\n")
            fh.write("\n")


def launches(csv_path, out):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    per = collections.OrderedDict()
    order = []
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        t = float(r[iv].replace(",", ""))
        order.append((r[ik], t))
        per.setdefault(r[ik], []).append(t)
    unit = rows[1][hdr.index("Metric Unit")]
    tot = sum(sum(v) for v in per.values())
    with open(out, "w") as fh:
        fh.write(f"ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), times in {unit}; cold-cache, "
                 "serialised -- compare SHARES, not absolutes\n\n| kernel | launches | total | share | mean |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            fh.write(f"| `{k[:110]}` | {len(v)} | {sum(v):.1f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.2f} |\n")
        fh.write("\nlaunch order (first 40):\n\n")
        for k, t in order[:40]:
            fh.write(f"- {t:.2f} {unit}  `{k[:100]}`\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
