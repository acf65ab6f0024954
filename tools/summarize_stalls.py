#!/usr/bin/env python
"""Warp-state sampling of an ncu `--set full --import-source on` capture, condensed for profiles/.

    python tools/summarize_stalls.py gpurun_out/x_pose2pose2.ncu-rep [more.ncu-rep ...] > profiles/r02_stall_sampling.md

Per report: the first launch's share of warp time per stall reason (source page, all samples) and the 12 SASS instructions
that collected the most samples with their dominant reason."""
import csv
import subprocess
import sys


def one(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:eval_kernel",
                          "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    name = rows[0][1] if rows and len(rows[0]) > 1 else rep
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    # the page is emitted once per view: keep the first copy
    first = data[0][0]
    for i in range(1, len(data)):
        if data[i][0] == first:
            data = data[:i]
            break
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def I(r, k):
        try:
            return int(r[ix[k]])
        except ValueError:
            return 0
    total = sum(I(r, "# Samples") for r in data)
    insts = sum(I(r, "Instructions Executed") for r in data)
    print(f"## `{name}`\n")
    print(f"{total} samples, {insts} warp instructions executed, {len(data)} SASS instructions\n")
    print("| stall reason | share of warp time |\n|---|---:|")
    for s in sorted(stalls, key=lambda s: -sum(I(r, s) for r in data)):
        v = sum(I(r, s) for r in data)
        if v * 200 >= total:
            print(f"| {s[6:]} | {100 * v / total:.1f} % |")
    print("\n| samples | executed | instruction | dominant reasons |\n|---:|---:|---|---|")
    for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:12]:
        top = sorted(((I(r, s), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"| {I(r, '# Samples')} | {I(r, 'Instructions Executed')} | `{r[ix['Source']].strip()[:70]}` | "
              + ", ".join(f"{n} {w}" for n, w in top if n) + " |")
    print()


if __name__ == "__main__":
    print("# ncu warp-state sampling of the hot kernels (source page of the `--set full` captures; cold-cache, serialised launches)\n")
    for rep in sys.argv[1:]:
        one(rep)
