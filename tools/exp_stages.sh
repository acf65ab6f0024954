#!/bin/bash
# One gpurun call: pipeline depth (ROME_B200_MAX_STAGES) x tile size on the bench workloads (L1 left for local memory).
set -u
P=${1:-exp}
O=gpurun_out
mkdir -p $O
for cfg in "8 6" "8 3" "8 2" "9 3" "9 2"; do
  set -- $cfg
  ROME_B200_TILE=$1 ROME_B200_MAX_STAGES=$2 timeout 200 python bench.py --no-cpu --no-parity --e2e-steps 4 > $O/${P}_bench_T_t$1_s$2.json 2> $O/${P}_bench_T_t$1_s$2.err
done
for S in 6 3 2; do
  ROME_B200_MAX_STAGES=$S timeout 200 python tools/bench_families.py > $O/${P}_families_s$S.jsonl 2> $O/${P}_families_s$S.err
done
python - <<EOF
import json,glob
for f in sorted(glob.glob("$O/${P}_bench_T_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][0]); r=d["roofline"]
        print(f.split("_bench_T_")[1], "ms_per_step", round(d["ms_per_step"]*1e3,2), "us_per_launch", round(r["us_per_launch"],2), "frac", round(r["frac"],3), "overlapped", round(r.get("us_per_launch_overlapped",0),2))
    except Exception as e: print(f, e)
for f in sorted(glob.glob("$O/${P}_families_s*.jsonl")):
    for l in open(f):
        x=json.loads(l); print(f.split("_families_")[1], x["config"][:12], x["family"], "fused_ser", round(x["fused_sample_serialized"]["us_per_launch"],2), round(x["fused_sample_serialized"]["frac_hbm"],3), "fused_ovl", round(x["fused_sample"]["us_per_launch"],2), "supplied_ser", round(x["supplied_meas_serialized"]["us_per_launch"],2))
EOF
