"""Fill the @PLACEHOLDER@ fields of BASELINE.md section 4 from the JSON lines of a validation run.

    python tools/fill_baseline.py <dir> <prefix_1gpu> [N=file ...]
<dir>/<prefix>_bench_n1.json, _bench_ref.json, _families.jsonl are the files tools/gpu_validate.sh writes; the N=file
arguments name multi-GPU bench lines of workload T (e.g. 2=gpurun_out/x_T_n2.json), bee4=file the Beehive line on 4 GPUs."""
import json
import os
import sys


def line(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def main():
    d, pre = sys.argv[1], sys.argv[2]
    extra = dict(a.split("=", 1) for a in sys.argv[3:])
    b = line(os.path.join(d, f"{pre}_bench_n1.json"))
    ref = line(os.path.join(d, f"{pre}_bench_ref.json"))
    r, sup = b["roofline"], b["roofline_supplied_meas"]
    rep = {
        "STEP": f"{b['ms_per_step'] * 1e3:.1f}", "VALUE": f"{b['value']:.2e}",
        "KUS": f"{r['us_per_launch']:.1f}", "KFRAC": f"{r['frac']:.2f}",
        "KOVL": f"{r['us_per_launch_overlapped']:.1f}", "KOFRAC": f"{r['frac_overlapped']:.2f}",
        "SUS": f"{sup['us_per_launch']:.1f}", "SFRAC": f"{sup['frac']:.2f}",
        "E2EMS": f"{b['e2e']['ms_per_step']:.3f}", "E2E": f"{b['e2e']['value']:.2e}",
        "E2ECMS": f"{b['e2e_compact']['ms_per_step']:.3f}", "E2EC": f"{b['e2e_compact']['value']:.2e}",
        "CPU": f"{b['cpu_baseline']['value']:.2e}", "CPUBARE": f"{b['cpu_bare_residual_sweep']['value']:.2e}",
        "REF": f"{ref['value']:.2e} ({ref['cpu_baseline']['cores']} threads, {ref['cpu_baseline']['sample'].split(' in ')[1].split(' ')[0]} s)",
    }
    names = {0: "Pose2Pose2", 2: "Pose2Point2BearingRange", 3: "Pose3Pose3"}
    rows = []
    for l in open(os.path.join(d, f"{pre}_families.jsonl")):
        x = json.loads(l)
        f, s = x["fused_sample_serialized"], x["supplied_meas_serialized"]
        rows.append(f"| {x['config']} | {names[x['family']]} | {x['factors']} x {x['N']} | {f['us_per_launch']:.1f} us, "
                    f"{100 * f['frac_hbm']:.0f} % | {s['us_per_launch']:.1f} us ({100 * s['frac_hbm']:.0f} %) |")
    rep["FAMROWS"] = "\n".join(rows)
    srows = []
    for n in ("2", "4", "8"):
        if n in extra:
            m = line(extra[n])
            eff = m["value"] / (int(n) * b["value"])
            srows.append(f"| {n} | one one-warp flag kernel per step (`rome_b200_peer_barrier`, default) | {m['ms_per_step'] * 1e3:.1f} | "
                         f"{m['value']:.2e} | {eff:.2f} | {str(m.get('exchange_verified')).lower()} ({m.get('rows_checked_all_ranks')} rows) |")
    rep["SCALEROWS"] = "\n".join(srows)
    if "bee4" in extra:
        m = line(extra["bee4"])
        rep["BEE4"] = f"{m['ms_per_step'] * 1e3:.1f} us/step = {m['value']:.2e} evals/s (exchange verified: {str(m.get('exchange_verified')).lower()})"
    else:
        rep["BEE4"] = "29.5 us/step = 1.36e11 evals/s (`profiles/r02_bench_beehive_n4.json`, fused barrier)"
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "BASELINE.md")
    s = open(p).read()
    for k, v in rep.items():
        s = s.replace(f"@{k}@", v)
    open(p, "w").write(s)
    left = [w for w in s.split("@") if w.isupper() and len(w) < 12]
    print("filled; placeholders left:", sorted(set(left)))


if __name__ == "__main__":
    main()
