#!/usr/bin/env python
"""bench.py -- factor-particle residual evals/s of the RoME hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.

Workloads (BASELINE.json configs; --workload):
  manhattan_shaped_10k_se2_N100 (default, the configuration the north-star target is quoted on): synthetic
      Manhattan-world SE(2) graph, 10 000 Pose2 PER GPU x N=100 particles, 9 999 odometry + 2 000 loop-closure Pose2Pose2
      per 10 000 poses and one PriorPose2 (rome_b200.workloads.manhattan_arrays, seeds 2 / 1) -- WEAK scaling: with
      --gpus G the graph has G x 10 000 poses.
  beehive_N200  (config 4): Beehive2D, 10 000 Pose2 + lattice landmarks, Pose2Pose2 + Pose2Point2BearingRange +
      PriorPose2, N=200 -- STRONG scaling (the one graph is split over the GPUs).
  se3_chain_10k (config 5): SE(3) helix chain, 10 000 Pose3, 9 999 + 1 000 Pose3Pose3 + PriorPose3, N=100 -- STRONG.

A STEP is one pass of the hot path over the whole graph: for every factor x every particle, getSample (in-kernel Philox)
+ residual + per-factor statistics, one fused kernel launch per factor family.  Steps are DEPENDENT (as Gibbs sweeps
are): the first launch of a step waits for the previous step to complete; the other family kernels of the same sweep
read the same particles and write other buffers, so they carry ROME_B200_INDEPENDENT and overlap it.

--gpus G > 1 (one process per GPU): owner-sharded.  Variables are split into contiguous ranges (bounds chosen so that
every rank evaluates the same number of factors), a factor lives on the owner of its first variable, and only what cut
edges need crosses NVLink: the forward-proposal rows of cut factors are written by the evaluating kernel's own TMA bulk
stores straight into a receive buffer of the target variable's owner (rome_b200_set_proposal_destinations), halo
particle blocks are pushed once (rome_b200_push_halo; in a solve: after every belief update), and a GPU-side flag
barrier over peer memory closes every step (rome_b200_peer_barrier: one one-warp kernel; --barrier selects the others).  After the timed region every rank
recomputes, from the same seeds, the rows its peers should have delivered and compares them bit for bit
(`exchange_verified`).

Timing: W warm-up steps, then EXACTLY K steps replayed from one CUDA graph inside one CUDA-event pair on the launch
stream, bracketed by a barrier + torch.cuda.synchronize(); max over ranks.  L2: the K steps rotate through `sets`
independent copies of the whole working set (> 2x the 126 MB L2), so no step finds its inputs in L2.

`e2e`: the same step through the public host API with HOST buffers every step: pinned Float64 particles in the
reference layout -> rome_b200_set_particles (H2D + layout kernel) -> rome_b200_eval_host_async (kernels + D2H of
residual rows and statistics).  `e2e_compact`: anchored float32 upload (rome_b200_set_particles_anchored) and only the
statistics downloaded.
`cpu_baseline` / `--impl reference`: the float64 C restatement (oracle/) of the SAME step (getSample + residual +
statistics) on all host cores (kind "port": the Julia reference cannot run in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "factor_particle_residual_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "manhattan_shaped_10k_se2_N100"
NPOSES, NPART = 10000, 100
STEP_TEXT = "getSample + residual + per-factor statistics for every factor x particle of every family of the graph"


# --------------------------------------------------------------------------------------------------------
def make_workload(name, G=1):
    """global workload arrays (rome_b200.workloads) + scaling kind + the dominant family"""
    import rome_b200 as rb
    from rome_b200 import workloads as W
    if name == WORKLOAD:
        return W.manhattan_arrays(NPOSES * G, seed=2, N=NPART, particle_seed=1), "weak"
    if name == "beehive_N200":
        fg = rb.generateGraph_Beehive(10000, N=200)
        rb.seed_particles(fg, N=200, seed=3)
        return W.graph_arrays(fg, 200), "strong"
    if name == "se3_chain_10k":
        fg = rb.generateGraph_Pose3Chain(10000, loops=1000)
        rb.seed_particles(fg, N=100, seed=4)
        return W.graph_arrays(fg, 100), "strong"
    raise SystemExit(f"unknown workload {name}")


def build_workload(copies=1):
    """round-1 layout of the default workload (kept for tests/tools): dict(poses, ip, iq, mu, cov, pr_ip, pr_mu, pr_cov)"""
    import rome_b200 as rb
    from rome_b200 import workloads as W
    w = W.manhattan_arrays(NPOSES * copies, seed=2, N=NPART, particle_seed=1)
    p2, pr = w["families"][rb.POSE2POSE2], w["families"][rb.PRIORPOSE2]
    return dict(poses=w["particles"][rb.POSE2], ip=p2["i0"], iq=p2["i1"], mu=p2["a"], cov=p2["b"],
                pr_ip=pr["i0"], pr_mu=pr["a"], pr_cov=pr["b"])


def cholesky_of(fam, f):
    """[nF][dm][dm] lower Cholesky factors of a family's beliefs (BearingRange: diag(sig_b, sig_r)) and the means"""
    import rome_b200 as rb
    if fam == rb.BEARINGRANGE:
        a, b = np.asarray(f["a"]), np.asarray(f["b"])
        Lc = np.zeros((len(a), 2, 2))
        Lc[:, 0, 0], Lc[:, 1, 1] = a[:, 1], b[:, 1]
        return np.column_stack([a[:, 0], b[:, 0]]), Lc
    return np.asarray(f["a"]), np.linalg.cholesky(np.asarray(f["b"]))


class ClockSampler:
    """samples SM clock / throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        for bit, n in names.items():
            if r & bit:
                self.reasons.add(n)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:
                return
            time.sleep(0.0005)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
            self._t = None
            self._stop = threading.Event()

    def summary(self, window):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "window": window}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "window": window}


def ncu_traffic(kernel_substr):
    """DRAM bytes per launch of the dominant kernel from the newest committed ncu summary (profiles/), or None"""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_metrics.csv")), reverse=True):
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        try:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if kernel_substr in r[1]:
                return float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]], os.path.basename(path)
    return None, None


# ---- CPU legs (oracle port; the only places bench.py executes oracle/) ------------------------------------------
FAM_ORACLE = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4}  # rome_b200 family id -> rome_oracle_step family id (same numbering)


def cpu_step_prepare(w, max_factors=None, nthreads=0):
    """the GPU step on the CPU: rome_oracle_step (getSample + residual + statistics) for every family of the workload.
    Returns (step, evals, sample_text); step(seed) -> threads used.  max_factors bounds the sample (first factors of
    every family, proportionally)."""
    from oracle import oracle as O
    import rome_b200 as rb
    O.build()
    total = sum(len(f["i0"]) for f in w["families"].values())
    frac = 1.0 if not max_factors or max_factors >= total else max_factors / total
    calls, evals, parts = [], 0, []
    for fam, f in w["families"].items():
        vt0, vt1 = rb.FAMILY[fam][0], rb.FAMILY[fam][1]
        n = len(f["i0"]) if frac == 1.0 else max(1, int(len(f["i0"]) * frac))
        mu, Lc = cholesky_of(fam, f)
        call, _, _ = O.step_prepare(FAM_ORACLE[fam], f["i0"][:n], None if f["i1"] is None else f["i1"][:n],
                                    w["particles"][vt0], None if vt1 is None else w["particles"][vt1], mu[:n], Lc[:n],
                                    nthreads)
        calls.append(call)
        evals += n * w["N"]
        parts.append(f"{n} of {len(f['i0'])} family-{fam} factors")

    def step(seed=0):
        nt = 1
        for c in calls:
            nt = c(seed)
        return nt

    step(0)  # warm-up: output pages touched, threads started
    return step, evals, " + ".join(parts) + f" x {w['N']} particles"


def cpu_rate(step, evals, seconds):
    reps, t0, nt = 0, time.perf_counter(), 1
    while True:
        nt = step(reps)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return reps * evals / dt, nt, reps, dt


def cpu_bare_sweep_rate(w, seconds=3.0):
    """round-1's figure, kept as a secondary: the bare Pose2Pose2 residual loop on prepared samples (no sampling, no
    statistics)"""
    from oracle import oracle as O
    import rome_b200 as rb
    f = w["families"].get(rb.POSE2POSE2)
    if f is None:
        return None
    rng = np.random.default_rng(5)
    F, N = len(f["i0"]), w["N"]
    mu, Lc = cholesky_of(rb.POSE2POSE2, f)
    meas = O._f64(mu[:, None, :] + np.einsum("fij,fnj->fni", Lc, rng.normal(size=(F, N, 3))))
    ip, iq, poses = O._i32(f["i0"]), O._i32(f["i1"]), O._f64(w["particles"][rb.POSE2])
    res = np.empty((F, N, 3))
    lib = O.lib()
    nt = os.cpu_count() or 1

    def sweep(_=0):
        return lib.rome_oracle_sweep_pose2pose2(F, N, O._ip(ip), O._ip(iq), O._dp(poses), O._dp(meas), O._dp(res), nt)
    sweep()
    v, nt, reps, dt = cpu_rate(sweep, F * N, seconds)
    return {"value": v, "unit": UNIT, "cores": nt, "sample": f"{reps} bare Pose2Pose2 residual sweeps ({F} x {N}) in {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port of the same step) on this box's host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, scaling = make_workload(args.workload, args.gpus)
    nt = os.cpu_count() or 1  # explicit: torch.distributed.run exports OMP_NUM_THREADS=1
    total = sum(len(f["i0"]) for f in w["families"].values())
    step, evals, sample = cpu_step_prepare(w, nthreads=nt)
    t0 = time.perf_counter()
    step(1)
    one = time.perf_counter() - t0
    budget = 100.0  # seconds for warm-up + timed steps
    if one * (args.steps + args.warmup) > budget:  # bound the per-step sample so that the run ends within minutes
        keep = max(256, int(total * budget / (one * (args.steps + args.warmup))))
        step, evals, sample = cpu_step_prepare(w, max_factors=keep, nthreads=nt)
    for k in range(args.warmup):
        step(k)
    n_timed, used = 0, 1
    t0 = time.perf_counter()
    while n_timed < args.steps or time.perf_counter() - t0 < 2.0:  # at least K steps and at least 2 s
        used = step(100 + n_timed)
        n_timed += 1
    dt = time.perf_counter() - t0
    value = n_timed * evals / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / n_timed, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, w, args.gpus, scaling),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": f"{n_timed} steps of [{sample}] in {dt:.2f} s (Xoshiro256++ / polar-method getSample "
                                       f"+ residual + statistics, OpenMP over factors)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "Julia reference cannot run here (no julia, unvendored deps); this is the float64 C restatement of the same step"}
    print(json.dumps(line))


def config_of(name, w, G, scaling):
    import rome_b200 as rb
    fams = {str(fam): len(f["i0"]) for fam, f in w["families"].items()}
    return {"workload": name, "variables": {str(vt): int(p.shape[0]) for vt, p in w["particles"].items()},
            "particles": w["N"], "npad": rb.npad(w["N"]), "factors": fams,
            "evals_per_step": int(sum(fams.values()) * w["N"]), "step": STEP_TEXT,
            "gpus": G, "scaling": scaling}


# --------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, choices=[WORKLOAD, "beehive_N200", "se3_chain_10k"])
    ap.add_argument("--sets", type=int, default=0, help="independent working-set copies rotated through (0: enough for > 2x L2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity check of a sampled subset against the oracle")
    ap.add_argument("--barrier", default=os.environ.get("ROME_B200_BENCH_BARRIER", "flags"),
                    choices=["fused", "flags", "flags2", "nccl", "none"],
                    help="G>1, the rank barrier that closes a step: flags (default) = ONE one-warp kernel that publishes this "
                         "rank's epoch into every peer's flag array over NVLink and polls its own (rome_b200_peer_barrier); "
                         "flags2 = the same as two kernels (peer_signal + peer_wait); fused = carried by the evaluation "
                         "kernels themselves (first launch waits, last launch publishes); nccl = a 4-byte all-reduce")
    ap.add_argument("--e2e-steps", type=int, default=40)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout must carry exactly one JSON line: park fd 1 on stderr while libraries (NCCL prints its version banner to
    # stdout) initialise and run; it is restored just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import rome_b200 as rb
    from rome_b200 import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: librome_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G, multi = world, world > 1
    wg, scaling = make_workload(args.workload, G)       # the GLOBAL graph (every rank builds the same arrays)
    N, Np = wg["N"], rb.npad(wg["N"])
    sh = W.sharding_of(wg, G)
    lv = W.local_view(wg, sh, rank)                     # what this rank holds: owned + halo variables, its factors
    fam_order = sorted(lv["families"], key=lambda f: -len(lv["families"][f]["i0"]))
    dom = max(wg["families"], key=lambda f: len(wg["families"][f]["i0"]) * rb.BYTES_PER_EVAL_SAMPLED[f])
    evals_per_step = sum(len(f["i0"]) for f in wg["families"].values()) * N
    evals_rank = sum(len(f["i0"]) for f in lv["families"].values()) * N
    F0 = rb.SAMPLE | rb.RESIDUAL | rb.STATS

    def set_bytes():
        b = 0
        for vt, p in lv["particles"].items():
            b += p.shape[0] * (48 + Np * rb.VAR_DIM[vt] * 4)
        for fam, f in lv["families"].items():
            b += len(f["i0"]) * (Np * rb.FAMILY[fam][3] * 4 + rb.FAMILY[fam][4] * 4 + 64)
        return b
    S = args.sets or int(min(64, max(12, -(-300e6 // max(set_bytes(), 1)))))
    if multi:  # the same number of working sets on every rank (their handles are exchanged set by set)
        t = torch.tensor([S], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        S = int(t.item())
    stream = torch.cuda.Stream()

    # ---- working sets ---------------------------------------------------------------------------------------------
    sets = []
    with torch.cuda.stream(stream):
        for s in range(S):
            c = rb.Context(local)
            c.use_torch_stream()
            for vt, p in lv["particles"].items():
                q = p + 1e-4 * s
                n_own = lv["loc"]["own"][vt][1] - lv["loc"]["own"][vt][0]
                q[n_own:] = 0.0   # halo slots: filled by their owners' rome_b200_push_halo
                c.set_particles(vt, q)
            bufs = {}
            for fam, f in lv["families"].items():
                W.upload_family(c, fam, f["i0"], f["i1"], f["a"], f["b"])
                nF, dr, ns, dfwd = len(f["i0"]), rb.FAMILY[fam][3], rb.FAMILY[fam][4], rb.FAMILY[fam][6]
                bufs[fam] = dict(res=torch.zeros((nF, Np, dr), device="cuda"), stats=torch.zeros((nF, ns), device="cuda"))
                if multi and dfwd:
                    bufs[fam]["recv"] = c.malloc_device(max(1, len(f["recv"])) * Np * dfwd * 4)
            sets.append((c, bufs))
        stream.synchronize()

    # ---- owner-sharded exchange plumbing: CUDA IPC handles of every receive buffer and particle store --------------
    state = peer_slots = None
    dummy_fwd = {}
    if multi:
        c0 = sets[0][0]
        mine = []
        for c, bufs in sets:
            mine.append({"recv": {fam: c.ipc_export(b["recv"]) for fam, b in bufs.items() if "recv" in b},
                         "store": {vt: c.ipc_export(c.particles_device(vt)[0]) for vt in lv["particles"]}})
        state = c0.peer_state_alloc()
        everyone = [None] * G
        dist.all_gather_object(everyone, {"sets": mine, "state": c0.ipc_export(state)})
        peer_slots = []
        for p in range(G):
            if p != rank:
                peer_slots.append(c0.ipc_import(everyone[p]["state"]) + 4 * (rank if rank < p else rank - 1))
        for si, (c, bufs) in enumerate(sets):
            recv_ptr = {p: {fam: c.ipc_import(h) for fam, h in everyone[p]["sets"][si]["recv"].items()}
                        for p in range(G) if p != rank}
            store_ptr = {p: {vt: c.ipc_import(h) for vt, h in everyone[p]["sets"][si]["store"].items()}
                         for p in range(G) if p != rank}
            for fam, f in lv["families"].items():
                dfwd = rb.FAMILY[fam][6]
                if not dfwd or f["n_cut"] == 0:
                    continue
                rowb = Np * dfwd * 4
                ptrs = [0] * f["cut_first"] + [recv_ptr[int(d)][fam] + int(r) * rowb
                                               for d, r in zip(f["dst_rank"], f["dst_row"])]
                ptrs += [0] * (len(f["i0"]) - len(ptrs))
                c.set_proposal_destinations(fam, 0, ptrs)
                c.set_barrier_range(fam, f["cut_first"], f["n_cut"])   # the rank barrier is only passed before the cut block
                if fam not in dummy_fwd:  # default rows of prop_fwd are never written for the cut range; one shared buffer
                    dummy_fwd[fam] = torch.zeros((len(f["i0"]), Np, dfwd), device="cuda")
            for vt, pushes in lv["loc"]["push"].items():
                src, dst = [], []
                bb = c.particles_device(vt)[1]
                for reader, local_vars, slots in pushes:
                    src += [int(v) for v in local_vars]
                    dst += [store_ptr[reader][vt] + int(sl) * bb for sl in slots]
                if src:
                    c.set_halo_plan(vt, src, dst)
            c.set_step_barrier(state, peer_slots)
        dist.barrier()
        token = torch.zeros(1, device="cuda")

    def rank_barrier(c):
        """stream-ordered barrier between the ranks (after everything enqueued so far on the launch stream)"""
        if args.barrier == "flags":
            c.peer_barrier(state, peer_slots)
        elif args.barrier in ("flags2", "fused", "none"):
            c.peer_signal(state, peer_slots)
            c.peer_wait(state, G - 1)
        else:
            dist.all_reduce(token)

    halo_checked = 0
    if multi:
        with torch.cuda.stream(stream):
            for c, _ in sets:
                c.use_torch_stream()
                for vt in lv["particles"]:
                    c.push_halo(vt)
            rank_barrier(sets[0][0])
            stream.synchronize()
        dist.barrier()
        # the halo blocks that arrived are, byte for byte, the blocks their owners hold
        halo_ok = True
        with torch.cuda.stream(stream):
            for si in (0, S - 1):
                c = sets[si][0]
                for vt, p in wg["particles"].items():
                    hal = lv["loc"]["halo"][vt]
                    if len(hal) == 0:
                        continue
                    chk = rb.Context(local)
                    chk.use_torch_stream()
                    chk.set_particles(vt, p[hal] + 1e-4 * si)
                    ptr, bb, _, nv, _, _ = c.particles_device(vt)
                    n_own = lv["loc"]["own"][vt][1] - lv["loc"]["own"][vt][0]
                    got = np.empty(len(hal) * bb, np.uint8)
                    want = np.empty(len(hal) * bb, np.uint8)
                    c.memcpy_d2h(got, ptr + n_own * bb)
                    chk.memcpy_d2h(want, chk.particles_device(vt)[0])
                    halo_ok = halo_ok and bool(np.array_equal(got, want))
                    halo_checked += len(hal)
                    chk.close()

    # ---- the step ---------------------------------------------------------------------------------------------------
    def step(k, fams=None, barrier=True, indep_all=False):
        """one pass of the hot path over this rank's factors: ONE launch per family.  The first launch is dependent (it
        waits for the previous step), the others read the same particles and write other buffers: INDEPENDENT.  G > 1:
        families with cut factors run with PROPOSAL_FWD | ROUTED_ONLY -- only the cut factors produce forward rows, and
        those go straight into the owners' receive buffers; the rank barrier rides on the step's first / last launch."""
        c, bufs = sets[k % S]
        todo = [fam for fam in (fams or fam_order) if len(lv["families"][fam]["i0"]) > 0]
        is_routed = [bool(multi and rb.FAMILY[fam][6] and lv["families"][fam]["n_cut"] > 0) for fam in todo]
        # every launch with cut factors passes the barrier before its cut tiles; the LAST of them publishes this rank's
        # epoch when its grid has finished (its epilogue runs after the launches before it have completed, so none of
        # them is still polling; launches behind it touch nothing a peer can see).  A rank without cut factors signals
        # from its first launch.
        signal_idx = max([i for i, r in enumerate(is_routed) if r], default=0)
        for idx, fam in enumerate(todo):
            f, b = lv["families"][fam], bufs[fam]
            fl = F0
            kw = dict(res=b["res"], stats=b["stats"])
            if is_routed[idx]:
                fl |= rb.PROPOSAL_FWD | rb.ROUTED_ONLY
                kw["prop_fwd"] = dummy_fwd[fam]  # default rows are never written: every cut factor has a destination
            if idx > 0 or indep_all:
                fl |= rb.INDEPENDENT
            if multi and barrier and args.barrier == "fused":
                fl |= (rb.BARRIER_WAIT if is_routed[idx] else 0) | (rb.BARRIER_SIGNAL if idx == signal_idx else 0)
            c.eval(fam, fl, seed=7, stream_id=k, **kw)
        if multi and barrier and args.barrier in ("flags", "flags2", "nccl"):   # "none": experiment only -- steps of the ranks uncoupled
            rank_barrier(c)

    def capture(fn, **kw):
        gr = torch.cuda.CUDAGraph()
        before = sum(c.launch_count for c, _ in sets)
        with torch.cuda.graph(gr, stream=stream):
            for c, _ in sets:
                c.use_torch_stream()
            for k in range(args.steps):
                fn(args.warmup + k, **kw)
        launches = sum(c.launch_count for c, _ in sets) - before
        for c, _ in sets:
            c.use_torch_stream()
        if multi:
            dist.barrier()  # replays carry the ranks' device-side barrier: start them together
        gr.replay()  # untimed replay: graph upload + warm instruction caches
        stream.synchronize()
        return gr, launches

    def timed(gr, sample_clocks=False):
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sample_clocks:
            clocks.start()
        e0.record(stream)
        gr.replay()
        e1.record(stream)
        stream.synchronize()
        if sample_clocks:
            clocks.stop()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def allmax(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)
    if multi:
        dist.barrier()   # the ranks enter the first fused-barrier step together (the device-side wait gives up after ~2 s)
    with torch.cuda.stream(stream):
        for c, _ in sets:
            c.use_torch_stream()
        for k in range(args.warmup):
            step(k)
        stream.synchronize()
        if multi:
            dist.barrier()
        g, launches_per_replay = capture(step)
        ms = timed(g, sample_clocks=True)
        window = "timed"
        # timed region too short to sample the clocks: keep sampling under identical replays.  Whether and how often is
        # decided COLLECTIVELY (every replay carries the ranks' barrier: same replay count on every rank)
        need = torch.tensor([1.0 if len(clocks.samples) < 5 else 0.0, ms], device="cuda", dtype=torch.float64)
        if multi:
            dist.all_reduce(need, op=dist.ReduceOp.MAX)
        if need[0].item() > 0:
            reps = max(1, min(200, int(300.0 / max(need[1].item(), 1e-3))))
            clocks.start()
            for _ in range(reps):
                g.replay()
                stream.synchronize()
            clocks.stop()
            window = "timed+identical replays"
    ms = allmax(ms)
    value = evals_per_step * args.steps / (ms * 1e-3)

    # ---- exchange verification (untimed): recompute what every peer should have delivered ---------------------------
    verify = None
    if multi:
        k_last = args.warmup + args.steps - 1
        si = k_last % S
        c, bufs = sets[si]
        ok, rows_checked, worst = True, 0, 0.0
        with torch.cuda.stream(stream):
            for fam, f in lv["families"].items():
                dfwd = rb.FAMILY[fam][6]
                if not dfwd or len(f["recv"]) == 0:
                    continue
                got = np.empty((len(f["recv"]), Np, dfwd), np.float32)
                c.memcpy_d2h(got, bufs[fam]["recv"])
                pos = 0
                F = sh.families[fam]
                for src in range(G):
                    if src == rank:
                        continue
                    rows = f["recv"][F["rank"][f["recv"]] == src]
                    if len(rows) == 0:
                        continue
                    lsrc = W.local_view(wg, sh, src, fill_halo=True)   # the source rank's local problem, rebuilt here
                    chk = rb.Context(local)
                    chk.use_torch_stream()
                    for vt, p in lsrc["particles"].items():
                        q = p.copy()
                        n_own = lsrc["loc"]["own"][vt][1] - lsrc["loc"]["own"][vt][0]
                        q[:n_own] += 1e-4 * si
                        # halo slots of the source hold what their owners pushed: the owners' own values of set si
                        q[n_own:] += 1e-4 * si
                        chk.set_particles(vt, q)
                    fs = lsrc["families"][fam]
                    W.upload_family(chk, fam, fs["i0"], fs["i1"], fs["a"], fs["b"])
                    nFs = len(fs["i0"])
                    out = dict(res=torch.zeros((nFs, Np, rb.FAMILY[fam][3]), device="cuda"),
                               stats=torch.zeros((nFs, rb.FAMILY[fam][4]), device="cuda"),
                               prop_fwd=torch.zeros((nFs, Np, dfwd), device="cuda"))
                    # the same launch the source rank issued (same kernel variant, same factor numbering), its cut rows
                    # routed into a local buffer instead of the peers' memory
                    routed = torch.zeros((max(1, fs["n_cut"]), Np, dfwd), device="cuda")
                    rowb = Np * dfwd * 4
                    lp = [0] * fs["cut_first"] + [routed.data_ptr() + j * rowb for j in range(fs["n_cut"])]
                    lp += [0] * (nFs - len(lp))
                    chk.set_proposal_destinations(fam, 0, lp)
                    chk.eval(fam, F0 | rb.PROPOSAL_FWD | rb.ROUTED_ONLY, seed=7, stream_id=k_last, **out)
                    stream.synchronize()
                    mine_rows = np.nonzero(fs["dst_rank"] == rank)[0]
                    want = routed[torch.as_tensor(mine_rows, device="cuda")].cpu().numpy()
                    # rows arrive in (source rank, global id) order; the source's cut factors to one destination are
                    # sorted by global id too
                    seg = got[pos:pos + len(rows)]
                    pos += len(rows)
                    same = seg.shape == want.shape and np.array_equal(seg[:, :N], want[:, :N])
                    ok = ok and same
                    if not same:
                        bad = np.inf
                        if seg.shape == want.shape:
                            dd = np.abs(seg[:, :N].astype(np.float64) - want[:, :N])
                            bad = float(np.nanmax(dd)) if np.isfinite(dd).any() else np.inf
                            print(f"[rank {rank}] exchange mismatch from rank {src}: {int((dd > 0).any(axis=(1, 2)).sum())} of "
                                  f"{len(rows)} rows differ, {int(np.isnan(seg).sum())} NaN received, worst {bad}", file=sys.stderr)
                        worst = max(worst, bad)
                    rows_checked += len(rows)
                    chk.close()
        gave_up = bool(sets[0][0].peer_gave_up(state)) if args.barrier != "nccl" else False
        words = np.zeros(16, np.uint32)
        sets[0][0].memcpy_d2h(words, state)
        print(f"[rank {rank}] barrier state: signals {words[8]}, passes {words[14]}, mean cycles per pass "
              f"{words[13] / max(1, words[14]):.0f}, publish cycles per signal {words[15] / max(1, words[8]):.0f} "
              f"(wrapping 32-bit sums)", file=sys.stderr)
        flag = torch.tensor([1.0 if (ok and halo_ok and not gave_up) else 0.0, float(rows_checked), worst, float(halo_checked)],
                            device="cuda", dtype=torch.float64)
        mn = flag.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm = flag.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        mx = flag.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        gu = torch.tensor([1.0 if gave_up else 0.0, 0.0 if halo_ok else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(gu, op=dist.ReduceOp.MAX)
        verify = {"exchange_verified": bool(mn[0].item() > 0.5), "barrier_gave_up": bool(gu[0].item() > 0),
                  "halo_mismatch": bool(gu[1].item() > 0), "rows_checked_all_ranks": int(sm[1].item()),
                  "max_abs_diff": float(mx[2].item()), "halo_blocks_checked_all_ranks": int(sm[3].item()),
                  "how": "after the timed replay every rank rebuilds each peer's local problem from the seeds, recomputes the "
                         "forward-proposal rows that peer's cut factors address to it (same seed / stream id as the last timed "
                         "step on that working set) and compares them bit for bit with its receive buffer; halo particle "
                         "blocks are compared byte for byte with a fresh pack of their owners' particles"}

    # ---- the dominant kernel alone (roofline): K dependent launches, live CUDA-event timing ------------------------
    def kernel_only(k, indep=False):
        step(k, fams=[dom], barrier=False, indep_all=indep)

    with torch.cuda.stream(stream):
        gk, _ = capture(kernel_only)
        kms = allmax(timed(gk)) / args.steps
        gko, _ = capture(kernel_only, indep=True)
        kms_over = allmax(timed(gko)) / args.steps
        pms = None
        if not multi and dom == rb.POSE2POSE2:  # same kernel with the measurement supplied from HBM (48 B/eval)
            nFd = len(lv["families"][dom]["i0"])
            meas_sets = [torch.randn((nFd, Np, 3), device="cuda") * 0.05 for _ in range(S)]

            def kernel_supplied(k):
                c, bufs = sets[k % S]
                c.eval(dom, rb.RESIDUAL | rb.STATS, meas=meas_sets[k % S], res=bufs[dom]["res"], stats=bufs[dom]["stats"])
            gp, _ = capture(kernel_supplied)
            pms = timed(gp) / args.steps
            del meas_sets, gp

    # ---- e2e through the host API (per rank, host buffers, copies inside the timed region) ----
    # Two contexts on their own streams alternate, so the H2D upload of step k+1 overlaps the D2H download of step k
    # (PCIe is full duplex); every step uploads this rank's particles and downloads its residual rows + statistics.
    lanes = []
    host_parts = W.local_view(wg, sh, rank, fill_halo=True)["particles"]  # the host holds its halo variables' values too
    for j in range(2):
        c = sets[j][0]
        c.set_stream(None)
        ln = dict(c=c, parts={vt: torch.from_numpy(p + 1e-4 * j).pin_memory() for vt, p in host_parts.items()}, out={})
        ln["anch"] = {vt: torch.from_numpy(np.ascontiguousarray(p[:, 0, :])).pin_memory() for vt, p in host_parts.items()}
        ln["offs"] = {vt: torch.from_numpy((p - p[:, :1, :]).astype(np.float32)).pin_memory() for vt, p in host_parts.items()}
        for fam, f in lv["families"].items():
            nF = max(1, len(f["i0"]))
            ln["out"][fam] = dict(res=torch.zeros((nF, Np, rb.FAMILY[fam][3]), dtype=torch.float32).pin_memory(),
                                  stats=torch.zeros((nF, rb.FAMILY[fam][4]), dtype=torch.float32).pin_memory())
        lanes.append(ln)

    def e2e_step(k, compact=False):
        ln = lanes[k % 2]
        c = ln["c"]
        c.synchronize()  # results of this lane's previous step (two steps ago) are complete and readable
        for vt in ln["parts"]:
            if compact:
                c.set_particles_anchored(vt, ln["anch"][vt], ln["offs"][vt])
            else:
                c.set_particles(vt, ln["parts"][vt])
        for fam in fam_order:
            o = ln["out"][fam]
            if len(lv["families"][fam]["i0"]) == 0:
                continue
            if compact:
                c.eval_host(fam, rb.SAMPLE | rb.STATS, seed=9, stream_id=k, stats=o["stats"], sync=False)
            else:
                c.eval_host(fam, F0, seed=9, stream_id=k, res=o["res"], stats=o["stats"], sync=False)

    def e2e_run(compact):
        for k in range(4):
            e2e_step(k, compact)
        for ln in lanes:
            ln["c"].synchronize()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            e2e_step(k, compact)
        for ln in lanes:
            ln["c"].synchronize()
        torch.cuda.synchronize()
        return allmax(time.perf_counter() - t0)

    # destinations off for the host-API legs (no exchange there: the device-resident exchange is not a host-buffer path)
    if multi:
        for c, _ in sets[:2]:
            for fam in lv["families"]:
                if rb.FAMILY[fam][6]:
                    c.set_proposal_destinations(fam, 0, [])
    edt = e2e_run(False)
    assert all(float(ln["out"][dom]["stats"].abs().sum()) > 0 and float(ln["out"][dom]["res"].abs().sum()) > 0 for ln in lanes)
    edt_c = e2e_run(True)
    assert all(float(ln["out"][dom]["stats"].abs().sum()) > 0 for ln in lanes)
    h2d = sum(p.shape[0] * N * rb.VAR_DIM[vt] * 8 for vt, p in lv["particles"].items())
    d2h = sum(len(f["i0"]) * (Np * rb.FAMILY[fam][3] + rb.FAMILY[fam][4]) * 4 for fam, f in lv["families"].items())
    h2d_c = sum(p.shape[0] * (N * 4 + 8) * rb.VAR_DIM[vt] for vt, p in lv["particles"].items())
    d2h_c = sum(len(f["i0"]) * rb.FAMILY[fam][4] * 4 for fam, f in lv["families"].items())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        names = {rb.POSE2POSE2: "FamPose2Pose2", rb.BEARINGRANGE: "FamBearingRange", rb.POSE3POSE3: "FamPose3Pose3"}
        traffic, traffic_src = ncu_traffic(names.get(dom, "eval_kernel"))
        bpe = rb.BYTES_PER_EVAL_SAMPLED[dom]
        fd = lv["families"][dom]
        n_dom = fd["n_interior"] + (fd["n_cut"] if not (multi and rb.FAMILY[dom][6]) else 0)
        n_dom_cut = len(fd["i0"]) - n_dom   # cut factors also write a forward-proposal row (+4 dfwd bytes per eval)
        alg_bytes = (n_dom * bpe + n_dom_cut * (bpe + 4 * rb.FAMILY[dom][6])) * N
        ach = alg_bytes / (kms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64 per factor / f32 per particle arithmetic on anchored float32 I/O (Float64 anchor + float32 offset)",
            "data": "synthetic",
            "config": dict(config_of(args.workload, wg, G, scaling),
                           sets=S, l2=f"{S} rotating working-set copies (> 2x L2); K steps replayed from one CUDA graph",
                           steps_are="dependent (each step's first launch waits for the previous step, like Gibbs sweeps); "
                                     "only the family kernels WITHIN a step overlap",
                           factors_this_rank={str(f): len(v["i0"]) for f, v in lv["families"].items()},
                           parallelism=(f"owner-sharded x{G}: variables in contiguous ranges, factors on the owner of their "
                                        f"first variable; rank 0 holds {sum(v['n_cut'] for v in lv['families'].values())} cut "
                                        f"factors; exchange = cut factors' proposal rows by the kernel's own TMA stores into "
                                        f"the owners' receive buffers + {args.barrier} barrier per step") if multi else "single GPU"),
            "roofline": {"bound": "hbm", "kernel": f"eval_kernel<{names.get(dom, dom)}, sample=true>", "achieved": ach,
                         "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "traffic_note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/{traffic_src}; "
                                         "below the algorithmic bytes: neighbouring factors share particle blocks in L2 and "
                                         "written rows are still L2-resident when the launch ends",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "bytes_per_eval": bpe, "evals_per_launch": len(fd["i0"]) * N if not multi else (n_dom + n_dom_cut) * N,
                         "us_per_launch": kms * 1e3, "launches": "dependent (serialized): every launch waits for its predecessor",
                         "us_per_launch_overlapped": kms_over * 1e3, "frac_overlapped": alg_bytes / (kms_over * 1e-3) / 1e9 / peak},
            "e2e": {"value": evals_per_step * args.e2e_steps / edt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * edt / args.e2e_steps, "steps": args.e2e_steps,
                    "api": "rome_b200_set_particles(pinned host f64, reference layout) + rome_b200_eval_host_async(pinned host "
                           "f32 residual rows + statistics), two contexts alternating so upload and download overlap"
                           + ("; bytes are per rank" if multi else "")},
            "e2e_compact": {"value": evals_per_step * args.e2e_steps / edt_c, "unit": UNIT, "h2d_bytes_per_step": h2d_c,
                            "d2h_bytes_per_step": d2h_c, "ms_per_step": 1e3 * edt_c / args.e2e_steps,
                            "api": "rome_b200_set_particles_anchored(f64 anchors + f32 offsets) + rome_b200_eval_host_async("
                                   "SAMPLE|STATS): same evaluations, only the per-factor statistics come back"},
            "gpu_launches": launches_per_replay,
            "clocks": clocks.summary(window),
        }
        if pms is not None:
            pach = len(fd["i0"]) * N * rb.BYTES_PER_EVAL[dom] / (pms * 1e-3) / 1e9
            line["roofline_supplied_meas"] = {"kernel": f"eval_kernel<{names.get(dom, dom)}, sample=false>", "achieved": pach,
                                              "peak": peak, "unit": "GB/s", "frac": pach / peak,
                                              "bytes_per_eval": rb.BYTES_PER_EVAL[dom], "us_per_launch": pms * 1e3,
                                              "launches": "dependent (serialized)"}
        if verify:
            line.update(verify)
        if not args.no_parity:
            try:
                line["parity"] = parity_check(rb, sets[2 % S][0], lv, wg, sh, rank, dom, N)
            except Exception as e:  # noqa: BLE001
                line["parity"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_cpu and G == 1:
            stepc, evc, sample = cpu_step_prepare(wg, nthreads=os.cpu_count() or 1)
            v, nt, reps, dt = cpu_rate(stepc, evc, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": nt, "kind": "port",
                                    "sample": f"{reps} steps of [{sample}] in {dt:.1f} s (the same step: getSample + residual + "
                                              f"statistics; float64 C port of the reference arithmetic, OpenMP)"}
            try:
                line["cpu_bare_residual_sweep"] = cpu_bare_sweep_rate(wg)
            except Exception as e:  # noqa: BLE001
                line["cpu_bare_residual_sweep"] = {"error": f"{type(e).__name__}: {e}"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if multi:
        # CUDA graphs must die before the process group; then leave without the (occasionally hanging) communicator
        # teardown -- every rank has finished its work at the barrier.
        del g, gk, gko
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def parity_check(rb, c, lv, wg, sh, rank, dom, N, nfac=256):
    """in-run parity of a sampled subset (oracle as the CHECKER, untimed): the dominant family's first `nfac` local
    factors, fused getSample with the samples written back, residuals against the float64 oracle on the caller's
    original Float64 particles: |gpu - ref| <= 1e-5 max(|ref|, 0.1)"""
    from oracle import oracle as O
    from rome_b200 import workloads as W
    full = W.local_view(wg, sh, rank, fill_halo=True)
    c.set_stream(None)
    for vt, p in full["particles"].items():
        c.set_particles(vt, p)
    f = full["families"][dom]
    m = min(nfac, len(f["i0"]))
    fl = rb.SAMPLE | rb.WRITE_MEAS | rb.RESIDUAL
    out = c.alloc_host_outputs(dom, fl)
    c.set_proposal_destinations(dom, 0, []) if rb.FAMILY[dom][6] else None
    c.eval_host(dom, fl, seed=11, stream_id=3, first=0, count=m, **out)
    mu, _ = cholesky_of(dom, f)
    meas = rb.offsets_to_meas(out["meas_out"][:m], mu[:m], N)
    res = rb.rows_to_particle_major(out["res"][:m], N)
    vt0, vt1 = rb.FAMILY[dom][0], rb.FAMILY[dom][1]
    P0 = full["particles"][vt0]
    if dom == rb.POSE2POSE2:
        ref, ang = O.sweep_pose2pose2(f["i0"][:m], f["i1"][:m], P0, meas), (2,)
    elif dom == rb.BEARINGRANGE:
        ref, ang = O.sweep_bearingrange(f["i0"][:m], f["i1"][:m], P0, full["particles"][vt1], meas), (0,)
    elif dom == rb.POSE3POSE3:
        ref, ang = O.sweep_pose3pose3(f["i0"][:m], f["i1"][:m], P0, meas), ()
    else:
        return {"skipped": f"no oracle sweep wired for family {dom}"}
    d = res - ref
    for a in ang:
        d[..., a] = O.np_wrap(d[..., a])
    rel = float((np.abs(d) / np.maximum(np.abs(ref), 0.1)).max())
    return {"checked_evals": int(m * N), "family": int(dom), "max_abs_err": float(np.abs(d).max()),
            "max_rel_err_floor_0.1": rel, "ok": bool(rel < 1e-5), "checker": "oracle (float64 C port), original Float64 inputs"}


if __name__ == "__main__":
    main()
