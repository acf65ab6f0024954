"""Control-flow dry run of bench.py's GPU arm on the CPU: torch.cuda and rome_b200.Context are replaced by inert
stand-ins (tests/bench_dryrun_runner.py), so `main()` walks its whole path -- working-set construction, CUDA-graph
capture and timing calls, clock sampling, the N>1 exchange set-up and collectives (gloo, two ranks), e2e lanes, CPU legs,
JSON assembly -- and must print exactly one JSON line carrying every key of the bench contract.  No arithmetic is
checked here (there is no device); the point is that the script itself cannot fail on a name, a shape, a missing key or
a mismatched collective when the driver runs it on a GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "bench_dryrun_runner.py")
COMMON = ["--steps", "6", "--warmup", "3", "--sets", "2", "--e2e-steps", "3"]   # (the parity leg records an error: no device)


def _check_line(stdout, n_gpus):
    lines = [ln for ln in stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and stdout.strip() == lines[0], stdout   # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["metric"] == "factor_particle_residual_evals_per_sec" and d["n_gpus"] == n_gpus and d["steps"] == 6
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["config"]["workload"] == "manhattan_shaped_10k_se2_N100" and "model" not in d["config"]
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "hbm"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    # bytes are per rank: its owned variables (+ halo variables when the graph is split)
    assert d["e2e"]["h2d_bytes_per_step"] >= 9000 * 100 * 3 * 8 and d["e2e_compact"]["h2d_bytes_per_step"] < d["e2e"]["h2d_bytes_per_step"]
    # value = evals of ALL ranks per step * K / elapsed: 12 000 factors x 100 particles per rank, 6 steps in the stub's 9 ms
    assert d["config"]["evals_per_step"] == n_gpus * 12000 * 100
    assert abs(d["value"] - n_gpus * 12000 * 100 * 6 / 9e-3) < 1e-3 * d["value"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    return d


@pytest.mark.parametrize("extra", [["--no-cpu"], ["--cpu-seconds", "0.2"]])
def test_bench_single_gpu_dry_run(extra):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, RUNNER] + COMMON + extra, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check_line(r.stdout, 1)
    assert d["gpu_launches"] == 12
    if "--no-cpu" in extra:
        assert "cpu_baseline" not in d
    else:
        assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
        assert "error" in d["parity"]   # the in-run parity leg has no device here: recorded as failed, the line survives


@pytest.mark.parametrize("extra", [[], ["--barrier", "flags2"], ["--barrier", "fused"], ["--barrier", "nccl"]])
def test_bench_two_rank_dry_run(extra):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, RUNNER, "--gpus", "2"] + COMMON + extra, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (_, err) in zip(procs, outs):
        assert p.returncode == 0, err[-2000:]
    d = _check_line(outs[0][0], 2)
    assert outs[1][0].strip() == ""   # only rank 0 prints
    # rank 0, per step: Pose2Pose2 (interior + routed cut factors in one launch), PriorPose2, and the closing barrier: ONE
    # one-warp kernel by default (rome_b200_peer_barrier), two with flags2 (signal + wait), none with the barrier fused
    # into the evaluation launches or carried by NCCL
    assert d["gpu_launches"] == {"flags2": 24, "fused": 12, "nccl": 12}.get(extra[-1] if extra else "", 18)
    assert d["exchange_verified"] is True and d["rows_checked_all_ranks"] > 0 and d["halo_blocks_checked_all_ranks"] > 0
    assert "cpu_baseline" not in d    # rank 0 at N=1 only
