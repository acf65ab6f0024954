"""The factor-family kernel SOURCES run on the CPU (tests/host_kernels: host build of the source text of csrc/fam_*.cu,
se3_common.cuh and the pack / unpack kernels behind a 32-lane warp emulator) through the very parity tests the GPU
suite runs -- same inputs, same oracle, same tolerances -- by handing those test functions an emulated context.
What this covers without a GPU: the families' arithmetic with its warp-uniform fast-path votes, masking of padding
lanes, the halving-butterfly statistics, closed-form proposals, Jacobians, the in-kernel sampler's counters, the
particle-store layout kernels and every particle-count regime (N = 1 ... 2000).  What only the GPU suite covers: the
TMA/mbarrier pipeline, streams and graphs, the product kernel, the C ABI's argument checking, MUFU approximations."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "host_kernels"))


@pytest.fixture(scope="module")
def hk_so(tmp_path_factory):
    """one host build of the kernel sources per test session (about half a minute of g++)"""
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import build as hk_build
    return hk_build.build(str(tmp_path_factory.mktemp("host_kernels")))


@pytest.fixture(scope="module")
def ectx(hk_so):
    from emu import EmulatedContext
    return EmulatedContext(hk_so)


def _raw():
    import test_gpu_parity_raw as T
    return T


@pytest.mark.parametrize("N", [100, 37, 200])
def test_pose2pose2(ectx, N):
    _raw().test_pose2pose2_parity(ectx, N)


def test_priorpose2_bearingrange_priorpose3(ectx):
    T = _raw()
    T.test_priorpose2_parity(ectx)
    T.test_bearingrange_parity(ectx)
    T.test_priorpose3_parity(ectx)


@pytest.mark.parametrize("N", [100, 64])
def test_pose3pose3(ectx, N):
    _raw().test_pose3pose3_parity(ectx, N)


def test_layout_roundtrip_nan_and_wide_headings(ectx):
    T = _raw()
    T.test_particle_roundtrip(ectx)
    T.test_nan_propagates(ectx)
    T.test_wide_heading_spread_takes_general_sincos_path(ectx)


@pytest.mark.parametrize("family", ["pose2pose2", "priorpose2", "bearingrange", "pose3pose3", "priorpose3"])
def test_fused_sampling(ectx, family):
    _raw().test_fused_sampling_matches_supplied_and_host_twin(ectx, family)


@pytest.mark.parametrize("N", [1, 8, 33, 129, 500, 2000])
def test_particle_counts(ectx, N):
    _raw().test_particle_count_edge_cases(ectx, N)


# ---- the remaining families and the reference-named host API, on the same emulated device ---------------------------
@pytest.fixture()
def emulated_api(ectx, monkeypatch):
    """route every context the host API creates to the emulated device (Context(0), DeviceGraph's default, default_context)"""
    import rome_b200 as rb
    from rome_b200 import graph as G
    make = lambda device=0: ectx  # noqa: E731 -- one shared emulated device, like default_context()
    monkeypatch.setattr(rb, "Context", make)
    monkeypatch.setattr(G, "Context", make)
    monkeypatch.setattr(G, "default_context", make)
    monkeypatch.setattr(rb, "default_context", make)
    return ectx


@pytest.mark.parametrize("N", [100, 37])
def test_next_families_2d(ectx, N):
    import test_gpu_next_families as T
    T.test_point2_gaussian_families(ectx, N)
    T.test_scalar_families(ectx, N)


@pytest.mark.parametrize("N", [100, 37])
def test_next_families_3d(ectx, N, golden_dir):
    import test_gpu_next_families_3d as T
    T.test_point3_families(ectx, N)
    T.test_pose3_partial_families(ectx, N)
    T.test_pose3_ternary_families(ectx, N)
    T.test_known_answers_3d(ectx, golden_dir)


def test_ternary_families_through_the_graph_api(emulated_api):
    import test_gpu_next_families_3d as T
    T.test_rotoffset_known_answer_and_graph_api(emulated_api)


def test_next_families_through_the_graph_api(emulated_api, golden_dir):
    import test_gpu_next_families as T
    T.test_bearing_known_answers(golden_dir)
    T.test_graph_api_with_point_factors()


def test_deconvolution(emulated_api):
    import test_gpu_deconv as T
    T.test_deconv_pose2_families(emulated_api)
    T.test_deconv_pose3_families(emulated_api)
    T.test_approxdeconv_reproduces_the_measurement_belief(emulated_api)


def test_reference_known_answers_through_the_host_api(emulated_api, golden_dir):
    import json
    import test_gpu_graph_api as T
    ka = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    T.test_known_answers_pose2pose2(ka)
    T.test_known_answers_bearingrange(ka)
    T.test_known_answers_pose3pose3(ka)
    T.test_priors_zero_at_measurement()


def test_canonical_graphs_and_fixtures(emulated_api, golden_dir):
    import test_gpu_graph_api as T
    T.test_hexagonal_graph_all_families()
    T.test_manhattan500_fixture_parity(golden_dir)
    T.test_manhattan_g2o_full_graph(golden_dir)
    T.test_approxconv_and_samplefactor()
    T.test_beehive_and_pose3_chain_parity()


def test_parametric_solves(emulated_api):
    import test_gpu_parametric as T
    T.test_square_loop(emulated_api)
    T.test_information_weighting(emulated_api)
    T.test_bearing_range_triangulation(emulated_api)
    T.test_pose3_loop(emulated_api)
    T.test_analytic_and_finite_difference_jacobians_give_the_same_solution(emulated_api)


def test_accumulated_factor_means(emulated_api):
    import test_gpu_zz_accumulate as T
    T.test_accumulate_factor_means_reference_case()
    T.test_accumulate_along_hexagon_and_back()
    for N in (1, 5, 17):
        T.test_statistics_with_fewer_particles_than_lanes(N, nF=250)


def test_full_size_baseline_workloads(emulated_api):
    """BASELINE configs at FULL size on the emulated device (the oracle sweeps them in milliseconds, the emulator in
    seconds): T = 10 000 Pose2 / 11 999 Pose2Pose2 + prior, N = 100 (the bench workload); C4 = Beehive-shaped 3 000 poses
    with bearing-range sightings, N = 200; C5 = SE(3) chain of 3 000 poses + 300 loops, N = 100.  Per config: residuals
    against the oracle on the written-back samples (1e-5 relative, floor 1e-2 as in the GPU suite), forward proposals
    are roots of the residual, statistics equal the sums over the live particles."""
    import bench
    import rome_b200 as rb
    from oracle import oracle as O
    c, N = emulated_api, 100
    w = bench.build_workload(1)
    c.set_particles(rb.POSE2, w["poses"])
    c.set_factors_pose2pose2(w["ip"], w["iq"], w["mu"], w["cov"])
    fl = rb.SAMPLE | rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD | rb.WRITE_MEAS
    out = c.alloc_host_outputs(rb.POSE2POSE2, fl)
    c.eval_host(rb.POSE2POSE2, fl, seed=7, **out)
    seen = rb.dequantized_particles(c.get_anchors(rb.POSE2), c.get_offsets(rb.POSE2), N)
    meas = rb.offsets_to_meas(out["meas_out"], w["mu"], N)
    res = rb.rows_to_particle_major(out["res"], N)
    ref = O.sweep_pose2pose2(w["ip"], w["iq"], seen, meas)
    d = res - ref
    d[..., 2] = O.np_wrap(d[..., 2])
    # default arithmetic (float32 per particle): 1e-5 relative with floor 0.1 -- and on this workload the error itself
    # stays below 3e-7 (unit steps: lever arm 1 m, offsets of ~0.1 m; measured 1.2e-7 over 3.6e6 components)
    assert (np.abs(d) / np.maximum(np.abs(ref), 1e-1)).max() < 1e-5 and np.abs(d).max() < 3e-7, np.abs(d).max()
    o64 = c.alloc_host_outputs(rb.POSE2POSE2, rb.RESIDUAL | rb.PRECISE)
    c.eval_host(rb.POSE2POSE2, rb.RESIDUAL | rb.PRECISE, meas=out["meas_out"], **o64)
    d64 = rb.rows_to_particle_major(o64["res"], N) - ref
    d64[..., 2] = O.np_wrap(d64[..., 2])
    assert (np.abs(d64) / np.maximum(np.abs(ref), 1e-7)).max() < 1e-5   # Float64 chain: pure relative down to 1e-7
    prop = rb.rows_to_particle_major(out["prop_fwd"], N) + c.get_anchors(rb.POSE2)[w["iq"]][:, None, :]
    # the proposal zeroes the residual: evaluate the oracle with q := proposal (one factor at a time would be slow; stack)
    stacked = np.concatenate([seen, prop])
    r0 = O.sweep_pose2pose2(w["ip"], np.arange(len(w["ip"]), dtype=np.int32) + len(seen), stacked, meas)
    r0[..., 2] = O.np_wrap(r0[..., 2])
    assert np.abs(r0).max() < 2e-5   # float32 rounding of proposal coordinates of magnitude <= ~150 m
    assert np.allclose(out["stats"][:, 0:3], res.sum(1), rtol=1e-3, atol=2e-3)
    # the bench's second family
    c.set_factors_priorpose2(w["pr_ip"], w["pr_mu"], w["pr_cov"])
    o = c.alloc_host_outputs(rb.PRIORPOSE2, rb.SAMPLE | rb.RESIDUAL | rb.STATS | rb.WRITE_MEAS)
    c.eval_host(rb.PRIORPOSE2, rb.SAMPLE | rb.RESIDUAL | rb.STATS | rb.WRITE_MEAS, seed=7, **o)
    rp = O.sweep_priorpose2(w["pr_ip"], seen, rb.offsets_to_meas(o["meas_out"], w["pr_mu"], N))
    assert np.abs(rb.rows_to_particle_major(o["res"], N) - rp).max() < 1e-6

    # C4 / C5 through the graph API, like tests/test_gpu_graph_api.py but at thousands of poses
    fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
    bh = rb.generateGraph_Beehive(3000, N=200)
    rb.seed_particles(bh, N=200, seed=3)
    dg = rb.DeviceGraph(bh)
    poses = np.stack([v.val for v in dg.by_type[rb.POSE2]])
    points = np.stack([v.val for v in dg.by_type[rb.POINT2]])
    o = dg.eval(rb.BEARINGRANGE, fl, seed=1)
    facs = dg.by_family[rb.BEARINGRANGE]
    ip = [bh[f.variableOrderSymbols[0]].index for f in facs]
    il = [bh[f.variableOrderSymbols[1]].index for f in facs]
    assert len(facs) == 3001
    ref = O.sweep_bearingrange(ip, il, poses, points, o["meas"])
    d = o["res"] - ref
    d[..., 0] = O.np_wrap(d[..., 0])
    assert (np.abs(d) / np.maximum(np.abs(ref), 1e-1)).max() < 1e-5   # default arithmetic: floor 0.1
    p3 = rb.generateGraph_Pose3Chain(3000, loops=300)
    rb.seed_particles(p3, seed=4)
    dg3 = rb.DeviceGraph(p3)
    o = dg3.eval(rb.POSE3POSE3, fl, seed=1)
    facs = dg3.by_family[rb.POSE3POSE3]
    assert len(facs) >= 3200
    ip = [p3[f.variableOrderSymbols[0]].index for f in facs]
    iq = [p3[f.variableOrderSymbols[1]].index for f in facs]
    poses3 = np.stack([v.val for v in dg3.by_type[rb.POSE3]])
    ref = O.sweep_pose3pose3(ip, iq, poses3, o["meas"])
    assert (np.abs(o["res"] - ref) / np.maximum(np.abs(ref), 1e-2)).max() < 1e-5


def test_product_kernel_and_device_resident_sweeps(emulated_api):
    """the belief-update kernel (csrc/product_kernels.cu: one block of 4 warps per variable, shared bandwidths and pair
    CDF, __syncthreads / warp scans) under the block-level thread emulator, through the GPU suite's statistical tests:
    analytic Gaussian products, heading wrap, multi-modal selection against the NumPy twin, plan errors, and the
    reference's Hexagonal acceptance boxes after three device-resident sweeps"""
    import test_gpu_product as T
    T.test_product_of_gaussians_point2(emulated_api)
    T.test_product_heading_wraps(emulated_api)
    T.test_product_multimodal_matches_numpy_twin(emulated_api)
    for case in ("pair_point2_N10", "unaligned_pair_point2_N11", "pair_pose2_N13", "triple_point2_N6", "far_apart_pose2_N10", "wide_headings_pose2_N10"):
        T.test_product_matches_exact_mixture(emulated_api, case)
    T.test_product_pose3_on_manifold(emulated_api)
    T.test_plan_errors(emulated_api)
    T.test_hexagonal_solve_reference_boxes(emulated_api)
    T.test_sweeps_on_pose3_chain_and_beehive(emulated_api)


# ---- the library's OWN kernels under the emulator: persistent blocks, producer warp, TMA bulk copies, mbarrier rings ---
@pytest.fixture(scope="module", params=[1, 3, 7], ids=lambda g: f"grid{g}")
def pctx(request, hk_so):
    """eval_kernel / eval_kernel_w themselves (csrc/eval_pipeline.cuh), planned by plan_launch and dispatched by
    launch_family, on 1, 3 or 7 blocks: few blocks mean many tiles per block, i.e. the stage rings wrap many times"""
    from emu import EmulatedContext
    return EmulatedContext(hk_so, pipeline=True, grid_cap=request.param)


def test_pipeline_kernels_parity(pctx):
    T = _raw()
    T.test_pose2pose2_parity(pctx, 100)
    assert pctx.last_plan["threads"] == 288 and pctx.last_plan["pipeline"] == 0     # 8 consumer warps + producer warp
    T.test_pose2pose2_parity(pctx, 37)
    T.test_priorpose2_parity(pctx)
    T.test_bearingrange_parity(pctx)
    T.test_pose3pose3_parity(pctx, 64)
    T.test_priorpose3_parity(pctx)
    T.test_wide_heading_spread_takes_general_sincos_path(pctx)


@pytest.mark.parametrize("family", ["pose2pose2", "bearingrange", "pose3pose3", "priorpose3"])
def test_pipeline_kernels_fused_sampling(pctx, family):
    _raw().test_fused_sampling_matches_supplied_and_host_twin(pctx, family)
    if family in ("pose3pose3", "priorpose3"):   # sampled SE(3): the per-warp pipeline (12 warps, no producer warp)
        assert pctx.last_plan["pipeline"] in (0, 1)


@pytest.mark.parametrize("N", [1, 33, 500, 2000])
def test_pipeline_kernels_tile_variants(pctx, N):
    """N = 500 switches to the 2-factor tile, N = 2000 to the 1-factor tile (shared-memory budget)"""
    _raw().test_particle_count_edge_cases(pctx, N)
    assert pctx.last_plan["ft"] == {1: 8, 33: 8, 500: 2, 2000: 1}[N]


def test_pipeline_kernels_ranges_and_next_families(pctx):
    T = _raw()
    T.test_empty_range_and_zero_factors(pctx)
    T.test_too_many_particles_is_an_error(pctx)
    import test_gpu_next_families as T2
    import test_gpu_next_families_3d as T3
    T2.test_point2_gaussian_families(pctx, 37)
    T2.test_scalar_families(pctx, 100)
    T3.test_point3_families(pctx, 37)
    T3.test_pose3_partial_families(pctx, 100)


def test_pipeline_fuzz_against_direct_family_arithmetic(hk_so):
    """random (family, N, factor count, sub-range, flags, block count): the pipeline kernels must write exactly what the
    family arithmetic fed directly writes -- bit for bit -- inside [first, first + count) and nothing outside it.
    Exercises tile tails (factor counts that are not multiples of the tile), chunked id prefetch, ring wrap-around,
    globally indexed output rows and measurement blocks, and every tile variant."""
    import rome_b200 as rb
    from emu import EmulatedContext
    rng = np.random.default_rng(2024)
    direct = EmulatedContext(hk_so)
    fams = [rb.POSE2POSE2, rb.PRIORPOSE2, rb.BEARINGRANGE, rb.POSE3POSE3, rb.PRIORPOSE3, rb.POINT2POINT2, rb.POSE2POINT2,
            rb.POSE2POINT2RANGE, rb.POINT3POINT3, rb.POSE3POSE3XYYAW, rb.POSE3POSE3UNITTRANS, rb.POSE3POSE3ROTOFFSET,
            rb.POSE3POSE3TRANSFORM]
    for trial in range(286):
        fam = fams[trial % len(fams)]
        vt0, vt1, dm, dr, ns, dj, dfwd, dbwd = rb.FAMILY[fam]
        N = int(rng.choice([1, 5, 31, 32, 33, 100, 104, 200, 333, 700]))
        nv = int(rng.integers(2, 9))
        nF = int(rng.integers(1, 60))
        first = int(rng.integers(0, nF))
        count = int(rng.integers(1, nF - first + 1))
        pipe = EmulatedContext(hk_so, pipeline=True, grid_cap=int(rng.integers(1, 6)))
        parts = {t: rng.normal(size=(nv, N, rb.VAR_DIM[t])) * 0.3 + rng.normal(size=(nv, 1, rb.VAR_DIM[t])) * 2
                 for t in {vt0, vt1, rb.FAMILY_VT2.get(fam)} - {None}}
        i0 = rng.integers(0, nv, nF).astype(np.int32)
        i1 = rng.integers(0, nv, nF).astype(np.int32)
        i2 = rng.integers(0, nv, nF).astype(np.int32)
        for c in (direct, pipe):
            for t, p in parts.items():
                c.set_particles(t, p)
            if fam == rb.BEARINGRANGE:
                c.set_factors_bearingrange(i0, i1, np.column_stack([np.linspace(-1, 1, nF), np.full(nF, 0.1)]),
                                           np.column_stack([np.linspace(3, 9, nF), np.full(nF, 0.5)]))
            elif fam == rb.POSE2POINT2RANGE:
                c.set_factors_scalar(fam, i0, i1, np.column_stack([np.linspace(3, 9, nF), np.full(nF, 0.3)]))
            else:
                A = np.random.default_rng(trial).normal(size=(nF, dm, dm)) * 0.1
                cov = A @ np.swapaxes(A, 1, 2) + 0.01 * np.eye(dm)
                if fam in rb.FAMILY_VT2:
                    c.set_factors_ternary(fam, i0, i1, i2, np.random.default_rng(trial).normal(size=(nF, dm)), cov)
                else:
                    c.set_factors_gaussian(fam, i0, None if vt1 is None else i1, np.random.default_rng(trial).normal(size=(nF, dm)), cov)
        flags = rb.RESIDUAL
        if rng.random() < 0.7:
            flags |= rb.STATS
        if dfwd and rng.random() < 0.6:
            flags |= rb.PROPOSAL_FWD
        if dbwd and vt1 is not None and rng.random() < 0.4:
            flags |= rb.PROPOSAL_BWD
        if dj and rng.random() < 0.3:
            flags |= rb.JACOBIAN
        sample = rng.random() < 0.5
        if sample:
            flags |= rb.SAMPLE | (rb.WRITE_MEAS if rng.random() < 0.5 else 0)
        meas = None if sample else (rng.normal(size=(nF, rb.npad(N), dm)) * 0.05).astype(np.float32)
        outs = []
        for c in (direct, pipe):
            out = c.alloc_host_outputs(fam, flags)
            for a in out.values():
                a.fill(-777.0)   # sentinel: rows outside the evaluated range must stay untouched
            c.eval_host(fam, flags, seed=trial, stream_id=3, first=first, count=count, meas=meas, **out)
            outs.append(out)
        what = f"trial {trial}: family {fam} N {N} nF {nF} range [{first}, {first + count}) flags {flags} plan {pipe.last_plan}"
        for key in outs[0]:
            a, b = outs[0][key], outs[1][key]
            live = slice(first, first + count)
            assert np.array_equal(a[live], b[live], equal_nan=True), (what, key)
            assert np.all(b[:first] == -777.0) and np.all(b[first + count:] == -777.0), (what, key, "wrote outside the range")


@pytest.mark.parametrize("world", [2, 8])
def test_fused_all_gather_indexing_on_emulated_ranks(hk_so, world):
    """SURVEY 8e on emulated ranks: the factor list is split into `world` contiguous ranges (sharding.shard_range), every
    rank evaluates its range with the other ranks' proposal buffers set as peers (rome_b200_set_peer_proposals), the
    kernels' bulk stores deliver each row to every buffer -- afterwards EVERY rank must hold the proposals of ALL
    factors, identical to a single-rank evaluation (the real 2-GPU run is tests/test_gpu_multi.py; 8 ranks were only
    ever timed on the device, never compared)."""
    import rome_b200 as rb
    from rome_b200 import sharding
    from emu import EmulatedContext
    rng = np.random.default_rng(9)
    N, nv, nF = 100, 40, 203
    poses = rng.normal(size=(nv, 1, 3)) * [20, 20, 1] + rng.normal(size=(nv, N, 3)) * [0.1, 0.1, 0.02]
    ip = rng.integers(0, nv - 1, nF).astype(np.int32)
    iq = (ip + 1).astype(np.int32)
    mu = rng.normal(size=(nF, 3)) * [5, 1, 0.5]
    cov = np.tile(np.diag([0.01, 0.01, 0.001]), (nF, 1, 1))
    Np, c = rb.npad(N), sharding.shard_size(nF, world)
    flags = rb.SAMPLE | rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD
    ranks = [EmulatedContext(hk_so, pipeline=True, grid_cap=2) for _ in range(world)]
    bufs = [np.zeros((world * c, Np, 3), np.float32) for _ in range(world)]
    for r, ctx in enumerate(ranks):
        ctx.set_particles(rb.POSE2, poses)
        ctx.set_factors_pose2pose2(ip, iq, mu, cov)
        ctx.set_peer_proposals(rb.POSE2POSE2, [bufs[p].ctypes.data for p in range(world) if p != r])
    full = np.zeros((world * c, Np, 3), np.float32)
    single = EmulatedContext(hk_so, pipeline=True, grid_cap=2)
    single.set_particles(rb.POSE2, poses)
    single.set_factors_pose2pose2(ip, iq, mu, cov)
    res, st = np.zeros((nF, Np, 3), np.float32), np.zeros((nF, 16), np.float32)
    single.eval_host(rb.POSE2POSE2, flags, seed=5, res=res, stats=st, prop_fwd=full)
    for r, ctx in enumerate(ranks):
        first, count = sharding.shard_range(nF, r, world)
        ctx.eval_host(rb.POSE2POSE2, flags, seed=5, first=first, count=count, res=res, stats=st, prop_fwd=bufs[r])
    for r in range(world):
        assert np.array_equal(bufs[r][:nF], full[:nF]), r
        assert not bufs[r][nF:].any()


@pytest.mark.parametrize("N", [1, 3, 8, 17, 24])
def test_small_particle_counts_do_not_read_beyond_their_blocks(hk_so, N):
    """Npad < 32: the dead lanes of the (only) slot group must not pick up whatever lies behind the particle block -- with
    shared memory poisoned by NaN patterns, statistics (zero-masked sums: NaN * 0 = NaN) and every live output stay finite
    and the statistics equal the sums over the live particles.  Found by the fuzz above; all families share the slot loop
    (eval_pipeline.cuh ROME_SLOT_LOOP) or the SE(3) pair iterator (se3_common.cuh pair_of)."""
    import ctypes
    import rome_b200 as rb
    from emu import EmulatedContext
    rng = np.random.default_rng(N)
    c = EmulatedContext(hk_so, pipeline=True, grid_cap=2)
    nv, nF = 6, 19
    for fam in (rb.POSE2POSE2, rb.PRIORPOSE2, rb.BEARINGRANGE, rb.POSE3POSE3, rb.PRIORPOSE3, rb.POINT2POINT2, rb.POSE2POINT2RANGE,
                rb.POINT3POINT3, rb.POSE3POSE3XYYAW, rb.POSE3POSE3UNITTRANS):
        vt0, vt1, dm, dr, ns, dj, dfwd, dbwd = rb.FAMILY[fam]
        for t in {vt0, vt1} - {None}:
            c.set_particles(t, rng.normal(size=(nv, N, rb.VAR_DIM[t])) * 0.2 + rng.normal(size=(nv, 1, rb.VAR_DIM[t])) * 3)
        i0, i1 = rng.integers(0, nv, nF).astype(np.int32), rng.integers(0, nv, nF).astype(np.int32)
        if fam == rb.BEARINGRANGE:
            c.set_factors_bearingrange(i0, i1, np.column_stack([np.zeros(nF), np.full(nF, 0.1)]), np.column_stack([np.full(nF, 5.0), np.full(nF, 0.5)]))
        elif fam == rb.POSE2POINT2RANGE:
            c.set_factors_scalar(fam, i0, i1, np.column_stack([np.full(nF, 5.0), np.full(nF, 0.3)]))
        else:
            c.set_factors_gaussian(fam, i0, None if vt1 is None else i1, rng.normal(size=(nF, dm)), np.tile(0.01 * np.eye(dm), (nF, 1, 1)))
        for sample in (False, True):
            flags = rb.RESIDUAL | rb.STATS | (rb.PROPOSAL_FWD if dfwd else 0) | (rb.SAMPLE if sample else 0)
            out = c.alloc_host_outputs(fam, flags)
            meas = None if sample else (rng.normal(size=(nF, rb.npad(N), dm)) * 0.05).astype(np.float32)
            c._hk.hk_poison_smem(0xFF)
            c.eval_host(fam, flags, seed=2, meas=meas, **out)
            res = out["res"][:, :N].astype(np.float64)
            assert np.isfinite(res).all() and np.isfinite(out["stats"]).all(), (fam, sample, c.last_plan)
            assert np.allclose(out["stats"][:, :dr], res.sum(1), rtol=1e-4, atol=1e-5), (fam, sample)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gpu_side_rank_barrier_on_emulated_ranks(tmp_path_factory, world):
    """rome_b200_peer_signal / rome_b200_peer_wait (csrc/peer_kernels.cu) with `world` emulated ranks, 60 rounds, ranks
    delayed at random (some fall several kernels behind, others run ahead): no rank leaves wait k before every peer has
    issued signal k, nobody gives up, every epoch and every flag slot ends at the round count.  On the device this was
    (on the device: tests/test_gpu_multi.py at 2 GPUs, bench.py --barrier flags at 4 and 8)."""
    import ctypes as C
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import build as hk_build
    lib = C.CDLL(hk_build.build_peer(str(tmp_path_factory.mktemp("host_peer"))))
    rounds = 60
    rng = np.random.default_rng(world)
    for pattern in range(4):
        lag = np.zeros((world, rounds), np.int32)
        if pattern == 1:
            lag = rng.integers(0, 4, (world, rounds)).astype(np.int32)
        elif pattern == 2:   # one straggler: everybody else has to wait for it every round
            lag[world - 1] = 25
        elif pattern == 3:   # bursts: a rank stalls for a long time once in a while
            lag = (rng.random((world, rounds)) < 0.1).astype(np.int32) * 200
        lag = np.ascontiguousarray(lag)
        viol = C.c_int(-1)
        state = np.zeros((world, 16), np.uint32)
        for fn in (lib.hk_peer_barrier_rounds, lib.hk_peer_barrier_rounds_merged):   # two kernels / rome_b200_peer_barrier
            rc = fn(world, rounds, lag.ctypes.data_as(C.POINTER(C.c_int)), C.byref(viol),
                    state.ctypes.data_as(C.POINTER(C.c_uint32)))
            assert rc == 0 and viol.value == 0, (pattern, rc, viol.value)
            assert np.all(state[:, :world - 1] == rounds) and np.all(state[:, world - 1:8] == 0)   # one slot per peer
            assert np.all(state[:, 8] == rounds) and np.all(state[:, 9] == rounds) and np.all(state[:, 10] == 0)


@pytest.mark.parametrize("max_delay", [1, 3, 17, 60])
def test_pipeline_with_asynchronous_copies(hk_so, max_delay):
    """the pipeline kernels with bulk copies that are NOT instantaneous: every load / store is carried out a pseudo-random
    1 .. max_delay scheduler rounds after it was issued (loads complete their mbarrier bytes only then, stores read their
    shared-memory source only then unless the issuing thread waits for its bulk group).  Results must still be exactly
    those of the directly fed family arithmetic: the `full` / `empty` hand-over, the wait that protects a warp's output
    slice before it is rewritten and the final wait before the block exits are what this exercises (a kernel without
    the slice wait passes with instantaneous copies and fails here)."""
    import rome_b200 as rb
    from emu import EmulatedContext
    rng = np.random.default_rng(max_delay)
    direct = EmulatedContext(hk_so)
    fams = list(rb.FAMILY)
    for trial in range(96):
        fam = fams[trial % len(fams)]
        vt0, vt1, dm, dr, ns, dj, dfwd, dbwd = rb.FAMILY[fam]
        N = int(rng.choice([1, 8, 31, 33, 100, 104, 200, 333, 700]))
        nv, nF = int(rng.integers(1, 7)), int(rng.integers(1, 70))
        first = int(rng.integers(0, nF))
        count = int(rng.integers(1, nF - first + 1))
        pipe = EmulatedContext(hk_so, pipeline=True, grid_cap=int(rng.integers(1, 5)))
        parts = {t: rng.normal(size=(nv, N, rb.VAR_DIM[t])) * 0.3 + rng.normal(size=(nv, 1, rb.VAR_DIM[t])) * 3
                 for t in {vt0, vt1, rb.FAMILY_VT2.get(fam)} - {None}}
        i0, i1 = rng.integers(0, nv, nF).astype(np.int32), rng.integers(0, nv, nF).astype(np.int32)
        i2 = rng.integers(0, nv, nF).astype(np.int32)
        for c in (direct, pipe):
            for t, p in parts.items():
                c.set_particles(t, p)
            if fam == rb.BEARINGRANGE:
                c.set_factors_bearingrange(i0, i1, np.column_stack([np.linspace(-1, 1, nF), np.full(nF, 0.1)]),
                                           np.column_stack([np.linspace(3, 9, nF), np.full(nF, 0.5)]))
            elif fam in (rb.POSE2POINT2RANGE, rb.POINT2POINT2RANGE, rb.POSE2POINT2BEARING):
                c.set_factors_scalar(fam, i0, i1, np.column_stack([np.linspace(1, 3, nF), np.full(nF, 0.3)]))
            else:
                A = np.random.default_rng(trial).normal(size=(nF, dm, dm)) * 0.1
                if fam in rb.FAMILY_VT2:
                    c.set_factors_ternary(fam, i0, i1, i2, np.random.default_rng(trial).normal(size=(nF, dm)),
                                          A @ np.swapaxes(A, 1, 2) + 0.01 * np.eye(dm))
                else:
                    c.set_factors_gaussian(fam, i0, None if vt1 is None else i1, np.random.default_rng(trial).normal(size=(nF, dm)),
                                           A @ np.swapaxes(A, 1, 2) + 0.01 * np.eye(dm))
        flags = rb.RESIDUAL | (rb.STATS if rng.random() < 0.6 else 0) | (rb.PROPOSAL_FWD if dfwd and rng.random() < 0.7 else 0)
        sample = rng.random() < 0.5
        flags |= rb.SAMPLE if sample else 0
        meas = None if sample else (rng.normal(size=(nF, rb.npad(N), dm)) * 0.05).astype(np.float32)
        outs = []
        for c in (direct, pipe):
            out = c.alloc_host_outputs(fam, flags)
            for a in out.values():
                a.fill(-777.0)
            if c is pipe:
                c._hk.hk_poison_smem(0xFF)
                c._hk.hk_set_async(max_delay, trial)
            try:
                c.eval_host(fam, flags, seed=trial, stream_id=3, first=first, count=count, meas=meas, **out)
            finally:
                c._hk.hk_set_async(0, 0)
            outs.append(out)
        for key in outs[0]:
            assert np.array_equal(outs[0][key], outs[1][key], equal_nan=True), (trial, fam, N, nF, first, count, flags, key, pipe.last_plan)


@pytest.mark.parametrize("grid", [1, 3, 7])
def test_deferred_tiles_do_not_change_results(hk_so, grid):
    """A launch that takes part in the rank barrier visits the tiles of its peer-dependent factor range FIRST and signals
    as soon as they have landed (TileOrder / early_signal, csrc/eval_pipeline.cuh; no peers here, so only the counting
    runs): every factor is still evaluated exactly once with the same result, every launch signals exactly once -- both
    pipelines (producer warp: Pose2Pose2 / BearingRange, per-warp rings: Pose3Pose3), ranges at the start, in the middle,
    at the end, covering everything, empty, and a partial last tile"""
    import ctypes as C
    import rome_b200 as rb
    from emu import EmulatedContext
    T = _raw()
    c = EmulatedContext(hk_so, pipeline=True, grid_cap=grid)
    rng = np.random.default_rng(50 + grid)
    N, nvars, nF = 40, 60, 203   # 203 = 25 tiles of 8 + 3, 16 tiles of 12 + 11
    poses, ip, iq = T.make_pose2_graph(rng, nvars, nF, N)
    c.set_particles(rb.POSE2, poses)
    c.set_factors_pose2pose2(ip, iq, rng.normal(size=(nF, 3)), T.rand_cov(rng, nF, 3, [0.1, 0.1, 0.02]))
    p3 = T.make_pose3(rng, nvars, N)
    c.set_particles(rb.POSE3, p3)
    c.set_factors_pose3pose3(ip, iq, rng.normal(size=(nF, 6)) * 0.2, T.rand_cov(rng, nF, 6, [0.1] * 3 + [0.01] * 3))
    hk = c._hk
    hk.hk_set_barrier_range.argtypes = [C.c_int, C.c_int]
    hk.hk_barrier_state.restype = C.POINTER(C.c_uint32)
    signals_before = hk.hk_barrier_state()[8]
    for fam in (rb.POSE2POSE2, rb.POSE3POSE3):
        fl = rb.SAMPLE | rb.RESIDUAL | rb.STATS
        hk.hk_set_barrier_range(0, 2 ** 31 - 1)
        ref = c.alloc_host_outputs(fam, fl)
        c.eval_host(fam, fl, seed=3, **ref)
        for lo, hi in ((0, 16), (72, 131), (96, 97), (180, 203), (0, 203), (50, 50), (199, 400)):
            for first, count in ((0, -1), (24, 150)):
                hk.hk_set_barrier_range(lo, hi)
                out = c.alloc_host_outputs(fam, fl)
                c.eval_host(fam, fl | rb.BARRIER_WAIT | rb.BARRIER_SIGNAL, seed=3, first=first, count=count, **out)
                a, b = first, nF if count < 0 else first + count
                assert np.array_equal(out["res"][a:b], ref["res"][a:b]) and np.array_equal(out["stats"][a:b], ref["stats"][a:b]), (fam, lo, hi, first)
                assert not out["res"][:a].any() and not out["res"][b:].any()
    hk.hk_set_barrier_range(0, 2 ** 31 - 1)
    st = hk.hk_barrier_state()
    assert st[8] - signals_before == 2 * 7 * 2 and st[12] == 0 and st[10] == 0   # one signal per launch, counter back at zero, no give-up


def test_fullsize_gpu_tests_logic_on_the_emulated_device(ectx):
    """tests/test_gpu_fullsize.py's heading-across-the-cut case and its checker, run against the emulated device (the
    10 000-pose cases themselves are covered by test_full_size_baseline_workloads above)"""
    import test_gpu_fullsize as T
    T.test_heading_spread_across_the_branch_cut(ectx)
