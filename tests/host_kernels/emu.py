"""`EmulatedContext`: the subset of rome_b200.Context the parity tests use, served by tests/host_kernels' host build
of the kernel SOURCES (warp emulator) instead of librome_b200.so.  TEST INFRASTRUCTURE for the CPU suite: it exists so
that the kernels' arithmetic -- fast-path selection, masking of padding lanes, statistics reductions, proposals, the
pack / unpack layout kernels -- can be checked against the oracle without a GPU.  The TMA/mbarrier pipeline, streams,
CUDA graphs, the product kernel and the C ABI's argument checking are NOT emulated (those are `-m gpu` tests)."""
import ctypes as C

import numpy as np

import rome_b200 as rb
from rome_b200 import _lib as L
from rome_b200.engine import FAMILY, FAMILY_VT2, VAR_DIM

_WRAP = {L.POSE2: 2, L.POINT2: -1, L.POSE3: -1, L.POINT3: -1, L.ROTATION3: -2}
_ROW = {
    "se2": np.dtype([("ip", "i4"), ("iq", "i4"), ("mu", "f8", 3), ("L", "f4", 6), ("pad", "f4", 2)]),
    "br": np.dtype([("ip", "i4"), ("iq", "i4"), ("mu_b", "f8"), ("mu_r", "f8"), ("sig_b", "f4"), ("sig_r", "f4")]),
    "se3": np.dtype([("ip", "i4"), ("iq", "i4"), ("mu", "f8", 6), ("L", "f4", 21), ("ir", "i4"), ("pad", "f4", 4)]),
    "pt2": np.dtype([("ip", "i4"), ("iq", "i4"), ("mu", "f8", 2), ("L", "f4", 3), ("pad", "f4", 3)]),
    "s1": np.dtype([("ip", "i4"), ("iq", "i4"), ("mu", "f8"), ("sigma", "f4"), ("pad", "f4", 3)]),
}


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class EmulatedContext:
    """pipeline=False: every factor's inputs are handed to the family arithmetic directly (fast: a warp of fibers per
    factor).  pipeline=True: the library's OWN kernels run -- eval_kernel / eval_kernel_w with their persistent blocks,
    producer warp, TMA bulk copies and mbarrier rings (host semantics in hk_shim.h), planned by the library's plan_launch
    and dispatched by its launch_family -- on at most `grid_cap` blocks."""

    def __init__(self, so_path, pipeline=False, grid_cap=3):
        self.pipeline, self.grid_cap, self.last_plan = pipeline, grid_cap, None
        self._hk = C.CDLL(so_path)
        self._hk.hk_eval_pipeline.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                              C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                              C.POINTER(C.c_float), C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_int), C.c_int,
                                              C.c_void_p]
        self._peers = {}
        self._hk.hk_eval.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint64, C.c_uint32,
                                     C.c_int]
        self._hk.hk_set_v2.argtypes = [C.c_void_p]
        self._hk.hk_pack.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
        self._hk.hk_unpack.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
        for kind, dt in _ROW.items():  # the numpy row layouts must be the C structs'
            fam = {"se2": L.POSE2POSE2, "br": L.BEARINGRANGE, "se3": L.POSE3POSE3, "pt2": L.POINT2POINT2, "s1": L.POSE2POINT2RANGE}[kind]
            assert dt.itemsize == self._hk.hk_row_bytes(fam), kind
        self._store, self._shape, self._rows = {}, {}, {}
        self.launch_count = 0
        self.device = 0

    # -- variables --------------------------------------------------------------------------------------------------
    def set_particles(self, vartype, coords):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        nvars, N, d = coords.shape
        assert d == VAR_DIM[vartype]
        Np = rb.npad(N)
        store = np.zeros((nvars, self._hk.hk_block_bytes(d, Np)), np.uint8)
        assert self._hk.hk_pack(d, _WRAP[vartype], nvars, N, Np, coords.ctypes.data, store.ctypes.data) == 0
        self._store[vartype], self._shape[vartype] = store, (nvars, N, Np)

    def particles_device(self, vartype):
        nvars, N, Np = self._shape[vartype]
        d = VAR_DIM[vartype]
        return self._store[vartype].ctypes.data, self._hk.hk_block_bytes(d, Np), self._hk.hk_header_bytes(d), nvars, N, Np

    def get_particles(self, vartype):
        nvars, N, Np = self._shape[vartype]
        d = VAR_DIM[vartype]
        out = np.empty((nvars, N, d))
        assert self._hk.hk_unpack(d, _WRAP[vartype], nvars, N, Np, self._store[vartype].ctypes.data, out.ctypes.data) == 0
        return out

    def get_anchors(self, vartype):
        d = VAR_DIM[vartype]
        return np.ascontiguousarray(self._store[vartype][:, :d * 8]).view(np.float64).reshape(-1, d).copy()

    def get_offsets(self, vartype):
        d, Np = VAR_DIM[vartype], self._shape[vartype][2]
        hb = self._hk.hk_header_bytes(d)
        return np.ascontiguousarray(self._store[vartype][:, hb:]).view(np.float32).reshape(-1, Np, d).copy()

    # -- factors (rows as rome_b200_api.cu builds them: Float64 Cholesky, float32 lower triangle row-major) ----------
    def _gauss(self, family, kind, i0, i1, mu, cov, i2=None):
        mu, cov = np.asarray(mu, np.float64), np.asarray(cov, np.float64)
        nF, d = mu.shape
        rows = np.zeros(nF, _ROW[kind])
        rows["ip"] = np.asarray(i0, np.int32)
        rows["iq"] = -1 if i1 is None else np.asarray(i1, np.int32)
        if i2 is not None:
            rows["ir"] = np.asarray(i2, np.int32)
        rows["mu"] = mu
        Lc = np.linalg.cholesky(0.5 * (cov + np.swapaxes(cov, 1, 2))) if nF else np.zeros((0, d, d))
        rows["L"] = np.stack([Lc[:, i, j] for i in range(d) for j in range(i + 1)], 1).astype(np.float32) if nF else 0
        self._rows[family] = rows

    def set_factors_pose2pose2(self, ip, iq, mu, cov):
        self._gauss(L.POSE2POSE2, "se2", ip, iq, mu, cov)

    def set_factors_priorpose2(self, ip, mu, cov):
        self._gauss(L.PRIORPOSE2, "se2", ip, None, mu, cov)

    def set_factors_pose3pose3(self, ip, iq, mu, cov):
        self._gauss(L.POSE3POSE3, "se3", ip, iq, mu, cov)

    def set_factors_priorpose3(self, ip, mu, cov):
        self._gauss(L.PRIORPOSE3, "se3", ip, None, mu, cov)

    def set_factors_point2(self, family, i0, i1, mu, cov):
        self._gauss(family, "pt2", i0, i1, mu, cov)

    def set_factors_gaussian(self, family, i0, i1, mu, cov):
        if family in FAMILY_VT2:  # as rome_b200_set_factors_gaussian
            raise rb.RomeB200Error(L.BAD_ARG, "family has a third variable: use rome_b200_set_factors_ternary")
        kind = {64: "se2", 160: "se3", 48: "pt2"}[self._hk.hk_row_bytes(family)]
        self._gauss(family, kind, i0, i1, mu, cov)

    def set_factors_ternary(self, family, i0, i1, i2, mu, cov):
        assert family in FAMILY_VT2
        self._gauss(family, "se3", i0, i1, mu, cov, i2)

    def set_factors_bearingrange(self, ip, il, bearing, rng):
        bearing, rng = np.asarray(bearing, np.float64), np.asarray(rng, np.float64)
        rows = np.zeros(len(ip), _ROW["br"])
        rows["ip"], rows["iq"] = np.asarray(ip, np.int32), np.asarray(il, np.int32)
        rows["mu_b"], rows["sig_b"], rows["mu_r"], rows["sig_r"] = bearing[:, 0], bearing[:, 1], rng[:, 0], rng[:, 1]
        self._rows[L.BEARINGRANGE] = rows

    def set_factors_scalar(self, family, i0, i1, belief):
        belief = np.asarray(belief, np.float64)
        rows = np.zeros(len(i0), _ROW["s1"])
        rows["ip"], rows["iq"], rows["mu"], rows["sigma"] = np.asarray(i0, np.int32), np.asarray(i1, np.int32), belief[:, 0], belief[:, 1]
        self._rows[family] = rows

    def num_factors(self, family):
        return len(self._rows[family])

    alloc_host_outputs = rb.Context.alloc_host_outputs

    # -- the hot path -------------------------------------------------------------------------------------------------
    def eval_host(self, family, flags, *, seed=0, stream_id=0, first=0, count=-1, meas=None, meas_out=None, res=None,
                  prop_fwd=None, prop_bwd=None, stats=None, jac=None, sync=True):
        vt0, vt1 = FAMILY[family][0], FAMILY[family][1]
        rows = self._rows[family]
        nvars, N, Np = self._shape[vt0]
        if count < 0:
            count = len(rows) - first

        def bad(msg, code=L.BAD_ARG):  # the argument checks of rome_b200_api.cu:check_eval that the parity tests rely on
            raise rb.RomeB200Error(code, msg)
        _, _, dm, dr, ns, dj, dfwd, dbwd = FAMILY[family]
        if first < 0 or count < 0 or first + count > len(rows):
            bad("factor range out of bounds")
        if not flags & L.SAMPLE and meas is None:
            bad("meas is NULL and SAMPLE is not set")
        if flags & L.WRITE_MEAS and (not flags & L.SAMPLE or meas_out is None):
            bad("WRITE_MEAS needs SAMPLE and meas_out")
        if flags & L.DECONV and (family > L.PRIORPOSE3 or flags & L.WRITE_MEAS or meas_out is None):
            bad("DECONV needs a closed form, meas_out and excludes WRITE_MEAS")
        if flags & L.RESIDUAL and res is None:
            bad("res is NULL")
        if flags & L.PROPOSAL_FWD and (dfwd == 0 or prop_fwd is None):
            bad("no closed-form forward proposal / prop_fwd is NULL")
        if flags & L.PROPOSAL_BWD and (dbwd == 0 or prop_bwd is None):
            bad("no closed-form backward proposal / prop_bwd is NULL")
        if flags & L.STATS and stats is None:
            bad("stats is NULL")
        if flags & L.JACOBIAN and (dj == 0 or jac is None):
            bad("no Jacobian output / jac is NULL")
        if count == 0:
            return
        assert rows["ip"].max() < nvars and (vt1 is None or (rows["iq"].max() < self._shape[vt1][0] and self._shape[vt1][1] == N))
        # compile-time flag variants exist where the planner picks the 8-factor tile (or the per-warp pipeline), as in
        # plan_launch / launch_sample of the library
        out_flags = (flags if dfwd else flags & ~L.PROPOSAL_FWD) & ~(L.SAMPLE | L.INDEPENDENT | L.ROUTED_ONLY |
                                                                       L.BARRIER_WAIT | L.BARRIER_SIGNAL)
        hot = 1 if out_flags == (L.RESIDUAL | L.STATS) else 2 if out_flags == (L.RESIDUAL | L.STATS | L.PROPOSAL_FWD) else 0
        plan = rb.plan_query(family, flags, N)
        variant = hot if (plan["warps"] in (8, 4) or plan["pipeline"] == 1) else 0
        for name, a in (("meas", meas), ("meas_out", meas_out), ("res", res), ("prop_fwd", prop_fwd), ("prop_bwd", prop_bwd),
                        ("stats", stats), ("jac", jac)):
            assert a is None or (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous), name
        v1 = self._store[vt1] if vt1 is not None else self._store[vt0]
        vt2 = FAMILY_VT2.get(family)
        if vt2 is not None:
            assert rows["ir"].max() < self._shape[vt2][0] and self._shape[vt2][1] == N
        self._hk.hk_set_v2(C.c_void_p(self._store[vt2].ctypes.data if vt2 is not None else None))
        if self.pipeline:
            info = (C.c_int * 6)()
            peers = self._peers.get(family, [])
            parr = (C.c_void_p * max(1, len(peers)))(*peers)
            rc = self._hk.hk_eval_pipeline(family, flags, first, count, N, Np, rows.ctypes.data, self._store[vt0].ctypes.data,
                                           v1.ctypes.data, _fp(meas), _fp(meas_out), _fp(res), _fp(prop_fwd), _fp(prop_bwd),
                                           _fp(stats), _fp(jac), seed, stream_id, self.grid_cap, info, len(peers), parr)
            if rc == -3:
                raise rb.RomeB200Error(L.SHAPE_MISMATCH, "N is too large for the shared-memory pipeline of this family")
            self.last_plan = dict(zip(("ft", "variant", "stages", "pipeline", "grid", "threads"), info))
            assert rc == 0, f"emulated kernel failed: {rc} (700 = stopped making progress), plan {self.last_plan}"
            assert (self.last_plan["ft"], self.last_plan["stages"], self.last_plan["pipeline"]) == \
                (plan["warps"], plan["stages"], plan["pipeline"])   # the same plan the shipped library reports
        else:
            rc = self._hk.hk_eval(family, flags, first, count, N, Np, rows.ctypes.data, self._store[vt0].ctypes.data,
                                  v1.ctypes.data, _fp(meas), _fp(meas_out), _fp(res), _fp(prop_fwd), _fp(prop_bwd), _fp(stats),
                                  _fp(jac), seed, stream_id, variant)
            assert rc == 0
        self.launch_count += 1

    # the device-pointer entry point: "device memory" of the emulated device is host memory (malloc_device below)
    def eval(self, family, flags, *, seed=0, stream_id=0, first=0, count=-1, meas=None, meas_out=None, res=None,
             prop_fwd=None, prop_bwd=None, stats=None, jac=None):
        def view(p, d):
            if p is None or isinstance(p, np.ndarray):
                return p
            nF, Np = len(self._rows[family]), self._shape[FAMILY[family][0]][2]
            n = nF * (Np * d if d > 0 else -d)
            return np.ctypeslib.as_array(C.cast(int(p), C.POINTER(C.c_float)), shape=(n,))
        _, _, dm, dr, ns, dj, dfwd, dbwd = FAMILY[family]
        self.eval_host(family, flags & ~L.INDEPENDENT, seed=seed, stream_id=stream_id, first=first, count=count,
                       meas=view(meas, dm), meas_out=view(meas_out, dm), res=view(res, dr), prop_fwd=view(prop_fwd, dfwd),
                       prop_bwd=view(prop_bwd, dbwd), stats=view(stats, -ns), jac=view(jac, dj))

    def set_peer_proposals(self, family, peer_ptrs):
        """fused all-gather: forward-proposal rows are also stored into these buffers (pipeline mode only)"""
        assert self.pipeline and len(peer_ptrs) <= 7
        self._peers[family] = [int(p) for p in peer_ptrs]

    # -- "device" memory ------------------------------------------------------------------------------------------------
    def malloc_device(self, nbytes):
        a = np.zeros(max(1, (int(nbytes) + 3) // 4), np.float32)
        self._mem = getattr(self, "_mem", {})
        self._mem[a.ctypes.data] = a
        return a.ctypes.data

    def free_device(self, ptr):
        self._mem.pop(int(ptr), None)

    def memcpy_h2d(self, dst_ptr, src):
        src = np.ascontiguousarray(src)
        C.memmove(int(dst_ptr), src.ctypes.data, src.nbytes)

    def memcpy_d2h(self, dst, src_ptr):
        C.memmove(dst.ctypes.data, int(src_ptr), dst.nbytes)

    # -- belief update (product of proposal densities) --------------------------------------------------------------------
    def set_product_plan(self, vartype, var_offsets, src_buf, src_row):
        off, sb, sr = (np.ascontiguousarray(a, dtype=np.int32) for a in (var_offsets, src_buf, src_row))
        if len(off) < 1 or off[0] != 0 or np.any(np.diff(off) < 0) or off[-1] != len(sb) or len(sb) != len(sr):
            raise rb.RomeB200Error(L.BAD_ARG, "bad plan arrays")
        if np.any(np.diff(off) > L.MAX_PRODUCT_SOURCES):
            raise rb.RomeB200Error(L.BAD_ARG, "too many proposals for one variable")
        if len(sb) and (sb.min() < 0 or sb.max() >= 16 or sr.min() < 0):
            raise rb.RomeB200Error(L.BAD_ARG, "source buffer index / row out of range")
        self._plan = getattr(self, "_plan", {})
        self._plan[vartype] = (off, sb, sr)

    def product(self, vartype, bufs, *, seed=0, stream_id=0, gibbs_iters=0, reanchor=True, bw_out=None, manifold=True):
        off, sb, sr = self._plan[vartype]
        nvars, N, Np = self._shape[vartype]
        if len(off) - 1 != nvars:
            raise rb.RomeB200Error(L.NOT_SET, "no product plan for this variable type (or its size differs)")
        if len(sb) and sb.max() >= len(bufs):
            raise rb.RomeB200Error(L.BAD_ARG, "the plan refers to more proposal buffers than were passed")
        d = VAR_DIM[vartype]
        arr = (C.c_void_p * max(1, len(bufs)))(*[int(b) for b in bufs])
        self._hk.hk_product.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_float, C.c_int]
        rc = self._hk.hk_product(d, _WRAP[vartype], self._store[vartype].ctypes.data, off.ctypes.data, sb.ctypes.data,
                                 sr.ctypes.data, len(bufs), arr, None if bw_out is None else int(bw_out), nvars, N, Np,
                                 gibbs_iters if gibbs_iters > 0 else 2, seed, stream_id,
                                 float((4.0 / ((d + 2.0) * N)) ** (1.0 / (d + 4.0))), 1 if manifold else 0)
        assert rc == 0
        self.launch_count += 1
        if reanchor:
            self.reanchor(vartype)

    def reanchor(self, vartype):
        nvars, N, Np = self._shape[vartype]
        self._hk.hk_reanchor.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        assert self._hk.hk_reanchor(VAR_DIM[vartype], _WRAP[vartype], self._store[vartype].ctypes.data, nvars, N, Np) == 0
        self.launch_count += 1

    def synchronize(self):
        pass

    def close(self):
        pass
