"""Thin object wrapper over the C ABI (include/rome_b200.h) plus the layout helpers a host caller
needs (reference Float64 particle-major arrays <-> anchored float32 particle-major device rows).

All compute goes through librome_b200.so; nothing here evaluates a factor on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import (BEARINGRANGE, JACOBIAN, POINT2, POINT2POINT2, POINT2POINT2RANGE, POINT3, POINT3POINT3, POSE2,
                   POSE2POINT2, POSE2POINT2BEARING, POSE2POINT2RANGE, POSE2POSE2, POSE3, POSE3POSE3,
                   POSE3POSE3ROTATION, POSE3POSE3ROTOFFSET, POSE3POSE3TRANSFORM, POSE3POSE3UNITTRANS, POSE3POSE3XYYAW,
                   PRIORPOINT2, PRIORPOINT3, PRIORPOSE2, ROTATION3,
                   PRIORPOSE3, PROPOSAL_BWD, PROPOSAL_FWD, RESIDUAL, SAMPLE, STATS, WRITE_MEAS, Buffers, RomeB200Error)

VAR_DIM = {POSE2: 3, POINT2: 2, POSE3: 6, POINT3: 3, ROTATION3: 3}
# family -> (vartype of first variable, vartype of second variable or None, dm, dr, nstats, dj, dfwd, dbwd)
FAMILY = {
    POSE2POSE2: (POSE2, POSE2, 3, 3, 16, 4, 3, 3),
    PRIORPOSE2: (POSE2, None, 3, 3, 16, 0, 3, 0),
    BEARINGRANGE: (POSE2, POINT2, 2, 2, 16, 4, 2, 3),
    POSE3POSE3: (POSE3, POSE3, 6, 6, 32, 36, 6, 6),
    PRIORPOSE3: (POSE3, None, 6, 6, 32, 9, 6, 0),
    # next-row families
    PRIORPOINT2: (POINT2, None, 2, 2, 16, 0, 2, 0),
    POINT2POINT2: (POINT2, POINT2, 2, 2, 16, 0, 2, 0),
    POSE2POINT2: (POSE2, POINT2, 2, 2, 16, 0, 2, 0),
    POSE2POINT2RANGE: (POSE2, POINT2, 1, 1, 16, 0, 0, 0),
    POINT2POINT2RANGE: (POINT2, POINT2, 1, 1, 16, 0, 0, 0),
    POSE2POINT2BEARING: (POSE2, POINT2, 1, 1, 16, 0, 0, 0),
    PRIORPOINT3: (POINT3, None, 3, 3, 16, 0, 3, 0),
    POINT3POINT3: (POINT3, POINT3, 3, 3, 16, 0, 3, 0),
    POSE3POSE3XYYAW: (POSE3, POSE3, 3, 3, 16, 0, 0, 0),
    POSE3POSE3ROTATION: (POSE3, POSE3, 3, 3, 16, 0, 0, 0),
    POSE3POSE3UNITTRANS: (POSE3, POSE3, 6, 6, 32, 0, 0, 0),
    # families with a third variable (FAMILY_VT2): src/factors/Pose3Pose3.jl:57-95
    POSE3POSE3ROTOFFSET: (POSE3, POSE3, 6, 6, 32, 0, 6, 0),
    POSE3POSE3TRANSFORM: (POSE3, POSE3, 6, 6, 32, 0, 6, 0),
}
# vartype of the THIRD variable of a family
FAMILY_VT2 = {POSE3POSE3ROTOFFSET: ROTATION3, POSE3POSE3TRANSFORM: POSE3}
# algorithmic bytes per factor-particle eval with this layout (DESIGN.md "bytes per eval"):
# read both variables' offsets + the measurement offsets, write the residual (float32 each)
BYTES_PER_EVAL = {POSE2POSE2: 48, PRIORPOSE2: 36, BEARINGRANGE: 36, POSE3POSE3: 96, PRIORPOSE3: 72,
                  PRIORPOINT2: 24, POINT2POINT2: 32, POSE2POINT2: 36, POSE2POINT2RANGE: 28, POINT2POINT2RANGE: 24,
                  POSE2POINT2BEARING: 28, PRIORPOINT3: 36, POINT3POINT3: 48, POSE3POSE3XYYAW: 72,
                  POSE3POSE3ROTATION: 72, POSE3POSE3UNITTRANS: 96, POSE3POSE3ROTOFFSET: 108, POSE3POSE3TRANSFORM: 120}
BYTES_PER_EVAL_SAMPLED = {POSE2POSE2: 36, PRIORPOSE2: 24, BEARINGRANGE: 28, POSE3POSE3: 72, PRIORPOSE3: 48,
                          PRIORPOINT2: 16, POINT2POINT2: 24, POSE2POINT2: 28, POSE2POINT2RANGE: 24,
                          POINT2POINT2RANGE: 20, POSE2POINT2BEARING: 24, PRIORPOINT3: 24, POINT3POINT3: 36,
                          POSE3POSE3XYYAW: 60, POSE3POSE3ROTATION: 60, POSE3POSE3UNITTRANS: 72,
                          POSE3POSE3ROTOFFSET: 84, POSE3POSE3TRANSFORM: 96}


def plan_query(family: int, flags: int, N: int):
    """launch geometry on a B200 for (family, flags, N): dict(warps, stages, ctas_per_sm, smem_bytes, pipeline); no GPU needed"""
    lib = L.load()
    v = [C.c_int() for _ in range(5)]
    rc = lib.rome_b200_plan_query(family, flags, N, *[C.byref(x) for x in v])
    if rc != 0:
        raise RomeB200Error(rc, "N is too large for the shared-memory pipeline of this family" if rc == L.SHAPE_MISMATCH
                            else "bad family / N")
    return dict(zip(("warps", "stages", "ctas_per_sm", "smem_bytes", "pipeline"), (x.value for x in v)))


def npad(N: int) -> int:
    return (N + 7) // 8 * 8


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    raise TypeError(f"cannot take a pointer of {type(x)}")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """One rome_b200_ctx: particle stores, factor tables and the stream work is issued on."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.rome_b200_create(device, C.byref(h))
        if rc != 0:
            raise RomeB200Error(rc, self._lib.rome_b200_last_error(None).decode())
        self._h = h
        self.device = device

    # -- plumbing ------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise RomeB200Error(rc, self._lib.rome_b200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rome_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        self._ck(self._lib.rome_b200_set_stream(self._h, cuda_stream_ptr))

    def use_torch_stream(self):
        import torch
        # torch's default stream has handle 0, which the C ABI reads as "the ctx's own stream": pass cudaStreamLegacy
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream or 1)

    def synchronize(self):
        self._ck(self._lib.rome_b200_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.rome_b200_launch_count(self._h))

    # -- variables -----------------------------------------------------------------------------
    def set_particles(self, vartype: int, coords):
        """coords: Float64 [nvars][N][d] (reference layout); numpy array or pinned torch tensor."""
        if isinstance(coords, np.ndarray):
            coords = _f64(coords)
        nvars, N, d = coords.shape
        if d != VAR_DIM[vartype]:
            raise ValueError("coordinate dimension does not match the variable type")
        self._ck(self._lib.rome_b200_set_particles(self._h, vartype, nvars, N, _ptr(coords)))
        self._keep = coords  # keep alive until the async copy is consumed

    def set_particles_anchored(self, vartype: int, anchors, offsets):
        """anchors Float64 [nvars][d], offsets float32 [nvars][N][d] (numpy arrays or pinned torch tensors): the
        device's own representation, half the upload bytes of set_particles"""
        if isinstance(anchors, np.ndarray):
            anchors = _f64(anchors)
        if isinstance(offsets, np.ndarray):
            offsets = np.ascontiguousarray(offsets, dtype=np.float32)
        nvars, N, d = offsets.shape
        if d != VAR_DIM[vartype] or tuple(anchors.shape) != (nvars, d):
            raise ValueError("anchor / offset shapes do not match the variable type")
        self._ck(self._lib.rome_b200_set_particles_anchored(self._h, vartype, nvars, N, _ptr(anchors), _ptr(offsets)))
        self._keep = (anchors, offsets)

    def get_particles(self, vartype: int) -> np.ndarray:
        nvars, N = self.particles_device(vartype)[3:5]
        out = np.empty((nvars, N, VAR_DIM[vartype]))
        self._ck(self._lib.rome_b200_get_particles(self._h, vartype, out.ctypes.data))
        return out

    def particles_device(self, vartype: int):
        """(store pointer, block bytes, header bytes, nvars, N, Npad) of the device particle store"""
        ps = C.c_void_p()
        bb, hb, nv, N, Np = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self._lib.rome_b200_particles_device(self._h, vartype, C.byref(ps), C.byref(bb), C.byref(hb),
                                                      C.byref(nv), C.byref(N), C.byref(Np)))
        return ps.value, bb.value, hb.value, nv.value, N.value, Np.value

    def _store_bytes(self, vartype: int):
        ps, bb, hb, nvars, N, Np = self.particles_device(vartype)
        self.synchronize()
        raw = np.empty((nvars, bb), dtype=np.uint8)
        self._ck(self._lib.rome_b200_memcpy_d2h(self._h, raw.ctypes.data, ps, raw.nbytes))
        return raw, hb, Np

    def get_anchors(self, vartype: int) -> np.ndarray:
        """Float64 anchors [nvars][d] read back from the device store"""
        raw, hb, _ = self._store_bytes(vartype)
        d = VAR_DIM[vartype]
        return np.ascontiguousarray(raw[:, :d * 8]).view(np.float64).reshape(-1, d).copy()

    def get_offsets(self, vartype: int) -> np.ndarray:
        """float32 offsets [nvars][Npad][d] read back from the device store"""
        raw, hb, Np = self._store_bytes(vartype)
        d = VAR_DIM[vartype]
        return np.ascontiguousarray(raw[:, hb:]).view(np.float32).reshape(-1, Np, d).copy()

    def adopt_proposal(self, vartype: int, var: int, d_prop, factor: int):
        self._ck(self._lib.rome_b200_adopt_proposal(self._h, vartype, var, _ptr(d_prop), factor))

    # -- factors -------------------------------------------------------------------------------
    def _dp(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def _ip(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_int32))

    def set_factors_pose2pose2(self, ip, iq, mu, cov):
        ip, iq, mu, cov = _i32(ip), _i32(iq), _f64(mu), _f64(cov)
        self._ck(self._lib.rome_b200_set_factors_pose2pose2(self._h, len(ip), self._ip(ip), self._ip(iq),
                                                            self._dp(mu), self._dp(cov)))

    def set_factors_priorpose2(self, ip, mu, cov):
        ip, mu, cov = _i32(ip), _f64(mu), _f64(cov)
        self._ck(self._lib.rome_b200_set_factors_priorpose2(self._h, len(ip), self._ip(ip), self._dp(mu),
                                                            self._dp(cov)))

    def set_factors_bearingrange(self, ip, il, bearing, rng):
        ip, il, bearing, rng = _i32(ip), _i32(il), _f64(bearing), _f64(rng)
        self._ck(self._lib.rome_b200_set_factors_bearingrange(self._h, len(ip), self._ip(ip), self._ip(il),
                                                              self._dp(bearing), self._dp(rng)))

    def set_factors_pose3pose3(self, ip, iq, mu, cov):
        ip, iq, mu, cov = _i32(ip), _i32(iq), _f64(mu), _f64(cov)
        self._ck(self._lib.rome_b200_set_factors_pose3pose3(self._h, len(ip), self._ip(ip), self._ip(iq),
                                                            self._dp(mu), self._dp(cov)))

    def set_factors_priorpose3(self, ip, mu, cov):
        ip, mu, cov = _i32(ip), _f64(mu), _f64(cov)
        self._ck(self._lib.rome_b200_set_factors_priorpose3(self._h, len(ip), self._ip(ip), self._dp(mu),
                                                            self._dp(cov)))

    def set_factors_point2(self, family, i0, i1, mu, cov):
        """PriorPoint2 (i1=None) / Point2Point2 / Pose2Point2: MvNormal(mu[2], cov[2][2])"""
        i0, mu, cov = _i32(i0), _f64(mu), _f64(cov)
        i1 = None if i1 is None else _i32(i1)
        self._ck(self._lib.rome_b200_set_factors_point2(self._h, family, len(i0), self._ip(i0),
                                                        None if i1 is None else self._ip(i1), self._dp(mu),
                                                        self._dp(cov)))

    def set_factors_gaussian(self, family, i0, i1, mu, cov):
        """any family with one MvNormal(mu[dm], cov[dm][dm]) belief; i1=None for priors"""
        i0, mu, cov = _i32(i0), _f64(mu), _f64(cov)
        i1 = None if i1 is None else _i32(i1)
        self._ck(self._lib.rome_b200_set_factors_gaussian(self._h, family, len(i0), self._ip(i0),
                                                          None if i1 is None else self._ip(i1), self._dp(mu),
                                                          self._dp(cov)))

    def set_factors_ternary(self, family, i0, i1, i2, mu, cov):
        """Pose3Pose3RotOffset (i2: Rotation3 variables) / Pose3Pose3Transform (i2: Pose3 variables): MvNormal(mu[6], cov)"""
        i0, i1, i2, mu, cov = _i32(i0), _i32(i1), _i32(i2), _f64(mu), _f64(cov)
        self._ck(self._lib.rome_b200_set_factors_ternary(self._h, family, len(i0), self._ip(i0), self._ip(i1),
                                                         self._ip(i2), self._dp(mu), self._dp(cov)))

    def set_factors_scalar(self, family, i0, i1, belief):
        """Pose2Point2Range / Point2Point2Range / Pose2Point2Bearing: Normal(mean, sigma) rows [nF][2]"""
        i0, i1, belief = _i32(i0), _i32(i1), _f64(belief)
        self._ck(self._lib.rome_b200_set_factors_scalar(self._h, family, len(i0), self._ip(i0), self._ip(i1),
                                                        self._dp(belief)))

    def num_factors(self, family: int) -> int:
        return self._lib.rome_b200_num_factors(self._h, family)

    # -- hot path ------------------------------------------------------------------------------
    @staticmethod
    def _buffers(meas, meas_out, res, prop_fwd, prop_bwd, stats, jac) -> Buffers:
        return Buffers(_ptr(meas), _ptr(meas_out), _ptr(res), _ptr(prop_fwd), _ptr(prop_bwd), _ptr(stats), _ptr(jac))

    def eval(self, family, flags, *, seed=0, stream_id=0, first=0, count=-1, meas=None, meas_out=None, res=None,
             prop_fwd=None, prop_bwd=None, stats=None, jac=None):
        """Asynchronous launch on the ctx stream; buffers are DEVICE pointers / CUDA tensors."""
        b = self._buffers(meas, meas_out, res, prop_fwd, prop_bwd, stats, jac)
        self._ck(self._lib.rome_b200_eval(self._h, family, flags, seed, stream_id, first, count, C.byref(b)))

    def eval_host(self, family, flags, *, seed=0, stream_id=0, first=0, count=-1, meas=None, meas_out=None,
                  res=None, prop_fwd=None, prop_bwd=None, stats=None, jac=None, sync=True):
        """Call with HOST buffers (numpy float32 / pinned tensors).  sync=False only enqueues (pinned buffers
        required); call synchronize() before reading the outputs."""
        b = self._buffers(meas, meas_out, res, prop_fwd, prop_bwd, stats, jac)
        fn = self._lib.rome_b200_eval_host if sync else self._lib.rome_b200_eval_host_async
        self._ck(fn(self._h, family, flags, seed, stream_id, first, count, C.byref(b)))

    def alloc_host_outputs(self, family, flags):
        """numpy float32 output arrays, shaped for all factors of the family, for eval_host."""
        vt0, _, dm, dr, ns, dj, dfwd, dbwd = FAMILY[family]
        nF = self.num_factors(family)
        Np = self.particles_device(vt0)[5]
        out = {}
        if flags & (WRITE_MEAS | L.DECONV):
            out["meas_out"] = np.zeros((nF, Np, dm), np.float32)
        if flags & RESIDUAL:
            out["res"] = np.zeros((nF, Np, dr), np.float32)
        if flags & PROPOSAL_FWD:
            out["prop_fwd"] = np.zeros((nF, Np, dfwd), np.float32)
        if flags & PROPOSAL_BWD:
            out["prop_bwd"] = np.zeros((nF, Np, dbwd), np.float32)
        if flags & STATS:
            out["stats"] = np.zeros((nF, ns), np.float32)
        if flags & JACOBIAN:
            out["jac"] = np.zeros((nF, Np, dj), np.float32)
        return out

    # -- belief update: product of proposal densities -----------------------------
    def set_product_plan(self, vartype: int, var_offsets, src_buf, src_row):
        """CSR plan: variable v multiplies sources [var_offsets[v], var_offsets[v+1]); source = (buffer index, row)"""
        off, sb, sr = _i32(var_offsets), _i32(src_buf), _i32(src_row)
        self._ck(self._lib.rome_b200_set_product_plan(self._h, vartype, len(off) - 1, self._ip(off), self._ip(sb),
                                                      self._ip(sr)))

    def product(self, vartype: int, bufs, *, seed=0, stream_id=0, gibbs_iters=0, reanchor=True, bw_out=None, manifold=True):
        """bufs: device pointers / CUDA tensors of the proposal buffers the plan indexes; updates the particle store.
        manifold: Pose3 rotations are multiplied in the tangent space at the anchor rotation (ROME_B200_PRODUCT_MANIFOLD)"""
        arr = (C.c_void_p * max(1, len(bufs)))(*[_ptr(b) for b in bufs])
        self._ck(self._lib.rome_b200_product(self._h, vartype, len(bufs), arr, seed, stream_id, gibbs_iters,
                                             (L.PRODUCT_REANCHOR if reanchor else 0) | (L.PRODUCT_MANIFOLD if manifold else 0),
                                             _ptr(bw_out)))

    def reanchor(self, vartype: int):
        self._ck(self._lib.rome_b200_reanchor(self._h, vartype))

    # -- multi-GPU: fused all-gather of forward proposals ---------------------------------------
    def set_peer_proposals(self, family: int, peer_ptrs):
        """peer_ptrs: device pointers (ints) to the peers' identically shaped prop_fwd buffers; [] clears"""
        arr = (C.c_void_p * max(1, len(peer_ptrs)))(*peer_ptrs)
        self._ck(self._lib.rome_b200_set_peer_proposals(self._h, family, len(peer_ptrs), arr))

    # -- multi-GPU, owner-sharded: per-factor proposal destinations + halo particle blocks ---------------
    def set_proposal_destinations(self, family: int, direction: int, row_ptrs):
        """row_ptrs[f]: device address (int, 0 = default row) the forward (0) / backward (1) proposal row of factor f
        is written to; [] clears"""
        n = len(row_ptrs)
        arr = (C.c_void_p * max(1, n))(*[int(p) or None for p in row_ptrs])
        self._ck(self._lib.rome_b200_set_proposal_destinations(self._h, family, direction, n, arr))

    def set_step_barrier(self, state_ptr: int, peer_slot_ptrs):
        """rank barrier fused into eval launches flagged BARRIER_WAIT / BARRIER_SIGNAL; [] clears"""
        arr = (C.c_void_p * max(1, len(peer_slot_ptrs)))(*peer_slot_ptrs)
        self._ck(self._lib.rome_b200_set_step_barrier(self._h, state_ptr or None, arr, len(peer_slot_ptrs)))

    def set_barrier_range(self, family: int, first: int, count: int):
        """only factors [first, first + count) depend on the peers: BARRIER_WAIT is passed right before the first of them
        is fetched (count < 0: every factor, the default)"""
        self._ck(self._lib.rome_b200_set_barrier_range(self._h, family, first, count))

    def set_owned_variables(self, vartype: int, n_owned: int):
        """product / reanchor update only variables [0, n_owned) (the rest are halo copies); -1: all"""
        self._ck(self._lib.rome_b200_set_owned_variables(self._h, vartype, n_owned))

    def set_halo_plan(self, vartype: int, src_vars, dst_block_ptrs):
        src = _i32(src_vars)
        arr = (C.c_void_p * max(1, len(src)))(*[int(p) for p in dst_block_ptrs])
        self._ck(self._lib.rome_b200_set_halo_plan(self._h, vartype, len(src), self._ip(src), arr))

    def push_halo(self, vartype: int):
        self._ck(self._lib.rome_b200_push_halo(self._h, vartype))

    # -- GPU-side barrier between ranks over NVLink peer memory (closes the fused exchange) ---------------
    def peer_state_alloc(self) -> int:
        """zeroed device state buffer (flag slots + epochs) for peer_signal / peer_wait"""
        p = self.malloc_device(L.PEER_STATE_WORDS * 4)
        self.memcpy_h2d(p, np.zeros(L.PEER_STATE_WORDS, np.uint32))
        return p

    def peer_signal(self, state_ptr: int, peer_slot_ptrs):
        arr = (C.c_void_p * max(1, len(peer_slot_ptrs)))(*peer_slot_ptrs)
        self._ck(self._lib.rome_b200_peer_signal(self._h, state_ptr, arr, len(peer_slot_ptrs)))

    def peer_wait(self, state_ptr: int, n_slots: int):
        slots = np.arange(n_slots, dtype=np.int32)
        self._ck(self._lib.rome_b200_peer_wait(self._h, state_ptr, self._ip(slots), n_slots))

    def peer_barrier(self, state_ptr: int, peer_slot_ptrs):
        """peer_signal + peer_wait in one launch (rome_b200_peer_barrier)"""
        arr = (C.c_void_p * max(1, len(peer_slot_ptrs)))(*peer_slot_ptrs)
        self._ck(self._lib.rome_b200_peer_barrier(self._h, state_ptr, arr, len(peer_slot_ptrs)))

    def peer_gave_up(self, state_ptr: int) -> bool:
        v = C.c_int()
        self._ck(self._lib.rome_b200_peer_status(self._h, state_ptr, C.byref(v)))
        return bool(v.value)

    def malloc_device(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self._lib.rome_b200_malloc_device(self._h, nbytes, C.byref(p)))
        return p.value

    def free_device(self, ptr: int):
        self._ck(self._lib.rome_b200_free_device(self._h, ptr))

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self._lib.rome_b200_ipc_export(self._h, ptr, buf))
        return buf.raw

    def ipc_import(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._ck(self._lib.rome_b200_ipc_import(self._h, handle, C.byref(p)))
        return p.value

    def memcpy_h2d(self, dst_ptr: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._ck(self._lib.rome_b200_memcpy_h2d(self._h, dst_ptr, src.ctypes.data, src.nbytes))

    def memcpy_d2h(self, dst: np.ndarray, src_ptr: int):
        self._ck(self._lib.rome_b200_memcpy_d2h(self._h, dst.ctypes.data, src_ptr, dst.nbytes))

    # -- CUDA graphs ---------------------------------------------------------------------------
    def graph_begin(self):
        self._ck(self._lib.rome_b200_graph_begin(self._h))

    def graph_end(self) -> int:
        g = C.c_int()
        self._ck(self._lib.rome_b200_graph_end(self._h, C.byref(g)))
        return g.value

    def graph_launch(self, graph_id: int):
        self._ck(self._lib.rome_b200_graph_launch(self._h, graph_id))


# ----------------------------------------------------------------------------------------------
# layout helpers (host side, numpy).  Device rows are particle-major like the reference's own arrays
# ([nF][Npad][d] vs [nF][N][d]), so these only pad/slice, subtract/add the Float64 anchor and change dtype.
# ----------------------------------------------------------------------------------------------
def meas_to_offsets(meas, mu, Npad=None) -> np.ndarray:
    """Float64 samples [nF][N][dm] (reference `sampleFactor` output, coordinates) -> float32 offsets from
    the factor mean, [nF][Npad][dm]."""
    meas, mu = _f64(meas), _f64(mu)
    nF, N, dm = meas.shape
    Np = Npad or npad(N)
    out = np.zeros((nF, Np, dm), np.float32)
    out[:, :N, :] = meas - mu[:, None, :]
    return out


def offsets_to_meas(off, mu, N) -> np.ndarray:
    return np.asarray(off, dtype=np.float64)[:, :N, :] + _f64(mu)[:, None, :]


def rows_to_particle_major(rows, N) -> np.ndarray:
    """device rows [nF][Npad][d] float32 -> [nF][N][d] Float64"""
    return np.asarray(rows, dtype=np.float64)[:, :N, :].copy()


def dequantized_particles(anchors, offsets, N, wrap_dim=None) -> np.ndarray:
    """The exact Float64 values the kernels see: anchor + float32 offset, [nvars][N][d]."""
    return _f64(anchors)[:, None, :] + np.asarray(offsets, dtype=np.float64)[:, :N, :]
