"""ctypes binding of librome_b200.so (include/rome_b200.h) -- the same C ABI a Julia `ccall`
binds.  There is no fallback: a missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "librome_b200.so")

# enums of include/rome_b200.h
OK, BAD_ARG, CUDA_ERROR, SHAPE_MISMATCH, NOT_SET, NO_DEVICE = 0, -1, -2, -3, -4, -5
POSE2, POINT2, POSE3, POINT3, ROTATION3 = 0, 1, 2, 3, 4
POSE2POSE2, PRIORPOSE2, BEARINGRANGE, POSE3POSE3, PRIORPOSE3 = 0, 1, 2, 3, 4
PRIORPOINT2, POINT2POINT2, POSE2POINT2, POSE2POINT2RANGE, POINT2POINT2RANGE, POSE2POINT2BEARING = 5, 6, 7, 8, 9, 10
PRIORPOINT3, POINT3POINT3, POSE3POSE3XYYAW, POSE3POSE3ROTATION, POSE3POSE3UNITTRANS = 11, 12, 13, 14, 15
POSE3POSE3ROTOFFSET, POSE3POSE3TRANSFORM = 16, 17   # families with a third variable
RESIDUAL, PROPOSAL_FWD, PROPOSAL_BWD, STATS, SAMPLE, WRITE_MEAS, JACOBIAN, INDEPENDENT = 1, 2, 4, 8, 16, 32, 64, 128
DECONV = 256
PRECISE = 512
ROUTED_ONLY, BARRIER_WAIT, BARRIER_SIGNAL = 1024, 2048, 4096
PRODUCT_REANCHOR, MAX_PRODUCT_SOURCES, MAX_PRODUCT_BUFFERS = 1, 32, 16
PRODUCT_MANIFOLD = 2
PEER_STATE_WORDS = 16

# every symbol include/rome_b200.h declares (tests check the library exports each one)
SYMBOLS = [
    "rome_b200_version", "rome_b200_create", "rome_b200_destroy", "rome_b200_last_error", "rome_b200_set_stream",
    "rome_b200_synchronize", "rome_b200_family_dims", "rome_b200_vartype_dim", "rome_b200_npad", "rome_b200_plan_query",
    "rome_b200_set_particles", "rome_b200_set_particles_anchored", "rome_b200_get_particles", "rome_b200_particles_device", "rome_b200_adopt_proposal",
    "rome_b200_set_factors_pose2pose2", "rome_b200_set_factors_priorpose2", "rome_b200_set_factors_bearingrange",
    "rome_b200_set_factors_pose3pose3", "rome_b200_set_factors_priorpose3", "rome_b200_set_factors_point2",
    "rome_b200_set_factors_scalar", "rome_b200_set_factors_gaussian", "rome_b200_set_factors_ternary", "rome_b200_num_factors",
    "rome_b200_eval", "rome_b200_eval_host", "rome_b200_eval_host_async", "rome_b200_set_product_plan", "rome_b200_product",
    "rome_b200_reanchor", "rome_b200_set_peer_proposals", "rome_b200_set_proposal_destinations",
    "rome_b200_set_step_barrier", "rome_b200_set_barrier_range", "rome_b200_set_owned_variables", "rome_b200_set_halo_plan", "rome_b200_push_halo", "rome_b200_ipc_export",
    "rome_b200_ipc_import", "rome_b200_ipc_close", "rome_b200_peer_signal", "rome_b200_peer_wait", "rome_b200_peer_barrier", "rome_b200_peer_status", "rome_b200_graph_begin", "rome_b200_graph_end",
    "rome_b200_graph_launch", "rome_b200_malloc_device", "rome_b200_free_device", "rome_b200_malloc_host",
    "rome_b200_free_host", "rome_b200_memcpy_h2d", "rome_b200_memcpy_d2h", "rome_b200_launch_count",
]


class Buffers(C.Structure):
    """struct rome_b200_buffers"""
    _fields_ = [("meas", C.c_void_p), ("meas_out", C.c_void_p), ("res", C.c_void_p), ("prop_fwd", C.c_void_p),
                ("prop_bwd", C.c_void_p), ("stats", C.c_void_p), ("jac", C.c_void_p)]


class RomeB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rome_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                          " (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    vp, i, u32, u64, dp, ip32 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.rome_b200_version.restype = i
    lib.rome_b200_create.argtypes = [i, C.POINTER(vp)]
    lib.rome_b200_destroy.argtypes = [vp]
    lib.rome_b200_last_error.restype = C.c_char_p
    lib.rome_b200_last_error.argtypes = [vp]
    lib.rome_b200_set_stream.argtypes = [vp, vp]
    lib.rome_b200_synchronize.argtypes = [vp]
    lib.rome_b200_family_dims.argtypes = [i] + [C.POINTER(i)] * 4
    lib.rome_b200_vartype_dim.argtypes = [i]
    lib.rome_b200_npad.argtypes = [i]
    lib.rome_b200_plan_query.argtypes = [i, u32, i] + [C.POINTER(i)] * 5
    lib.rome_b200_set_particles.argtypes = [vp, i, i, i, vp]
    lib.rome_b200_set_particles_anchored.argtypes = [vp, i, i, i, vp, vp]
    lib.rome_b200_get_particles.argtypes = [vp, i, vp]
    lib.rome_b200_particles_device.argtypes = [vp, i, C.POINTER(vp)] + [C.POINTER(i)] * 5
    lib.rome_b200_adopt_proposal.argtypes = [vp, i, i, vp, i]
    lib.rome_b200_set_factors_pose2pose2.argtypes = [vp, i, ip32, ip32, dp, dp]
    lib.rome_b200_set_factors_priorpose2.argtypes = [vp, i, ip32, dp, dp]
    lib.rome_b200_set_factors_bearingrange.argtypes = [vp, i, ip32, ip32, dp, dp]
    lib.rome_b200_set_factors_pose3pose3.argtypes = [vp, i, ip32, ip32, dp, dp]
    lib.rome_b200_set_factors_priorpose3.argtypes = [vp, i, ip32, dp, dp]
    lib.rome_b200_set_factors_point2.argtypes = [vp, i, i, ip32, ip32, dp, dp]
    lib.rome_b200_set_factors_scalar.argtypes = [vp, i, i, ip32, ip32, dp]
    lib.rome_b200_set_factors_gaussian.argtypes = [vp, i, i, ip32, ip32, dp, dp]
    lib.rome_b200_set_factors_ternary.argtypes = [vp, i, i, ip32, ip32, ip32, dp, dp]
    lib.rome_b200_num_factors.argtypes = [vp, i]
    lib.rome_b200_eval.argtypes = [vp, i, u32, u64, u32, i, i, C.POINTER(Buffers)]
    lib.rome_b200_eval_host.argtypes = [vp, i, u32, u64, u32, i, i, C.POINTER(Buffers)]
    lib.rome_b200_eval_host_async.argtypes = [vp, i, u32, u64, u32, i, i, C.POINTER(Buffers)]
    lib.rome_b200_set_product_plan.argtypes = [vp, i, i, ip32, ip32, ip32]
    lib.rome_b200_product.argtypes = [vp, i, i, C.POINTER(vp), u64, u32, i, u32, vp]
    lib.rome_b200_reanchor.argtypes = [vp, i]
    lib.rome_b200_set_peer_proposals.argtypes = [vp, i, i, C.POINTER(vp)]
    lib.rome_b200_set_proposal_destinations.argtypes = [vp, i, i, i, C.POINTER(vp)]
    lib.rome_b200_set_step_barrier.argtypes = [vp, vp, C.POINTER(vp), i]
    lib.rome_b200_set_barrier_range.argtypes = [vp, i, i, i]
    lib.rome_b200_set_owned_variables.argtypes = [vp, i, i]
    lib.rome_b200_set_halo_plan.argtypes = [vp, i, i, ip32, C.POINTER(vp)]
    lib.rome_b200_push_halo.argtypes = [vp, i]
    lib.rome_b200_ipc_export.argtypes = [vp, vp, C.c_char_p]
    lib.rome_b200_ipc_import.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.rome_b200_ipc_close.argtypes = [vp, vp]
    lib.rome_b200_peer_signal.argtypes = [vp, vp, C.POINTER(vp), i]
    lib.rome_b200_peer_wait.argtypes = [vp, vp, ip32, i]
    lib.rome_b200_peer_barrier.argtypes = [vp, vp, C.POINTER(vp), i]
    lib.rome_b200_peer_status.argtypes = [vp, vp, C.POINTER(i)]
    lib.rome_b200_graph_begin.argtypes = [vp]
    lib.rome_b200_graph_end.argtypes = [vp, C.POINTER(i)]
    lib.rome_b200_graph_launch.argtypes = [vp, i]
    lib.rome_b200_malloc_device.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.rome_b200_free_device.argtypes = [vp, vp]
    lib.rome_b200_malloc_host.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.rome_b200_free_host.argtypes = [vp, vp]
    lib.rome_b200_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    lib.rome_b200_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    lib.rome_b200_launch_count.restype = u64
    lib.rome_b200_launch_count.argtypes = [vp]
    _lib = lib
    return lib
