"""ctypes binding of libkmpc.so (include/kmpc.h).  There is no CPU implementation behind this module: if the CUDA
library is missing or no GPU is present, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("KMPC_LIB") or os.path.join(_HERE, "libkmpc.so")  # KMPC_LIB: alternative build (kernel tuning only)

KMPC_VERSION = 200
LAYOUT_INSTANCE_MAJOR, LAYOUT_BATCH_MINOR = 0, 1
COST_README, COST_CODE_LITERAL = 0, 1
NO_BOUND = 1e19

# symbols include/kmpc.h declares (checked by tests/test_abi.py)
SYMBOLS = ["kmpc_version", "kmpc_workspace_bytes", "kmpc_create", "kmpc_destroy", "kmpc_last_error", "kmpc_solve",
           "kmpc_solve_tracks", "kmpc_solve_host", "kmpc_solve_host_into", "kmpc_host_sync", "kmpc_pinned_alloc", "kmpc_pinned_free",
           "kmpc_host_result", "kmpc_shared_buffer_create", "kmpc_shared_buffer_open", "kmpc_shared_buffer_close", "kmpc_enable_peer",
           "kmpc_agent_handoff", "kmpc_closed_loop", "kmpc_environment_loop", "kmpc_select_obstacles", "kmpc_predict_tracks",
           "kmpc_map_distance", "kmpc_map_to_circles", "kmpc_set_queue_order", "kmpc_set_timing", "kmpc_get_stats", "kmpc_measure_fp64_peak"]
IPC_HANDLE_BYTES = 64


class KmpcConfig(C.Structure):
    _fields_ = [("N", C.c_int32), ("O_max", C.c_int32), ("cost_mode", C.c_int32), ("goal_k_lo", C.c_int32),
                ("goal_k_hi", C.c_int32), ("max_iter", C.c_int32), ("B_max", C.c_int32), ("layout", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32), ("T", C.c_double), ("W", C.c_double * 3),
                ("Wv_neg", C.c_double), ("Wv_pos", C.c_double), ("Ww", C.c_double), ("lo", C.c_double * 4),
                ("hi", C.c_double * 4), ("tol", C.c_double)]


class KmpcStats(C.Structure):
    _fields_ = [("last_kernel_ms", C.c_double), ("launches", C.c_int64), ("slots", C.c_int32), ("blocks", C.c_int32),
                ("threads_per_block", C.c_int32), ("sm_count", C.c_int32), ("trips", C.c_int64), ("warp_path", C.c_int32),
                ("reserved", C.c_int32)]


class KmpcError(RuntimeError):
    pass


_lib = None


def load():
    """Load libkmpc.so.  Raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise KmpcError(f"{SO_PATH} not found: build it with `python -m kiss_mpc_b200.build` (nvcc, sm_100a). "
                        "kiss_mpc_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (device or host)
    L.kmpc_version.restype = C.c_int
    L.kmpc_workspace_bytes.restype = C.c_size_t
    L.kmpc_workspace_bytes.argtypes = [C.POINTER(KmpcConfig)]
    L.kmpc_create.restype = C.c_int
    L.kmpc_create.argtypes = [C.POINTER(KmpcConfig), C.POINTER(vp)]
    L.kmpc_destroy.restype = None
    L.kmpc_destroy.argtypes = [vp]
    L.kmpc_last_error.restype = C.c_char_p
    L.kmpc_last_error.argtypes = [vp]
    solve_args = [vp, C.c_int, dp, dp, dp, dp, dp, C.c_int, C.c_double, dp, C.c_double, dp, dp, dp, ip, ip]
    L.kmpc_solve.restype = C.c_int
    L.kmpc_solve.argtypes = solve_args + [vp]
    L.kmpc_solve_tracks.restype = C.c_int
    L.kmpc_solve_tracks.argtypes = solve_args + [vp]
    L.kmpc_solve_host.restype = C.c_int
    L.kmpc_solve_host.argtypes = solve_args
    L.kmpc_solve_host_into.restype = C.c_int
    L.kmpc_solve_host_into.argtypes = solve_args
    L.kmpc_host_sync.restype = C.c_int
    L.kmpc_host_sync.argtypes = [vp]
    L.kmpc_pinned_alloc.restype = C.c_int
    L.kmpc_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.kmpc_pinned_free.restype = C.c_int
    L.kmpc_pinned_free.argtypes = [vp]
    L.kmpc_shared_buffer_create.restype = C.c_int
    L.kmpc_shared_buffer_create.argtypes = [vp, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    L.kmpc_shared_buffer_open.restype = C.c_int
    L.kmpc_shared_buffer_open.argtypes = [vp, C.c_void_p, C.POINTER(C.c_void_p)]
    L.kmpc_shared_buffer_close.restype = C.c_int
    L.kmpc_shared_buffer_close.argtypes = [vp, vp, C.c_int]
    L.kmpc_enable_peer.restype = C.c_int
    L.kmpc_enable_peer.argtypes = [vp, C.c_int]
    L.kmpc_host_result.restype = C.c_int
    L.kmpc_host_result.argtypes = [vp] + [C.POINTER(C.c_void_p)] * 5
    L.kmpc_agent_handoff.restype = C.c_int
    L.kmpc_agent_handoff.argtypes = [vp, C.c_int, dp, dp, dp, dp, vp]
    L.kmpc_closed_loop.restype = C.c_int
    L.kmpc_closed_loop.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, dp, ip, ip, ip, C.c_double, C.c_double, vp]
    L.kmpc_environment_loop.restype = C.c_int
    L.kmpc_environment_loop.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, C.c_int, dp, dp, C.c_int, C.c_int, dp, dp, dp, dp, C.c_int, C.c_int,
                                        C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, dp, ip, ip, ip, ip, ip,
                                        C.c_double, C.c_double, vp]
    L.kmpc_select_obstacles.restype = C.c_int
    L.kmpc_select_obstacles.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, dp, ip, ip, dp, vp]
    L.kmpc_predict_tracks.restype = C.c_int
    L.kmpc_predict_tracks.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, dp, dp, dp, C.c_double, C.c_int, C.c_double, C.c_double, dp, vp]
    L.kmpc_map_distance.restype = C.c_int
    L.kmpc_map_distance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.kmpc_map_to_circles.restype = C.c_int
    L.kmpc_map_to_circles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    L.kmpc_set_queue_order.restype = C.c_int
    L.kmpc_set_queue_order.argtypes = [vp, C.c_int]
    L.kmpc_set_timing.restype = C.c_int
    L.kmpc_set_timing.argtypes = [vp, C.c_int]
    L.kmpc_get_stats.restype = C.c_int
    L.kmpc_get_stats.argtypes = [vp, C.POINTER(KmpcStats)]
    L.kmpc_measure_fp64_peak.restype = C.c_int
    L.kmpc_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    if L.kmpc_version() != KMPC_VERSION:
        raise KmpcError(f"libkmpc.so version {L.kmpc_version()} != binding version {KMPC_VERSION}: rebuild")
    _lib = L
    return L


def check(rc: int, handle=None, what: str = "kmpc"):
    if rc != 0:
        msg = load().kmpc_last_error(handle)
        raise KmpcError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
