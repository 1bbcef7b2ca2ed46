"""ctypes binding of libldrones.so (include/ldrones.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

from ._abi import LD_ABI_VERSION, LdGaParams, LdParams, LdPeerTable, LdPidParams, LdRecord

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDRONES_LIB: development override (A/B builds of the kernel library); the product always loads the in-tree build
LIB_PATH = os.environ.get("LDRONES_LIB") or os.path.join(_HERE, "libldrones.so")

# every symbol include/ldrones.h declares
EXPORTS = (
    "ld_abi_version",
    "ld_last_error",
    "ld_device_count",
    "ld_wave_size",
    "ld_fly_mlp_f64_device",
    "ld_fly_mlp_f64_host",
    "ld_fly_mlp_f64_from_host",
    "ld_fly_mlp_f64_from_device",
    "ld_fly_mlp_f64_host_submit",
    "ld_fly_mlp_f64_host_wait",
    "ld_brain_forward_host",
    "ld_activation_host",
    "ld_fly_mlp_f64_trace_device",
    "ld_reduce_records_device",
    "ld_release_workspace",
    "ld_measure_fp64_peak",
    "ld_debug_math_device",
    "ld_debug_fly_profile_device",
    "ld_fly_pid_f64_device",
    "ld_fly_pid_f64_host",
    "ld_ga_step_device",
    "ld_ga_step_p2p_device",
    "ld_fly_mlp_f32_device",
    "ld_fly_mlp_f32_host",
    "ld_drone_step_host",
    "ld_track_best_device",
)


class LdError(RuntimeError):
    """A libldrones entry point returned a negative ld_status."""

    def __init__(self, code, message):
        super().__init__(f"libldrones error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load libldrones.so (built in-tree by learning_drones_b200.build). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m learning_drones_b200.build` "
            "(nvcc, sm_100a). learning_drones_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, dp, i64 = C.c_void_p, C.c_void_p, C.c_int64
    L.ld_abi_version.restype = C.c_int
    L.ld_last_error.restype = C.c_char_p
    L.ld_device_count.restype = C.c_int
    L.ld_wave_size.restype = C.c_int64
    L.ld_fly_mlp_f64_device.argtypes = [C.POINTER(LdParams), dp, i64, vp, dp, vp, vp]
    L.ld_fly_mlp_f64_device.restype = C.c_int
    L.ld_fly_mlp_f64_host.argtypes = [C.POINTER(LdParams), dp, i64, vp, dp]
    L.ld_fly_mlp_f64_host.restype = C.c_int
    L.ld_fly_mlp_f64_from_host.argtypes = [C.POINTER(LdParams), dp, dp, i64, vp, dp]
    L.ld_fly_mlp_f64_from_host.restype = C.c_int
    L.ld_fly_mlp_f64_from_device.argtypes = [C.POINTER(LdParams), dp, dp, i64, vp, dp, vp, vp]
    L.ld_fly_mlp_f64_from_device.restype = C.c_int
    L.ld_fly_mlp_f64_host_submit.argtypes = [C.POINTER(LdParams), dp, dp, i64, vp, dp, C.POINTER(C.c_int64)]
    L.ld_fly_mlp_f64_host_submit.restype = C.c_int
    L.ld_fly_mlp_f64_host_wait.argtypes = [i64]
    L.ld_fly_mlp_f64_host_wait.restype = C.c_int
    L.ld_brain_forward_host.argtypes = [C.POINTER(C.c_int32), C.c_int32, dp, i64, dp, i64, dp, C.c_double]
    L.ld_brain_forward_host.restype = C.c_int
    L.ld_activation_host.argtypes = [C.c_int, dp, i64, dp, C.c_double, C.c_double]
    L.ld_activation_host.restype = C.c_int
    L.ld_fly_mlp_f32_device.argtypes = [C.POINTER(LdParams), dp, i64, vp, dp, vp, vp]
    L.ld_fly_mlp_f32_device.restype = C.c_int
    L.ld_fly_mlp_f32_host.argtypes = [C.POINTER(LdParams), dp, i64, vp, dp]
    L.ld_fly_mlp_f32_host.restype = C.c_int
    L.ld_drone_step_host.argtypes = [C.POINTER(LdParams), C.c_int, dp, dp, dp, vp, i64]
    L.ld_drone_step_host.restype = C.c_int
    L.ld_fly_mlp_f64_trace_device.argtypes = [C.POINTER(LdParams), dp, vp, dp, vp]
    L.ld_fly_mlp_f64_trace_device.restype = C.c_int
    L.ld_reduce_records_device.argtypes = [vp, i64, vp, vp]
    L.ld_reduce_records_device.restype = C.c_int
    L.ld_track_best_device.argtypes = [vp, vp, C.POINTER(LdPeerTable), i64, vp, vp, vp, vp]
    L.ld_track_best_device.restype = C.c_int
    L.ld_release_workspace.restype = None
    L.ld_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ld_measure_fp64_peak.restype = C.c_int
    L.ld_debug_math_device.argtypes = [C.c_int, vp, vp, vp, vp, i64, vp]
    L.ld_debug_math_device.restype = C.c_int
    L.ld_debug_fly_profile_device.argtypes = [C.POINTER(LdParams), dp, i64, vp, vp, vp, vp, vp]
    L.ld_debug_fly_profile_device.restype = C.c_int
    L.ld_fly_pid_f64_device.argtypes = [C.POINTER(LdParams), C.POINTER(LdPidParams), dp, i64, vp, dp, vp]
    L.ld_fly_pid_f64_device.restype = C.c_int
    L.ld_fly_pid_f64_host.argtypes = [C.POINTER(LdParams), C.POINTER(LdPidParams), dp, i64, vp, dp]
    L.ld_fly_pid_f64_host.restype = C.c_int
    L.ld_ga_step_device.argtypes = [C.POINTER(LdGaParams), dp, vp, i64, dp, i64, i64, vp]
    L.ld_ga_step_device.restype = C.c_int
    L.ld_ga_step_p2p_device.argtypes = [C.POINTER(LdGaParams), C.POINTER(LdPeerTable), vp, i64, dp, i64, i64, vp]
    L.ld_ga_step_p2p_device.restype = C.c_int
    if L.ld_abi_version() != LD_ABI_VERSION:
        raise ImportError(f"libldrones ABI {L.ld_abi_version()} != python binding {LD_ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise LdError(rc, load().ld_last_error().decode(errors="replace"))
