"""Loader of ``marlon_b200/libcbx.so`` (the C ABI of include/cbx.h) through ctypes.

There is no CPU fallback: if the shared library is missing or fails to load, importing anything that needs
it raises.  ``load()`` does not need a GPU (the symbol/ABI tests run on CPU); creating a batch does.
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcbx.so")
_LIB = None

# every symbol declared in include/cbx.h
SYMBOLS = [
    "cbx_last_error", "cbx_abi_version", "cbx_scenario_create", "cbx_scenario_destroy", "cbx_config_default",
    "cbx_batch_create", "cbx_batch_destroy", "cbx_batch_reset", "cbx_batch_step", "cbx_batch_step_host", "cbx_batch_step_i16", "cbx_batch_step_host_i16",
    "cbx_batch_sample_actions", "cbx_batch_views", "cbx_batch_stats_reset", "cbx_export_words",
    "cbx_batch_export_state", "cbx_batch_launch_count", "cbx_batch_enable_timing", "cbx_batch_step_kernel_ms",
    "cbx_abi_sizeof", "cbx_batch_step_ex", "cbx_batch_reset_ex", "cbx_batch_phase_cycles", "cbx_batch_notify_reset", "cbx_batch_kernel_info", "cbx_batch_create_multi", "cbx_batch_export_words", "cbx_gae",
    "cbx_batch_host_prepare", "cbx_batch_fetch_host_layout", "cbx_batch_fetch_host", "cbx_batch_step_host_ex", "cbx_batch_tile_counter", "cbx_scenario_set_firewall_tables",
]


class CbxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"cbx error {code}: {message}")
        self.code = code
        self.message = message


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("CBX_LIB") or SO_PATH  # CBX_LIB: an experiment build of the same sources (marlon_b200/build.py)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m marlon_b200.build` "
                          "(nvcc, sm_100a). marlon_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, i64, i32p = C.c_void_p, C.c_int64, C.c_void_p
    L.cbx_last_error.restype = C.c_char_p
    L.cbx_abi_version.restype = C.c_int
    L.cbx_scenario_create.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.cbx_scenario_destroy.argtypes = [vp]
    L.cbx_scenario_set_firewall_tables.argtypes = [vp, vp, C.c_size_t]
    L.cbx_config_default.argtypes = [C.POINTER(_abi.Config)]
    L.cbx_batch_create.argtypes = [vp, i64, C.POINTER(_abi.Config), C.c_int, C.POINTER(vp)]
    L.cbx_batch_destroy.argtypes = [vp]
    L.cbx_batch_reset.argtypes = [vp, vp, vp]
    L.cbx_batch_step.argtypes = [vp, i32p, i32p, C.POINTER(_abi.Tape), vp]
    L.cbx_batch_reset_ex.argtypes = [vp, vp, C.c_int, vp]
    L.cbx_batch_step_ex.argtypes = [vp, i32p, i32p, C.POINTER(_abi.Tape), C.c_int, vp]
    L.cbx_batch_step_host.argtypes = [vp, i32p, i32p, vp, C.c_size_t, vp]
    L.cbx_batch_step_i16.argtypes = [vp, vp, vp, vp]
    L.cbx_batch_step_host_i16.argtypes = [vp, vp, vp, vp, C.c_size_t, vp]
    L.cbx_batch_sample_actions.argtypes = [vp, i32p, i32p, C.c_uint64, vp]
    L.cbx_batch_views.argtypes = [vp, C.POINTER(_abi.Views)]
    L.cbx_batch_stats_reset.argtypes = [vp, vp]
    L.cbx_export_words.restype = i64
    L.cbx_export_words.argtypes = [vp, C.POINTER(_abi.Config)]
    L.cbx_batch_export_state.argtypes = [vp, i64, i64, vp, vp]
    L.cbx_batch_launch_count.restype = i64
    L.cbx_batch_launch_count.argtypes = [vp]
    L.cbx_batch_enable_timing.argtypes = [vp, C.c_int]
    L.cbx_batch_step_kernel_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.cbx_batch_notify_reset.argtypes = [vp, vp, C.c_int, C.c_double, vp]
    L.cbx_batch_phase_cycles.argtypes = [vp, C.c_int, vp]
    L.cbx_batch_kernel_info.argtypes = [vp, vp]
    L.cbx_batch_create_multi.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
    L.cbx_batch_export_words.restype = C.c_int64
    L.cbx_batch_export_words.argtypes = [vp]
    L.cbx_gae.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, i64, vp, vp, vp]
    L.cbx_batch_host_prepare.argtypes = [vp]
    L.cbx_batch_tile_counter.argtypes = [vp, vp]
    L.cbx_batch_step_host_ex.argtypes = [vp, vp, vp, C.c_int, vp, C.c_size_t, C.c_int, vp]
    L.cbx_batch_fetch_host_layout.restype = C.c_int64
    L.cbx_batch_fetch_host_layout.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int64)]
    L.cbx_batch_fetch_host.argtypes = [vp, C.c_uint32, vp, C.c_size_t, vp]
    L.cbx_abi_sizeof.restype = C.c_size_t
    L.cbx_abi_sizeof.argtypes = [C.c_int]
    if L.cbx_abi_version() != _abi.ABI_VERSION:
        raise ImportError(f"libcbx.so ABI {L.cbx_abi_version()} != python mirror {_abi.ABI_VERSION}")
    if L.cbx_abi_sizeof(0) != C.sizeof(_abi.Config) or L.cbx_abi_sizeof(1) != C.sizeof(_abi.Views):
        raise ImportError("struct layout mismatch between include/cbx.h and marlon_b200/_abi.py")
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise CbxError(rc, load().cbx_last_error().decode(errors="replace"))
