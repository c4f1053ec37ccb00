"""ctypes front-end of the CPU oracle (``oracle/cbx_oracle.c``).  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module; ``marlon_b200`` never does.  The oracle exposes the same arrays as the CUDA
library's ``cbx_views`` (host memory here), so tests compare array for array.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from marlon_b200 import _abi  # noqa: E402  (struct definitions of include/cbx.h)

_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcbx_oracle.so")
    src = os.path.join(_HERE, "cbx_oracle.c")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "cbx.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libcbx_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_abi.Config), C.c_int64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_views.argtypes = [C.c_void_p, C.POINTER(_abi.Views)]
        L.orc_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_reset_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_step_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_notify_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.orc_stats_reset.argtypes = [C.c_void_p]
        L.orc_export_words.restype = C.c_int64
        L.orc_export_words.argtypes = [C.c_void_p]
        L.orc_export_state.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_firewall_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_sample_actions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        for f in (L.orc_l1_local, L.orc_l1_remote, L.orc_l1_connect):
            f.restype = C.c_double
        L.orc_l1_local.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_l1_remote.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_l1_connect.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


class OracleBatch:
    """Host-side batch with the same surface as ``marlon_b200.batch.Batch`` (arrays are numpy)."""

    def __init__(self, compiled, cfg: _abi.Config, n_envs: int):
        self._lib = lib()
        self.compiled = compiled
        self.cfg = cfg
        self.n_envs = int(n_envs)
        blob = np.ascontiguousarray(compiled.blob, dtype=np.uint32)
        self._h = self._lib.orc_create(blob.ctypes.data, blob.size, C.byref(cfg), self.n_envs)
        if not self._h:
            raise ValueError("oracle rejected the scenario/config (node count > maximum_node_count or credentials > maximum_total_credentials?)")
        if getattr(compiled, "fw_ext", None) is not None:  # rule-list alias groups (the `live` defender binding edits them)
            ext = np.ascontiguousarray(compiled.fw_ext, dtype=np.uint32)
            if self._lib.orc_set_firewall_tables(self._h, ext.ctypes.data, ext.size):
                raise ValueError("oracle rejected the firewall tables")
        v = _abi.Views()
        self._lib.orc_views(self._h, C.byref(v))
        self.views = v
        self.arrays = {}
        for name, (shape, dt) in _abi.view_specs(v, cfg).items():
            ptr = getattr(v, name)
            if not ptr:
                continue
            full = (self.n_envs,) + tuple(shape)
            count = int(np.prod(full)) if full else 1
            if count == 0:
                self.arrays[name] = np.zeros(full, dtype=dt)
                continue
            buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(C.cast(ptr, C.c_void_p).value)
            self.arrays[name] = np.frombuffer(buf, dtype=dt).reshape(full)
        sb = (C.c_double * _abi.STAT_COUNT).from_address(C.cast(v.episode_stats, C.c_void_p).value)
        self.stats = np.frombuffer(sb, dtype=np.float64)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.orc_destroy(h)

    def reset(self, mask=None, who=3):
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self._lib.orc_reset_ex(self._h, None if mask is None else mask.ctypes.data, who)

    def step(self, attacker_actions, defender_actions=None, scan_u=None, detect_u=None, who=3):
        aa = None if attacker_actions is None else np.ascontiguousarray(attacker_actions, dtype=np.int32)
        da = None if defender_actions is None else np.ascontiguousarray(defender_actions, dtype=np.int32)
        su = None if scan_u is None else np.ascontiguousarray(scan_u, dtype=np.float64)
        du = None if detect_u is None else np.ascontiguousarray(detect_u, dtype=np.float64)
        self._lib.orc_step_ex(self._h, None if aa is None else aa.ctypes.data, None if da is None else da.ctypes.data,
                              None if su is None else su.ctypes.data, None if du is None else du.ctypes.data, who)

    def notify_reset(self, who, last_reward=0.0, mask=None):
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self._lib.orc_notify_reset(self._h, None if mask is None else mask.ctypes.data, who, float(last_reward))

    def stats_reset(self):
        self._lib.orc_stats_reset(self._h)

    def sample_actions(self, seed=0, step=None):
        """CyberBattleEnv.sample_valid_action's distribution (+ a uniform defender action) for the current state of every env;
        the same draws as ``Batch.sample_actions`` (its k-th call uses step k, counted from 0)."""
        if step is None:
            step = getattr(self, "_sample_step", 0)
            self._sample_step = step + 1
        marlon = self.cfg.mode == _abi.MODE_MARLON
        att = np.zeros((self.n_envs, 10 if marlon else 5), dtype=np.int32)
        dfn = np.zeros((self.n_envs, 12), dtype=np.int32) if marlon and self.cfg.def_enabled else None
        self._lib.orc_sample_actions(self._h, att.ctypes.data, None if dfn is None else dfn.ctypes.data,
                                     int(seed) & 0xFFFFFFFFFFFFFFFF, int(step) & 0xFFFFFFFF)
        return att, dfn

    def export_state(self, begin=0, end=None):
        end = self.n_envs if end is None else end
        w = self._lib.orc_export_words(self._h)
        out = np.zeros((end - begin, w), dtype=np.int32)
        self._lib.orc_export_state(self._h, begin, end, out.ctypes.data)
        return out

    # L1 (AgentActions) entry points for the commandcontrol known-answer test
    def l1_local(self, env, node, v):
        o = C.c_int(0)
        return self._lib.orc_l1_local(self._h, env, node, v, C.byref(o)), o.value

    def l1_remote(self, env, src, tgt, v):
        o = C.c_int(0)
        return self._lib.orc_l1_remote(self._h, env, src, tgt, v, C.byref(o)), o.value

    def l1_connect(self, env, src, tgt, port, secret):
        o = C.c_int(0)
        return self._lib.orc_l1_connect(self._h, env, src, tgt, port, secret, C.byref(o)), o.value
