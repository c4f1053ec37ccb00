"""``Batch``: n independent environments resident in HBM, stepped by one fused kernel launch.

Every observation / reward / flag array is a zero-copy ``torch.Tensor`` view of library-owned device memory
(``__cuda_array_interface__``; the tensors export DLPack like any torch tensor).  Work is enqueued on torch's
current stream of the batch's device, so it orders with the caller's torch ops without extra synchronisation.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _abi, _lib
from .scenario import CompiledScenario

_TYPESTR = {"int32": "<i4", "int8": "|i1", "uint8": "|u1", "uint32": "<u4", "float32": "<f4", "float64": "<f8"}


def narrow_actions(actions, out: Optional[np.ndarray] = None) -> np.ndarray:
    """MultiDiscrete actions as the policy returns them (int64 from SB3, ``baseline_marlon_agent.py:118-131``) -> int16
    elements for ``Batch.step_host`` / ``cbx_batch_step_host_i16``.  `out` may be a page-locked array
    (``torch.empty(shape, dtype=torch.int16, pin_memory=True).numpy()``) that the kernel then reads in place.  Values that
    do not fit raise instead of wrapping: no component of either action space comes near 32 768 (the largest is
    ``maximum_total_credentials``), so an overflow means a corrupted action, not a big one."""
    a = np.asarray(actions)
    if a.dtype.kind not in "iu":
        raise TypeError(f"integer actions expected, got {a.dtype}")
    if a.size and (int(a.max()) > np.iinfo(np.int16).max or int(a.min()) < np.iinfo(np.int16).min):
        raise OverflowError("action component outside the int16 range")
    if out is None:
        return np.ascontiguousarray(a, dtype=np.int16)
    if out.dtype != np.int16 or out.shape != a.shape:
        raise ValueError(f"out must be int16 of shape {a.shape}")
    np.copyto(out, a, casting="unsafe")
    return out


class _DevArray:
    def __init__(self, ptr, shape, dtype):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": _TYPESTR[dtype], "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class Batch:
    def __init__(self, compiled, cfg: _abi.Config, n_envs, device: int = 0):
        """`compiled`: one CompiledScenario, or a list of them with `n_envs` the list of envs per scenario (envs grouped by
        scenario, every group but the last a multiple of 32: ``cbx_batch_create_multi``; e.g. CyberBattleRandom networks)."""
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("marlon_b200.Batch needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._torch = torch
        self._L = _lib.load()
        multi = isinstance(compiled, (list, tuple))
        self.scenarios = list(compiled) if multi else [compiled]
        self.envs_per_scenario = [int(x) for x in n_envs] if multi else [int(n_envs)]
        if len(self.scenarios) != len(self.envs_per_scenario):
            raise ValueError("one env count per scenario")
        self.compiled, self.cfg, self.n_envs, self.device = self.scenarios[0], cfg, sum(self.envs_per_scenario), int(device)
        self.torch_device = torch.device("cuda", self.device)
        self._scns = []
        self._h = C.c_void_p()
        try:
            for comp in self.scenarios:
                blob = comp.tobytes()
                h = C.c_void_p()
                _lib.check(self._L.cbx_scenario_create(blob, len(blob), C.byref(h)))
                self._scns.append(h)
                if getattr(comp, "fw_ext", None) is not None:  # rule-list alias groups: what the `live` defender binding edits
                    ext = np.ascontiguousarray(comp.fw_ext, dtype="<u4").tobytes()
                    _lib.check(self._L.cbx_scenario_set_firewall_tables(h, ext, len(ext)))
            with torch.cuda.device(self.device):
                torch.cuda.init()
                if multi:
                    arr = (C.c_void_p * len(self._scns))(*[h.value for h in self._scns])
                    cnt = (C.c_int64 * len(self._scns))(*self.envs_per_scenario)
                    _lib.check(self._L.cbx_batch_create_multi(arr, len(self._scns), cnt, C.byref(cfg), self.device, C.byref(self._h)))
                else:
                    _lib.check(self._L.cbx_batch_create(self._scns[0], self.n_envs, C.byref(cfg), self.device, C.byref(self._h)))
        except Exception:
            for h in self._scns:
                self._L.cbx_scenario_destroy(h)
            self._scns = []
            raise
        self._scn = self._scns[0]
        v = _abi.Views()
        _lib.check(self._L.cbx_batch_views(self._h, C.byref(v)))
        self.views = v
        self.tensors: Dict[str, "torch.Tensor"] = {}
        for name, (shape, dt) in _abi.view_specs(v, cfg).items():
            ptr = C.cast(getattr(v, name), C.c_void_p).value
            if not ptr:
                continue
            full = (self.n_envs,) + tuple(shape)
            if int(np.prod(full)) == 0:
                self.tensors[name] = torch.zeros(full, dtype=getattr(torch, dt), device=self.torch_device)
            else:
                self.tensors[name] = torch.as_tensor(_DevArray(ptr, full, dt), device=self.torch_device)
        self.stats_tensor = torch.as_tensor(
            _DevArray(C.cast(v.episode_stats, C.c_void_p).value, (_abi.STAT_COUNT,), "float64"), device=self.torch_device)
        self.att_width = 10 if cfg.mode == _abi.MODE_MARLON else 5
        self._keep = []

    # ------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._torch.cuda.synchronize(self.device)
            self.tensors.clear()
            self._L.cbx_batch_destroy(self._h)
            self._h = None
        for h in getattr(self, "_scns", []):
            self._L.cbx_scenario_destroy(h)
        self._scns, self._scn = [], None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, dtype, width=None):
        """-> contiguous device tensor of `dtype` (accepts numpy / lists / torch tensors)."""
        torch = self._torch
        if a is None:
            return None
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.ascontiguousarray(a))
        a = a.to(device=self.torch_device, dtype=dtype, non_blocking=True).contiguous()
        if width is not None and tuple(a.shape) != (self.n_envs, width):
            raise ValueError(f"expected shape {(self.n_envs, width)}, got {tuple(a.shape)}")
        return a

    # ------------------------------------------------------------------------------------------------
    WHO_ATTACKER, WHO_DEFENDER, WHO_BOTH = 1, 2, 3

    def reset(self, mask=None, who: int = 3):
        """Reset the envs selected by `mask` (all when None) and write their reset observations.  In MARLon mode `who`
        selects AttackerEnvWrapper.reset() (1), DefenderEnvWrapper.reset() (2) or both, attacker first (3)."""
        torch = self._torch
        m = self._dev(mask, torch.uint8) if mask is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_reset_ex(self._h, C.c_void_p(m.data_ptr()) if m is not None else None, int(who), self._stream()))
        self._keep = [m]

    def notify_reset(self, who: int, last_reward: float = 0.0, mask=None):
        """EnvironmentEventSource.notify_reset from outside: raise reset_request on the selected wrappers."""
        torch = self._torch
        m = self._dev(mask, torch.uint8) if mask is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_notify_reset(self._h, C.c_void_p(m.data_ptr()) if m is not None else None, int(who),
                                                      float(last_reward), self._stream()))
        self._keep = [m]

    def step(self, attacker_actions, defender_actions=None, scan_u=None, detect_u=None, who: int = 3):
        """One env-step for every env. Actions: int32 [n,10] (MARLon) or [n,5] (CyberBattleEnv), defender [n,12].
        `who`: both halves of the MARLon pair step (3), only the attacker's (1) or only the defender's (2)."""
        torch = self._torch
        if (who == 3 and scan_u is None and torch.is_tensor(attacker_actions) and attacker_actions.dtype == torch.int16
                and (defender_actions is None or (torch.is_tensor(defender_actions) and defender_actions.dtype == torch.int16))):
            a = self._dev(attacker_actions, torch.int16, self.att_width)  # compact action elements
            d = self._dev(defender_actions, torch.int16, 12) if defender_actions is not None else None
            with torch.cuda.device(self.device):
                _lib.check(self._L.cbx_batch_step_i16(self._h, C.c_void_p(a.data_ptr()),
                                                      C.c_void_p(d.data_ptr()) if d is not None else None, self._stream()))
            self._keep = [a, d]
            return
        a = self._dev(attacker_actions, torch.int32, self.att_width) if attacker_actions is not None else None
        d = self._dev(defender_actions, torch.int32, 12) if defender_actions is not None else None
        tape = None
        su = du = None
        if scan_u is not None:
            su, du = self._dev(scan_u, torch.float64), self._dev(detect_u, torch.float64)
            tape = _abi.Tape(C.cast(C.c_void_p(su.data_ptr()), C.POINTER(C.c_double)),
                             C.cast(C.c_void_p(du.data_ptr()), C.POINTER(C.c_double)))
        with torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_step_ex(self._h, C.c_void_p(a.data_ptr()) if a is not None else None,
                                                 C.c_void_p(d.data_ptr()) if d is not None else None,
                                                 C.byref(tape) if tape is not None else None, int(who), self._stream()))
        self._keep = [a, d, su, du]  # keep inputs alive until the next call (stream-ordered use)

    def host_prepare(self):
        """Allocate everything the host-buffer calls need (the library's staging, the page-locked result ring) now, so that no
        ``step_host`` call pays for it; ``step_host`` calls it on first use.  A timed loop calls this, and warms up, first."""
        if getattr(self, "_host_out", None) is not None:
            return
        n = self.n_envs
        with self._torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_host_prepare(self._h))
        self._host_out = [self._torch.empty(n * 12, dtype=self._torch.uint8, pin_memory=True).numpy() for _ in range(3)]
        self._host_views = [{
            "att_reward": o[: 4 * n].view(np.float32), "def_reward": o[4 * n: 8 * n].view(np.float32),
            "att_terminated": o[8 * n: 9 * n], "att_truncated": o[9 * n: 10 * n],
            "def_terminated": o[10 * n: 11 * n], "def_truncated": o[11 * n: 12 * n]} for o in self._host_out]
        self._host_turn = 0

    def step_host(self, attacker_actions: np.ndarray, defender_actions: Optional[np.ndarray] = None, sync: bool = True):
        """The same step through HOST buffers; synchronous.  Page-locked action arrays (``torch.empty(..., pin_memory=True)``)
        are read by the kernel in place over PCIe; pageable ones go through the library's pinned staging.  The rewards and
        done flags land in one of three page-locked result buffers used in rotation: the returned arrays are views of it and
        stay valid until the third call after this one (copy them to keep them longer).  With both agents stepping and
        page-locked actions the kernel writes the results into that buffer itself (no copy back).  int16 action arrays
        are taken as they are (``cbx_batch_step_host_i16``: half the PCIe bytes); anything else is converted to int32.
        Observations stay in HBM (``self.tensors``); ``fetch_host`` brings selected ones to the host.
        ``sync=False`` (page-locked action arrays only) just enqueues the step on the current stream: the results are valid
        after the caller has synchronised it -- a host loop can overlap this batch's step with another batch's copies."""
        i16 = (getattr(attacker_actions, "dtype", None) == np.int16
               and (defender_actions is None or getattr(defender_actions, "dtype", None) == np.int16))
        dt = np.int16 if i16 else np.int32
        a = np.ascontiguousarray(attacker_actions, dtype=dt)
        d = None if defender_actions is None else np.ascontiguousarray(defender_actions, dtype=dt)
        n = self.n_envs
        # the library reads n * width elements from these pointers: a wrong shape would run past the caller's buffer
        if a.shape != (n, self.att_width):
            raise ValueError(f"attacker actions: expected shape {(n, self.att_width)}, got {a.shape}")
        if d is not None and d.shape != (n, 12):
            raise ValueError(f"defender actions: expected shape {(n, 12)}, got {d.shape}")
        if getattr(self, "_host_out", None) is None:
            self.host_prepare()
        out, views = self._host_out[self._host_turn], self._host_views[self._host_turn]
        self._host_turn = (self._host_turn + 1) % 3
        with self._torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_step_host_ex(self._h, a.ctypes.data, None if d is None else d.ctypes.data, 2 if i16 else 4,
                                                      out.ctypes.data, out.nbytes, 0 if sync else 1, self._stream()))
        if not sync:
            self._keep = [a, d]
        return dict(views)

    def fetch_host(self, fields: int = _abi.F_OBS_FACTORED | _abi.F_RESULTS, out: Optional[np.ndarray] = None, sync: bool = True):
        """Selected observation arrays of the last step in HOST memory with one call (``cbx_batch_fetch_host``): what a
        host-side policy reads back each step.  -> dict name -> numpy view of a page-locked block (`out`, or one of two
        internal blocks used in turn: a view stays valid until the second ``fetch_host`` after it).  ``sync=False`` only
        enqueues the copies on the current stream; synchronise before reading."""
        offs = (C.c_int64 * _abi.F_COUNT)()
        total = int(self._L.cbx_batch_fetch_host_layout(self._h, int(fields), offs))
        if out is None:
            ring = getattr(self, "_fetch_ring", None)
            if ring is None or ring[0].nbytes < total:
                ring = self._fetch_ring = [self._torch.empty(max(total, 256), dtype=self._torch.uint8, pin_memory=True).numpy()
                                           for _ in range(2)]
                self._fetch_turn = 0
            out = ring[self._fetch_turn]
            self._fetch_turn ^= 1
        elif out.nbytes < total:
            raise ValueError(f"out needs {total} bytes")
        with self._torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_fetch_host(self._h, int(fields), out.ctypes.data, out.nbytes, self._stream()))
            if sync:
                self._torch.cuda.current_stream(self.device).synchronize()
        n, res = self.n_envs, {}
        specs = _abi.view_specs(self.views, self.cfg)
        flat = out.reshape(-1).view(np.uint8)
        for k, name in enumerate(_abi.F_NAMES):
            o = int(offs[k])
            if o < 0:
                continue
            if name == "results":
                blk = flat[o:o + 12 * n]
                res.update(att_reward=blk[:4 * n].view(np.float32), def_reward=blk[4 * n:8 * n].view(np.float32),
                           att_terminated=blk[8 * n:9 * n], att_truncated=blk[9 * n:10 * n],
                           def_terminated=blk[10 * n:11 * n], def_truncated=blk[11 * n:12 * n])
                continue
            shape, dt = specs[name]
            nb = int(np.prod((n,) + tuple(shape))) * np.dtype(dt).itemsize
            res[name] = flat[o:o + nb].view(dt).reshape((n,) + tuple(shape))
        return res

    def sample_actions(self, seed: int = 0, attacker_out=None, defender_out=None):
        """Uniformly sampled VALID attacker actions (and uniform defender actions) for the current state, on device."""
        torch = self._torch
        if attacker_out is None:
            attacker_out = torch.empty((self.n_envs, self.att_width), dtype=torch.int32, device=self.torch_device)
        need_def = self.cfg.mode == _abi.MODE_MARLON and self.cfg.def_enabled
        if need_def and defender_out is None:
            defender_out = torch.empty((self.n_envs, 12), dtype=torch.int32, device=self.torch_device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_sample_actions(self._h, C.c_void_p(attacker_out.data_ptr()),
                                                        C.c_void_p(defender_out.data_ptr()) if need_def else None,
                                                        int(seed) & 0xFFFFFFFFFFFFFFFF, self._stream()))
        return attacker_out, (defender_out if need_def else None)

    def export_state(self, begin: int = 0, end: Optional[int] = None) -> np.ndarray:
        end = self.n_envs if end is None else end
        w = self._L.cbx_batch_export_words(self._h)
        out = np.zeros((end - begin, w), dtype=np.int32)
        with self._torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_export_state(self._h, begin, end, out.ctypes.data, self._stream()))
        return out

    def stats(self) -> np.ndarray:
        return self.stats_tensor.cpu().numpy().copy()

    def stats_reset(self):
        with self._torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_stats_reset(self._h, self._stream()))

    def enable_timing(self, on: bool = True):
        _lib.check(self._L.cbx_batch_enable_timing(self._h, int(on)))

    def step_kernel_ms(self):
        ms, n = C.c_double(0), C.c_int64(0)
        _lib.check(self._L.cbx_batch_step_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # slots 0-6: fused kernel (thread 0 of every CTA); slots 8-12: pipelined kernel (lane 0 of every logic / encoder warp)
    PHASES = ["prologue", "state_load", "attacker_logic", "terminal_obs", "defender_logic_desc", "encode", "state_store", "_",
              "pipe_logic_wait_slot", "pipe_logic_load", "pipe_logic_play", "pipe_logic_fields_writeback", "pipe_encoder_wait",
              "pipe_enc_wait_tma_read", "pipe_enc_build_rows", "pipe_enc_issue_bulk"]

    def phase_cycles(self, enable: bool = True):
        """Per-phase SM cycles of the step kernel since the last call (summed over CTAs); switches counting on/off."""
        out = (C.c_uint64 * 16)()
        _lib.check(self._L.cbx_batch_phase_cycles(self._h, int(enable), out))
        return {k: int(out[i]) for i, k in enumerate(self.PHASES)}

    def kernel_info(self) -> dict:
        """Which kernel a step launches and its launch shape (cbx_batch_kernel_info)."""
        out = (C.c_int32 * 8)()
        _lib.check(self._L.cbx_batch_kernel_info(self._h, out))
        keys = ["pipelined", "ctas", "threads", "smem_bytes", "logic_warps", "encoder_warps", "encoder_variant", "tma"]
        d = {k: int(out[i]) for i, k in enumerate(keys)}
        d["tile_order"] = "dynamic" if d["tma"] & 2 else "static"
        d["overlapped_launches"] = bool(d["tma"] & 4)
        # L2 cache policies on the bulk copies: 1 state evict_last, 2 masks / 4 other observations + actions evict_first, 8 tables
        d["l2_policies"] = (d["tma"] >> 4) & 15
        d["tma"] &= 1
        d["name"] = {1: "cbx_pipe_kernel", 2: "cbx_wide_kernel"}.get(d["pipelined"], "cbx_step_kernel")
        return d

    def tile_counter(self):
        """(tickets handed out, CTAs finished) of the dynamic tile order: (0, 0) between launches."""
        out = (C.c_int32 * 2)()
        with self._torch.cuda.device(self.device):
            _lib.check(self._L.cbx_batch_tile_counter(self._h, out))
        return int(out[0]), int(out[1])

    @property
    def launch_count(self) -> int:
        return int(self._L.cbx_batch_launch_count(self._h))

    def numpy(self, name: str) -> np.ndarray:
        return self.tensors[name].cpu().numpy()
