"""A ``marlon_b200.batch.Batch`` look-alike over the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Lets the HOST-side Python of the package (the gym-shaped ``CyberBattleEnv`` view, the MARLon wrappers, the universe and the
SB3 ``VecEnv`` adapter) run where there is no GPU -- the dev container, the only place the reference can be imported -- so
that the reference's own callers can be pointed at it (tests/test_reference_callers.py).  Step semantics come from
``oracle/cbx_oracle.c``; that the CUDA library produces the same arrays is what the ``-m gpu`` parity tests establish.
"""
import numpy as np

from marlon_b200 import _abi
from oracle import OracleBatch


class OracleBackedBatch(OracleBatch):
    WHO_ATTACKER, WHO_DEFENDER, WHO_BOTH = 1, 2, 3

    def __init__(self, compiled, cfg, n_envs, device=0):
        super().__init__(compiled, cfg, int(n_envs))
        import torch

        self.device, self.torch_device = device, torch.device("cpu")
        self.att_width = 10 if cfg.mode == _abi.MODE_MARLON else 5
        self.tensors = {k: torch.from_numpy(v) for k, v in self.arrays.items()}  # share the oracle's memory
        self.stats_tensor = torch.from_numpy(self.stats)
        self.launch_count = 0

    @staticmethod
    def _np(a, dtype=np.int32):
        if a is None:
            return None
        if hasattr(a, "detach"):
            a = a.detach().cpu().numpy()
        return np.ascontiguousarray(a, dtype=dtype)

    def step(self, attacker_actions, defender_actions=None, scan_u=None, detect_u=None, who=3):
        super().step(self._np(attacker_actions), self._np(defender_actions), scan_u, detect_u, who)
        self.launch_count += 1

    def reset(self, mask=None, who=3):
        super().reset(self._np(mask, np.uint8), who)

    def notify_reset(self, who, last_reward=0.0, mask=None):
        super().notify_reset(who, last_reward, self._np(mask, np.uint8))

    def numpy(self, name):
        return self.arrays[name]

    def stats_reset(self):
        super().stats_reset()

    def fetch_host(self, fields=_abi.F_OBS_FACTORED | _abi.F_RESULTS, out=None, sync=True):
        res = {}
        for k, name in enumerate(_abi.F_NAMES):
            if not (fields >> k) & 1:
                continue
            if name == "results":
                for r in ("att_reward", "def_reward", "att_terminated", "att_truncated", "def_terminated", "def_truncated"):
                    if r in self.arrays:
                        res[r] = self.arrays[r].copy()
            elif name in self.arrays:
                res[name] = self.arrays[name].copy()
        return res

    def close(self):
        pass
