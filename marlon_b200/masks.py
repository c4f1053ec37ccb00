"""Dense action masks from their factored form.

``CyberBattleEnv.__update_action_mask`` (cyberbattle_env.py:643-677) fills three dense int8 arrays, but every entry is a
product of a few per-env quantities:

    local  [s, v]        = owned[s] and the vulnerability v exists on the node at discovery index s   (:654-660)
    remote [s, t, r]     = owned[s] and t < n_discovered                                              (:668)
    connect[s, t, p, c]  = owned[s] and t < n_discovered and c < n_cached_credentials                 (:672-677)

With ``mask_mode="factored"`` (required where the dense connect mask would be megabytes per env, e.g. Chain-100) the batch
emits only ``owned_bits`` (bit s = owned[s]), ``discovered_node_count`` and ``credential_cache_length``; this module
materialises the dense arrays on demand, for a few envs at a time (legacy single-env views, tests).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import scenario as _scn


def local_presence_table(compiled) -> np.ndarray:
    """[n_nodes, L] uint8: local vulnerability v is defined on node i (library or the node's own dict, ENV:660)."""
    blob = compiled.blob
    n, L, R = int(blob[_scn.H_N_NODES]), int(blob[_scn.H_N_LOCAL]), int(blob[_scn.H_N_REMOTE])
    off = int(blob[_scn.H_OFF_VULN])
    tab = blob[off: off + n * (L + R) * _scn.VULN_WORDS].reshape(n, L + R, _scn.VULN_WORDS)
    return (tab[:, :L, 0] & 1).astype(np.uint8)


def dense_masks_from_factored(owned_bits: np.ndarray, n_discovered: np.ndarray, n_cached: np.ndarray, discovery_order: np.ndarray,
                              present_local: np.ndarray, N: int, R: int, P: int, C: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (local [n,N,L], remote [n,N,N,R], connect [n,N,N,P,C]) int8.

    owned_bits: uint32 [n, ceil(N/32)]; n_discovered / n_cached: int [n]; discovery_order: int [n, >=n_discovered] node index
    per discovery index (-1 beyond), e.g. the first section of ``Batch.export_state``; present_local: ``local_presence_table``."""
    owned_bits = np.asarray(owned_bits, dtype=np.uint32)
    n = owned_bits.shape[0]
    s = np.arange(N)
    owned = ((owned_bits[:, s // 32] >> (s % 32).astype(np.uint32)) & 1).astype(bool)  # [n, N]
    nd = np.asarray(n_discovered).reshape(n, 1)
    nc = np.asarray(n_cached).reshape(n, 1)
    L = present_local.shape[1]
    order = np.full((n, N), -1, dtype=np.int64)
    k = min(N, discovery_order.shape[1])
    order[:, :k] = np.asarray(discovery_order)[:, :k]
    present = np.where((order >= 0)[:, :, None], present_local[np.clip(order, 0, present_local.shape[0] - 1)], 0)  # [n, N, L]
    local = (owned[:, :, None] & present.astype(bool)).astype(np.int8)
    t_ok = s[None, :] < nd  # [n, N]
    remote = np.broadcast_to((owned[:, :, None] & t_ok[:, None, :])[:, :, :, None], (n, N, N, R)).astype(np.int8)
    c_ok = np.arange(C)[None, :] < nc  # [n, C]
    connect = np.broadcast_to((owned[:, :, None] & t_ok[:, None, :])[:, :, :, None, None] & c_ok[:, None, None, None, :],
                              (n, N, N, P, C)).astype(np.int8)
    return local, remote, connect
