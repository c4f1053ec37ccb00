"""ctypes mirror of include/cbx.h (structs and enums). Keep in lock-step with the header;
``tests/test_abi.py`` checks sizes/offsets against the compiled library (``cbx_abi_sizeof``)."""
import ctypes as C

ABI_VERSION = 2

MODE_CYBERBATTLE, MODE_MARLON = 0, 1
KIND_LOCAL, KIND_REMOTE, KIND_CONNECT = 0, 1, 2
KIND_NAMES = {KIND_LOCAL: "local_vulnerability", KIND_REMOTE: "remote_vulnerability", KIND_CONNECT: "connect"}
KIND_WIDTH = {KIND_LOCAL: 2, KIND_REMOTE: 3, KIND_CONNECT: 4}
MASK_DENSE, MASK_FACTORED = 0, 1
BUILTIN_NONE, BUILTIN_SCAN_AND_REIMAGE = 0, 1
DEF_BINDING_STALE, DEF_BINDING_LIVE = 0, 1

(STAT_EPISODES, STAT_ATT_RETURN, STAT_ATT_RETURN_SQ, STAT_EP_LEN, STAT_EP_LEN_SQ, STAT_DEF_RETURN,
 STAT_DEF_RETURN_SQ, STAT_ATT_VALID, STAT_ATT_INVALID, STAT_DEF_VALID, STAT_DEF_INVALID, STAT_ATT_WINS,
 STAT_SLA_BREACHES, STAT_TIMEOUTS, STAT_ENV_STEPS) = range(15)
STAT_COUNT = 16
STAT_NAMES = ["episodes", "att_return", "att_return_sq", "ep_len", "ep_len_sq", "def_return", "def_return_sq",
              "att_valid", "att_invalid", "def_valid", "def_invalid", "att_wins", "sla_breaches", "timeouts",
              "env_steps", "reserved"]

(RES_NONE, RES_EXPLOIT_FAILED, RES_LEAKED_CREDENTIALS, RES_LEAKED_NODES, RES_LATERAL_MOVE, RES_CUSTOMER_DATA,
 RES_PROBE_SUCCEEDED, RES_PROBE_FAILED, RES_ESCALATION, RES_OUT_OF_BOUND) = range(10)
E_NONE, E_SOURCE_NOT_OWNED, E_TARGET_NOT_DISCOVERED, E_CREDENTIAL_NOT_GATHERED, E_STEP_AFTER_DONE = range(5)

X_HEADER_WORDS = 16
X_NAMES = ["stepcount", "done", "n_discovered", "n_cached", "att_timesteps", "def_timesteps", "att_reset_request",
           "def_reset_request", "has_breached_sla", "att_valid", "att_invalid", "def_valid", "def_invalid",
           "live_imaging_count", "shadow_imaging_count", "prev_shadow_imaging_count"]


# cbx_batch_fetch_host field bits (CBX_F_*), in packing order, with the view each one copies
F_NAMES = ["scalars", "leaked_credentials", "credential_cache_matrix", "discovered_nodes_properties", "nodes_privilegelevel",
           "owned_bits", "local_vulnerability", "remote_vulnerability", "connect", "def_infected_nodes", "def_incoming_firewall",
           "def_outgoing_firewall", "def_services_status", "results"]
F_COUNT = 14
F_RESULTS = 1 << 13
F_OBS_FACTORED = 0x3F | (0xF << 9)
F_OBS_DENSE = F_OBS_FACTORED | (0x7 << 6)


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("mode", C.c_int32),
        ("maximum_node_count", C.c_int32),
        ("maximum_total_credentials", C.c_int32),
        ("maximum_discoverable_credentials_per_action", C.c_int32),
        ("throws_on_invalid_actions", C.c_int32),
        ("has_attacker_goal", C.c_int32),
        ("goal_own_atleast", C.c_int32),
        ("goal_reward", C.c_double),
        ("goal_low_availability", C.c_double),
        ("goal_own_atleast_percent", C.c_double),
        ("defender_goal_eviction", C.c_int32),
        ("builtin_defender", C.c_int32),
        ("maintain_sla", C.c_double),
        ("winning_reward", C.c_double),
        ("losing_reward", C.c_double),
        ("scan_probability", C.c_double),
        ("scan_capacity", C.c_int32),
        ("scan_frequency", C.c_int32),
        ("seed", C.c_uint64),
        ("env_index_base", C.c_int64),
        ("kind_of_index", C.c_int32 * 3),
        ("att_max_timesteps", C.c_int32),
        ("att_invalid_action_reward_modifier", C.c_double),
        ("def_enabled", C.c_int32),
        ("def_max_timesteps", C.c_int32),
        ("def_reset_on_constraint_broken", C.c_int32),
        ("auto_reset", C.c_int32),
        ("def_invalid_action_reward", C.c_double),
        ("def_loss_reward", C.c_double),
        ("def_sla_worsening_penalty_scale", C.c_double),
        ("mask_mode", C.c_int32),
        ("emit_terminal_obs", C.c_int32),
        ("def_binding", C.c_int32),
        ("reserved0", C.c_int32),
    ]


_P32 = C.POINTER(C.c_int32)
_P8 = C.POINTER(C.c_int8)
_PU8 = C.POINTER(C.c_uint8)


class Views(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64),
        ("N", C.c_int32), ("L", C.c_int32), ("R", C.c_int32), ("P", C.c_int32), ("C", C.c_int32),
        ("LEAK", C.c_int32), ("n_props", C.c_int32), ("n_nodes", C.c_int32), ("n_services", C.c_int32),
        ("owned_words", C.c_int32),
        ("scalars", _P32),
        ("leaked_credentials", _P32),
        ("credential_cache_matrix", _P32),
        ("discovered_nodes_properties", _P32),
        ("nodes_privilegelevel", _P32),
        ("local_vulnerability", _P8),
        ("remote_vulnerability", _P8),
        ("connect", _P8),
        ("owned_bits", C.POINTER(C.c_uint32)),
        ("def_infected_nodes", _P8),
        ("def_incoming_firewall", _P8),
        ("def_outgoing_firewall", _P8),
        ("def_services_status", _P8),
        ("att_reward", C.POINTER(C.c_float)),
        ("def_reward", C.POINTER(C.c_float)),
        ("att_terminated", _PU8),
        ("att_truncated", _PU8),
        ("def_terminated", _PU8),
        ("def_truncated", _PU8),
        ("att_info", _P32),
        ("network_availability", C.POINTER(C.c_double)),
        ("episode_stats", C.POINTER(C.c_double)),
        ("term_scalars", _P32),
        ("term_leaked_credentials", _P32),
        ("term_credential_cache_matrix", _P32),
        ("term_discovered_nodes_properties", _P32),
        ("term_nodes_privilegelevel", _P32),
        ("term_local_vulnerability", _P8),
        ("term_remote_vulnerability", _P8),
        ("term_connect", _P8),
        ("term_def_infected_nodes", _P8),
    ]


class Tape(C.Structure):
    _fields_ = [("scan_u", C.POINTER(C.c_double)), ("detect_u", C.POINTER(C.c_double))]


def view_specs(v: Views, cfg: Config):
    """name -> (field, per-env shape, numpy dtype string) for every array of a Views struct."""
    N, L, R, P, Cc, LEAK = v.N, v.L, v.R, v.P, v.C, v.LEAK
    specs = {
        "scalars": ((8,), "int32"),
        "leaked_credentials": ((4 * LEAK,), "int32"),
        "credential_cache_matrix": ((2 * Cc,), "int32"),
        "discovered_nodes_properties": ((N * v.n_props,), "int32"),
        "nodes_privilegelevel": ((N,), "int32"),
        "local_vulnerability": ((N, L), "int8"),
        "remote_vulnerability": ((N, N, R), "int8"),
        "connect": ((N, N, P, Cc), "int8"),
        "owned_bits": ((v.owned_words,), "uint32"),
        "def_infected_nodes": ((v.n_nodes,), "int8"),
        "def_incoming_firewall": ((6 * v.n_nodes,), "int8"),
        "def_outgoing_firewall": ((6 * v.n_nodes,), "int8"),
        "def_services_status": ((v.n_services,), "int8"),
        "att_reward": ((), "float32"),
        "def_reward": ((), "float32"),
        "att_terminated": ((), "uint8"),
        "att_truncated": ((), "uint8"),
        "def_terminated": ((), "uint8"),
        "def_truncated": ((), "uint8"),
        "att_info": ((8,), "int32"),
        "network_availability": ((), "float64"),
    }
    for k in ["scalars", "leaked_credentials", "credential_cache_matrix", "discovered_nodes_properties",
              "nodes_privilegelevel", "local_vulnerability", "remote_vulnerability", "connect", "def_infected_nodes"]:
        specs["term_" + k] = specs[k]
    return specs
