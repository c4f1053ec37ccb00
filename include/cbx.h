/*
 * cbx.h -- C ABI of the B200-native batched CyberBattleSim / MARLon environment step.
 *
 * The reference (zsh239040/MARLon, pure Python) has no FFI: its seam is the Python object
 * protocol of CyberBattleEnv / AttackerEnvWrapper / DefenderEnvWrapper.  This header is the
 * boundary a maintainer would bind (ctypes) to replace, per entry point:
 *
 *   cbx_scenario_create   <- model.Environment + Identifiers as built by
 *                            cyberbattle/samples/toyctf/toy_ctf.py:196-200,
 *                            samples/chainpattern/chainpattern.py:198-243,
 *                            simulation/generate_network.py:266-294 (compiled once to tables)
 *   cbx_batch_create      <- CyberBattleEnv.__init__ (_env/cyberbattle_env.py:470-566),
 *                            AttackerEnvWrapper.__init__ (marlon/.../attack_wrapper.py:34-106),
 *                            DefenderEnvWrapper.__init__ (marlon/.../defend_wrapper.py:34-102)
 *   cbx_batch_reset       <- CyberBattleEnv.reset (cyberbattle_env.py:1187-1209),
 *                            AttackerEnvWrapper.reset (attack_wrapper.py:400-468),
 *                            DefenderEnvWrapper.reset (defend_wrapper.py:414-477)
 *   cbx_batch_step        <- CyberBattleEnv.step (cyberbattle_env.py:1145-1185),
 *                            AttackerEnvWrapper.step (attack_wrapper.py:255-398),
 *                            DefenderEnvWrapper.step (defend_wrapper.py:197-327) +
 *                            LearningDefender.executeAction (marlon/defender_agents/defender.py:31-107),
 *                            ScanAndReimageCompromisedMachines.step (_env/defender.py:42-55),
 *                            SB3 DummyVecEnv auto-reset (third party, not vendored)
 *   cbx_batch_views       <- the observation dicts built by
 *                            __observation_reward_from_action_result (cyberbattle_env.py:859-933),
 *                            transform_observation (attack_wrapper.py:474-522),
 *                            DefenderEnvWrapper.observe (defend_wrapper.py:492-534)
 *   cbx_batch_export_state<- nothing (parity instrumentation: the digest of SURVEY.md section C)
 *
 * Conventions: every call returns 0 on success, a negative cbx_status otherwise; the message is
 * available from cbx_last_error() (thread-local).  The library owns all device memory of a batch
 * for the batch's lifetime; the caller owns action / tape buffers and keeps them alive until the
 * stream work that reads them has completed.  Nothing here synchronises the stream unless stated.
 * A batch is not re-entrant; distinct batches are independent.  There is NO CPU path: device < 0
 * or a missing CUDA device is an error.
 */
#ifndef CBX_H_
#define CBX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBX_ABI_VERSION 2

typedef enum {
  CBX_OK = 0,
  CBX_ERR_INVALID = -1,     /* bad argument / malformed scenario blob */
  CBX_ERR_UNSUPPORTED = -2, /* valid in the reference but outside what the tables can express */
  CBX_ERR_CUDA = -3,        /* CUDA runtime error (message carries cudaGetErrorString) */
  CBX_ERR_NODEVICE = -4     /* no CUDA device: there is no CPU fallback */
} cbx_status;

/* ---------------------------------------------------------------------------------------------
 * Scenario wire format ("compiled tables"): a little-endian array of 32-bit words.
 * Produced by marlon_b200/scenario.py; consumed by the CUDA library and by the oracle.
 * ------------------------------------------------------------------------------------------- */
#define CBX_SCN_MAGIC 0x31584243u /* "CBX1" */
#define CBX_SCN_VERSION 3u

enum { /* word indices inside the header */
  CBX_H_MAGIC = 0,
  CBX_H_VERSION,
  CBX_H_TOTAL_WORDS,
  CBX_H_N_NODES,    /* real node count n (<= 255) */
  CBX_H_N_PORTS,    /* len(identifiers.ports)  (<= 32) */
  CBX_H_N_PROPS,    /* len(identifiers.properties) (<= 64) */
  CBX_H_N_LOCAL,    /* len(identifiers.local_vulnerabilities) */
  CBX_H_N_REMOTE,   /* len(identifiers.remote_vulnerabilities) */
  CBX_H_N_SECRETS,  /* distinct CredentialID strings */
  CBX_H_N_TRIPLES,  /* distinct CachedCredential (node, port, credential) triples */
  CBX_H_N_SERVICES, /* total listening services over all nodes (MARLon defender obs) */
  CBX_H_MAX_LEAK,   /* largest LeakedCredentials list of any vulnerability */
  CBX_H_FLAGS,      /* bit0: some vulnerability escalates privileges */
  CBX_H_OFF_NODE,   /* n * CBX_NODE_WORDS */
  CBX_H_OFF_AUTH,   /* n * n_ports * ceil(n_secrets/32): secrets accepted by a RUNNING service on (node, port) */
  CBX_H_OFF_VULN,   /* n * (n_local + n_remote) * CBX_VULN_WORDS, local ids first */
  CBX_H_OFF_PAYLOAD,
  CBX_H_N_PAYLOAD,
  CBX_H_OFF_TRIPLE, /* n_triples * 3 : node index, port index, secret id */
  CBX_H_WORDS = 24
};

/* firewall extension tables (cbx_scenario_set_firewall_tables; `live` defender binding) */
#define CBX_FWX_MAGIC 0x46584243u /* "CBXF" */
enum { CBX_FX_MAGIC = 0,
       CBX_FX_N_NAMES,  /* port names that matter: RDP, SSH, HTTPS, HTTP, su, sudo (defend_wrapper.py:31), then the attacker's ports */
       CBX_FX_N_GROUPS, /* distinct firewall rule LIST objects (several (node, direction) pairs may share one) */
       CBX_FX_RESERVED,
       CBX_FX_WORDS };  /* then: name index of every attacker port [n_ports] | per node: incoming group | outgoing group << 16 [n]
                           | per group: names with a rule, names whose first rule ALLOWs [2 * groups] (the initial per-env state) */

#define CBX_NODE_WORDS 8
enum { /* per-node record */
  CBX_N_VALUE = 0,   /* NodeInfo.value (int32) */
  CBX_N_FLAGS,       /* bit0 reimagable | bit1 agent_installed at reset | bits2-3 privilege at reset | bits8-15 #services */
  CBX_N_PROPS_LO,    /* property bitset over identifiers.properties */
  CBX_N_PROPS_HI,
  CBX_N_FW_OUT,      /* bit p: outgoing rules let identifiers.ports[p] through (first matching rule decides) */
  CBX_N_FW_IN,       /* same for incoming rules */
  CBX_N_LISTEN,      /* bit p: some service is named identifiers.ports[p] */
  CBX_N_DEFOBS       /* bits0-5 incoming rule exists for RDP,SSH,HTTPS,HTTP,su,sudo | bits8-13 same for outgoing |
                        bits16-31 index of this node's first service in the flattened services list */
};

#define CBX_VULN_WORDS 4
enum { /* per (node, vulnerability) record */
  CBX_V_FLAGS = 0,   /* bit0 present | bits1-3 outcome kind | bits4-5 escalation level |
                        bits8-23 precondition truth table indexed by the node's dynamic privilege_N tag bits */
  CBX_V_COST,        /* float32 bits */
  CBX_V_PAYLOAD_OFF, /* word offset into the payload section */
  CBX_V_PAYLOAD_CNT  /* items: triple ids (LeakedCredentials), node indices (LeakedNodesId), 2 words lo/hi (ProbeSucceeded) */
};

enum { /* outcome kinds (model.py:118-195) */
  CBX_OUT_EXPLOIT_FAILED = 0,
  CBX_OUT_LEAKED_CREDENTIALS = 1,
  CBX_OUT_LEAKED_NODES = 2,
  CBX_OUT_LATERAL_MOVE = 3,
  CBX_OUT_CUSTOMER_DATA = 4,
  CBX_OUT_PROBE_SUCCEEDED = 5,
  CBX_OUT_PROBE_FAILED = 6,
  CBX_OUT_ESCALATION = 7
};

/* ---------------------------------------------------------------------------------------------
 * Batch configuration
 * ------------------------------------------------------------------------------------------- */
enum { CBX_MODE_CYBERBATTLE = 0, /* raw CyberBattleEnv.step on gym-style actions */
       CBX_MODE_MARLON = 1 };    /* AttackerEnvWrapper.step then (optionally) DefenderEnvWrapper.step */

enum { CBX_KIND_LOCAL = 0, CBX_KIND_REMOTE = 1, CBX_KIND_CONNECT = 2 };

enum { CBX_MASK_DENSE = 0,     /* int8 [N,L], [N,N,R], [N,N,P,C] written every step (reference layout) */
       CBX_MASK_FACTORED = 1 };/* owned bitset + counts only; dense masks not materialised */

enum { CBX_BUILTIN_NONE = 0, CBX_BUILTIN_SCAN_AND_REIMAGE = 1 };

typedef struct cbx_config {
  int32_t abi_version; /* CBX_ABI_VERSION */
  int32_t mode;        /* CBX_MODE_* */
  /* EnvironmentBounds (cyberbattle_env.py:172-224) */
  int32_t maximum_node_count;
  int32_t maximum_total_credentials;
  int32_t maximum_discoverable_credentials_per_action;
  /* CyberBattleEnv ctor (cyberbattle_env.py:470-485) */
  int32_t throws_on_invalid_actions; /* 1: invalid action -> per-env error code, state untouched but for stepcount */
  int32_t has_attacker_goal;
  int32_t goal_own_atleast;
  double goal_reward;
  double goal_low_availability;
  double goal_own_atleast_percent;
  int32_t defender_goal_eviction;
  int32_t builtin_defender;          /* CBX_BUILTIN_* */
  double maintain_sla;               /* DefenderConstraint.maintain_sla */
  double winning_reward;
  double losing_reward;
  double scan_probability;           /* ScanAndReimageCompromisedMachines(probability, scan_capacity, scan_frequency) */
  int32_t scan_capacity;
  int32_t scan_frequency;
  uint64_t seed;                     /* Philox key for the built-in defender's draws */
  int64_t env_index_base;            /* global index of this batch's env 0: Philox counters use base + local index, so a
                                        run sharded over several GPUs draws exactly what the unsharded run draws */
  /* AttackerEnvWrapper (attack_wrapper.py:34-42); MultiDiscrete layout [3, slice0.., slice1.., slice2..] */
  int32_t kind_of_index[3];          /* CBX_KIND_* selected by action[0] = 0,1,2 (gymnasium 0.29 sorts: connect, local, remote) */
  int32_t att_max_timesteps;
  double att_invalid_action_reward_modifier;
  /* DefenderEnvWrapper (defend_wrapper.py:34-45); reference_stale binding */
  int32_t def_enabled;
  int32_t def_max_timesteps;
  int32_t def_reset_on_constraint_broken;
  int32_t auto_reset;                /* 1: SB3 DummyVecEnv semantics (reset in the same step, terminal obs kept aside) */
  double def_invalid_action_reward;
  double def_loss_reward;
  double def_sla_worsening_penalty_scale;
  int32_t mask_mode;                 /* CBX_MASK_* */
  int32_t emit_terminal_obs;         /* 1: keep a second observation buffer holding the pre-reset observation of done envs */
  int32_t def_binding;               /* CBX_DEF_BINDING_*: which environment the LearningDefender acts on */
  int32_t reserved0;
} cbx_config;

enum { CBX_DEF_BINDING_STALE = 0, /* the reference AS EXECUTED: DefenderEnvWrapper / LearningDefender keep the actuator and
                                     environment they saw at construction (defend_wrapper.py:51, defender.py:29-30), which every
                                     CyberBattleEnv.reset() replaces: re-imaging, blocking and allowing hit a dead copy */
       CBX_DEF_BINDING_LIVE = 1 };/* the binding refreshed at every CyberBattleEnv.reset(): the defender re-images nodes of, and
                                     edits the firewall rule lists of, the environment the attacker plays in (rule lists shared
                                     between nodes stay shared; allow always appends to INCOMING, B.3; stop / start
                                     service stay no-ops, B.4).  Needs cbx_scenario_set_firewall_tables. */

/* ---------------------------------------------------------------------------------------------
 * Views: device pointers of every per-env output array (row-major, batch dimension first).
 * They stay valid for the lifetime of the batch; contents are defined after the stream work of
 * the last reset/step has completed.
 * ------------------------------------------------------------------------------------------- */
typedef struct cbx_views {
  int64_t n_envs;
  int32_t N, L, R, P, C, LEAK, n_props, n_nodes, n_services, owned_words;
  /* attacker observation, AttackerEnvWrapper normalised form */
  int32_t* scalars;          /* [n,8] newly_discovered_nodes_count, lateral_move, customer_data_found, probe_result,
                                      escalation, credential_cache_length, discovered_node_count, (blank-observation flag) */
  int32_t* leaked_credentials;        /* [n, 4*LEAK] */
  int32_t* credential_cache_matrix;   /* [n, 2*C] */
  int32_t* discovered_nodes_properties; /* [n, N*n_props] */
  int32_t* nodes_privilegelevel;      /* [n, N] */
  int8_t* local_vulnerability;        /* [n, N, L]        (NULL in factored mode) */
  int8_t* remote_vulnerability;       /* [n, N, N, R]     (NULL in factored mode) */
  int8_t* connect;                    /* [n, N, N, P, C]  (NULL in factored mode) */
  uint32_t* owned_bits;               /* [n, owned_words] bit s: discovery index s holds the agent (both modes) */
  /* defender observation (defend_wrapper.py:162-172,492-534); NULL unless def_enabled */
  int8_t* def_infected_nodes;         /* [n, n_nodes] */
  int8_t* def_incoming_firewall;      /* [n, 6*n_nodes] */
  int8_t* def_outgoing_firewall;      /* [n, 6*n_nodes] */
  int8_t* def_services_status;        /* [n, n_services] */
  /* step results */
  float* att_reward;                  /* [n] */
  float* def_reward;                  /* [n] */
  uint8_t* att_terminated;            /* [n] */
  uint8_t* att_truncated;
  uint8_t* def_terminated;
  uint8_t* def_truncated;
  int32_t* att_info;                  /* [n,8] cyber reward bits(f32), raw reward bits (f32, pre-clip), outcome code, error code,
                                               stepcount, intercepted flag, episode length at done, reserved */
  double* network_availability;       /* [n] info["network_availability"] of the live env */
  double* episode_stats;              /* [CBX_STAT_COUNT] per-GPU partial sums, accumulated since cbx_batch_stats_reset */
  /* terminal observations (emit_terminal_obs): same shapes as the main ones, written only for envs done this step */
  int32_t* term_scalars;
  int32_t* term_leaked_credentials;
  int32_t* term_credential_cache_matrix;
  int32_t* term_discovered_nodes_properties;
  int32_t* term_nodes_privilegelevel;
  int8_t* term_local_vulnerability;
  int8_t* term_remote_vulnerability;
  int8_t* term_connect;
  int8_t* term_def_infected_nodes;
} cbx_views;

enum { /* episode_stats slots: sums over finished episodes on this GPU */
  CBX_STAT_EPISODES = 0,
  CBX_STAT_ATT_RETURN,
  CBX_STAT_ATT_RETURN_SQ,
  CBX_STAT_EP_LEN,
  CBX_STAT_EP_LEN_SQ,
  CBX_STAT_DEF_RETURN,
  CBX_STAT_DEF_RETURN_SQ,
  CBX_STAT_ATT_VALID,
  CBX_STAT_ATT_INVALID,
  CBX_STAT_DEF_VALID,
  CBX_STAT_DEF_INVALID,
  CBX_STAT_ATT_WINS,
  CBX_STAT_SLA_BREACHES,
  CBX_STAT_TIMEOUTS,
  CBX_STAT_ENV_STEPS,
  CBX_STAT_COUNT = 16
};

enum { /* att_info[.,2] outcome code: what ActionResult.outcome was (None = 0) */
  CBX_RES_NONE = 0,
  CBX_RES_EXPLOIT_FAILED = 1, /* model.ExploitFailed (precondition false, or a trap vulnerability) */
  CBX_RES_LEAKED_CREDENTIALS = 2,
  CBX_RES_LEAKED_NODES = 3,
  CBX_RES_LATERAL_MOVE = 4,
  CBX_RES_CUSTOMER_DATA = 5,
  CBX_RES_PROBE_SUCCEEDED = 6,
  CBX_RES_PROBE_FAILED = 7,
  CBX_RES_ESCALATION = 8,
  CBX_RES_OUT_OF_BOUND = 9    /* OutOfBoundIndexError swallowed by CyberBattleEnv.step: blank observation */
};

enum { /* att_info[.,3] error code when throws_on_invalid_actions (reference raises ValueError) */
  CBX_E_NONE = 0,
  CBX_E_SOURCE_NOT_OWNED = 1,
  CBX_E_TARGET_NOT_DISCOVERED = 2,
  CBX_E_CREDENTIAL_NOT_GATHERED = 3,
  CBX_E_STEP_AFTER_DONE = 4   /* RuntimeError("new episode must be started with env.reset()") */
};

/* RNG tape for the built-in defender (parity runs): draws the reference consumed, SURVEY.md A.5.
 * scan_u[i*cap+k]   = k-th random.random() of random.choices(...) for env i this step (NaN: no scan this step)
 * detect_u[i*cap+k] = numpy.random.random() drawn for slot k (NaN: not drawn)                                 */
typedef struct cbx_tape {
  const double* scan_u;   /* device pointer [n, scan_capacity] */
  const double* detect_u; /* device pointer [n, scan_capacity] */
} cbx_tape;

/* Canonical, layout-independent state dump used by the parity tests. One record per env:
 * int32 words, see CBX_X_* ; then per node records; then lists.  cbx_export_words(scenario, cfg) gives the size. */
enum {
  CBX_X_STEPCOUNT = 0,
  CBX_X_DONE,
  CBX_X_N_DISCOVERED,
  CBX_X_N_CACHED,
  CBX_X_ATT_TIMESTEPS,
  CBX_X_DEF_TIMESTEPS,
  CBX_X_ATT_RESET_REQUEST,
  CBX_X_DEF_RESET_REQUEST,
  CBX_X_HAS_BREACHED_SLA,
  CBX_X_ATT_VALID,
  CBX_X_ATT_INVALID,
  CBX_X_DEF_VALID,
  CBX_X_DEF_INVALID,
  CBX_X_LIVE_IMAGING_COUNT,   /* nodes not Running in the live env */
  CBX_X_SHADOW_IMAGING_COUNT, /* nodes not Running in the defender's stale copy (availability = (n-k)/n) */
  CBX_X_PREV_SHADOW_IMAGING_COUNT,
  CBX_X_HEADER_WORDS = 16
  /* followed by, with n = n_nodes:
   *   discovered order   [n]   node index at discovery position, -1 padded
   *   agent_installed    [n]
   *   privilege_level    [n]
   *   live countdown     [n]   0 = Running, k>0 = Imaging with k ticks left (16 right after reimage_node)
   *   shadow countdown   [n]
   *   ever_owned         [n]
   *   discovered props lo[n], hi[n]
   *   attacked bits      [n]   bit (2v) ever attacked, bit (2v+1) attacked since last reimage; v = local ids then remote ids
   *   tags               [n]   dynamic privilege_N tag bits
   *   credential cache   [C]   triple ids, -1 padded
   *   gathered secrets   [ceil(n_secrets/32)] */
};

typedef struct cbx_scenario cbx_scenario;
typedef struct cbx_batch cbx_batch;

const char* cbx_last_error(void);
int cbx_abi_version(void);

int cbx_scenario_create(const void* tables, size_t nbytes, cbx_scenario** out);
int cbx_scenario_destroy(cbx_scenario* s);
/* Firewall extension tables of a scenario (marlon_b200/scenario.py FWX_* layout: port-name indices, the rule-list alias group
 * of every (node, direction), two bits per (group, name)): required by CBX_DEF_BINDING_LIVE, where firewall rule lists
 * (model.FirewallConfiguration, model.py:240-262) become per-env state. */
int cbx_scenario_set_firewall_tables(cbx_scenario* s, const void* words, size_t nbytes);

int cbx_config_default(cbx_config* cfg);
int cbx_batch_create(const cbx_scenario* s, int64_t n_envs, const cbx_config* cfg, int device, cbx_batch** out);

/* One batch over SEVERAL scenarios of one Identifiers family (configs[4]: CyberBattleRandom networks generated by
 * simulation/generate_network.py:266-294, a different network per group of envs).  Envs are grouped by scenario:
 * envs_per_scenario[k] consecutive envs play scenarios[k]; every group but the last must be a multiple of 32 envs.
 * The ports / vulnerability ids / properties must agree; node, credential and service counts may differ: the per-env
 * state and the n-sized observation arrays (defender observation, cbx_batch_export_state) are laid out for the largest
 * scenario and zero-padded, the game's arithmetic (availability, scan targets, ownership ratios) uses each scenario's
 * own counts.  Bounds (maximum_node_count, ...) come from cfg and must hold the largest scenario. */
int cbx_batch_create_multi(const cbx_scenario* const* scenarios, int n_scenarios, const int64_t* envs_per_scenario,
                           const cbx_config* cfg, int device, cbx_batch** out);
int cbx_batch_destroy(cbx_batch* b);

/* Reset envs whose mask byte is non-zero (all when mask == NULL; device pointer [n]) and write their
 * reset observation.  In MARLon mode this is attacker.reset() followed by defender.reset(). */
int cbx_batch_reset(cbx_batch* b, const uint8_t* mask_or_null, void* cuda_stream);

/* One env-step for every env.  attacker_actions: device int32.
 *   CBX_MODE_CYBERBATTLE: [n,5]  = kind (CBX_KIND_*), then up to 4 coordinates as CyberBattleEnv.step takes them
 *   CBX_MODE_MARLON:      [n,10] = AttackerEnvWrapper MultiDiscrete action
 * defender_actions: device int32 [n,12] DefenderEnvWrapper MultiDiscrete action, column 0 < 0 = empty action
 *   (NULL when def_enabled == 0).  tape: NULL -> Philox draws keyed (seed, env, stepcount, slot). */
int cbx_batch_step(cbx_batch* b, const int32_t* attacker_actions, const int32_t* defender_actions,
                   const cbx_tape* tape_or_null, void* cuda_stream);

/* The two halves of the MARLon pair step as separate calls, for callers that drive the two wrappers one after the
 * other like the reference does (marl_algorithm.py:43-49): who = CBX_WHO_ATTACKER, CBX_WHO_DEFENDER or both.
 * cbx_batch_reset_ex(.., CBX_WHO_ATTACKER) is AttackerEnvWrapper.reset(), (.., CBX_WHO_DEFENDER) DefenderEnvWrapper.reset(). */
enum { CBX_WHO_ATTACKER = 1, CBX_WHO_DEFENDER = 2 };
int cbx_batch_step_ex(cbx_batch* b, const int32_t* attacker_actions, const int32_t* defender_actions,
                      const cbx_tape* tape_or_null, int who, void* cuda_stream);
int cbx_batch_reset_ex(cbx_batch* b, const uint8_t* mask_or_null, int who, void* cuda_stream);

/* EnvironmentEventSource.notify_reset(last_reward) delivered from outside (environment_event_source.py:30-38; used by
 * marl_algorithm.run_episode, marl_algorithm.py:176-178): sets reset_request on the selected wrappers of the masked envs and,
 * for the defender, records last_reward (defend_wrapper.py:479-482). */
int cbx_batch_notify_reset(cbx_batch* b, const uint8_t* mask_or_null, int who, double last_reward, void* cuda_stream);

/* Same step on HOST buffers: pinned staging + H2D of the actions, the step, D2H of rewards and done flags
 * (att_reward, def_reward, 4 flag arrays -> host_out, layout: float[n], float[n], uint8[4][n]); synchronises. */
int cbx_batch_step_host(cbx_batch* b, const int32_t* host_attacker_actions, const int32_t* host_defender_actions,
                        void* host_out, size_t host_out_bytes, void* cuda_stream);

/* The general form: elem_bytes = 4 (int32 action elements) or 2 (int16); flags = CBX_HOST_NOSYNC: only enqueue -- the call
 * returns before the buffers are used, all three must be page-locked and stay untouched until the caller has synchronised
 * the stream (lets a host loop overlap this batch's step with another batch's copies). */
enum { CBX_HOST_NOSYNC = 1 };
int cbx_batch_step_host_ex(cbx_batch* b, const void* host_attacker_actions, const void* host_defender_actions, int elem_bytes,
                           void* host_out, size_t host_out_bytes, int flags, void* cuda_stream);

/* Allocate what the host-buffer calls need -- page-locked staging for pageable caller buffers, device action buffers for
 * the explicit-copy mode -- NOW instead of inside the first cbx_batch_step_host call (which otherwise does it, all or
 * nothing).  A loop that is timed calls this (and a few untimed steps) first.  Idempotent. */
int cbx_batch_host_prepare(cbx_batch* b);

/* Observation arrays of the last step to HOST memory in one call: what a host-side policy (the reference's
 * obs_as_tensor(self._last_obs) consumer, baseline_marlon_agent.py:113-116) reads back each step.  `fields` selects arrays
 * (CBX_F_*); they are packed one after the other, each [n_envs, per-env size] row-major in its reference dtype, in the
 * order of the CBX_F_* bits, every array starting on a 256-byte boundary.  cbx_batch_fetch_host_layout fills
 * offsets[k] (byte offset of the array of bit k, -1 if not selected or not materialised) and returns the total size.
 * The copies are enqueued on the stream (page-locked host_out: asynchronous); nothing is synchronised. */
enum {
  CBX_F_SCALARS = 1 << 0,      /* int32 [n,8] */
  CBX_F_LEAKED = 1 << 1,       /* int32 [n,4*LEAK] */
  CBX_F_CACHEM = 1 << 2,       /* int32 [n,2*C] */
  CBX_F_PROPS = 1 << 3,        /* int32 [n,N*n_props] */
  CBX_F_PRIV = 1 << 4,         /* int32 [n,N] */
  CBX_F_OWNED_BITS = 1 << 5,   /* uint32 [n,owned_words]: with scalars[5], scalars[6] the factored action masks */
  CBX_F_LOCAL = 1 << 6,        /* int8 dense masks (dense mode only) */
  CBX_F_REMOTE = 1 << 7,
  CBX_F_CONNECT = 1 << 8,
  CBX_F_DEF_INFECTED = 1 << 9, /* int8 defender observation (def_enabled only) */
  CBX_F_DEF_FW_IN = 1 << 10,
  CBX_F_DEF_FW_OUT = 1 << 11,
  CBX_F_DEF_SERVICES = 1 << 12,
  CBX_F_RESULTS = 1 << 13,     /* float att_reward[n], float def_reward[n], uint8 att_terminated[n], att_truncated[n],
                                  def_terminated[n], def_truncated[n] (cbx_batch_step_host's layout) */
  CBX_F_COUNT = 14,
  CBX_F_OBS_FACTORED = 0x3F | (0xF << 9)  /* every small field + factored masks + defender observation */
};
int64_t cbx_batch_fetch_host_layout(const cbx_batch* b, uint32_t fields, int64_t* offsets /* [CBX_F_COUNT] */);
int cbx_batch_fetch_host(cbx_batch* b, uint32_t fields, void* host_out, size_t host_out_bytes, void* cuda_stream);

/* The same two calls with int16 action elements (every component of both MultiDiscrete spaces is far below 32768):
 * half the bytes over PCIe, which is what bounds cbx_batch_step_host.  The reference hands int64 numpy arrays from the
 * policy to env.step (baseline_marlon_agent.py:118-131), so an adapter narrows them either way. */
int cbx_batch_step_i16(cbx_batch* b, const int16_t* attacker_actions, const int16_t* defender_actions, void* cuda_stream);
int cbx_batch_step_host_i16(cbx_batch* b, const int16_t* host_attacker_actions, const int16_t* host_defender_actions,
                            void* host_out, size_t host_out_bytes, void* cuda_stream);

/* Fill device action buffers with VALID attacker actions drawn the way CyberBattleEnv.sample_valid_action draws them
 * (cyberbattle_env.py:959-1047: whole proposals -- kind, then coordinates uniform over their ranges, kind 1 = local and kind 0 =
 * remote as in the reference -- are redrawn until the action mask admits one) and uniform defender actions; Philox4x32-10 keyed
 * by (seed; env, call number).  The oracle's orc_sample_actions produces the same actions from the same state. */
int cbx_batch_sample_actions(cbx_batch* b, int32_t* attacker_actions, int32_t* defender_actions, uint64_t seed,
                             void* cuda_stream);

int cbx_batch_views(cbx_batch* b, cbx_views* out);
int cbx_batch_stats_reset(cbx_batch* b, void* cuda_stream);

int64_t cbx_export_words(const cbx_scenario* s, const cbx_config* cfg);
/* Canonical state of envs [env_begin, env_end) into a HOST buffer; synchronises the stream. */
/* words per env written by cbx_batch_export_state for this batch (multi-scenario batches: the largest scenario's) */
int64_t cbx_batch_export_words(const cbx_batch* b);
int cbx_batch_export_state(cbx_batch* b, int64_t env_begin, int64_t env_end, int32_t* host_out, void* cuda_stream);

/* Number of kernels this library has launched on behalf of the batch so far (bench.py's gpu_launches). */
int64_t cbx_batch_launch_count(const cbx_batch* b);
/* Average duration (ms) of the step kernel over the launches recorded since the last call, measured with CUDA events
 * on the launch stream when timing is enabled.  When consecutive launches overlap (cbx_batch_kernel_info bit 2) an event
 * between two launches would serialise them: the events then bracket the whole run of step launches since timing was
 * enabled (or since the last call) and the mean is that span divided by the launches in it; the bracket closes when timing
 * is switched off (an event on the launch stream, nothing synchronised) or, failing that, at cbx_batch_step_kernel_ms. */
int cbx_batch_enable_timing(cbx_batch* b, int enabled);
int cbx_batch_step_kernel_ms(cbx_batch* b, double* mean_ms, int64_t* launches);

/* Generalised advantage estimation over a device-resident rollout: what stable-baselines3's
 * RolloutBuffer.compute_returns_and_advantage computes when MARLon's on_rollout_end calls it
 * (marlon/baseline_models/multiagent/baseline_marlon_agent.py:276-284).  All pointers are DEVICE pointers; rewards / values /
 * episode_starts / advantages / returns are [n_steps, n_envs] row-major, last_values / last_dones are [n_envs].
 * advantages[t] = delta_t + gamma * lambda * (1 - start_{t+1}) * advantages[t+1],
 * delta_t = r_t + gamma * V_{t+1} * (1 - start_{t+1}) - V_t, with V_T = last_values and start_T = last_dones. */
int cbx_gae(const float* rewards, const float* values, const uint8_t* episode_starts, const float* last_values,
            const uint8_t* last_dones, double gamma, double gae_lambda, int n_steps, int64_t n_envs, float* advantages,
            float* returns, void* cuda_stream);

/* Which kernel one step launches (reported by bench.py next to the roofline): out8[0] = 1 pipelined kernel (cbx_pipe_kernel:
 * logic warps ahead of TMA-storing encoder warps) / 2 warp-per-tile kernel for large state (cbx_wide_kernel) / 0 fused kernel
 * (cbx_step_kernel); [1] CTAs; [2] threads per CTA;
 * [3] dynamic shared memory bytes per CTA; [4] logic warps; [5] encoder warps; [6] encoder variant (0 generic, 1 warp per
 * env, 2 ToyCtf(12,10) static, 3 Chain-10(12,12) static); [7] bit 0 TMA staging enabled, bit 1 dynamic tile order (tiles after a warp's first are drawn
 * from a global ticket counter instead of a fixed stride), bit 2 consecutive launches overlap (programmatic dependent launch, accesses
 * ordered tile by tile through completion counters in HBM), bits 4-7 the L2 cache policies on the pipelined kernel's bulk
 * copies (1 state tiles evict_last, 2 dense masks evict_first, 4 other observations and the actions evict_first, 8 scenario
 * tables evict_last; chosen from the batch size, CBX_L2_HINTS overrides). */
int cbx_batch_kernel_info(const cbx_batch* b, int32_t* out8);

/* Parity instrumentation of the dynamic tile order; both words must read 0 between launches.  Warp-per-tile kernel: the
 * global counters {tickets handed out, CTAs finished} (the last CTA of a launch resets them).  Pipelined kernel: {tickets
 * the last launch drew from its own counter - the number of tiles, 0}: every logic warp draws until one ticket is past the
 * end, which makes exactly n_tiles draws per launch.  Synchronises the device. */
int cbx_batch_tile_counter(cbx_batch* b, int32_t* out2);

/* Instrumentation: per-phase SM cycle counters of the step kernel, summed over CTAs (thread 0 of each CTA):
 * [0] prologue [1] state-tile load [2] attacker logic [3] terminal observations [4] reset/defender logic + descriptors
 * [5] observation/mask encode [6] state write-back.  enable!=0 switches counting on; out16 (may be NULL) receives and clears. */
int cbx_batch_phase_cycles(cbx_batch* b, int enable, uint64_t* out16);

/* sizeof(cbx_config) (0), sizeof(cbx_views) (1), sizeof(cbx_tape) (2): lets a binding check its struct mirrors. */
size_t cbx_abi_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* CBX_H_ */
