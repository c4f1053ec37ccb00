/*
 * cbx_oracle.c -- CPU restatement of the reference's environment step.  TEST INFRASTRUCTURE.
 *
 * This file is the parity oracle: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product (marlon_b200/) never does; it has no
 * CPU path.  The oracle deliberately keeps the reference's *object shape* (per-node records, explicit
 * ordered lists, linear searches, a logical clock where the reference compares datetime.now() stamps)
 * instead of the bit-packed layout of the CUDA kernels, so the two implementations share no logic.
 *
 * Pinned against the reference itself: the .npz tapes under tests/golden are recorded from the unmodified
 * reference (oracle/gen_golden.py, run where /root/reference exists) and tests/test_oracle_golden.py
 * replays every one of them through this file, including the reference's own fixtures
 * (cyberbattle_env_test.py:43-98 Chain-10 solve; commandcontrol_test.py:14-73, total reward 389.0).
 *
 * Each function cites the reference lines it follows; paths are relative to /root/reference:
 *   ACT  = src/CyberBattleSim/cyberbattle/simulation/actions.py
 *   ENV  = src/CyberBattleSim/cyberbattle/_env/cyberbattle_env.py
 *   DEF  = src/CyberBattleSim/cyberbattle/_env/defender.py
 *   ATT  = marlon/baseline_models/env_wrappers/attack_wrapper.py
 *   DWR  = marlon/baseline_models/env_wrappers/defend_wrapper.py
 *   LDF  = marlon/defender_agents/defender.py
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cbx.h"

#define MAXN 256
#define MAXV 64

/* ---- scenario view ------------------------------------------------------------------------ */
typedef struct {
  const uint32_t* w;
  int n, P, nprops, L, R, nsecrets, ntriples, nservices, Ws;
  const uint32_t *node, *auth, *vuln, *payload, *triple;
} scn_t;

static int scn_parse(scn_t* s, const uint32_t* w, size_t nwords) {
  if (nwords < CBX_H_WORDS || w[CBX_H_MAGIC] != CBX_SCN_MAGIC || w[CBX_H_VERSION] != CBX_SCN_VERSION) return -1;
  if (w[CBX_H_TOTAL_WORDS] > nwords) return -1;
  s->w = w;
  s->n = (int)w[CBX_H_N_NODES];
  s->P = (int)w[CBX_H_N_PORTS];
  s->nprops = (int)w[CBX_H_N_PROPS];
  s->L = (int)w[CBX_H_N_LOCAL];
  s->R = (int)w[CBX_H_N_REMOTE];
  s->nsecrets = (int)w[CBX_H_N_SECRETS];
  s->ntriples = (int)w[CBX_H_N_TRIPLES];
  s->nservices = (int)w[CBX_H_N_SERVICES];
  s->Ws = (s->nsecrets + 31) / 32;
  s->node = w + w[CBX_H_OFF_NODE];
  s->auth = w + w[CBX_H_OFF_AUTH];
  s->vuln = w + w[CBX_H_OFF_VULN];
  s->payload = w + w[CBX_H_OFF_PAYLOAD];
  s->triple = w + w[CBX_H_OFF_TRIPLE];
  if (s->n > MAXN || s->L + s->R > MAXV) return -1;
  return 0;
}
static inline const uint32_t* scn_node(const scn_t* s, int i) { return s->node + (size_t)i * CBX_NODE_WORDS; }
static inline const uint32_t* scn_vuln(const scn_t* s, int i, int v) {
  return s->vuln + ((size_t)i * (s->L + s->R) + v) * CBX_VULN_WORDS;
}
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* ---- Philox4x32-10 (Salmon et al., SC'11), the counter-based stream of the built-in defender ---- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* 53-bit uniform in [0,1) from two words, the construction CPython's random.random() uses */
static inline double u53(uint32_t a, uint32_t b) { return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0; }

void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

/* ---- per-env state, shaped like the reference's objects ------------------------------------- */
typedef struct {
  /* model.NodeInfo, mutable part */
  uint8_t agent_installed, privilege_level, running /* status == Running */;
  uint8_t tags;           /* privilege_N strings appended to NodeInfo.properties (ACT:378) */
  int64_t last_reimaging; /* 0 = None */
  /* actions.NodeTrackingInformation (ACT:114-124) */
  uint8_t tracked;        /* node_id in AgentActions._discovered_nodes */
  int64_t last_owned_at;  /* 0 = None */
  uint64_t discovered_properties;
  int64_t last_attack[MAXV]; /* 0 = key absent; index = local ids then remote ids */
  /* DefenderAgentActions.node_reimaging_progress (ACT:689): -1 = key absent */
  int progress;
} onode_t;

typedef struct {
  onode_t* nodes;           /* live model.Environment + actuators */
  int* actuator_order;      /* OrderedDict order of AgentActions._discovered_nodes */
  int n_actuator;
  int* discovered;          /* CyberBattleEnv.__discovered_nodes */
  int n_discovered;
  int* cache;               /* CyberBattleEnv.__credential_cache (triple ids) */
  int n_cache;
  uint8_t* gathered;        /* AgentActions._gathered_credentials */
  int64_t clock;            /* logical datetime.now() */
  int stepcount, done;
  double episode_reward_sum; /* numpy.sum(__episode_rewards) */
  double availability;      /* live DefenderAgentActions.network_availability */
  /* LIVE defender binding: model.FirewallConfiguration rule lists as the LearningDefender edits them (LDF:50-69), one list
     object per alias group (several (node, direction) pairs may share one); a rule = (port-name index, ALLOW?) */
  struct fwlist { int n; uint8_t name[96]; uint8_t allow[96]; } * fw;
  /* the stale copy the MARLon defender acts on: initial environment + its own actuator */
  uint8_t* sh_running;
  int* sh_progress;
  double sh_availability;
  /* AttackerEnvWrapper */
  int att_timesteps, att_reset_request, att_valid, att_invalid, att_last_valid, att_last_invalid;
  int n_cyber_rewards, n_rewards;
  double last_cyber_reward, last_reward;
  double att_return;
  /* DefenderEnvWrapper */
  int def_timesteps, def_reset_request, has_breached, def_valid, def_invalid, def_last_valid, def_last_invalid;
  double last_attacker_reward, prev_availability, def_return;
} oenv_t;

typedef struct orc_batch {
  uint32_t* blob;
  scn_t s;
  cbx_config cfg;
  int64_t n;
  int N, C, LEAK, OW;
  oenv_t* envs;
  cbx_views v; /* HOST pointers */
  double stats[CBX_STAT_COUNT];
  int slice_of_kind[3]; /* column where kind k's coordinates start inside the MARLon MultiDiscrete action */
  /* firewall extension tables (cbx.h CBX_FX_*), NULL unless given; `live` = cfg.def_binding == LIVE with a MARLon defender */
  uint32_t* fwx;
  int live, n_fw_groups;
} orc_batch;

/* ---- firewall rule lists (live binding) ---- */
static int fw_group_of(const orc_batch* b, int node, int incoming) {
  uint32_t g = b->fwx[CBX_FX_WORDS + b->s.P + node];
  return incoming ? (int)(g & 0xFFFFu) : (int)(g >> 16);
}
/* ACT:504-515 __is_passing_firewall_rules: the first rule of that port name decides; no rule -> blocked */
static int fw_list_passes(const struct fwlist* l, int name) {
  for (int k = 0; k < l->n; ++k)
    if (l->name[k] == name) return l->allow[k];
  return 0;
}
static int fw_list_has(const struct fwlist* l, int name) {
  for (int k = 0; k < l->n; ++k)
    if (l->name[k] == name) return 1;
  return 0;
}
static void fw_lists_init(const orc_batch* b, oenv_t* e) {
  const uint32_t* g0 = b->fwx + CBX_FX_WORDS + b->s.P + b->s.n;
  const int names = (int)b->fwx[CBX_FX_N_NAMES];
  for (int g = 0; g < b->n_fw_groups; ++g) {
    struct fwlist* l = &e->fw[g];
    l->n = 0;
    for (int nm = 0; nm < names; ++nm)
      if (g0[2 * g] >> nm & 1u) { l->name[l->n] = (uint8_t)nm; l->allow[l->n] = (uint8_t)(g0[2 * g + 1] >> nm & 1u); l->n++; }
  }
}

static int64_t now(oenv_t* e) { return ++e->clock; }

/* ---- observation helpers -------------------------------------------------------------------- */
typedef struct { /* one env's slices of the view arrays */
  int32_t *scalars, *leaked, *cachem, *props, *priv;
  int8_t *local, *remote, *connect;
  uint32_t* owned;
} obs_t;

static obs_t obs_main(orc_batch* b, int64_t i) {
  obs_t o;
  const cbx_views* v = &b->v;
  o.scalars = v->scalars + i * 8;
  o.leaked = v->leaked_credentials + i * 4 * b->LEAK;
  o.cachem = v->credential_cache_matrix + i * 2 * b->C;
  o.props = v->discovered_nodes_properties + i * b->N * b->s.nprops;
  o.priv = v->nodes_privilegelevel + i * b->N;
  o.local = v->local_vulnerability ? v->local_vulnerability + i * b->N * b->s.L : NULL;
  o.remote = v->remote_vulnerability ? v->remote_vulnerability + i * (int64_t)b->N * b->N * b->s.R : NULL;
  o.connect = v->connect ? v->connect + i * (int64_t)b->N * b->N * b->s.P * b->C : NULL;
  o.owned = v->owned_bits + i * b->OW;
  return o;
}
static obs_t obs_term(orc_batch* b, int64_t i) {
  obs_t o;
  const cbx_views* v = &b->v;
  o.scalars = v->term_scalars + i * 8;
  o.leaked = v->term_leaked_credentials + i * 4 * b->LEAK;
  o.cachem = v->term_credential_cache_matrix + i * 2 * b->C;
  o.props = v->term_discovered_nodes_properties + i * b->N * b->s.nprops;
  o.priv = v->term_nodes_privilegelevel + i * b->N;
  o.local = v->term_local_vulnerability ? v->term_local_vulnerability + i * b->N * b->s.L : NULL;
  o.remote = v->term_remote_vulnerability ? v->term_remote_vulnerability + i * (int64_t)b->N * b->N * b->s.R : NULL;
  o.connect = v->term_connect ? v->term_connect + i * (int64_t)b->N * b->N * b->s.P * b->C : NULL;
  o.owned = NULL;
  return o;
}

static int find_external_index(const oenv_t* e, int node) { /* ENV:603-605 */
  for (int k = 0; k < e->n_discovered; ++k)
    if (e->discovered[k] == node) return k;
  return -1;
}

/* ENV:753-773 __get_blank_observation */
static void blank_observation(orc_batch* b, oenv_t* e, obs_t* o) {
  const scn_t* s = &b->s;
  memset(o->scalars, 0, 8 * sizeof(int32_t));
  o->scalars[6] = e->n_discovered; /* discovered_node_count=len(self.__discovered_nodes) */
  memset(o->leaked, 0, sizeof(int32_t) * 4 * b->LEAK);
  memset(o->cachem, 0, sizeof(int32_t) * 2 * b->C);
  for (int k = 0; k < b->N * s->nprops; ++k) o->props[k] = 2;
  memset(o->priv, 0, sizeof(int32_t) * b->N);
  if (o->local) {
    memset(o->local, 0, (size_t)b->N * s->L);
    memset(o->remote, 0, (size_t)b->N * b->N * s->R);
    memset(o->connect, 0, (size_t)b->N * b->N * s->P * b->C);
  }
  if (o->owned) memset(o->owned, 0, sizeof(uint32_t) * b->OW);
}

/* ENV:643-677 __update_action_mask */
static void update_action_mask(orc_batch* b, oenv_t* e, obs_t* o) {
  const scn_t* s = &b->s;
  const int N = b->N, L = s->L, R = s->R, P = s->P, C = b->C;
  for (int si = 0; si < e->n_discovered; ++si) {
    int src = e->discovered[si];
    if (!e->nodes[src].agent_installed) continue;
    if (o->owned) o->owned[si / 32] |= 1u << (si % 32);
    if (!o->local) continue;
    for (int v = 0; v < L; ++v)
      if (scn_vuln(s, src, v)[CBX_V_FLAGS] & 1u) o->local[si * L + v] = 1;
    for (int ti = 0; ti < e->n_discovered; ++ti) {
      for (int r = 0; r < R; ++r) o->remote[((int64_t)si * N + ti) * R + r] = 1;
      for (int p = 0; p < P; ++p)
        for (int c = 0; c < e->n_cache && c < C; ++c) o->connect[(((int64_t)si * N + ti) * P + p) * C + c] = 1;
    }
  }
}

/* ENV:811-830 property matrix (rows in ACTUATOR discovery order), ENV:840-857 privilege array (ENV order) */
static void property_and_privilege(orc_batch* b, oenv_t* e, obs_t* o) {
  const scn_t* s = &b->s;
  memset(o->props, 0, sizeof(int32_t) * b->N * s->nprops);
  for (int k = 0; k < e->n_actuator && k < b->N; ++k) {
    uint64_t dp = e->nodes[e->actuator_order[k]].discovered_properties;
    for (int p = 0; p < s->nprops; ++p)
      if (dp >> p & 1) o->props[k * s->nprops + p] = 1;
  }
  memset(o->priv, 0, sizeof(int32_t) * b->N);
  for (int k = 0; k < e->n_discovered && k < b->N; ++k) o->priv[k] = e->nodes[e->discovered[k]].privilege_level;
}

/* ---- AgentActions ---------------------------------------------------------------------------- */
typedef struct {
  double reward;
  int outcome;     /* CBX_RES_* */
  int level;       /* escalation level */
  const uint32_t* vrec; /* vulnerability record for Leaked* payloads */
  int error;       /* CBX_E_* (only when throws_on_invalid_actions) */
} result_t;

static void mark_tracked(oenv_t* e, int node) {
  if (!e->nodes[node].tracked) {
    e->nodes[node].tracked = 1;
    e->actuator_order[e->n_actuator++] = node;
  }
}

/* ACT:517-522 */
static int is_currently_owned(const oenv_t* e, int node, int64_t* last_owned_at) {
  const onode_t* nd = &e->nodes[node];
  *last_owned_at = nd->tracked ? nd->last_owned_at : 0;
  return *last_owned_at != 0 && (nd->last_reimaging == 0 || *last_owned_at >= nd->last_reimaging);
}

/* ACT:234-249 (all properties of the node) */
static int mark_properties(oenv_t* e, int node, uint64_t props) {
  onode_t* nd = &e->nodes[node];
  mark_tracked(e, node);
  int before = __builtin_popcountll(nd->discovered_properties);
  nd->discovered_properties |= props;
  return __builtin_popcountll(nd->discovered_properties) - before;
}

/* ACT:251-275 __mark_node_as_owned */
static int mark_node_as_owned(const scn_t* s, oenv_t* e, int node, int privilege, int64_t* last_owned_at) {
  onode_t* nd = &e->nodes[node];
  int owned = is_currently_owned(e, node, last_owned_at);
  if (!owned) {
    mark_tracked(e, node);
    nd->agent_installed = 1;
    if (privilege > nd->privilege_level) nd->privilege_level = (uint8_t)privilege; /* model.escalate */
    const uint32_t* rec = scn_node(s, node);
    mark_properties(e, node, (uint64_t)rec[CBX_N_PROPS_LO] | ((uint64_t)rec[CBX_N_PROPS_HI] << 32));
    nd->last_owned_at = now(e);
  }
  return owned;
}

/* ACT:325-423 __process_outcome */
static result_t process_outcome(orc_batch* b, oenv_t* e, int node, int v, int is_local, double failed_penalty) {
  const scn_t* s = &b->s;
  onode_t* nd = &e->nodes[node];
  result_t r = {0.0, CBX_RES_NONE, 0, NULL, 0};
  if (!nd->running) { r.reward = 0.0; return r; }              /* MACHINE_NOT_RUNNING */
  const uint32_t* rec = scn_vuln(s, node, v);
  if (!(rec[CBX_V_FLAGS] & 1u)) { r.reward = -5.0; return r; } /* SUPSPICIOUSNESS */
  int kind = (rec[CBX_V_FLAGS] >> 1) & 7;
  int level = (rec[CBX_V_FLAGS] >> 4) & 3;
  int truth = (rec[CBX_V_FLAGS] >> (8 + nd->tags)) & 1;         /* _check_prerequisites, ACT:158-171 */
  if (!truth) { r.reward = failed_penalty; r.outcome = CBX_RES_EXPLOIT_FAILED; return r; }
  double reward = 0;
  r.vrec = rec;
  switch (kind) {
    case CBX_OUT_ESCALATION: {
      r.outcome = CBX_RES_ESCALATION; r.level = level;
      if (nd->tags >> level & 1) { r.reward = -1.0; return r; } /* REPEAT, ACT:370-371 */
      int64_t last;
      mark_node_as_owned(s, e, node, level, &last);
      if (!last) reward += (double)(int32_t)scn_node(s, node)[CBX_N_VALUE];
      nd->tags |= (uint8_t)(1u << level);
    } break;
    case CBX_OUT_LATERAL_MOVE: {
      r.outcome = CBX_RES_LATERAL_MOVE;
      int64_t last;
      mark_node_as_owned(s, e, node, 1, &last);
      if (!last) reward += (double)(int32_t)scn_node(s, node)[CBX_N_VALUE];
    } break;
    case CBX_OUT_PROBE_SUCCEEDED: {
      r.outcome = CBX_RES_PROBE_SUCCEEDED;
      const uint32_t* pl = s->payload + rec[CBX_V_PAYLOAD_OFF];
      reward += 2 * mark_properties(e, node, (uint64_t)pl[0] | ((uint64_t)pl[1] << 32));
    } break;
    case CBX_OUT_LEAKED_CREDENTIALS: r.outcome = CBX_RES_LEAKED_CREDENTIALS; break;
    case CBX_OUT_LEAKED_NODES: r.outcome = CBX_RES_LEAKED_NODES; break;
    case CBX_OUT_CUSTOMER_DATA: r.outcome = CBX_RES_CUSTOMER_DATA; break;
    case CBX_OUT_PROBE_FAILED: r.outcome = CBX_RES_PROBE_FAILED; break;
    default: r.outcome = CBX_RES_EXPLOIT_FAILED; break;
  }
  mark_tracked(e, node); /* ACT:393-394 */
  if (nd->last_attack[v]) {                                   /* ACT:396-407 */
    if (nd->last_reimaging == 0 || nd->last_attack[v] >= nd->last_reimaging) reward += -1.0;
  } else {
    reward += 7.0;
  }
  nd->last_attack[v] = now(e);
  (void)is_local; /* the key (vulnerability_id, local_or_remote) is unique per v because ids are per-type */
  /* ACT:277-310 __mark_discovered_entities */
  int new_nodes = 0, new_creds = 0;
  if (kind == CBX_OUT_LEAKED_CREDENTIALS) {
    const uint32_t* pl = s->payload + rec[CBX_V_PAYLOAD_OFF];
    for (uint32_t k = 0; k < rec[CBX_V_PAYLOAD_CNT]; ++k) {
      const uint32_t* t = s->triple + 3 * pl[k];
      if (!e->nodes[t[0]].tracked) { mark_tracked(e, (int)t[0]); new_nodes++; }
      if (!e->gathered[t[2]]) { e->gathered[t[2]] = 1; new_creds++; }
    }
  } else if (kind == CBX_OUT_LEAKED_NODES) {
    const uint32_t* pl = s->payload + rec[CBX_V_PAYLOAD_OFF];
    for (uint32_t k = 0; k < rec[CBX_V_PAYLOAD_CNT]; ++k)
      if (!e->nodes[pl[k]].tracked) { mark_tracked(e, (int)pl[k]); new_nodes++; }
  }
  reward += new_nodes * 5.0;
  reward += new_creds * 3.0;
  reward -= (double)u2f(rec[CBX_V_COST]);
  r.reward = reward;
  return r;
}

static result_t invalid_action(orc_batch* b, int code) { /* throws_on_invalid_actions switch */
  result_t r = {-1.0, CBX_RES_NONE, 0, NULL, 0};
  if (b->cfg.throws_on_invalid_actions) r.error = code;
  return r;
}

/* ACT:473-502 */
static result_t exploit_local(orc_batch* b, oenv_t* e, int node, int v) {
  if (!e->nodes[node].agent_installed) return invalid_action(b, CBX_E_SOURCE_NOT_OWNED);
  return process_outcome(b, e, node, v, 1, -20.0);
}
/* ACT:425-471 */
static result_t exploit_remote(orc_batch* b, oenv_t* e, int src, int tgt, int v) {
  if (!e->nodes[src].agent_installed) return invalid_action(b, CBX_E_SOURCE_NOT_OWNED);
  if (!e->nodes[tgt].tracked) return invalid_action(b, CBX_E_TARGET_NOT_DISCOVERED);
  return process_outcome(b, e, tgt, b->s.L + v, 0, -50.0);
}
/* ACT:524-606 */
static result_t connect_to_remote(orc_batch* b, oenv_t* e, int src, int tgt, int port, int secret) {
  const scn_t* s = &b->s;
  result_t r = {0.0, CBX_RES_NONE, 0, NULL, 0};
  if (!e->nodes[src].agent_installed) return invalid_action(b, CBX_E_SOURCE_NOT_OWNED);
  if (!e->nodes[tgt].tracked) return invalid_action(b, CBX_E_TARGET_NOT_DISCOVERED);
  if (!e->gathered[secret]) return invalid_action(b, CBX_E_CREDENTIAL_NOT_GATHERED);
  if (b->live) {
    const int name = (int)b->fwx[CBX_FX_WORDS + port];
    if (!fw_list_passes(&e->fw[fw_group_of(b, src, 0)], name)) { r.reward = -10.0; return r; }
    if (!fw_list_passes(&e->fw[fw_group_of(b, tgt, 1)], name)) { r.reward = -10.0; return r; }
  } else {
    if (!(scn_node(s, src)[CBX_N_FW_OUT] >> port & 1)) { r.reward = -10.0; return r; }
    if (!(scn_node(s, tgt)[CBX_N_FW_IN] >> port & 1)) { r.reward = -10.0; return r; }
  }
  if (!(scn_node(s, tgt)[CBX_N_LISTEN] >> port & 1)) { r.reward = -10.0; return r; }
  if (!e->nodes[tgt].running) { r.reward = 0.0; return r; }
  const uint32_t* auth = s->auth + ((size_t)tgt * s->P + port) * s->Ws;
  if (!(auth[secret / 32] >> (secret % 32) & 1)) { r.reward = -10.0; return r; }
  int64_t last;
  int already = mark_node_as_owned(s, e, tgt, 1, &last);
  r.outcome = CBX_RES_LATERAL_MOVE;
  if (already) { r.reward = -1.0; return r; }
  r.reward = last == 0 ? (double)(int32_t)scn_node(s, tgt)[CBX_N_VALUE] : 0.0;
  return r;
}

/* ---- DefenderAgentActions -------------------------------------------------------------------- */
/* ACT:700-712 */
static void reimage_node_live(oenv_t* e, int node) {
  onode_t* nd = &e->nodes[node];
  nd->progress = 15;
  nd->agent_installed = 0;
  nd->privilege_level = 0;
  nd->running = 0;
  nd->last_reimaging = now(e);
}
/* ACT:714-746 */
static double tick_and_availability(int n, uint8_t* running, int* progress, size_t stride_r, size_t stride_p,
                                    const scn_t* s) {
  for (int i = 0; i < n; ++i) {
    int* p = (int*)((char*)progress + i * stride_p);
    uint8_t* run = (uint8_t*)((char*)running + i * stride_r);
    if (*p < 0) continue;
    if (*p > 0) *p -= 1;
    else { *run = 1; *p = -1; }
  }
  double total = 0, avail = 0;
  for (int i = 0; i < n; ++i) {
    uint8_t run = *(uint8_t*)((char*)running + i * stride_r);
    int nsvc = (int)(scn_node(s, i)[CBX_N_FLAGS] >> 8 & 0xFF);
    double tw = 0, rw = 0;
    for (int k = 0; k < nsvc; ++k) { tw += 1.0; rw += 1.0 * 1; } /* services never stop */
    double adj = run ? (1 + rw) / (1 + tw) : 0.0;
    total += 1.0;
    avail += adj * 1.0;
  }
  return avail / total;
}

/* DEF:42-55 ScanAndReimageCompromisedMachines.step */
static void scan_and_reimage(orc_batch* b, oenv_t* e, int64_t env_index, const double* scan_u, const double* detect_u) {
  const cbx_config* c = &b->cfg;
  if (e->stepcount % c->scan_frequency != 0) return;
  const int n = b->s.n, cap = c->scan_capacity;
  for (int k = 0; k < cap; ++k) {
    double us, ud;
    if (scan_u) {
      us = scan_u[env_index * cap + k];
      ud = detect_u[env_index * cap + k];
    } else {
      const int64_t ge = env_index + c->env_index_base;
      uint32_t ctr[4] = {(uint32_t)ge, (uint32_t)(ge >> 32), (uint32_t)e->stepcount, (uint32_t)k};
      uint32_t key[2] = {(uint32_t)c->seed, (uint32_t)(c->seed >> 32)}, o[4];
      philox4x32_10(ctr, key, o);
      us = u53(o[0], o[1]);
      ud = u53(o[2], o[3]);
    }
    int node = (int)floor(us * n); /* random.choices: population[floor(random() * n)] */
    onode_t* nd = &e->nodes[node];
    if (nd->running && nd->agent_installed) {
      int detected = ud <= c->scan_probability;
      if (detected && (scn_node(&b->s, node)[CBX_N_FLAGS] & 1u)) reimage_node_live(e, node);
    }
  }
}

/* ---- CyberBattleEnv --------------------------------------------------------------------------- */
static int owned_count(const orc_batch* b, const oenv_t* e) { /* ACT:317-319 */
  int c = 0;
  for (int i = 0; i < b->s.n; ++i) c += e->nodes[i].privilege_level >= 1;
  return c;
}
/* ENV:1080-1101 */
static int attacker_goal_reached(const orc_batch* b, const oenv_t* e) {
  const cbx_config* c = &b->cfg;
  if (!c->has_attacker_goal) return 0;
  if (e->episode_reward_sum < c->goal_reward) return 0;
  int oc = owned_count(b, e);
  if (oc < c->goal_own_atleast) return 0;
  if ((double)oc / (double)b->s.n < c->goal_own_atleast_percent) return 0;
  if (c->builtin_defender != CBX_BUILTIN_NONE && e->availability >= c->goal_low_availability) return 0;
  return 1;
}
/* ENV:1103-1116 */
static int defender_constraints_broken(const orc_batch* b, const oenv_t* e) {
  return b->cfg.builtin_defender != CBX_BUILTIN_NONE && e->availability < b->cfg.maintain_sla;
}
static int defender_goal_reached(const orc_batch* b, const oenv_t* e) {
  return b->cfg.defender_goal_eviction && owned_count(b, e) == 0;
}

/* ENV:375-394 __reset_environment + ENV:1187-1209 reset */
static void cyber_reset(orc_batch* b, oenv_t* e, obs_t* o) {
  const scn_t* s = &b->s;
  for (int i = 0; i < s->n; ++i) {
    onode_t* nd = &e->nodes[i];
    memset(nd, 0, sizeof(*nd));
    uint32_t f = scn_node(s, i)[CBX_N_FLAGS];
    nd->agent_installed = f >> 1 & 1;
    nd->privilege_level = f >> 2 & 3;
    nd->running = 1;
    nd->progress = -1;
  }
  e->n_actuator = e->n_discovered = e->n_cache = 0;
  memset(e->gathered, 0, (size_t)s->nsecrets);
  e->clock = 0;
  e->stepcount = 0;
  e->done = 0;
  e->episode_reward_sum = 0;
  e->availability = 1.0;
  if (b->live) {
    fw_lists_init(b, e);      /* deepcopy of the initial environment: its rule lists, sharing preserved */
    e->sh_availability = 1.0; /* the defender wrapper's actuator IS the fresh live one */
  }
  /* AgentActions.__init__ (ACT:149-152): owned nodes are marked owned (LocalUser) */
  for (int i = 0; i < s->n; ++i)
    if (e->nodes[i].agent_installed) { int64_t last; mark_node_as_owned(s, e, i, 1, &last); }
  for (int i = 0; i < s->n; ++i)
    if (e->nodes[i].agent_installed) e->discovered[e->n_discovered++] = i;
  if (o) {
    blank_observation(b, e, o);
    update_action_mask(b, e, o);
    property_and_privilege(b, e, o);
  }
}

/* ENV:707-751 __execute_action + ENV:584-601; returns 1 on OutOfBoundIndexError */
static int execute_action(orc_batch* b, oenv_t* e, int kind, const int32_t* a, result_t* out) {
  const scn_t* s = &b->s;
  if (kind == CBX_KIND_LOCAL) {
    if (a[0] < 0 || a[0] >= e->n_discovered) return 1;
    *out = exploit_local(b, e, e->discovered[a[0]], a[1]);
  } else if (kind == CBX_KIND_REMOTE) {
    if (a[0] < 0 || a[0] >= e->n_discovered) return 1;
    if (a[1] < 0 || a[1] >= e->n_discovered) return 1;
    *out = exploit_remote(b, e, e->discovered[a[0]], e->discovered[a[1]], a[2]);
  } else {
    if (a[3] < 0 || a[3] >= e->n_cache) { result_t r = {-1.0, CBX_RES_NONE, 0, NULL, 0}; *out = r; return 0; } /* ENV:736-737 */
    if (a[0] < 0 || a[0] >= e->n_discovered) return 1;
    if (a[1] < 0 || a[1] >= e->n_discovered) return 1;
    *out = connect_to_remote(b, e, e->discovered[a[0]], e->discovered[a[1]], a[2], (int)s->triple[3 * e->cache[a[3]] + 2]);
  }
  return 0;
}

/* ENV:859-933 __observation_reward_from_action_result */
static void observation_from_result(orc_batch* b, oenv_t* e, const result_t* r, obs_t* o) {
  const scn_t* s = &b->s;
  blank_observation(b, e, o);
  if (r->outcome == CBX_RES_LEAKED_NODES) {
    const uint32_t* pl = s->payload + r->vrec[CBX_V_PAYLOAD_OFF];
    int cnt = 0;
    for (uint32_t k = 0; k < r->vrec[CBX_V_PAYLOAD_CNT]; ++k)
      if (find_external_index(e, (int)pl[k]) < 0) { e->discovered[e->n_discovered++] = (int)pl[k]; cnt++; }
    o->scalars[0] = cnt;
  } else if (r->outcome == CBX_RES_LEAKED_CREDENTIALS) {
    const uint32_t* pl = s->payload + r->vrec[CBX_V_PAYLOAD_OFF];
    int cnt = 0, nleak = 0;
    int added[1024];
    for (uint32_t k = 0; k < r->vrec[CBX_V_PAYLOAD_CNT]; ++k) {
      int t = (int)pl[k];
      int node = (int)s->triple[3 * t];
      if (find_external_index(e, node) < 0) { e->discovered[e->n_discovered++] = node; cnt++; }
      int present = 0;
      for (int c = 0; c < e->n_cache; ++c) present |= e->cache[c] == t;
      if (!present) { e->cache[e->n_cache++] = t; if (nleak < 1024) added[nleak++] = e->n_cache - 1; }
    }
    o->scalars[0] = cnt;
    for (int k = 0; k < nleak && k < b->LEAK; ++k) { /* ENV:890-907 */
      int t = e->cache[added[k]];
      o->leaked[4 * k + 0] = 1;
      o->leaked[4 * k + 1] = added[k];
      o->leaked[4 * k + 2] = find_external_index(e, (int)s->triple[3 * t]);
      o->leaked[4 * k + 3] = (int)s->triple[3 * t + 1];
    }
  } else if (r->outcome == CBX_RES_LATERAL_MOVE) o->scalars[1] = 1;
  else if (r->outcome == CBX_RES_CUSTOMER_DATA) o->scalars[2] = 1;
  else if (r->outcome == CBX_RES_PROBE_SUCCEEDED) o->scalars[3] = 2;
  else if (r->outcome == CBX_RES_PROBE_FAILED) o->scalars[3] = 1;
  else if (r->outcome == CBX_RES_ESCALATION) o->scalars[4] = r->level;
  for (int c = 0; c < e->n_cache && c < b->C; ++c) { /* ENV:920-922 */
    int t = e->cache[c];
    o->cachem[2 * c] = find_external_index(e, (int)s->triple[3 * t]);
    o->cachem[2 * c + 1] = (int)s->triple[3 * t + 1];
  }
  o->scalars[5] = e->n_cache;
  o->scalars[6] = e->n_discovered;
  property_and_privilege(b, e, o);
  update_action_mask(b, e, o);
}

typedef struct { double reward, raw; int terminated, outcome, error, oob; } cstep_t;

/* ENV:1145-1185 CyberBattleEnv.step */
static cstep_t cyber_step(orc_batch* b, oenv_t* e, int64_t env_index, int kind, const int32_t* a, obs_t* o,
                          const double* scan_u, const double* detect_u) {
  cstep_t st = {0.0, 0.0, 0, CBX_RES_NONE, 0, 0};
  if (e->done) { st.error = CBX_E_STEP_AFTER_DONE; st.terminated = 1; return st; }
  e->stepcount += 1;
  result_t r;
  if (execute_action(b, e, kind, a, &r)) { /* OutOfBoundIndexError, ENV:1171-1174 */
    blank_observation(b, e, o);
    o->scalars[7] = 1;
    st.oob = 1;
    st.outcome = CBX_RES_OUT_OF_BOUND;
    st.reward = 0.0;
    e->episode_reward_sum += 0.0;
    return st;
  }
  if (r.error) { st.error = r.error; return st; } /* ValueError propagates out of step(); nothing else happened */
  observation_from_result(b, e, &r, o);
  double reward = r.reward;
  st.raw = r.reward;
  st.outcome = r.outcome;
  if (b->cfg.builtin_defender != CBX_BUILTIN_NONE) { /* ENV:1156-1158 */
    e->availability = tick_and_availability(b->s.n, &e->nodes[0].running, &e->nodes[0].progress, sizeof(onode_t),
                                            sizeof(onode_t), &b->s);
    scan_and_reimage(b, e, env_index, scan_u, detect_u);
  }
  if (attacker_goal_reached(b, e) || defender_constraints_broken(b, e)) { e->done = 1; reward = b->cfg.winning_reward; }
  else if (defender_goal_reached(b, e)) { e->done = 1; reward = b->cfg.losing_reward; }
  else reward = reward > 0.0 ? reward : 0.0;
  e->episode_reward_sum += reward;
  st.reward = reward;
  st.terminated = e->done;
  return st;
}

/* ---- MARLon wrappers -------------------------------------------------------------------------- */
static void defender_observe(orc_batch* b, oenv_t* e, int64_t i, int term) { /* DWR:492-534 (live env) */
  const scn_t* s = &b->s;
  int8_t* inf = (term ? b->v.term_def_infected_nodes : b->v.def_infected_nodes) + i * s->n;
  for (int k = 0; k < s->n; ++k) inf[k] = e->nodes[k].agent_installed ? 1 : 0;
  if (term) return;
  int8_t* fin = b->v.def_incoming_firewall + i * 6 * s->n;
  int8_t* fout = b->v.def_outgoing_firewall + i * 6 * s->n;
  int8_t* svc = b->v.def_services_status + i * s->nservices;
  for (int k = 0; k < s->n; ++k) {
    uint32_t d = scn_node(s, k)[CBX_N_DEFOBS];
    for (int r = 0; r < 6; ++r) {
      if (b->live) { /* DWR:506-517 on the env's own lists */
        fin[6 * k + r] = (int8_t)fw_list_has(&e->fw[fw_group_of(b, k, 1)], r);
        fout[6 * k + r] = (int8_t)fw_list_has(&e->fw[fw_group_of(b, k, 0)], r);
      } else { fin[6 * k + r] = d >> r & 1; fout[6 * k + r] = d >> (8 + r) & 1; }
    }
  }
  for (int k = 0; k < s->nservices; ++k) svc[k] = 1;
}

static void stats_episode(orc_batch* b, int who, double ret, int len, int valid, int invalid) {
  if (who == 0) {
    b->stats[CBX_STAT_EPISODES] += 1;
    b->stats[CBX_STAT_ATT_RETURN] += ret;
    b->stats[CBX_STAT_ATT_RETURN_SQ] += ret * ret;
    b->stats[CBX_STAT_EP_LEN] += len;
    b->stats[CBX_STAT_EP_LEN_SQ] += (double)len * len;
    b->stats[CBX_STAT_ATT_VALID] += valid;
    b->stats[CBX_STAT_ATT_INVALID] += invalid;
  } else {
    b->stats[CBX_STAT_DEF_RETURN] += ret;
    b->stats[CBX_STAT_DEF_RETURN_SQ] += ret * ret;
    b->stats[CBX_STAT_DEF_VALID] += valid;
    b->stats[CBX_STAT_DEF_INVALID] += invalid;
  }
}

/* ATT:400-468 AttackerEnvWrapper.reset */
static void attacker_reset(orc_batch* b, oenv_t* e, obs_t* o) {
  if (!e->att_reset_request) { /* EnvironmentEventSource.notify_reset -> on_reset of both observers */
    double last_reward = e->n_rewards > 0 ? e->last_reward : 0.0;
    e->att_reset_request = 1;
    e->def_reset_request = 1;
    e->last_attacker_reward = last_reward;
  }
  e->att_last_valid = e->att_valid;
  e->att_last_invalid = e->att_invalid;
  cyber_reset(b, e, o);
  e->att_reset_request = 0;
  e->att_valid = e->att_invalid = 0;
  e->att_timesteps = 0;
  e->n_cyber_rewards = e->n_rewards = 0;
  e->last_cyber_reward = e->last_reward = 0.0;
  e->att_return = 0.0;
}

/* DWR:414-477 DefenderEnvWrapper.reset */
static void defender_reset(orc_batch* b, oenv_t* e, int64_t i) {
  if (!e->def_reset_request) {
    e->att_reset_request = 1;
    e->def_reset_request = 1;
    e->last_attacker_reward = 0.0;
  }
  cyber_reset(b, e, NULL); /* the observation of this reset is dropped by the wrapper (DWR:449-453) */
  e->def_reset_request = 0;
  e->last_attacker_reward = 0.0; /* None */
  e->def_timesteps = 0;
  e->def_last_valid = e->def_valid;
  e->def_last_invalid = e->def_invalid;
  e->def_valid = e->def_invalid = 0;
  e->has_breached = 0;
  e->prev_availability = e->sh_availability;
  e->def_return = 0.0;
  defender_observe(b, e, i, 0);
}

/* DWR:329-412 is_defender_action_valid (reads the LIVE environment) */
static int defender_action_valid(const orc_batch* b, const oenv_t* e, const int32_t* a) {
  const scn_t* s = &b->s;
  /* values the MultiDiscrete space cannot produce (negative, node >= n, rule >= 6) count as an invalid action: the
     reference would raise IndexError for that env; a batch must not */
#define NODE_OK(x) ((x) >= 0 && (x) < s->n && e->nodes[(x)].running)
  switch (a[0]) {
    case 0: return NODE_OK(a[1]) && (scn_node(s, a[1])[CBX_N_FLAGS] & 1u);
    case 1:
      if (!(NODE_OK(a[2]) && a[3] >= 0 && a[3] < 6)) return 0;
      if (b->live) return fw_list_has(&e->fw[fw_group_of(b, a[2], a[4] != 0)], a[3]);
      return (scn_node(s, a[2])[CBX_N_DEFOBS] >> ((a[4] ? 0 : 8) + a[3]) & 1u);
    case 2: return NODE_OK(a[5]);
    case 3: return NODE_OK(a[8]) && a[9] >= 0 && a[9] < (int)(scn_node(s, a[8])[CBX_N_FLAGS] >> 8 & 0xFF);
    case 4: return NODE_OK(a[10]) && a[11] >= 0 && a[11] < (int)(scn_node(s, a[10])[CBX_N_FLAGS] >> 8 & 0xFF);
    default: return 0;
  }
#undef NODE_OK
}

static void defender_half_step(orc_batch* b, int64_t i, const int32_t* da);

static void marlon_pair_step(orc_batch* b, int64_t i, const int32_t* aa, const int32_t* da, const double* scan_u,
                             const double* detect_u, int who) {
  oenv_t* e = &b->envs[i];
  const cbx_config* c = &b->cfg;
  obs_t o = obs_main(b, i);
  if (!(who & CBX_WHO_ATTACKER)) {
    if (c->def_enabled && (who & CBX_WHO_DEFENDER)) defender_half_step(b, i, da);
    return;
  }
  int32_t* info = b->v.att_info + i * 8;
  memset(info, 0, 8 * sizeof(int32_t));
  /* ---------------- AttackerEnvWrapper.step, ATT:255-398 ---------------- */
  /* action[0] outside 0..2 or a negative coordinate: not producible by the MultiDiscrete space (the reference would raise
     for this env); treated as the wrapper-level invalid action */
  int kind_ok = aa[0] >= 0 && aa[0] <= 2;
  int kind = c->kind_of_index[kind_ok ? aa[0] : 0];
  const int32_t* coords = aa + b->slice_of_kind[kind];
  int in_range; /* ATT:233-253 */
  if (kind == CBX_KIND_LOCAL) in_range = coords[0] >= 0 && coords[0] < e->n_discovered;
  else in_range = coords[0] >= 0 && coords[0] < e->n_discovered && coords[1] >= 0 && coords[1] < e->n_discovered;
  in_range = in_range && kind_ok;
  double reward_modifier = 0.0, reward = 0.0, cyber_reward = 0.0;
  int terminated = 0, truncated = 0;
  if (!in_range) {
    e->att_invalid += 1;
    reward_modifier += c->att_invalid_action_reward_modifier;
    info[5] = 1;
    info[0] = (int32_t)f2u(0.0f);
    /* observation: the previous _last_transformed_observation, i.e. the buffers stay as they are */
  } else {
    e->att_valid += 1;
    cstep_t st = cyber_step(b, e, i, kind, coords, &o, scan_u, detect_u);
    reward = st.reward;
    cyber_reward = st.reward;
    terminated = st.terminated;
    info[0] = (int32_t)f2u((float)st.reward);
    info[1] = (int32_t)f2u((float)st.raw);
    info[2] = st.outcome;
    info[3] = st.error;
  }
  e->n_cyber_rewards += 1;
  e->last_cyber_reward = reward;
  e->att_timesteps += 1;
  if (e->att_reset_request) truncated = 1;
  if (e->att_timesteps >= c->att_max_timesteps) truncated = 1;
  reward = reward + reward_modifier;
  e->n_rewards += 1;
  e->last_reward = reward;
  e->att_return += reward;
  info[4] = e->stepcount;
  b->v.att_reward[i] = (float)reward;
  b->v.att_terminated[i] = (uint8_t)terminated;
  b->v.att_truncated[i] = (uint8_t)truncated;
  b->v.network_availability[i] = e->availability;
  b->stats[CBX_STAT_ENV_STEPS] += 1;
  if (terminated || truncated) {
    info[6] = e->att_timesteps;
    stats_episode(b, 0, e->att_return, e->att_timesteps, e->att_valid, e->att_invalid);
    if (terminated && cyber_reward == c->winning_reward) b->stats[CBX_STAT_ATT_WINS] += 1; /* ATT:330-332 */
    if (!terminated && e->att_timesteps >= c->att_max_timesteps) b->stats[CBX_STAT_TIMEOUTS] += 1;
    if (c->auto_reset) { /* DummyVecEnv.step_wait: keep the terminal observation, then reset */
      if (c->emit_terminal_obs) {
        obs_t t = obs_term(b, i);
        memcpy(t.scalars, o.scalars, 8 * sizeof(int32_t));
        memcpy(t.leaked, o.leaked, sizeof(int32_t) * 4 * b->LEAK);
        memcpy(t.cachem, o.cachem, sizeof(int32_t) * 2 * b->C);
        memcpy(t.props, o.props, sizeof(int32_t) * b->N * b->s.nprops);
        memcpy(t.priv, o.priv, sizeof(int32_t) * b->N);
        if (t.local) {
          memcpy(t.local, o.local, (size_t)b->N * b->s.L);
          memcpy(t.remote, o.remote, (size_t)b->N * b->N * b->s.R);
          memcpy(t.connect, o.connect, (size_t)b->N * b->N * b->s.P * b->C);
        }
      }
      attacker_reset(b, e, &o);
    }
  }
  if (!c->def_enabled || !(who & CBX_WHO_DEFENDER)) return;
  defender_half_step(b, i, da);
}

/* ---------------- DefenderEnvWrapper.step, DWR:197-327 ---------------- */
static void defender_half_step(orc_batch* b, int64_t i, const int32_t* da) {
  oenv_t* e = &b->envs[i];
  const cbx_config* c = &b->cfg;
  double dreward = 0.0;
  int dterm = 0, dtrunc = 0;
  int empty = da[0] < 0;
  int valid = empty ? 1 : defender_action_valid(b, e, da);
  if (!valid) { e->def_invalid += 1; dreward += c->def_invalid_action_reward; }
  else e->def_valid += 1;
  double cur;
  if (b->live) {
    /* LearningDefender.executeAction on the LIVE environment (binding refreshed at every CyberBattleEnv.reset), LDF:31-107 */
    e->availability = tick_and_availability(b->s.n, &e->nodes[0].running, &e->nodes[0].progress, sizeof(onode_t), sizeof(onode_t), &b->s);
    if (valid && !empty) {
      if (da[0] == 0) reimage_node_live(e, da[1]);
      else if (da[0] == 1) { /* block_traffic: every rule of that name leaves the selected list */
        struct fwlist* l = &e->fw[fw_group_of(b, da[2], da[4] != 0)];
        int m = 0;
        for (int k = 0; k < l->n; ++k)
          if (l->name[k] != da[3]) { l->name[m] = l->name[k]; l->allow[m] = l->allow[k]; m++; }
        l->n = m;
      } else if (da[0] == 2) { /* allow_traffic: no rule of that name in the selected list -> append ALLOW to INCOMING */
        if (!fw_list_has(&e->fw[fw_group_of(b, da[5], da[7] != 0)], da[6])) {
          struct fwlist* l = &e->fw[fw_group_of(b, da[5], 1)];
          if (l->n < 96) { l->name[l->n] = (uint8_t)da[6]; l->allow[l->n] = 1; l->n++; }
        }
      } /* stop / start service: DefenderAgentActions compares a ListeningService object with a name, never equal (B.4) */
    }
    cur = e->availability;
    e->sh_availability = cur; /* the wrapper's _actuator IS the live one */
  } else {
    /* LearningDefender.executeAction on the STALE copy, LDF:31-107 */
    e->sh_availability = tick_and_availability(b->s.n, e->sh_running, e->sh_progress, sizeof(uint8_t), sizeof(int), &b->s);
    if (valid && !empty && da[0] == 0) { e->sh_progress[da[1]] = 15; e->sh_running[da[1]] = 0; }
    cur = e->sh_availability;
  }
  double worsening = e->prev_availability - cur;
  if (e->n_cyber_rewards > 0) dreward += -1.0 * e->last_cyber_reward;
  if (cur < c->maintain_sla) {
    if (!e->has_breached) {
      dreward += c->def_loss_reward;
      if (c->def_reset_on_constraint_broken) dterm = 1;
      e->has_breached = 1;
      b->stats[CBX_STAT_SLA_BREACHES] += 1;
    } else if (worsening > 0) dreward += -c->def_sla_worsening_penalty_scale * worsening;
  } else e->has_breached = 0;
  e->prev_availability = cur;
  if (defender_goal_reached(b, e)) { dreward = c->winning_reward; dterm = 1; }
  defender_observe(b, e, i, 0);
  e->def_timesteps += 1;
  if (e->def_reset_request) { dtrunc = 1; dreward = -1.0 * e->last_attacker_reward; }
  else if (e->def_timesteps >= c->def_max_timesteps) dtrunc = 1;
  e->def_return += dreward;
  b->v.def_reward[i] = (float)dreward;
  b->v.def_terminated[i] = (uint8_t)dterm;
  b->v.def_truncated[i] = (uint8_t)dtrunc;
  if (dterm || dtrunc) {
    stats_episode(b, 1, e->def_return, e->def_timesteps, e->def_valid, e->def_invalid);
    if (c->auto_reset) {
      if (c->emit_terminal_obs) defender_observe(b, e, i, 1);
      defender_reset(b, e, i);
    }
  }
}

static void cyber_only_step(orc_batch* b, int64_t i, const int32_t* a, const double* scan_u, const double* detect_u) {
  oenv_t* e = &b->envs[i];
  obs_t o = obs_main(b, i);
  int32_t* info = b->v.att_info + i * 8;
  memset(info, 0, 8 * sizeof(int32_t));
  cstep_t st = cyber_step(b, e, i, a[0], a + 1, &o, scan_u, detect_u);
  info[0] = (int32_t)f2u((float)st.reward);
  info[1] = (int32_t)f2u((float)st.raw);
  info[2] = st.outcome;
  info[3] = st.error;
  info[4] = e->stepcount;
  b->v.att_reward[i] = (float)st.reward;
  b->v.att_terminated[i] = (uint8_t)st.terminated;
  b->v.att_truncated[i] = 0;
  b->v.network_availability[i] = e->availability;
  if (st.error == CBX_E_STEP_AFTER_DONE) return;
  b->stats[CBX_STAT_ENV_STEPS] += 1;
  e->att_return += st.reward;
  if (st.terminated) {
    info[6] = e->stepcount;
    stats_episode(b, 0, e->att_return, e->stepcount, 0, 0);
    if (st.reward == b->cfg.winning_reward) b->stats[CBX_STAT_ATT_WINS] += 1;
    if (b->cfg.auto_reset) {
      if (b->cfg.emit_terminal_obs) {
        obs_t t = obs_term(b, i);
        memcpy(t.scalars, o.scalars, 8 * sizeof(int32_t));
        memcpy(t.leaked, o.leaked, sizeof(int32_t) * 4 * b->LEAK);
        memcpy(t.cachem, o.cachem, sizeof(int32_t) * 2 * b->C);
        memcpy(t.props, o.props, sizeof(int32_t) * b->N * b->s.nprops);
        memcpy(t.priv, o.priv, sizeof(int32_t) * b->N);
        if (t.local) {
          memcpy(t.local, o.local, (size_t)b->N * b->s.L);
          memcpy(t.remote, o.remote, (size_t)b->N * b->N * b->s.R);
          memcpy(t.connect, o.connect, (size_t)b->N * b->N * b->s.P * b->C);
        }
      }
      cyber_reset(b, e, &o);
      e->att_return = 0.0;
    }
  }
}

/* ---- batch API (host arrays laid out exactly like cbx_views) --------------------------------- */
static void* zalloc(size_t n) { void* p = calloc(n ? n : 1, 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }

orc_batch* orc_create(const uint32_t* blob, size_t nwords, const cbx_config* cfg, int64_t n_envs) {
  orc_batch* b = (orc_batch*)zalloc(sizeof(*b));
  b->blob = (uint32_t*)zalloc(nwords * 4);
  memcpy(b->blob, blob, nwords * 4);
  if (scn_parse(&b->s, b->blob, nwords)) { free(b->blob); free(b); return NULL; }
  b->cfg = *cfg;
  b->n = n_envs;
  const scn_t* s = &b->s;
  b->N = cfg->maximum_node_count;
  b->C = cfg->maximum_total_credentials;
  b->LEAK = cfg->maximum_discoverable_credentials_per_action;
  b->OW = (b->N + 31) / 32;
  if (s->n > b->N || s->ntriples > b->C) { free(b->blob); free(b); return NULL; }
  /* MultiDiscrete layout [3, slice(kind_of_index[0]).., ...]: widths local 2, remote 3, connect 4 */
  int col = 1;
  for (int k = 0; k < 3; ++k) {
    int kind = cfg->kind_of_index[k];
    b->slice_of_kind[kind] = col;
    col += kind == CBX_KIND_LOCAL ? 2 : kind == CBX_KIND_REMOTE ? 3 : 4;
  }
  const int64_t n = n_envs;
  const int dense = cfg->mask_mode == CBX_MASK_DENSE;
  cbx_views* v = &b->v;
  v->n_envs = n; v->N = b->N; v->L = s->L; v->R = s->R; v->P = s->P; v->C = b->C; v->LEAK = b->LEAK;
  v->n_props = s->nprops; v->n_nodes = s->n; v->n_services = s->nservices; v->owned_words = b->OW;
  v->scalars = zalloc(n * 8 * 4);
  v->leaked_credentials = zalloc(n * 4 * b->LEAK * 4);
  v->credential_cache_matrix = zalloc(n * 2 * b->C * 4);
  v->discovered_nodes_properties = zalloc(n * b->N * s->nprops * 4);
  v->nodes_privilegelevel = zalloc(n * b->N * 4);
  if (dense) {
    v->local_vulnerability = zalloc(n * b->N * s->L);
    v->remote_vulnerability = zalloc(n * b->N * b->N * s->R);
    v->connect = zalloc(n * b->N * b->N * s->P * b->C);
  }
  v->owned_bits = zalloc(n * b->OW * 4);
  if (cfg->def_enabled) {
    v->def_infected_nodes = zalloc(n * s->n);
    v->def_incoming_firewall = zalloc(n * 6 * s->n);
    v->def_outgoing_firewall = zalloc(n * 6 * s->n);
    v->def_services_status = zalloc(n * s->nservices);
  }
  v->att_reward = zalloc(n * 4); v->def_reward = zalloc(n * 4);
  v->att_terminated = zalloc(n); v->att_truncated = zalloc(n); v->def_terminated = zalloc(n); v->def_truncated = zalloc(n);
  v->att_info = zalloc(n * 8 * 4);
  v->network_availability = zalloc(n * 8);
  v->episode_stats = b->stats;
  if (cfg->emit_terminal_obs) {
    v->term_scalars = zalloc(n * 8 * 4);
    v->term_leaked_credentials = zalloc(n * 4 * b->LEAK * 4);
    v->term_credential_cache_matrix = zalloc(n * 2 * b->C * 4);
    v->term_discovered_nodes_properties = zalloc(n * b->N * s->nprops * 4);
    v->term_nodes_privilegelevel = zalloc(n * b->N * 4);
    if (dense) {
      v->term_local_vulnerability = zalloc(n * b->N * s->L);
      v->term_remote_vulnerability = zalloc(n * b->N * b->N * s->R);
      v->term_connect = zalloc(n * b->N * b->N * s->P * b->C);
    }
    if (cfg->def_enabled) v->term_def_infected_nodes = zalloc(n * s->n);
  }
  b->envs = (oenv_t*)zalloc(sizeof(oenv_t) * n);
  for (int64_t i = 0; i < n; ++i) {
    oenv_t* e = &b->envs[i];
    e->nodes = zalloc(sizeof(onode_t) * s->n);
    e->actuator_order = zalloc(sizeof(int) * s->n);
    e->discovered = zalloc(sizeof(int) * s->n);
    e->cache = zalloc(sizeof(int) * (s->ntriples + 1));
    e->gathered = zalloc(s->nsecrets);
    e->sh_running = zalloc(s->n);
    e->sh_progress = zalloc(sizeof(int) * s->n);
    for (int k = 0; k < s->n; ++k) { e->sh_running[k] = 1; e->sh_progress[k] = -1; }
    e->sh_availability = 1.0;
    e->prev_availability = 1.0;
    e->availability = 1.0;
  }
  return b;
}

/* Firewall extension tables (marlon_b200/scenario.py FWX layout): call right after orc_create, before the first reset.  Turns
 * the live defender binding on when the config asks for it. */
int orc_set_firewall_tables(orc_batch* b, const uint32_t* words, size_t nwords) {
  if (!b || !words || nwords < CBX_FX_WORDS || words[CBX_FX_MAGIC] != CBX_FWX_MAGIC) return -1;
  b->fwx = (uint32_t*)zalloc(nwords * 4);
  memcpy(b->fwx, words, nwords * 4);
  b->n_fw_groups = (int)words[CBX_FX_N_GROUPS];
  b->live = b->cfg.mode == CBX_MODE_MARLON && b->cfg.def_enabled && b->cfg.def_binding == CBX_DEF_BINDING_LIVE;
  for (int64_t i = 0; i < b->n; ++i) {
    free(b->envs[i].fw);
    b->envs[i].fw = (struct fwlist*)zalloc(sizeof(struct fwlist) * (size_t)b->n_fw_groups);
    if (b->live) fw_lists_init(b, &b->envs[i]);
  }
  return 0;
}

void orc_destroy(orc_batch* b) {
  if (!b) return;
  for (int64_t i = 0; i < b->n; ++i) {
    oenv_t* e = &b->envs[i];
    free(e->nodes); free(e->actuator_order); free(e->discovered); free(e->cache); free(e->gathered);
    free(e->sh_running); free(e->sh_progress); free(e->fw);
  }
  free(b->fwx);
  cbx_views* v = &b->v;
  void* ptrs[] = {v->scalars, v->leaked_credentials, v->credential_cache_matrix, v->discovered_nodes_properties,
                  v->nodes_privilegelevel, v->local_vulnerability, v->remote_vulnerability, v->connect, v->owned_bits,
                  v->def_infected_nodes, v->def_incoming_firewall, v->def_outgoing_firewall, v->def_services_status,
                  v->att_reward, v->def_reward, v->att_terminated, v->att_truncated, v->def_terminated, v->def_truncated,
                  v->att_info, v->network_availability, v->term_scalars, v->term_leaked_credentials,
                  v->term_credential_cache_matrix, v->term_discovered_nodes_properties, v->term_nodes_privilegelevel,
                  v->term_local_vulnerability, v->term_remote_vulnerability, v->term_connect, v->term_def_infected_nodes};
  for (size_t k = 0; k < sizeof(ptrs) / sizeof(ptrs[0]); ++k) free(ptrs[k]);
  free(b->envs); free(b->blob); free(b);
}

void orc_views(orc_batch* b, cbx_views* out) { *out = b->v; }

void orc_reset_ex(orc_batch* b, const uint8_t* mask, int who) {
  for (int64_t i = 0; i < b->n; ++i) {
    if (mask && !mask[i]) continue;
    oenv_t* e = &b->envs[i];
    obs_t o = obs_main(b, i);
    int att = 1;
    if (b->cfg.mode == CBX_MODE_MARLON) {
      att = who & CBX_WHO_ATTACKER;
      if (att) attacker_reset(b, e, &o);
      if (b->cfg.def_enabled && (who & CBX_WHO_DEFENDER)) {
        defender_reset(b, e, i);
        b->v.def_reward[i] = 0; b->v.def_terminated[i] = b->v.def_truncated[i] = 0;
      }
    } else {
      cyber_reset(b, e, &o);
      e->att_return = 0.0;
    }
    if (att) {
      b->v.att_reward[i] = 0; b->v.att_terminated[i] = b->v.att_truncated[i] = 0;
      memset(b->v.att_info + i * 8, 0, 32);
    }
    if (b->cfg.mode != CBX_MODE_MARLON || (who & CBX_WHO_DEFENDER)) { b->v.def_reward[i] = 0; b->v.def_terminated[i] = b->v.def_truncated[i] = 0; }
    b->v.network_availability[i] = e->availability;
  }
}

/* EnvironmentEventSource.notify_reset delivered from outside (environment_event_source.py:30-38) */
void orc_notify_reset(orc_batch* b, const uint8_t* mask, int who, double last_reward) {
  for (int64_t i = 0; i < b->n; ++i) {
    if (mask && !mask[i]) continue;
    oenv_t* e = &b->envs[i];
    if (who & CBX_WHO_ATTACKER) e->att_reset_request = 1;
    if (who & CBX_WHO_DEFENDER) { e->def_reset_request = 1; e->last_attacker_reward = (double)(float)last_reward; }
  }
}

void orc_reset(orc_batch* b, const uint8_t* mask) { orc_reset_ex(b, mask, CBX_WHO_ATTACKER | CBX_WHO_DEFENDER); }

void orc_step_ex(orc_batch* b, const int32_t* aa, const int32_t* da, const double* scan_u, const double* detect_u, int who) {
  if (b->cfg.mode == CBX_MODE_MARLON) {
    for (int64_t i = 0; i < b->n; ++i)
      marlon_pair_step(b, i, aa ? aa + i * 10 : NULL, da ? da + i * 12 : NULL, scan_u, detect_u, who);
  } else {
    for (int64_t i = 0; i < b->n; ++i) cyber_only_step(b, i, aa + i * 5, scan_u, detect_u);
  }
}

void orc_step(orc_batch* b, const int32_t* aa, const int32_t* da, const double* scan_u, const double* detect_u) {
  orc_step_ex(b, aa, da, scan_u, detect_u, CBX_WHO_ATTACKER | CBX_WHO_DEFENDER);
}

void orc_stats_reset(orc_batch* b) { memset(b->stats, 0, sizeof(b->stats)); }

/* ---- CyberBattleEnv.sample_valid_action (ENV:959-1047), restated over a counter-based stream --------------------------
 * The reference draws proposals until the action mask admits one (ENV:1041-1047): kind uniform over [0, 1, 2] -- without 2
 * while the credential cache is empty (ENV:972-976) --, then, quirk B.9, kind 1 builds a LOCAL action (owned source, any local
 * vulnerability id) and kind 0 a REMOTE one (owned source, discovered target, any remote id); kind 2 a connect (owned source,
 * discovered target, any port, a cached credential; ENV:935-957).  "Owned" for the proposal is privilege >= LocalUser
 * (ENV:832-838); the mask wants agent_installed and, for a local action, the vulnerability on that node (ENV:643-677).  The whole
 * proposal -- kind included -- is redrawn after a rejection, so local actions are rarer than 1 / kinds.
 * Draws: Philox4x32-10, counter (env, step, 0x5A170000 + attempt) (+ 0x5A180000 + attempt for the credential), key = seed; an
 * integer in [0, n) is the high word of r * n.  The CUDA sampler (cbx_sample_kernel) does the same arithmetic: equal outputs.
 * The defender's action is uniform over its MultiDiscrete space (words of counter 0x5A190000, stepped by an LCG). */
static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

void orc_sample_actions(orc_batch* b, int32_t* att, int32_t* def, uint64_t seed, uint32_t step) {
  const scn_t* s = &b->s;
  const int marlon = b->cfg.mode == CBX_MODE_MARLON;
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int64_t i = 0; i < b->n; ++i) {
    const oenv_t* e = &b->envs[i];
    int owned[256], n_owned = 0;
    for (int k = 0; k < e->n_discovered; ++k)
      if (e->nodes[e->discovered[k]].privilege_level >= 1) owned[n_owned++] = k;
    const int nd = e->n_discovered, nc = e->n_cache;
    int kind = CBX_KIND_REMOTE, a[4] = {0, 0, 0, 0};
    for (uint32_t it = 0; it < 64; ++it) {
      uint32_t ctr[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), step, 0x5A170000u + it}, r[4];
      philox4x32_10(ctr, key, r);
      const uint32_t k = mulhi32(r[0], nc > 0 ? 3u : 2u);
      const int src = n_owned ? owned[mulhi32(r[1], (uint32_t)n_owned)] : 0;
      const int node = nd ? e->discovered[src] : 0;
      int valid = n_owned && e->nodes[node].agent_installed;
      a[0] = src; a[1] = a[2] = a[3] = 0;
      if (k == 1u) {
        kind = CBX_KIND_LOCAL;
        a[1] = (int)mulhi32(r[2], (uint32_t)s->L);
        valid = valid && (scn_vuln(s, node, a[1])[CBX_V_FLAGS] & 1u);
      } else if (k == 0u) {
        kind = CBX_KIND_REMOTE;
        a[1] = (int)mulhi32(r[2], (uint32_t)(nd > 0 ? nd : 1));
        a[2] = (int)mulhi32(r[3], (uint32_t)s->R);
      } else {
        uint32_t ctr2[4] = {ctr[0], ctr[1], step, 0x5A180000u + it}, q[4];
        philox4x32_10(ctr2, key, q);
        kind = CBX_KIND_CONNECT;
        a[1] = (int)mulhi32(r[2], (uint32_t)(nd > 0 ? nd : 1));
        a[2] = (int)mulhi32(r[3], (uint32_t)s->P);
        a[3] = (int)mulhi32(q[0], (uint32_t)nc);
      }
      if (valid) break;
    }
    if (marlon) {
      int32_t* o = att + i * 10;
      for (int k = 0; k < 10; ++k) o[k] = 0;
      for (int k = 0; k < 3; ++k)
        if (b->cfg.kind_of_index[k] == kind) o[0] = k;
      const int width = kind == CBX_KIND_LOCAL ? 2 : kind == CBX_KIND_REMOTE ? 3 : 4;
      for (int k = 0; k < width; ++k) o[b->slice_of_kind[kind] + k] = a[k];
      if (def) {
        uint32_t ctr3[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), step, 0x5A190000u}, w[4];
        philox4x32_10(ctr3, key, w);
        int32_t* d = def + i * 12;
        const uint32_t n = (uint32_t)s->n;
        uint32_t x = w[2], y = w[3];
        d[0] = (int32_t)mulhi32(w[1], 5u);
        d[1] = (int32_t)mulhi32(x, n); x = x * 1664525u + 1013904223u;
        d[2] = (int32_t)mulhi32(x, n); x = x * 1664525u + 1013904223u;
        d[3] = (int32_t)mulhi32(x, 6u); x = x * 1664525u + 1013904223u;
        d[4] = (int32_t)mulhi32(x, 2u); x = x * 1664525u + 1013904223u;
        d[5] = (int32_t)mulhi32(x, n);
        d[6] = (int32_t)mulhi32(y, 6u); y = y * 1664525u + 1013904223u;
        d[7] = (int32_t)mulhi32(y, 2u); y = y * 1664525u + 1013904223u;
        d[8] = (int32_t)mulhi32(y, n); y = y * 1664525u + 1013904223u;
        d[9] = (int32_t)mulhi32(y, 3u); y = y * 1664525u + 1013904223u;
        d[10] = (int32_t)mulhi32(y, n); y = y * 1664525u + 1013904223u;
        d[11] = (int32_t)mulhi32(y, 3u);
      }
    } else {
      int32_t* o = att + i * 5;
      o[0] = kind; o[1] = a[0]; o[2] = a[1]; o[3] = a[2]; o[4] = a[3];
    }
  }
}

int64_t orc_export_words(const orc_batch* b) {
  const scn_t* s = &b->s;
  return CBX_X_HEADER_WORDS + 10 * (int64_t)s->n + b->C + s->Ws;
}

void orc_export_state(orc_batch* b, int64_t begin, int64_t end, int32_t* out) {
  const scn_t* s = &b->s;
  const int n = s->n;
  const int64_t W = orc_export_words(b);
  for (int64_t i = begin; i < end; ++i) {
    const oenv_t* e = &b->envs[i];
    int32_t* x = out + (i - begin) * W;
    memset(x, 0, W * 4);
    x[CBX_X_STEPCOUNT] = e->stepcount; x[CBX_X_DONE] = e->done;
    x[CBX_X_N_DISCOVERED] = e->n_discovered; x[CBX_X_N_CACHED] = e->n_cache;
    x[CBX_X_ATT_TIMESTEPS] = e->att_timesteps; x[CBX_X_DEF_TIMESTEPS] = e->def_timesteps;
    x[CBX_X_ATT_RESET_REQUEST] = e->att_reset_request;
    x[CBX_X_DEF_RESET_REQUEST] = b->cfg.def_enabled ? e->def_reset_request : 0; /* no defender wrapper, no flag */
    x[CBX_X_HAS_BREACHED_SLA] = e->has_breached;
    x[CBX_X_ATT_VALID] = e->att_valid; x[CBX_X_ATT_INVALID] = e->att_invalid;
    x[CBX_X_DEF_VALID] = e->def_valid; x[CBX_X_DEF_INVALID] = e->def_invalid;
    int live = 0, sh = 0;
    for (int k = 0; k < n; ++k) { live += !e->nodes[k].running; sh += !e->sh_running[k]; }
    x[CBX_X_LIVE_IMAGING_COUNT] = live;
    x[CBX_X_SHADOW_IMAGING_COUNT] = (int32_t)llround((1.0 - e->sh_availability) * n);
    x[CBX_X_PREV_SHADOW_IMAGING_COUNT] = (int32_t)llround((1.0 - e->prev_availability) * n);
    (void)sh;
    int32_t* p = x + CBX_X_HEADER_WORDS;
    for (int k = 0; k < n; ++k) p[k] = k < e->n_discovered ? e->discovered[k] : -1;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = e->nodes[k].agent_installed;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = e->nodes[k].privilege_level;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = e->nodes[k].progress < 0 ? 0 : e->nodes[k].progress + 1;
    p += n;
    for (int k = 0; k < n; ++k) {
      const int pr = b->live ? e->nodes[k].progress : e->sh_progress[k]; /* the actuator the defender's wrapper is bound to */
      p[k] = pr < 0 ? 0 : pr + 1;
    }
    p += n;
    for (int k = 0; k < n; ++k) p[k] = e->nodes[k].tracked && e->nodes[k].last_owned_at != 0;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = (int32_t)(uint32_t)e->nodes[k].discovered_properties;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = (int32_t)(uint32_t)(e->nodes[k].discovered_properties >> 32);
    p += n;
    for (int k = 0; k < n; ++k) {
      uint32_t bits = 0;
      const onode_t* nd = &e->nodes[k];
      for (int v = 0; v < s->L + s->R && v < 16; ++v) {
        if (nd->last_attack[v]) bits |= 1u << (2 * v);
        if (nd->last_attack[v] && (nd->last_reimaging == 0 || nd->last_attack[v] >= nd->last_reimaging)) bits |= 2u << (2 * v);
      }
      p[k] = (int32_t)bits;
    }
    p += n;
    for (int k = 0; k < n; ++k) p[k] = e->nodes[k].tags;
    p += n;
    for (int k = 0; k < b->C; ++k) p[k] = k < e->n_cache ? e->cache[k] : -1;
    p += b->C;
    for (int k = 0; k < s->nsecrets; ++k)
      if (e->gathered[k]) p[k / 32] |= (int32_t)(1u << (k % 32));
  }
}

/* L1 entry points for the commandcontrol KAT (commandcontrol_test.py:14-73): raw AgentActions calls on node indices */
double orc_l1_local(orc_batch* b, int64_t i, int node, int v, int* outcome) {
  result_t r = exploit_local(b, &b->envs[i], node, v);
  *outcome = r.outcome;
  return r.reward;
}
double orc_l1_remote(orc_batch* b, int64_t i, int src, int tgt, int v, int* outcome) {
  result_t r = exploit_remote(b, &b->envs[i], src, tgt, v);
  *outcome = r.outcome;
  return r.reward;
}
double orc_l1_connect(orc_batch* b, int64_t i, int src, int tgt, int port, int secret, int* outcome) {
  result_t r = connect_to_remote(b, &b->envs[i], src, tgt, port, secret);
  *outcome = r.outcome;
  return r.reward;
}
