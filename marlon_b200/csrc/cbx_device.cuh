// cbx_device.cuh -- the game rules on the bit-packed state tile: one thread plays one env.
//
// Everything here runs on shared memory: `st` is this env's column of the S x 32 state tile, `tb` the scenario
// tables (both staged by the caller).  Reference semantics followed (paths relative to the reference root):
//   ACT = src/CyberBattleSim/cyberbattle/simulation/actions.py     ENV = .../_env/cyberbattle_env.py
//   DEF = .../_env/defender.py   ATT = marlon/baseline_models/env_wrappers/attack_wrapper.py
//   DWR = marlon/baseline_models/env_wrappers/defend_wrapper.py    LDF = marlon/defender_agents/defender.py
// The reference compares wall-clock stamps (ACT:402,521); here that is logical state: `installed` is
// "owned since the last re-image", two bits per (node, vulnerability) say "ever attacked" / "attacked since the
// last re-image" (SURVEY.md A.3).
#ifndef CBX_DEVICE_CUH_
#define CBX_DEVICE_CUH_

#include "cbx_layout.h"

namespace cbx {

// hdr flag bits
constexpr uint32_t HDR_DONE = 1u << 24, HDR_ATT_RR = 1u << 25, HDR_DEF_RR = 1u << 26, HDR_BREACHED = 1u << 27,
                   HDR_HAS_CYBER = 1u << 28, HDR_HAS_REWARD = 1u << 29;

// staging written by the game-logic thread of each env, word-major like the state tile
enum { STG_SCALARS = 0, STG_OBS_KIND = 8, STG_ATT_DONE = 9, STG_DEF_DONE = 10, STG_DEF_TERM_INST = 11 /* Wn words */ };
enum { OBS_NORMAL = 0, OBS_BLANK = 1, OBS_KEEP = 2 };

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

struct Result {
  double reward;
  int outcome;  // CBX_RES_*
  int level;
  int error;    // CBX_E_*
};

struct StepOut {
  double reward, raw;
  int terminated, outcome, error, oob;
};

// Independent state loads issued together by the loops that walk a per-node array: 4 where the state is read in place from
// global memory (cbx_wide.cu defines CBX_STATE_IN_PLACE), 1 where the state tile is in shared memory.
#ifdef CBX_STATE_IN_PLACE
constexpr int kLoadBatch = 4;
#else
constexpr int kLoadBatch = 1;
#endif

// One leaked-credential slot of the staging area (ENV:890-907: the observation row [1, cache index, target discovery index,
// port]) in one word: bit 0 used | cache index << 1 (15 bits) | target discovery index << 16 (8) | port << 24 (8);
// cbx_batch_create rejects bounds that do not fit.  leak_field(word, k) is element k of the row.
__device__ __forceinline__ uint32_t leak_pack(uint32_t cache_index, uint32_t target, uint32_t port) {
  return 1u | (cache_index << 1) | (target << 16) | (port << 24);
}
__device__ __forceinline__ uint32_t leak_field(uint32_t word, int k) {
  return k == 0 ? (word & 1u) : k == 1 ? ((word >> 1) & 0x7FFFu) : k == 2 ? ((word >> 16) & 0xFFu) : (word >> 24);
}

struct Ctx {
  uint32_t* st;          // this env's column of the state tile: word w at st[w * CBX_TILE]
  uint32_t* sg;          // this env's column of the staging area
  const uint32_t* tb;    // scenario tables in shared memory
  const cbx_layout* L;
  const cbx_config* cfg;
  int64_t env;           // global env index
  const uint32_t* fx = nullptr;  // firewall extension tables (cbx.h CBX_FX_*) under the LIVE defender binding, else nullptr

  __device__ __forceinline__ uint32_t& w(int off) const { return st[off * CBX_TILE]; }
  __device__ __forceinline__ uint32_t& g(int off) const { return sg[off * CBX_TILE]; }
  __device__ __forceinline__ bool bit(int off, int i) const { return (w(off + (i >> 5)) >> (i & 31)) & 1u; }
  __device__ __forceinline__ void setbit(int off, int i) const { w(off + (i >> 5)) |= 1u << (i & 31); }
  __device__ __forceinline__ void clrbit(int off, int i) const { w(off + (i >> 5)) &= ~(1u << (i & 31)); }
  __device__ __forceinline__ uint32_t byte(int off, int i) const { return (w(off + (i >> 2)) >> ((i & 3) * 8)) & 0xFFu; }
  __device__ __forceinline__ void setbyte(int off, int i, uint32_t v) const {
    uint32_t& x = w(off + (i >> 2));
    int sh = (i & 3) * 8;
    x = (x & ~(0xFFu << sh)) | (v << sh);
  }
  __device__ __forceinline__ uint32_t half(int off, int i) const { return (w(off + (i >> 1)) >> ((i & 1) * 16)) & 0xFFFFu; }
  __device__ __forceinline__ void sethalf(int off, int i, uint32_t v) const {
    uint32_t& x = w(off + (i >> 1));
    int sh = (i & 1) * 16;
    x = (x & ~(0xFFFFu << sh)) | (v << sh);
  }
  __device__ __forceinline__ int priv(int node) const { return (w(L->o_priv + (node >> 4)) >> ((node & 15) * 2)) & 3; }
  __device__ __forceinline__ void setpriv(int node, int p) const {
    uint32_t& x = w(L->o_priv + (node >> 4));
    int sh = (node & 15) * 2;
    x = (x & ~(3u << sh)) | ((uint32_t)p << sh);
  }
  __device__ __forceinline__ int tags(int node) const {
    return L->o_tags < 0 ? 0 : (int)((w(L->o_tags + (node >> 3)) >> ((node & 7) * 4)) & 15u);
  }
  __device__ __forceinline__ void settag(int node, int level) const { w(L->o_tags + (node >> 3)) |= 1u << ((node & 7) * 4 + level); }

  __device__ __forceinline__ int nd() const { return w(L->o_hdr) & 0xFF; }
  __device__ __forceinline__ int nc() const { return (w(L->o_hdr) >> 8) & 0xFFFF; }
  __device__ __forceinline__ bool flag(uint32_t f) const { return w(L->o_hdr) & f; }
  __device__ __forceinline__ void setflag(uint32_t f, bool on) const {
    uint32_t& x = w(L->o_hdr);
    x = on ? (x | f) : (x & ~f);
  }
  __device__ __forceinline__ float f32(int off) const { return __uint_as_float(w(off)); }
  __device__ __forceinline__ void setf32(int off, float v) const { w(off) = __float_as_uint(v); }

  // ---- scenario tables --------------------------------------------------------------------------------
  // Real dimensions of THIS env's scenario (the staged tables).  The layout (L->n, L->nsecrets, ...) is the maximum over the
  // scenarios of a multi-scenario batch (cbx_batch_create_multi); the game's arithmetic uses the scenario's own counts.
  __device__ __forceinline__ int n_nodes() const { return (int)tb[CBX_H_N_NODES]; }
  __device__ __forceinline__ int n_secrets() const { return (int)tb[CBX_H_N_SECRETS]; }
  __device__ __forceinline__ const uint32_t* node_rec(int node) const { return tb + tb[CBX_H_OFF_NODE] + node * CBX_NODE_WORDS; }
  __device__ __forceinline__ const uint32_t* vuln_rec(int node, int v) const {
    return tb + tb[CBX_H_OFF_VULN] + (node * (L->L + L->R) + v) * CBX_VULN_WORDS;
  }
  __device__ __forceinline__ const uint32_t* triple(int t) const { return tb + tb[CBX_H_OFF_TRIPLE] + 3 * t; }
  __device__ __forceinline__ const uint32_t* payload(const uint32_t* rec) const { return tb + tb[CBX_H_OFF_PAYLOAD] + rec[CBX_V_PAYLOAD_OFF]; }

  // ---- firewall rule lists as per-env state (live defender binding) ------------------------------
  // One list object may serve several (node, direction) pairs: the unit of state is the list (alias group), two words each:
  // port names that have a rule, names whose FIRST rule allows (ACT:504-515 asks nothing else of a list).
  __device__ __forceinline__ int fw_group(int node, bool incoming) const {
    const uint32_t g = fx[CBX_FX_WORDS + tb[CBX_H_N_PORTS] + node];
    return (int)(incoming ? (g & 0xFFFFu) : (g >> 16));
  }
  __device__ __forceinline__ bool fw_has_rule(int node, bool incoming, int name) const {
    return (w(L->o_fw + 2 * fw_group(node, incoming)) >> name) & 1u;
  }
  // ACT:504-515 __is_passing_firewall_rules for an attacker port: static pass bits, or the env's own lists when the defender
  // edits them
  __device__ __forceinline__ bool fw_passes(int node, bool incoming, int port) const {
    if (!fx) return (node_rec(node)[incoming ? CBX_N_FW_IN : CBX_N_FW_OUT] >> port) & 1u;
    const int g = fw_group(node, incoming), name = (int)fx[CBX_FX_WORDS + port];
    return ((w(L->o_fw + 2 * g) & w(L->o_fw + 2 * g + 1)) >> name) & 1u;  // a rule exists and the first one allows
  }
  // LDF:50-58 block_traffic: every rule of that name leaves the list the (node, direction) pair points at
  __device__ __forceinline__ void fw_block(int node, bool incoming, int name) const {
    const int g = fw_group(node, incoming);
    w(L->o_fw + 2 * g) &= ~(1u << name);
    w(L->o_fw + 2 * g + 1) &= ~(1u << name);
  }
  // LDF:60-69 allow_traffic: when the selected list has no rule of that name, an ALLOW rule is appended -- to the node's
  // INCOMING list whatever the direction (both arms of the conditional expression append to incoming); appended behind an
  // existing rule of that name it changes neither "has a rule" nor "first rule allows"
  __device__ __forceinline__ void fw_allow(int node, bool incoming, int name) const {
    if (fw_has_rule(node, incoming, name)) return;
    const int gi = fw_group(node, true);
    if ((w(L->o_fw + 2 * gi) >> name) & 1u) return;
    w(L->o_fw + 2 * gi) |= 1u << name;
    w(L->o_fw + 2 * gi + 1) |= 1u << name;
  }

  // ---- AgentActions -------------------------------------------------------------------------------------
  __device__ int discover(int node) const {  // ACT:227-232 + ENV:866-869: append to the discovery order
    if (byte(L->o_disc_idx, node) != 0xFFu) return 0;
    int k = nd();
    setbyte(L->o_disc_order, k, (uint32_t)node);
    setbyte(L->o_disc_idx, node, (uint32_t)k);
    w(L->o_hdr) = (w(L->o_hdr) & ~0xFFu) | (uint32_t)(k + 1);
    return 1;
  }
  __device__ int mark_props(int node, uint32_t lo, uint32_t hi) const {  // ACT:234-249
    uint32_t& a = w(L->o_props + node * L->PW);
    int added = __popc(lo & ~a);
    a |= lo;
    if (L->PW > 1) {
      uint32_t& b = w(L->o_props + node * L->PW + 1);
      added += __popc(hi & ~b);
      b |= hi;
    }
    return added;
  }
  // ACT:251-275; returns whether the node was already owned, *ever = owned at any time before (last_owned_at != None)
  __device__ bool mark_owned(int node, int privilege, bool* ever) const {
    bool owned = bit(L->o_installed, node);
    *ever = bit(L->o_everowned, node);
    if (!owned) {
      setbit(L->o_installed, node);
      if (privilege > priv(node)) setpriv(node, privilege);
      const uint32_t* r = node_rec(node);
      mark_props(node, r[CBX_N_PROPS_LO], r[CBX_N_PROPS_HI]);
      setbit(L->o_everowned, node);
    }
    return owned;
  }

  // ACT:325-423 + the observation side effects of ENV:863-907 (discovery order, credential cache, leaked slots)
  __device__ Result process_outcome(int node, int v, double failed_penalty) const {
    Result r = {0.0, CBX_RES_NONE, 0, 0};
    if (bit(L->o_notrunning, node)) return r;  // MACHINE_NOT_RUNNING = 0
    const uint32_t* rec = vuln_rec(node, v);
    uint32_t fl = rec[CBX_V_FLAGS];
    if (!(fl & 1u)) { r.reward = -5.0; return r; }  // SUPSPICIOUSNESS
    int kind = (fl >> 1) & 7, level = (fl >> 4) & 3;
    if (!((fl >> (8 + tags(node))) & 1u)) { r.reward = failed_penalty; r.outcome = CBX_RES_EXPLOIT_FAILED; return r; }
    double reward = 0.0;
    const int value = (int)node_rec(node)[CBX_N_VALUE];
    switch (kind) {
      case CBX_OUT_ESCALATION: {
        r.outcome = CBX_RES_ESCALATION; r.level = level;
        if ((tags(node) >> level) & 1) { r.reward = -1.0; return r; }
        bool ever;
        mark_owned(node, level, &ever);
        if (!ever) reward += (double)value;
        settag(node, level);
      } break;
      case CBX_OUT_LATERAL_MOVE: {
        r.outcome = CBX_RES_LATERAL_MOVE;
        bool ever;
        mark_owned(node, 1, &ever);
        if (!ever) reward += (double)value;
      } break;
      case CBX_OUT_PROBE_SUCCEEDED: {
        r.outcome = CBX_RES_PROBE_SUCCEEDED;
        const uint32_t* pl = payload(rec);
        reward += 2.0 * mark_props(node, pl[0], pl[1]);
      } break;
      case CBX_OUT_LEAKED_CREDENTIALS: r.outcome = CBX_RES_LEAKED_CREDENTIALS; break;
      case CBX_OUT_LEAKED_NODES: r.outcome = CBX_RES_LEAKED_NODES; break;
      case CBX_OUT_CUSTOMER_DATA: r.outcome = CBX_RES_CUSTOMER_DATA; break;
      case CBX_OUT_PROBE_FAILED: r.outcome = CBX_RES_PROBE_FAILED; break;
      default: r.outcome = CBX_RES_EXPLOIT_FAILED; break;
    }
    {  // ACT:396-407: +7 first time, -1 repeat since the last re-image, 0 when the previous attempt predates it
      uint32_t& a = w(L->o_attacked + node * L->AW + (v >> 4));
      int sh = (v & 15) * 2;
      if ((a >> sh) & 1u) { if ((a >> sh) & 2u) reward += -1.0; }
      else reward += 7.0;
      a |= 3u << sh;
    }
    int new_nodes = 0, new_creds = 0;
    if (kind == CBX_OUT_LEAKED_CREDENTIALS) {
      const uint32_t* pl = payload(rec);
      const int cnt = (int)rec[CBX_V_PAYLOAD_CNT];
      int slot = 0;
      for (int k = 0; k < cnt; ++k) {
        int t = (int)pl[k];
        const uint32_t* tr = triple(t);
        new_nodes += discover((int)tr[0]);
        if (!bit(L->o_gathered, (int)tr[2])) { setbit(L->o_gathered, (int)tr[2]); new_creds++; }
        if (!bit(L->o_cached, t)) {  // ENV:882-885
          setbit(L->o_cached, t);
          int c = nc();
          sethalf(L->o_cache, c, (uint32_t)t);
          w(L->o_hdr) = (w(L->o_hdr) & ~0xFFFF00u) | ((uint32_t)(c + 1) << 8);
          if (slot < L->LEAKS) {    // ENV:890-907; a list never has more than LEAKS entries (target index is final: a node keeps its discovery index)
            g(L->g_leaked + slot) = leak_pack((uint32_t)c, byte(L->o_disc_idx, (int)tr[0]), tr[1]);
            slot++;
          }
        }
      }
    } else if (kind == CBX_OUT_LEAKED_NODES) {
      const uint32_t* pl = payload(rec);
      const int cnt = (int)rec[CBX_V_PAYLOAD_CNT];
      for (int k = 0; k < cnt; ++k) new_nodes += discover((int)pl[k]);
    }
    g(STG_SCALARS + 0) = (uint32_t)new_nodes;
    reward += new_nodes * 5.0 + new_creds * 3.0;
    reward -= (double)__uint_as_float(rec[CBX_V_COST]);
    r.reward = reward;
    return r;
  }

  __device__ Result invalid(int code) const {
    Result r = {-1.0, CBX_RES_NONE, 0, cfg->throws_on_invalid_actions ? code : 0};
    return r;
  }
  __device__ Result exploit_local(int node, int v) const {  // ACT:473-502
    if (!bit(L->o_installed, node)) return invalid(CBX_E_SOURCE_NOT_OWNED);
    return process_outcome(node, v, -20.0);
  }
  __device__ Result exploit_remote(int src, int tgt, int v) const {  // ACT:425-471
    if (!bit(L->o_installed, src)) return invalid(CBX_E_SOURCE_NOT_OWNED);
    if (byte(L->o_disc_idx, tgt) == 0xFFu) return invalid(CBX_E_TARGET_NOT_DISCOVERED);
    return process_outcome(tgt, L->L + v, -50.0);
  }
  __device__ Result connect(int src, int tgt, int port, int secret) const {  // ACT:524-606
    Result r = {0.0, CBX_RES_NONE, 0, 0};
    if (!bit(L->o_installed, src)) return invalid(CBX_E_SOURCE_NOT_OWNED);
    if (byte(L->o_disc_idx, tgt) == 0xFFu) return invalid(CBX_E_TARGET_NOT_DISCOVERED);
    if (!bit(L->o_gathered, secret)) return invalid(CBX_E_CREDENTIAL_NOT_GATHERED);
    const uint32_t* rs = node_rec(src);
    const uint32_t* rt = node_rec(tgt);
    (void)rs;
    if (!fw_passes(src, false, port)) { r.reward = -10.0; return r; }
    if (!fw_passes(tgt, true, port)) { r.reward = -10.0; return r; }
    if (!((rt[CBX_N_LISTEN] >> port) & 1u)) { r.reward = -10.0; return r; }
    if (bit(L->o_notrunning, tgt)) { r.reward = 0.0; return r; }
    const int Ws = (n_secrets() + 31) >> 5;
    const uint32_t* auth = tb + tb[CBX_H_OFF_AUTH] + (tgt * L->P + port) * Ws;
    if (!((auth[secret >> 5] >> (secret & 31)) & 1u)) { r.reward = -10.0; return r; }
    bool ever;
    bool already = mark_owned(tgt, 1, &ever);
    r.outcome = CBX_RES_LATERAL_MOVE;
    if (already) { r.reward = -1.0; return r; }
    r.reward = ever ? 0.0 : (double)(int)rt[CBX_N_VALUE];
    return r;
  }

  // ---- DefenderAgentActions --------------------------------------------------------------------------------
  __device__ void reimage_live(int node) const {  // ACT:700-712
    setbyte(L->o_cd_live, node, 16);
    clrbit(L->o_installed, node);
    setpriv(node, 0);
    setbit(L->o_notrunning, node);
    for (int k = 0; k < L->AW; ++k) w(L->o_attacked + node * L->AW + k) &= 0x55555555u;  // "since last re-image" bits
  }
  // ACT:714-746: tick the countdowns, return how many nodes are not running (availability = (n - k) / n)
  __device__ int tick(int o_cd, int o_notrun /* -1 for the shadow copy */) const {
    int down = 0;
    const int words = (L->n + 3) >> 2;
    // kLoadBatch words are loaded before the arithmetic: with the state in place (cbx_wide_kernel) every load is an L2 round
    // trip, and issued one per iteration they took 10 us per tick at 26 words (+5 % env-steps/s with four in flight); with
    // the state tile in shared memory the plain loop is the shorter code
    for (int q0 = 0; q0 < words; q0 += kLoadBatch) {
      uint32_t xs[kLoadBatch];
#pragma unroll
      for (int j = 0; j < kLoadBatch; ++j) xs[j] = q0 + j < words ? w(o_cd + q0 + j) : 0u;
#pragma unroll
      for (int j = 0; j < kLoadBatch; ++j) {
        uint32_t x = xs[j];
        if (!x) continue;
        const int q = q0 + j;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint32_t c = (x >> (8 * b)) & 0xFFu;
          if (c) {
            c -= 1;
            x = (x & ~(0xFFu << (8 * b))) | (c << (8 * b));
            if (c) down++;
            else if (o_notrun >= 0) clrbit(o_notrun, q * 4 + b);
          }
        }
        w(o_cd + q) = x;
      }
    }
    return down;
  }
  __device__ double availability(int down) const { const int n = n_nodes(); return (double)(n - down) / (double)n; }

  __device__ void scan_and_reimage(int stepcount, const double* scan_u, const double* detect_u) const {  // DEF:42-55
    if (stepcount % cfg->scan_frequency != 0) return;
    const int cap = cfg->scan_capacity;
    for (int k = 0; k < cap; ++k) {
      double us, ud;
      if (scan_u) {
        us = scan_u[env * cap + k];
        ud = detect_u[env * cap + k];
      } else {
        uint32_t o[4];
        const int64_t ge = env + cfg->env_index_base;
        philox4x32_10((uint32_t)ge, (uint32_t)(ge >> 32), (uint32_t)stepcount, (uint32_t)k, (uint32_t)cfg->seed,
                      (uint32_t)(cfg->seed >> 32), o);
        us = u53(o[0], o[1]);
        ud = u53(o[2], o[3]);
      }
      int node = (int)floor(us * (double)n_nodes());
      if (!bit(L->o_notrunning, node) && bit(L->o_installed, node)) {
        bool detected = ud <= cfg->scan_probability;
        if (detected && (node_rec(node)[CBX_N_FLAGS] & 1u)) reimage_live(node);
      }
    }
  }

  // ---- CyberBattleEnv -----------------------------------------------------------------------------------------
  __device__ int owned_count() const {  // ACT:317-319
    int c = 0;
    const int words = (L->n + 15) >> 4;
    for (int q = 0; q < words; ++q) {
      uint32_t x = w(L->o_priv + q);
      c += __popc((x | (x >> 1)) & 0x55555555u);
    }
    return c;
  }
  __device__ double live_availability() const { return availability((int)(w(L->o_avail) & 0xFFu)); }
  __device__ bool attacker_goal_reached() const {  // ENV:1080-1101
    if (!cfg->has_attacker_goal) return false;
    if ((double)f32(L->o_ep_sum) < cfg->goal_reward) return false;
    int oc = owned_count();
    if (oc < cfg->goal_own_atleast) return false;
    if ((double)oc / (double)n_nodes() < cfg->goal_own_atleast_percent) return false;
    if (cfg->builtin_defender != CBX_BUILTIN_NONE && live_availability() >= cfg->goal_low_availability) return false;
    return true;
  }
  __device__ bool constraints_broken() const {  // ENV:1103-1110
    return cfg->builtin_defender != CBX_BUILTIN_NONE && live_availability() < cfg->maintain_sla;
  }
  __device__ bool defender_goal_reached() const { return cfg->defender_goal_eviction && owned_count() == 0; }  // ENV:1112-1116

  // ENV:375-394 + 1187-1209: the initial state was computed on the host; wrapper / shadow words are kept
  __device__ void cyber_reset(const uint32_t* init) const {
    for (int k = L->o_cyber_begin; k < L->S; ++k) w(k) = init[k];
    uint32_t flags = w(L->o_hdr) & (HDR_ATT_RR | HDR_DEF_RR | HDR_BREACHED | HDR_HAS_CYBER | HDR_HAS_REWARD);
    w(L->o_hdr) = (init[L->o_hdr] & 0x00FFFFFFu) | flags;
    // fresh DefenderAgentActions: availability 1.0 -- of the live env (bits 0-7) and, under the live binding, of the
    // actuator the defender's wrapper reads (bits 8-15: it is the same object there)
    w(L->o_avail) &= fx ? ~0xFFFFu : ~0xFFu;
  }
  __device__ void snapshot_for_obs() const {
    const int pw = (L->n + 15) >> 4;
    for (int k0 = 0; k0 < L->Wn + pw; k0 += kLoadBatch) {  // loads kLoadBatch at a time, then the staging stores
      uint32_t v[kLoadBatch];
#pragma unroll
      for (int j = 0; j < kLoadBatch; ++j) {
        const int k = k0 + j;
        v[j] = k < L->Wn ? w(L->o_installed + k) : k < L->Wn + pw ? w(L->o_priv + k - L->Wn) : 0u;
      }
#pragma unroll
      for (int j = 0; j < kLoadBatch; ++j) {
        const int k = k0 + j;
        if (k < L->Wn) g(L->g_inst + k) = v[j];
        else if (k < L->Wn + pw) g(L->g_priv + k - L->Wn) = v[j];
      }
    }
  }
  __device__ void stage_reset_obs() const {  // blank observation + masks/properties of the fresh state (ENV:1197-1200)
    for (int k = 0; k < 8; ++k) g(STG_SCALARS + k) = 0;
    g(STG_SCALARS + 6) = (uint32_t)nd();
    for (int k = 0; k < L->LEAKS; ++k) g(L->g_leaked + k) = 0;
    g(STG_OBS_KIND) = OBS_NORMAL;
    snapshot_for_obs();
  }

  // ENV:707-751 + 584-601; returns true on OutOfBoundIndexError
  __device__ bool execute_action(int kind, const int32_t* a, Result* out) const {
    const int ndisc = nd();
    // coordinates outside the action space (vulnerability / port index beyond the Identifiers) cannot come out of the gym
    // spaces; the reference would raise IndexError.  Here they take the OutOfBoundIndexError path instead of indexing tables.
    if (kind == CBX_KIND_LOCAL ? (a[1] < 0 || a[1] >= L->L) : kind == CBX_KIND_REMOTE ? (a[2] < 0 || a[2] >= L->R) : (a[2] < 0 || a[2] >= L->P))
      return true;
    if (kind == CBX_KIND_LOCAL) {
      if (a[0] < 0 || a[0] >= ndisc) return true;
      *out = exploit_local((int)byte(L->o_disc_order, a[0]), a[1]);
    } else if (kind == CBX_KIND_REMOTE) {
      if (a[0] < 0 || a[0] >= ndisc) return true;
      if (a[1] < 0 || a[1] >= ndisc) return true;
      *out = exploit_remote((int)byte(L->o_disc_order, a[0]), (int)byte(L->o_disc_order, a[1]), a[2]);
    } else {
      if (a[3] < 0 || a[3] >= nc()) { Result r = {-1.0, CBX_RES_NONE, 0, 0}; *out = r; return false; }  // ENV:736-737
      if (a[0] < 0 || a[0] >= ndisc) return true;
      if (a[1] < 0 || a[1] >= ndisc) return true;
      int t = (int)half(L->o_cache, a[3]);
      *out = connect((int)byte(L->o_disc_order, a[0]), (int)byte(L->o_disc_order, a[1]), a[2], (int)triple(t)[2]);
    }
    return false;
  }

  // ENV:1145-1185
  __device__ StepOut cyber_step(int kind, const int32_t* a, const double* scan_u, const double* detect_u) const {
    StepOut so = {0.0, 0.0, 0, CBX_RES_NONE, 0, 0};
    if (flag(HDR_DONE)) { so.error = CBX_E_STEP_AFTER_DONE; so.terminated = 1; g(STG_OBS_KIND) = OBS_KEEP; return so; }
    w(L->o_stepcount) += 1;
    for (int k = 0; k < 8; ++k) g(STG_SCALARS + k) = 0;
    for (int k = 0; k < L->LEAKS; ++k) g(L->g_leaked + k) = 0;
    Result r;
    if (execute_action(kind, a, &r)) {  // blank observation, reward 0, the built-in defender does not move
      g(STG_SCALARS + 6) = (uint32_t)nd();
      g(STG_SCALARS + 7) = 1;
      g(STG_OBS_KIND) = OBS_BLANK;
      so.oob = 1;
      so.outcome = CBX_RES_OUT_OF_BOUND;
      return so;
    }
    if (r.error) { so.error = r.error; g(STG_OBS_KIND) = OBS_KEEP; return so; }
    g(STG_OBS_KIND) = OBS_NORMAL;
    if (r.outcome == CBX_RES_LATERAL_MOVE) g(STG_SCALARS + 1) = 1;
    else if (r.outcome == CBX_RES_CUSTOMER_DATA) g(STG_SCALARS + 2) = 1;
    else if (r.outcome == CBX_RES_PROBE_SUCCEEDED) g(STG_SCALARS + 3) = 2;
    else if (r.outcome == CBX_RES_PROBE_FAILED) g(STG_SCALARS + 3) = 1;
    else if (r.outcome == CBX_RES_ESCALATION) g(STG_SCALARS + 4) = (uint32_t)r.level;
    g(STG_SCALARS + 5) = (uint32_t)nc();
    g(STG_SCALARS + 6) = (uint32_t)nd();
    so.raw = r.reward;
    so.outcome = r.outcome;
    // The observation (masks, privileges, properties) is encoded from the state as it is NOW: before the built-in
    // defender moves.  The encoder reads these snapshots instead of the post-defender state.
    snapshot_for_obs();
    double reward = r.reward;
    if (cfg->builtin_defender != CBX_BUILTIN_NONE) {  // ENV:1156-1158
      int down = tick(L->o_cd_live, L->o_notrunning);
      w(L->o_avail) = (w(L->o_avail) & ~0xFFu) | (uint32_t)down;
      scan_and_reimage((int)w(L->o_stepcount), scan_u, detect_u);
    }
    if (attacker_goal_reached() || constraints_broken()) { setflag(HDR_DONE, true); reward = cfg->winning_reward; }
    else if (defender_goal_reached()) { setflag(HDR_DONE, true); reward = cfg->losing_reward; }
    else reward = reward > 0.0 ? reward : 0.0;
    setf32(L->o_ep_sum, f32(L->o_ep_sum) + (float)reward);
    so.reward = reward;
    so.terminated = flag(HDR_DONE) ? 1 : 0;
    return so;
  }

  // ---- MARLon wrappers ---------------------------------------------------------------------------------------------
  __device__ void attacker_reset(const uint32_t* init) const {  // ATT:400-468
    if (!flag(HDR_ATT_RR)) {  // notify_reset(last_reward) -> both observers
      float last = flag(HDR_HAS_REWARD) ? f32(L->o_last_reward) : 0.0f;
      setflag(HDR_ATT_RR | HDR_DEF_RR, true);
      setf32(L->o_last_att, last);
    }
    cyber_reset(init);
    setflag(HDR_ATT_RR | HDR_HAS_CYBER | HDR_HAS_REWARD, false);
    w(L->o_att_valid) = 0; w(L->o_att_invalid) = 0; w(L->o_att_ts) = 0;
    setf32(L->o_last_cyber, 0.f); setf32(L->o_last_reward, 0.f); setf32(L->o_att_return, 0.f);
  }
  __device__ void defender_reset(const uint32_t* init) const {  // DWR:414-477
    if (!flag(HDR_DEF_RR)) { setflag(HDR_ATT_RR | HDR_DEF_RR, true); setf32(L->o_last_att, 0.f); }
    cyber_reset(init);
    setflag(HDR_DEF_RR | HDR_BREACHED, false);
    setf32(L->o_last_att, 0.f);
    w(L->o_def_ts) = 0; w(L->o_def_valid) = 0; w(L->o_def_invalid) = 0;
    uint32_t a = w(L->o_avail);
    // _prev_network_availability = the bound actuator's value: the stale copy's (it outlives resets), or, live, the fresh
    // actuator's 1.0 (bits 8-15 were just cleared with the env)
    w(L->o_avail) = (a & ~0xFF0000u) | (((a >> 8) & 0xFFu) << 16);
    setf32(L->o_def_return, 0.f);
  }
  __device__ bool defender_action_valid(const int32_t* a) const {  // DWR:329-412, on the LIVE env
    // node coordinates beyond the scenario's own nodes (padded action space) are invalid; so is anything negative or past the
    // six firewall rule names -- values the MultiDiscrete space cannot produce are an invalid action, never an index
    const int n = n_nodes();
    auto node_ok = [&](int x) { return x >= 0 && x < n && !bit(L->o_notrunning, x); };
    switch (a[0]) {
      case 0: return node_ok(a[1]) && (node_rec(a[1])[CBX_N_FLAGS] & 1u);
      case 1: return node_ok(a[2]) && a[3] >= 0 && a[3] < 6 &&
                     (fx ? fw_has_rule(a[2], a[4] != 0, a[3]) : (bool)((node_rec(a[2])[CBX_N_DEFOBS] >> ((a[4] ? 0 : 8) + a[3])) & 1u));
      case 2: return node_ok(a[5]);
      case 3: return node_ok(a[8]) && a[9] >= 0 && a[9] < (int)((node_rec(a[8])[CBX_N_FLAGS] >> 8) & 0xFFu);
      case 4: return node_ok(a[10]) && a[11] >= 0 && a[11] < (int)((node_rec(a[10])[CBX_N_FLAGS] >> 8) & 0xFFu);
      default: return false;
    }
  }
};

}  // namespace cbx
#endif  // CBX_DEVICE_CUH_
