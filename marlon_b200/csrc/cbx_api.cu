// cbx_api.cu -- host side of the C ABI declared in include/cbx.h: scenario upload, HBM layout, launches.
// No simulation logic lives here and there is no CPU path: every entry point that computes launches a kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "cbx_layout.h"

extern "C" {
cudaError_t cbx_launch_step(const cbx_params* p, int op, int grid, int smem_bytes, int use_tma, cudaStream_t stream);
cudaError_t cbx_launch_sample(const cbx_params* p, int32_t* att, int32_t* def, uint64_t seed, uint32_t step, cudaStream_t stream);
cudaError_t cbx_kernel_attrs(int smem_bytes, int use_tma, int fast, int* blocks_per_sm);
cudaError_t cbx_pipe_attrs(int enc, int smem_bytes, int live);
cudaError_t cbx_launch_pipe(const cbx_params* p, int op, int grid, cudaStream_t stream);
cudaError_t cbx_launch_gae(const float* rewards, const float* values, const uint8_t* episode_starts, const float* last_values,
                           const uint8_t* last_dones, float gamma, float lam, int T, int64_t n, float* advantages, float* returns,
                           cudaStream_t stream);
cudaError_t cbx_wide_attrs(int smem_bytes);
cudaError_t cbx_launch_wide(const cbx_params* p, int op, int grid, cudaStream_t stream);
}

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                                        \
  do {                                                                                                        \
    cudaError_t _e = (expr);                                                                                  \
    if (_e != cudaSuccess) return fail(CBX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                                   \
  } while (0)

cbx_fastdiv make_fastdiv(uint32_t d) {
  cbx_fastdiv f = {0, 0};
  if (d == 0) d = 1;
  uint32_t s = 0;
  while ((1u << (s + 1)) <= d && s < 31) ++s;  // s = floor(log2 d)
  if ((d & (d - 1)) == 0) { f.m = 0; f.s = s; return f; }
  unsigned __int128 num = ((unsigned __int128)1) << (32 + s);
  f.m = (uint32_t)((num + d - 1) / d);
  f.s = s;
  return f;
}

int align_up(int x, int a) { return (x + a - 1) / a * a; }

// Shared-memory plan of the pipelined kernel (cbx_pipe.cuh) for `wl` logic and `we` encoder warps; false when the
// configuration does not qualify (mask rows that no bulk copy can carry) or does not fit in 227 KiB.
bool plan_pipe(const cbx_params& p, int wl, int we, cbx_pipe_plan* Q) {
  const cbx_layout& L = p.lay;
  memset(Q, 0, sizeof(*Q));
  if (p.enc.warp_env < 1 || CBX_TILE != 32 || wl < 1 || we < 1 || (wl + we + 1) * 32 > 512) return false;
  const bool dense = L.sz_connect > 0;
  const int ROWR = L.N * L.R, ROWC = L.N * L.P * L.C;
  int gs = 1;
  if (dense) {
    if (ROWR % 8 || ROWC % 8 || (L.N * ROWR) % 16 || L.sz_local % 4) return false;
    if (ROWC % 16 == 0) gs = 1;
    else if (L.N % 2 == 0) gs = 2;
    else return false;
  }
  Q->wl = wl; Q->we = we; Q->nslot = 2 * wl; Q->gs = gs;
  const bool defobs = p.cfg.mode == CBX_MODE_MARLON && p.cfg.def_enabled;
  int64_t o = 0;
  auto up = [](int64_t x, int a) { return (x + a - 1) / a * a; };
  Q->tables = (int)o; o = up(o + p.table_words + ((L.S + 3) & ~3) + p.fwx_words, 32);
  Q->lut = (int)o; o += 512;
  Q->bars = (int)o; o = up(o + 2 * (1 + wl + 3 * Q->nslot), 32);
  Q->done_ring = (int)o; o += 96;  // overlapped launches: completed parts + tile index of 32 tracked tiles, publisher progress
  Q->zero = (int)o; o = up(o + (dense ? gs * ROWC / 4 : 0), 32);
  // static parts of the defender observation for a tile: [32][6n] in | [32][6n] out | [32][services]; under the live binding
  // the firewall rows are per-env state (images in the logic buffers below) and only the service rows remain
  const bool live = L.o_fw >= 0;
  Q->def_svc = live ? 0 : 2 * CBX_TILE * 6 * L.n;
  Q->def_static = (int)o; o = up(o + (defobs ? (Q->def_svc + CBX_TILE * L.nservices + 3) / 4 : 0), 32);
  // logic buffers: state tile | staging | field images [32][words per env] | actions (aliasing the property image when it
  // is large enough: the actions are consumed before the images are laid out)
  Q->lbufs = (int)o;
  {
    int64_t q = 0;
    q += (int64_t)L.S * CBX_TILE;
    Q->l_stage = (int)q; q += (int64_t)L.G * CBX_TILE;
    Q->i_scal = (int)q; q += 8 * CBX_TILE;
    Q->i_leak = (int)q; q += 4 * L.LEAK * CBX_TILE;
    Q->i_cachem = (int)q; q += 2 * L.C * CBX_TILE;
    Q->i_props = (int)q; q += (int64_t)L.N * L.nprops * CBX_TILE;
    Q->i_priv = (int)q; q += L.N * CBX_TILE;
    Q->i_local = (int)q; q += (int64_t)(L.sz_local / 4) * CBX_TILE;
    Q->i_fwin = Q->i_fwout = -1;
    if (live && defobs) {  // 192 n bytes each: whole 16-byte words
      Q->i_fwin = (int)q; q += CBX_TILE * 6 * L.n / 4;
      Q->i_fwout = (int)q; q += CBX_TILE * 6 * L.n / 4;
    }
    if (L.N * L.nprops >= 22) Q->l_acts = Q->i_props;
    else { Q->l_acts = (int)q; q += 22 * CBX_TILE; }
    q = up(q, 32);
    if (q * 4 * wl > 227 * 1024) return false;
    Q->lbuf_words = (int)q;
  }
  o += (int64_t)wl * Q->lbuf_words;
  Q->slots = (int)o;
  Q->s_hdr = p.enc.desc_words * CBX_TILE;
  Q->slot_words = Q->s_hdr + 32;
  o += (int64_t)Q->nslot * Q->slot_words;
  Q->wbufs = (int)o;
  {
    int64_t bo = 0;
    Q->b_remote = (int)bo; bo = up(bo + (dense ? L.N * ROWR : 0), 16);
    Q->b_conn = (int)bo; bo = up(bo + (dense ? (int64_t)(gs == 1 ? ROWC : 6 * ROWC) : 0), 16);
    Q->b_inf = (int)bo; bo = up(bo + (defobs ? CBX_TILE * L.n : 0), 16);
    if (bo * we > 227 * 1024) return false;
    Q->wbuf_words = (int)up(bo / 4, 32);
  }
  o += (int64_t)we * Q->wbuf_words;
  if (o * 4 > 227 * 1024) return false;
  Q->total_bytes = (int)(o * 4);
  return true;
}

// Shared-memory plan of the warp-per-tile kernel (cbx_wide.cuh): factored masks only.
bool plan_wide(const cbx_params& p, int sms, cbx_wide_plan* Q) {
  const cbx_layout& L = p.lay;
  memset(Q, 0, sizeof(*Q));
  if (p.enc.warp_env < 1 || CBX_TILE != 32 || L.sz_connect > 0) return false;
  // image words per env: the tile's actions (22), the credential cache two entries per word (chunks of at most 56 words), the
  // property bit stream of a chunk of nodes + their 2-bit privilege codes, the infected-node bytes, the defender's static rows
  auto cdiv = [](int a, int b) { return (a + b - 1) / b; };
  const int cache_words = cdiv(L.C, 2) < 56 ? cdiv(L.C, 2) : 56;
  int npc = (L.N + 3) & ~3;                                         // nodes per chunk: a multiple of 4, all of them if they fit
  while (npc > 4 && cdiv(npc * L.nprops, 32) + cdiv(npc, 16) > 64) npc -= 4;
  const int prop_words = cdiv(npc * L.nprops, 32) + cdiv(npc, 16);
  const int n6s = (6 * L.n + 4 + 3) & ~3, svs = (L.nservices + 4 + 3) & ~3;
  int iw = 22;
  if (cache_words > iw) iw = cache_words;
  if (prop_words > iw) iw = prop_words;
  if (cdiv(L.n, 4) > iw) iw = cdiv(L.n, 4);
  if (cdiv(2 * n6s + svs, 4 * CBX_TILE) > iw) iw = cdiv(2 * n6s + svs, 4 * CBX_TILE);
  Q->img_words = iw;
  Q->img_stride = iw | 1;
  Q->nodes_per_chunk = npc;
  int64_t o = 0;
  Q->warps = (int)o;
  int64_t q = 0;
  Q->w_stage = (int)q; q += (int64_t)L.G * CBX_TILE;
  Q->w_desc = (int)q; q += (int64_t)p.enc.desc_words * CBX_TILE;
  Q->w_img = (int)q; q += (int64_t)Q->img_stride * CBX_TILE;
  q = (q + 31) / 32 * 32;
  Q->warp_words = (int)q;
  // as many warps as fit while leaving L1 room for the scenario tables and the hot state lines; then the FEWEST warps that
  // still need the same number of rounds over the tiles (each warp takes whole tiles: with 4096 tiles, 11 to 13 warps per SM
  // all need three rounds and the extra warps only add contention, 14 need two)
  int nw = (int)((212 * 1024 / 4 - o) / q);
  if (nw > CBX_WIDE_WARPS) nw = CBX_WIDE_WARPS;
  if (nw < 4) return false;
  if (sms > 0 && p.n_tiles > 0) {
    const int64_t rounds = (p.n_tiles + (int64_t)sms * nw - 1) / ((int64_t)sms * nw);
    while (nw > 4 && (p.n_tiles + (int64_t)sms * (nw - 1) - 1) / ((int64_t)sms * (nw - 1)) == rounds) --nw;
  }
  Q->nwarps = nw;
  o += nw * q;
  Q->total_bytes = (int)(o * 4);
  return true;
}

}  // namespace

constexpr int kTicketRing = 1024;

struct cbx_scenario {
  std::vector<uint32_t> blob;
  std::vector<uint32_t> fwx;  // firewall extension tables (cbx_scenario_set_firewall_tables), empty if never set
  int n, P, nprops, L, R, nsecrets, ntriples, nservices, max_leak, flags;
};

struct cbx_batch {
  cbx_params p;
  int device, grid, smem_bytes, use_tma;
  int pipe_grid;  // pipelined kernel: CTAs (one per SM); the kernel is used when p.pipe.enabled
  int wide_grid;  // warp-per-tile kernel (large state, factored masks); used when p.wide.enabled
  std::vector<void*> allocs;
  uint32_t* d_tables;
  int64_t launches;
  uint32_t sample_step;
  // pinned staging for the host-buffer entry point
  int32_t *h_att, *h_def, *d_att, *d_def;
  uint8_t* h_out;
  int host_ready;  // cbx_batch_host_prepare has allocated the five buffers above
  // kernel timing
  int timing;
  std::vector<cudaEvent_t> ev;
  size_t ev_used;
  double ms_sum;
  int64_t ms_count;
  // overlapped launches: an event between two launches would serialise them, so the timed region is bracketed instead
  int region_open;
  int64_t region_count;
  cudaStream_t region_stream;
  const cbx_scenario* scn;       // dimensions the layout was computed from (== &vscn for a multi-scenario batch)
  cbx_scenario vscn;             // multi-scenario batch: the element-wise maximum of the scenarios' dimensions (no blob)
  int32_t* d_tile_scn;
  std::vector<uint32_t> init_state;
  int* ticket_ring;  // pipelined kernel: kTicketRing per-launch ticket counters (slot = seq % kTicketRing)
};

static void host_release(cbx_batch* b) {
  if (b->h_att) cudaFreeHost(b->h_att);
  if (b->h_def) cudaFreeHost(b->h_def);
  if (b->h_out) cudaFreeHost(b->h_out);
  if (b->d_att) cudaFree(b->d_att);
  if (b->d_def) cudaFree(b->d_def);
  b->h_att = b->h_def = b->d_att = b->d_def = nullptr;
  b->h_out = nullptr;
  b->host_ready = 0;
}


extern "C" {

const char* cbx_last_error(void) { return g_err.c_str(); }
int cbx_abi_version(void) { return CBX_ABI_VERSION; }

int cbx_scenario_create(const void* tables, size_t nbytes, cbx_scenario** out) {
  if (!tables || !out || nbytes < CBX_H_WORDS * 4 || nbytes % 4) return fail(CBX_ERR_INVALID, "scenario blob too small or misaligned");
  const uint32_t* w = (const uint32_t*)tables;
  if (w[CBX_H_MAGIC] != CBX_SCN_MAGIC) return fail(CBX_ERR_INVALID, "bad scenario magic");
  if (w[CBX_H_VERSION] != CBX_SCN_VERSION) return fail(CBX_ERR_INVALID, "scenario version %u, library expects %u", w[CBX_H_VERSION], CBX_SCN_VERSION);
  if ((size_t)w[CBX_H_TOTAL_WORDS] * 4 != nbytes || w[CBX_H_TOTAL_WORDS] % 4) return fail(CBX_ERR_INVALID, "scenario blob size mismatch");
  cbx_scenario* s = new cbx_scenario();
  s->blob.assign(w, w + nbytes / 4);
  s->n = (int)w[CBX_H_N_NODES]; s->P = (int)w[CBX_H_N_PORTS]; s->nprops = (int)w[CBX_H_N_PROPS];
  s->L = (int)w[CBX_H_N_LOCAL]; s->R = (int)w[CBX_H_N_REMOTE]; s->nsecrets = (int)w[CBX_H_N_SECRETS];
  s->ntriples = (int)w[CBX_H_N_TRIPLES]; s->nservices = (int)w[CBX_H_N_SERVICES]; s->max_leak = (int)w[CBX_H_MAX_LEAK];
  s->flags = (int)w[CBX_H_FLAGS];
  const size_t tot = w[CBX_H_TOTAL_WORDS];
  const int V = s->L + s->R, Ws = (s->nsecrets + 31) / 32;
  bool ok = s->n >= 1 && s->n <= 255 && s->P >= 1 && s->P <= 32 && s->nprops >= 1 && s->nprops <= 64 && s->L >= 1 && s->R >= 1 &&
            s->ntriples <= 65535 && (size_t)w[CBX_H_OFF_NODE] + (size_t)s->n * CBX_NODE_WORDS <= tot &&
            (size_t)w[CBX_H_OFF_AUTH] + (size_t)s->n * s->P * Ws <= tot &&
            (size_t)w[CBX_H_OFF_VULN] + (size_t)s->n * V * CBX_VULN_WORDS <= tot &&
            (size_t)w[CBX_H_OFF_PAYLOAD] + w[CBX_H_N_PAYLOAD] <= tot && (size_t)w[CBX_H_OFF_TRIPLE] + 3 * (size_t)s->ntriples <= tot;
  if (!ok) { delete s; return fail(CBX_ERR_INVALID, "scenario header out of range"); }
  *out = s;
  return CBX_OK;
}

int cbx_scenario_destroy(cbx_scenario* s) {
  delete s;
  return CBX_OK;
}

int cbx_scenario_set_firewall_tables(cbx_scenario* s, const void* words, size_t nbytes) {
  if (!s || !words || nbytes < 16 || nbytes % 16) return fail(CBX_ERR_INVALID, "firewall tables: null, too small or not a multiple of 16 bytes");
  const uint32_t* w = (const uint32_t*)words;
  if (w[CBX_FX_MAGIC] != CBX_FWX_MAGIC) return fail(CBX_ERR_INVALID, "bad firewall-table magic");
  const size_t need = (size_t)CBX_FX_WORDS + s->P + s->n + 2 * (size_t)w[CBX_FX_N_GROUPS];
  if (w[CBX_FX_N_NAMES] < 6 || w[CBX_FX_N_NAMES] > 32 || w[CBX_FX_N_GROUPS] < 1 || w[CBX_FX_N_GROUPS] > 2u * s->n || need * 4 > nbytes)
    return fail(CBX_ERR_INVALID, "firewall tables: header out of range");
  for (int p = 0; p < s->P; ++p)
    if (w[CBX_FX_WORDS + p] >= w[CBX_FX_N_NAMES]) return fail(CBX_ERR_INVALID, "firewall tables: port name index out of range");
  for (int i = 0; i < s->n; ++i) {
    const uint32_t g = w[CBX_FX_WORDS + s->P + i];
    if ((g & 0xFFFFu) >= w[CBX_FX_N_GROUPS] || (g >> 16) >= w[CBX_FX_N_GROUPS]) return fail(CBX_ERR_INVALID, "firewall tables: group out of range");
  }
  s->fwx.assign(w, w + nbytes / 4);
  return CBX_OK;
}

int cbx_config_default(cbx_config* c) {
  if (!c) return fail(CBX_ERR_INVALID, "null config");
  memset(c, 0, sizeof(*c));
  c->abi_version = CBX_ABI_VERSION;
  c->mode = CBX_MODE_CYBERBATTLE;
  c->maximum_node_count = 100;
  c->maximum_total_credentials = 1000;
  c->maximum_discoverable_credentials_per_action = 5;
  c->throws_on_invalid_actions = 1;
  c->has_attacker_goal = 1;
  c->goal_low_availability = 1.0;
  c->goal_own_atleast_percent = 1.0;
  c->defender_goal_eviction = 1;
  c->winning_reward = 5000.0;
  c->scan_frequency = 1;
  c->kind_of_index[0] = CBX_KIND_CONNECT; c->kind_of_index[1] = CBX_KIND_LOCAL; c->kind_of_index[2] = CBX_KIND_REMOTE;
  c->att_max_timesteps = 2000;
  c->att_invalid_action_reward_modifier = -1.0;
  c->def_max_timesteps = 100;
  c->def_reset_on_constraint_broken = 1;
  c->auto_reset = 1;
  c->def_loss_reward = -5000.0;
  c->def_sla_worsening_penalty_scale = 200.0;
  return CBX_OK;
}

static int compute_layout(const cbx_scenario* s, const cbx_config* cfg, int64_t n_envs, cbx_layout* L) {
  memset(L, 0, sizeof(*L));
  const bool live = cfg->mode == CBX_MODE_MARLON && cfg->def_enabled && cfg->def_binding == CBX_DEF_BINDING_LIVE;
  if (live && s->fwx.empty())
    return fail(CBX_ERR_INVALID, "the live defender binding needs the scenario's firewall tables (cbx_scenario_set_firewall_tables)");
  L->n = s->n; L->N = cfg->maximum_node_count; L->C = cfg->maximum_total_credentials;
  L->LEAK = cfg->maximum_discoverable_credentials_per_action;
  L->P = s->P; L->L = s->L; L->R = s->R; L->nprops = s->nprops; L->nsecrets = s->nsecrets; L->ntriples = s->ntriples;
  L->nservices = s->nservices;
  if (L->n > L->N) return fail(CBX_ERR_INVALID, "Network node count (%d) exceeds the specified limit of %d.", L->n, L->N);  // ENV:416-418
  if (s->max_leak > L->LEAK)
    return fail(CBX_ERR_INVALID, "Some action in the environment returns %d credentials which exceeds the maximum number of discoverable credentials of %d",
                s->max_leak, L->LEAK);  // ENV:430-435
  if (L->N > 255) return fail(CBX_ERR_UNSUPPORTED, "maximum_node_count %d > 255", L->N);
  if (L->C > 65535) return fail(CBX_ERR_UNSUPPORTED, "maximum_total_credentials %d > 65535", L->C);
  if (s->ntriples > L->C)
    return fail(CBX_ERR_UNSUPPORTED, "scenario has %d distinct credentials but maximum_total_credentials is %d (the reference would emit out-of-space observations)",
                s->ntriples, L->C);
  L->Wn = (L->n + 31) / 32;
  L->PW = (L->nprops + 31) / 32;
  L->AW = (2 * (L->L + L->R) + 31) / 32;
  L->OW = (L->N + 31) / 32;
  int o = 0;
  L->o_hdr = o++;
  L->o_att_ts = o++; L->o_def_ts = o++;
  L->o_att_valid = o++; L->o_att_invalid = o++; L->o_def_valid = o++; L->o_def_invalid = o++;
  L->o_last_cyber = o++; L->o_last_reward = o++; L->o_last_att = o++; L->o_att_return = o++; L->o_def_return = o++;
  L->o_avail = o++;
  L->o_cd_shadow = o; o += (L->n + 3) / 4;
  L->o_cyber_begin = o;
  L->o_stepcount = o++;
  L->o_ep_sum = o++;
  L->o_installed = o; o += L->Wn;
  L->o_everowned = o; o += L->Wn;
  L->o_notrunning = o; o += L->Wn;
  L->o_priv = o; o += (L->n + 15) / 16;
  if (s->flags & 1) { L->o_tags = o; o += (L->n + 7) / 8; } else L->o_tags = -1;
  L->o_fw = -1; L->n_fw_groups = 0;
  if (live) { L->n_fw_groups = (int)s->fwx[CBX_FX_N_GROUPS]; L->o_fw = o; o += 2 * L->n_fw_groups; }
  L->o_cd_live = o; o += (L->n + 3) / 4;
  L->o_disc_order = o; o += (L->n + 3) / 4;
  L->o_disc_idx = o; o += (L->n + 3) / 4;
  L->o_props = o; o += L->n * L->PW;
  L->o_attacked = o; o += L->n * L->AW;
  L->o_gathered = o; o += (L->nsecrets + 31) / 32;
  L->o_cached = o; o += (L->ntriples + 31) / 32 > 0 ? (L->ntriples + 31) / 32 : 1;
  L->o_cache = o; o += (L->ntriples + 1 + 1) / 2;
  L->S = o;
  int g = 11 + L->Wn;
  L->LEAKS = s->max_leak < L->LEAK ? s->max_leak : L->LEAK;
  L->g_leaked = g; g += L->LEAKS;  // one packed word per slot (cbx_device.cuh leak_pack)
  if (L->C > 32767 || L->P > 255) return fail(CBX_ERR_UNSUPPORTED, "maximum_total_credentials > 32767 or more than 255 ports");
  L->g_inst = g; g += L->Wn;
  L->g_priv = g; g += (L->n + 15) / 16;
  L->G = g;
  const int dense = cfg->mask_mode == CBX_MASK_DENSE;
  int64_t szc = (int64_t)L->N * L->N * L->P * L->C;
  if (dense && szc * (n_envs < CBX_TILE ? n_envs : CBX_TILE) >= (1ll << 31)) return fail(CBX_ERR_UNSUPPORTED, "dense connect mask of %lld bytes per env is too large; use factored masks", (long long)szc);
  L->sz_local = dense ? L->N * L->L : 0;
  L->sz_remote = dense ? L->N * L->N * L->R : 0;
  L->sz_connect = dense ? (int)szc : 0;
  return CBX_OK;
}

static void build_init_state(const cbx_scenario* s, const cbx_layout& L, std::vector<uint32_t>& st) {
  st.assign((size_t)((L.S + 3) & ~3), 0u);
  const uint32_t* w = s->blob.data();
  const uint32_t* node = w + w[CBX_H_OFF_NODE];
  auto setbyte = [&](int off, int i, uint32_t v) { st[off + i / 4] = (st[off + i / 4] & ~(0xFFu << ((i & 3) * 8))) | (v << ((i & 3) * 8)); };
  for (int i = 0; i < ((L.n + 3) / 4) * 4; ++i) setbyte(L.o_disc_idx, i, 0xFFu);
  int nd = 0;
  for (int i = 0; i < s->n; ++i) {  // the scenario's own nodes (L.n is the padded count in a multi-scenario batch)
    const uint32_t* r = node + (size_t)i * CBX_NODE_WORDS;
    uint32_t f = r[CBX_N_FLAGS];
    int priv = (f >> 2) & 3;
    if (f & 2u) {  // agent_installed: AgentActions.__init__ marks it owned at LocalUser and discovers its properties (ACT:149-152)
      st[L.o_installed + i / 32] |= 1u << (i % 32);
      st[L.o_everowned + i / 32] |= 1u << (i % 32);
      if (priv < 1) priv = 1;
      st[L.o_props + i * L.PW] = r[CBX_N_PROPS_LO];
      if (L.PW > 1) st[L.o_props + i * L.PW + 1] = r[CBX_N_PROPS_HI];
      setbyte(L.o_disc_order, nd, (uint32_t)i);  // ENV:392-394
      setbyte(L.o_disc_idx, i, (uint32_t)nd);
      nd++;
    }
    st[L.o_priv + i / 16] |= (uint32_t)priv << ((i % 16) * 2);
  }
  st[L.o_hdr] = (uint32_t)nd;
  if (L.o_fw >= 0) {  // live defender binding: the rule lists as the scenario defines them
    const uint32_t* g0 = s->fwx.data() + CBX_FX_WORDS + s->P + s->n;
    for (int k = 0; k < 2 * L.n_fw_groups; ++k) st[L.o_fw + k] = g0[k];
  }
}

static int create_impl(const cbx_scenario* const* scns, int n_scn, const int64_t* counts, int64_t n_envs, const cbx_config* cfg, int device,
                       cbx_batch** out);

int cbx_batch_create(const cbx_scenario* s, int64_t n_envs, const cbx_config* cfg, int device, cbx_batch** out) {
  if (!s) return fail(CBX_ERR_INVALID, "null argument");
  return create_impl(&s, 1, &n_envs, n_envs, cfg, device, out);
}

int cbx_batch_create_multi(const cbx_scenario* const* scenarios, int n_scenarios, const int64_t* envs_per_scenario, const cbx_config* cfg,
                           int device, cbx_batch** out) {
  if (!scenarios || !envs_per_scenario || n_scenarios < 1) return fail(CBX_ERR_INVALID, "null argument");
  int64_t total = 0;
  const cbx_scenario* s0 = scenarios[0];
  for (int k = 0; k < n_scenarios; ++k) {
    const cbx_scenario* s = scenarios[k];
    if (!s) return fail(CBX_ERR_INVALID, "scenario %d is null", k);
    if (envs_per_scenario[k] <= 0) return fail(CBX_ERR_INVALID, "scenario %d has no envs", k);
    if (k + 1 < n_scenarios && envs_per_scenario[k] % CBX_TILE)
      return fail(CBX_ERR_INVALID, "envs_per_scenario[%d] = %lld: every group but the last must be a multiple of %d envs (a tile shares one "
                  "scenario's tables)", k, (long long)envs_per_scenario[k], CBX_TILE);
    // one observation / action space for the whole batch: the identifier-derived dimensions must agree
    if (s->P != s0->P || s->L != s0->L || s->R != s0->R || s->nprops != s0->nprops || s->flags != s0->flags)
      return fail(CBX_ERR_INVALID, "scenario %d: ports / vulnerability ids / properties differ from scenario 0 (one batch, one Identifiers)", k);
    total += envs_per_scenario[k];
  }
  return create_impl(scenarios, n_scenarios, envs_per_scenario, total, cfg, device, out);
}

static int create_impl(const cbx_scenario* const* scns, int n_scn, const int64_t* counts, int64_t n_envs, const cbx_config* cfg, int device,
                       cbx_batch** out) {
  if (!cfg || !out) return fail(CBX_ERR_INVALID, "null argument");
  const cbx_scenario* s = scns[0];
  if (cfg->abi_version != CBX_ABI_VERSION) return fail(CBX_ERR_INVALID, "config abi_version %d, library %d", cfg->abi_version, CBX_ABI_VERSION);
  if (n_envs <= 0) return fail(CBX_ERR_INVALID, "n_envs must be positive");
  if (device < 0) return fail(CBX_ERR_NODEVICE, "device %d: this library has no CPU path", device);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(CBX_ERR_NODEVICE, "no CUDA device available (there is no CPU fallback)");
  if (device >= ndev) return fail(CBX_ERR_NODEVICE, "device %d out of range (%d devices)", device, ndev);
  if (cfg->mode != CBX_MODE_CYBERBATTLE && cfg->mode != CBX_MODE_MARLON) return fail(CBX_ERR_INVALID, "bad mode");
  if (cfg->builtin_defender == CBX_BUILTIN_SCAN_AND_REIMAGE && (cfg->scan_frequency <= 0 || cfg->scan_capacity < 0))
    return fail(CBX_ERR_INVALID, "scan_frequency must be positive");
  { int seen = 0; for (int k = 0; k < 3; ++k) { if (cfg->kind_of_index[k] < 0 || cfg->kind_of_index[k] > 2) return fail(CBX_ERR_INVALID, "kind_of_index"); seen |= 1 << cfg->kind_of_index[k]; }
    if (seen != 7) return fail(CBX_ERR_INVALID, "kind_of_index must be a permutation"); }
  CUDA_TRY(cudaSetDevice(device));
  cbx_batch* b = new cbx_batch();
  memset(&b->p, 0, sizeof(b->p));
  b->device = device; b->launches = 0; b->sample_step = 0; b->timing = 0; b->ev_used = 0; b->ms_sum = 0; b->ms_count = 0;
  b->region_open = 0; b->region_count = 0; b->region_stream = nullptr;
  b->h_att = b->h_def = b->d_att = b->d_def = nullptr; b->h_out = nullptr; b->host_ready = 0; b->scn = s; b->d_tile_scn = nullptr;
  size_t max_blob = s->blob.size();
  if (n_scn > 1) {  // padded layout: the maximum of every per-scenario dimension
    cbx_scenario& v = b->vscn;
    v.n = s->n; v.P = s->P; v.nprops = s->nprops; v.L = s->L; v.R = s->R; v.nsecrets = s->nsecrets; v.ntriples = s->ntriples;
    v.nservices = s->nservices; v.max_leak = s->max_leak; v.flags = s->flags;
    for (int k = 1; k < n_scn; ++k) {
      const cbx_scenario* q = scns[k];
      v.n = q->n > v.n ? q->n : v.n; v.nsecrets = q->nsecrets > v.nsecrets ? q->nsecrets : v.nsecrets;
      v.ntriples = q->ntriples > v.ntriples ? q->ntriples : v.ntriples; v.nservices = q->nservices > v.nservices ? q->nservices : v.nservices;
      v.max_leak = q->max_leak > v.max_leak ? q->max_leak : v.max_leak;
      max_blob = q->blob.size() > max_blob ? q->blob.size() : max_blob;
    }
    b->scn = s = &b->vscn;
  }
  int rc = compute_layout(s, cfg, n_envs, &b->p.lay);
  if (rc) { delete b; return rc; }
  cbx_layout& L = b->p.lay;
  b->p.cfg = *cfg;
  b->p.n_envs = n_envs;
  b->p.n_pad = (n_envs + CBX_TILE - 1) / CBX_TILE * CBX_TILE;
  b->p.n_tiles = (int)(b->p.n_pad / CBX_TILE);
  b->p.table_words = (int)max_blob;
  b->p.n_scenarios = n_scn;
  const bool live = L.o_fw >= 0;
  if (live && n_scn > 1) { delete b; return fail(CBX_ERR_UNSUPPORTED, "the live defender binding is single-scenario"); }
  b->p.fwx_words = live ? (int)scns[0]->fwx.size() : 0;
  b->p.table_stride = (int)max_blob + ((b->p.lay.S + 3) & ~3) + b->p.fwx_words;
  b->p.tile_scn = nullptr;
  int col = 1;
  for (int k = 0; k < 3; ++k) {
    int kind = cfg->kind_of_index[k];
    b->p.slice_of_kind[kind] = col;
    col += kind == CBX_KIND_LOCAL ? 2 : kind == CBX_KIND_REMOTE ? 3 : 4;
  }
  // encoder constants
  cbx_enc_consts& K = b->p.enc;
  K.d_leaked = make_fastdiv(4 * L.LEAK); K.d_cachem = make_fastdiv(2 * L.C); K.d_props = make_fastdiv(L.N * L.nprops);
  K.d_priv = make_fastdiv(L.N); K.d_nprops = make_fastdiv(L.nprops); K.d_L = make_fastdiv(L.L);
  K.d_local = make_fastdiv(L.N * L.L); K.d_remote = make_fastdiv(L.N * L.N * L.R);
  K.d_connect = make_fastdiv((uint32_t)((int64_t)L.N * L.N * L.P * L.C)); K.d_rowr = make_fastdiv(L.N * L.R);
  K.d_rowc = make_fastdiv(L.N * L.P * L.C); K.d_C = make_fastdiv(L.C); K.d_n = make_fastdiv(L.n);
  K.d_6n = make_fastdiv(6 * L.n); K.d_svc = make_fastdiv(L.nservices > 0 ? L.nservices : 1);
  K.desc_words = 8 + L.OW + L.Wn;
  K.debug_skip = 0;
#ifdef CBX_EXPERIMENTS
  { const char* dbg = getenv("CBX_DEBUG_SKIP"); K.debug_skip = dbg ? atoi(dbg) : 0; }
#endif
  auto gcd16 = [](int x) { int g = 16; while (x % g) g >>= 1; return g; };
  {
    const int row_r = L.N * L.R, row_c = L.N * L.P * L.C;
    K.tmpl_unit_r = gcd16(row_r);
    K.tmpl_unit_c = gcd16(row_c);
    const char* slow = getenv("CBX_GENERIC_ENCODER");
    bool ok = !(slow && slow[0] == '1');
    if (cfg->mask_mode == CBX_MASK_DENSE) {
      // 4-byte store granules at least, rows short enough for CBX_MAXG (4) granules per lane, word-sized local mask
      ok = ok && K.tmpl_unit_r >= 4 && K.tmpl_unit_c >= 4 && row_r / K.tmpl_unit_r <= 128 && row_c / K.tmpl_unit_c <= 128 &&
           L.sz_local % 4 == 0;
    }
    K.warp_env = ok ? 1 : 0;
    // statically specialised encoders for MARLon's canonical configurations (dense masks)
    const char* nospec = getenv("CBX_NO_STATIC_DIMS");
    if (ok && cfg->mask_mode == CBX_MASK_DENSE && !(nospec && nospec[0] == '1')) {
      auto is = [&](int N, int Lv, int R, int P, int C, int props, int leak, int n, int svc) {
        return L.N == N && L.L == Lv && L.R == R && L.P == P && L.C == C && L.nprops == props && L.LEAK == leak && L.n == n &&
               L.nservices == svc;
      };
      if (is(12, 3, 8, 7, 10, 10, 5, 10, 13)) K.warp_env = 2;       // ToyCtf (12, 10)
      else if (is(12, 5, 2, 8, 12, 14, 5, 12, 22)) K.warp_env = 3;  // Chain-10 (12, 12)
    }
  }
  // shared-memory plan
  cbx_smem_plan& pl = b->p.plan;
  int o = 0;
  pl.tables = o; o = align_up(o + b->p.table_words + ((L.S + 3) & ~3) + b->p.fwx_words, 32);
  pl.lut = o; o = align_up(o + 512, 32);
  pl.bars = o; o = align_up(o + 8, 32);
  pl.state = o; o = align_up(o + L.S * CBX_TILE, 32);
  pl.stage = o; o = align_up(o + L.G * CBX_TILE, 32);
  pl.desc = o; o = align_up(o + K.desc_words * CBX_TILE, 32);
  pl.acts = o; o = align_up(o + 22 * CBX_TILE, 32);
  pl.total_bytes = o * 4;
  b->smem_bytes = pl.total_bytes;
  if (b->smem_bytes > 227 * 1024) {
    delete b;
    return fail(CBX_ERR_UNSUPPORTED, "scenario needs %d bytes of shared memory per CTA (> 227 KiB)", pl.total_bytes);
  }
  b->use_tma = 1;
#ifdef CBX_EXPERIMENTS  // plain-copy staging instead of TMA bulk copies: debugging aid, not in the release library
  { const char* no_tma = getenv("CBX_NO_TMA"); b->use_tma = !(no_tma && no_tma[0] == '1'); }
#endif
  int bps = 0;
  { cudaError_t e = cbx_kernel_attrs(b->smem_bytes, b->use_tma, b->p.enc.warp_env, &bps);
    if (e != cudaSuccess) { delete b; return fail(CBX_ERR_CUDA, "kernel attributes: %s", cudaGetErrorString(e)); } }
  if (bps < 1) { delete b; return fail(CBX_ERR_CUDA, "step kernel does not fit on an SM"); }
  int sms = 0;
  { cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) { delete b; return fail(CBX_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); } }
  b->grid = sms * bps;
  if (b->grid > b->p.n_tiles) b->grid = b->p.n_tiles;
  const char* gridenv = getenv("CBX_GRID");
  if (gridenv && atoi(gridenv) > 0) b->grid = atoi(gridenv) < b->p.n_tiles ? atoi(gridenv) : b->p.n_tiles;
  // L2 policies of the bulk copies (see the pipelined kernel's plan below): on while the state is at most half of L2
  int l2_default = 0;
  { const char* lh = getenv("CBX_L2_HINTS");
    int l2_bytes = 0;
    cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, device);
    const int64_t state_bytes = (int64_t)b->p.n_pad * b->p.lay.S * 4;
    l2_default = lh ? atoi(lh) : (state_bytes * 2 <= (int64_t)l2_bytes ? 7 : 0); }
  b->p.l2_hints = l2_default & 1;  // the fused kernel: the state tile's bulk load / store only (its observation stores are plain)
  // pipelined kernel (logic warps ahead of TMA-storing encoder warps) when the configuration qualifies; CBX_PIPE=0 forces the
  // fused kernel, CBX_PIPE_WL / CBX_PIPE_WE / CBX_PIPE_CTAS (CTAs per SM) are tuning knobs
  b->p.pipe.enabled = 0; b->pipe_grid = 0;
  {
    const char* pe = getenv("CBX_PIPE");
    if (b->use_tma && n_scn == 1 && !(pe && pe[0] == '0')) {
      const char *ewl = getenv("CBX_PIPE_WL"), *ewe = getenv("CBX_PIPE_WE"), *ect = getenv("CBX_PIPE_CTAS");
      const int cand[4][2] = {{4, 8}, {3, 8}, {2, 8}, {2, 4}};
      cbx_pipe_plan Q;
      bool ok = false;
      if (ewl || ewe) ok = plan_pipe(b->p, ewl ? atoi(ewl) : 4, ewe ? atoi(ewe) : 8, &Q);
      for (int k = 0; k < 4 && !ok && !(ewl || ewe); ++k) ok = plan_pipe(b->p, cand[k][0], cand[k][1], &Q);
      if (ok && cbx_pipe_attrs(b->p.enc.warp_env, Q.total_bytes, live ? 1 : 0) == cudaSuccess) {
        Q.enabled = 1;
        { const char* lt = getenv("CBX_PIPE_LOGIC_TMA"); Q.logic_tma = !(lt && lt[0] == '0'); }
        int per_sm = ect ? atoi(ect) : 1;
        if (per_sm < 1) per_sm = 1;
        b->pipe_grid = sms * per_sm < b->p.n_tiles ? sms * per_sm : b->p.n_tiles;
        // dynamic tile order pays from ~24 tiles per CTA on (131 072 envs per GPU); CBX_PIPE_DYNAMIC=0/1 overrides
        // consecutive launches overlap (programmatic dependent launch + per-tile completion counters); CBX_PIPE_OVERLAP=0
        // restores fully serialised launches
        { const char* ov = getenv("CBX_PIPE_OVERLAP"); b->p.overlap = !(ov && ov[0] == '0'); }
        // Tile order.  Serialised launches: the dynamic order (global ticket counter) pays from ~24 tiles per CTA on (131 072
        // envs per GPU) and costs 2.5 % at the 14 tiles per CTA of 65 536 envs.  Overlapped launches: it wins at every size
        // (+6.5 % at 65 536 envs: a CTA that starts late -- its SM was still draining the previous launch -- takes fewer
        // tiles), so it is always on.  CBX_PIPE_DYNAMIC=0/1 overrides.
        { const char* dy = getenv("CBX_PIPE_DYNAMIC");
          Q.dynamic = dy ? atoi(dy) != 0 : (b->p.overlap || b->p.n_tiles >= 24 * b->pipe_grid); }
        // L2 cache policies on the bulk copies (cbx_params.l2_hints, decided above; CBX_L2_HINTS=0..15 overrides).  The per-env state is the
        // only data a step reads that an earlier step wrote: with evict_last on its tiles and evict_first on everything that
        // streams (observations, masks, actions) it stays in L2 under a write stream 50x its size, so a launch's first
        // state loads are L2 hits instead of HBM reads queued behind the previous launch's writes.  Measured (ToyCtf, B200,
        // profiles/r02_l2_policy_sweep.txt): roofline fraction 0.874 -> 0.972 at 65 536 envs, 0.931 -> 0.994 at 131 072,
        // level at 262 144 (59 MB of state), and -1 % at 1 048 576 envs, where the state (235 MB) cannot stay: policies on
        // while the state is at most half of L2.
        b->p.l2_hints = l2_default;
        b->p.pipe = Q;
      } else {
        cudaGetLastError();
      }
    }
  }

  // warp-per-tile kernel: large per-env state (or several scenarios) with factored masks; CBX_WIDE=0/1 overrides the choice
  b->p.wide.enabled = 0; b->wide_grid = 0;
  if (!b->p.pipe.enabled && b->use_tma && !live) {
    const char* we = getenv("CBX_WIDE");
    const bool want = we ? we[0] == '1' : (b->p.lay.S >= 128 || n_scn > 1);
    cbx_wide_plan Q;
    if (want && plan_wide(b->p, sms, &Q) && cbx_wide_attrs(Q.total_bytes) == cudaSuccess) {
      Q.enabled = 1;
      { const char* dy = getenv("CBX_WIDE_DYNAMIC"); Q.dynamic = dy ? atoi(dy) != 0 : 1; }  // never slower, +6 % on multi-scenario batches
      b->p.wide = Q;
      const int need = (b->p.n_tiles + Q.nwarps - 1) / Q.nwarps;
      b->wide_grid = sms < need ? sms : need;
    } else {
      cudaGetLastError();
    }
  }

  auto dalloc = [&](void** ptr, size_t bytes) -> cudaError_t {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) { b->allocs.push_back(*ptr); e = cudaMemset(*ptr, 0, bytes); }
    return e;
  };
#define ALLOC(field, type, per_env)                                                                                \
  do {                                                                                                             \
    void* _p = nullptr;                                                                                            \
    cudaError_t _e = dalloc(&_p, (size_t)b->p.n_pad * (size_t)(per_env) * sizeof(type));                          \
    if (_e != cudaSuccess) { int _rc = fail(CBX_ERR_CUDA, "cudaMalloc(%s): %s", #field, cudaGetErrorString(_e)); \
      cbx_batch_destroy(b); return _rc; }                                                                          \
    field = (type*)_p;                                                                                             \
  } while (0)
  {
    std::vector<uint32_t> tab;
    tab.reserve((size_t)n_scn * b->p.table_stride);
    for (int k = 0; k < n_scn; ++k) {  // per scenario: blob (zero-padded to the largest) then its initial per-env state
      std::vector<uint32_t> init;
      build_init_state(scns[k], L, init);
      if (k == 0) b->init_state = init;
      tab.insert(tab.end(), scns[k]->blob.begin(), scns[k]->blob.end());
      tab.resize(tab.size() + (max_blob - scns[k]->blob.size()), 0u);
      tab.insert(tab.end(), init.begin(), init.end());
      if (b->p.fwx_words) tab.insert(tab.end(), scns[k]->fwx.begin(), scns[k]->fwx.end());
    }
    void* pt = nullptr;
    cudaError_t e = dalloc(&pt, tab.size() * 4);
    if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "cudaMalloc(tables): %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
    b->d_tables = (uint32_t*)pt;
    e = cudaMemcpy(pt, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "upload tables: %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
    b->p.tables = b->d_tables;
    if (n_scn > 1) {
      std::vector<int32_t> ts((size_t)b->p.n_tiles);
      int64_t tile = 0;
      for (int k = 0; k < n_scn; ++k)
        for (int64_t q = 0; q < (counts[k] + CBX_TILE - 1) / CBX_TILE; ++q) ts[(size_t)tile++] = k;
      void* pts = nullptr;
      e = dalloc(&pts, ts.size() * 4);
      if (e == cudaSuccess) e = cudaMemcpy(pts, ts.data(), ts.size() * 4, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "upload tile scenarios: %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
      b->d_tile_scn = (int32_t*)pts;
      b->p.tile_scn = b->d_tile_scn;
    }
  }
  {
    void* pc = nullptr;
    cudaError_t e = dalloc(&pc, 64);
    if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "cudaMalloc(tile counter): %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
    b->p.tile_counter = (int*)pc;
    void* pd = nullptr;
    e = dalloc(&pd, (size_t)b->p.n_tiles * sizeof(uint32_t));
    if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "cudaMalloc(tile completion counters): %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
    b->p.tile_done = (uint32_t*)pd;
    b->p.seq = 0;
    void* pr = nullptr;
    e = dalloc(&pr, (size_t)kTicketRing * sizeof(int));
    if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "cudaMalloc(ticket ring): %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
    b->ticket_ring = (int*)pr;
    b->p.tickets = b->ticket_ring;
  }
  cbx_views& v = b->p.v;
  v.n_envs = n_envs; v.N = L.N; v.L = L.L; v.R = L.R; v.P = L.P; v.C = L.C; v.LEAK = L.LEAK; v.n_props = L.nprops;
  v.n_nodes = L.n; v.n_services = L.nservices; v.owned_words = L.OW;
  ALLOC(b->p.state, uint32_t, L.S);
  ALLOC(v.scalars, int32_t, 8);
  ALLOC(v.leaked_credentials, int32_t, 4 * L.LEAK);
  ALLOC(v.credential_cache_matrix, int32_t, 2 * L.C);
  ALLOC(v.discovered_nodes_properties, int32_t, L.N * L.nprops);
  ALLOC(v.nodes_privilegelevel, int32_t, L.N);
  const bool dense = cfg->mask_mode == CBX_MASK_DENSE;
  if (dense) {
    ALLOC(v.local_vulnerability, int8_t, L.sz_local);
    ALLOC(v.remote_vulnerability, int8_t, L.sz_remote);
    ALLOC(v.connect, int8_t, L.sz_connect);
  }
  ALLOC(v.owned_bits, uint32_t, L.OW);
  const bool defobs = cfg->mode == CBX_MODE_MARLON && cfg->def_enabled;
  if (defobs) {
    ALLOC(v.def_infected_nodes, int8_t, L.n);
    ALLOC(v.def_incoming_firewall, int8_t, 6 * L.n);
    ALLOC(v.def_outgoing_firewall, int8_t, 6 * L.n);
    ALLOC(v.def_services_status, int8_t, L.nservices);
  }
  {  // rewards and done flags share one block so that the host-buffer step reads them back with a single copy
    uint8_t* blk = nullptr;
    ALLOC(blk, uint8_t, 12);
    const size_t np = (size_t)b->p.n_pad;
    v.att_reward = (float*)blk; v.def_reward = (float*)(blk + 4 * np);
    v.att_terminated = blk + 8 * np; v.att_truncated = blk + 9 * np;
    v.def_terminated = blk + 10 * np; v.def_truncated = blk + 11 * np;
  }
  ALLOC(v.att_info, int32_t, 8);
  ALLOC(v.network_availability, double, 1);
  { void* ps = nullptr; cudaError_t e = dalloc(&ps, CBX_STAT_COUNT * sizeof(double));
    if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "cudaMalloc(stats): %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
    v.episode_stats = (double*)ps; }
  if (cfg->emit_terminal_obs) {
    ALLOC(v.term_scalars, int32_t, 8);
    ALLOC(v.term_leaked_credentials, int32_t, 4 * L.LEAK);
    ALLOC(v.term_credential_cache_matrix, int32_t, 2 * L.C);
    ALLOC(v.term_discovered_nodes_properties, int32_t, L.N * L.nprops);
    ALLOC(v.term_nodes_privilegelevel, int32_t, L.N);
    if (dense) {
      ALLOC(v.term_local_vulnerability, int8_t, L.sz_local);
      ALLOC(v.term_remote_vulnerability, int8_t, L.sz_remote);
      ALLOC(v.term_connect, int8_t, L.sz_connect);
    }
    if (defobs) ALLOC(v.term_def_infected_nodes, int8_t, L.n);
  }
#undef ALLOC
  // bring every env to its initial state (the reference resets before the first step as well)
  {
    b->p.reset_mask = nullptr;
    const int op0 = CBX_OP_RESET | CBX_OP_ATTACKER | CBX_OP_DEFENDER;
    b->p.seq += 1;
    b->p.tickets = b->ticket_ring + (b->p.seq % kTicketRing);
    cudaError_t e = b->p.pipe.enabled ? cbx_launch_pipe(&b->p, op0, b->pipe_grid, 0)
                    : b->p.wide.enabled ? cbx_launch_wide(&b->p, op0, b->wide_grid, 0)
                                        : cbx_launch_step(&b->p, op0, b->grid, b->smem_bytes, b->use_tma, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { int rc2 = fail(CBX_ERR_CUDA, "initial reset: %s", cudaGetErrorString(e)); cbx_batch_destroy(b); return rc2; }
    b->launches++;
  }
  *out = b;
  return CBX_OK;
}

int cbx_batch_destroy(cbx_batch* b) {
  if (!b) return CBX_OK;
  cudaSetDevice(b->device);
  for (void* p : b->allocs) cudaFree(p);
  host_release(b);
  for (cudaEvent_t e : b->ev) cudaEventDestroy(e);
  delete b;
  return CBX_OK;
}

static int timed_launch(cbx_batch* b, int op, cudaStream_t st) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool t = b->timing && !(op & CBX_OP_RESET);
  const bool region = t && b->p.pipe.enabled && b->p.overlap;  // bracket the run of launches instead of each launch
  if (t && !region) {
    if (b->ev_used + 2 > b->ev.size()) {
      for (int k = 0; k < 2; ++k) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); b->ev.push_back(e); }
    }
    e0 = b->ev[b->ev_used]; e1 = b->ev[b->ev_used + 1];
    b->ev_used += 2;
    CUDA_TRY(cudaEventRecord(e0, st));
  }
  if (region && b->region_open != 1) {
    if (b->ev.size() < 2) { for (int k = 0; k < 2; ++k) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); b->ev.push_back(e); } }
    CUDA_TRY(cudaEventRecord(b->ev[0], st));
    b->region_open = 1; b->region_count = 0; b->region_stream = st;
  }
  b->p.seq += 1;
  if (b->p.pipe.enabled) {
    // this launch's ticket counter: the next slot of the ring.  Entering a half, re-zero the OTHER half: a stream operation,
    // so every launch that used those slots has completed and none of the next 512 starts before they are zero again
    const int slot = (int)(b->p.seq % kTicketRing);
    if (slot % (kTicketRing / 2) == 0)
      CUDA_TRY(cudaMemsetAsync(b->ticket_ring + (slot + kTicketRing / 2) % kTicketRing, 0, (kTicketRing / 2) * sizeof(int), st));
    b->p.tickets = b->ticket_ring + slot;
  }
  if (b->p.pipe.enabled) CUDA_TRY(cbx_launch_pipe(&b->p, op, b->pipe_grid, st));
  else if (b->p.wide.enabled) CUDA_TRY(cbx_launch_wide(&b->p, op, b->wide_grid, st));
  else CUDA_TRY(cbx_launch_step(&b->p, op, b->grid, b->smem_bytes, b->use_tma, st));
  if (t && !region) CUDA_TRY(cudaEventRecord(e1, st));
  if (region) b->region_count++;
  b->launches++;
  return CBX_OK;
}

int cbx_batch_reset_ex(cbx_batch* b, const uint8_t* mask_or_null, int who, void* cuda_stream) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  if (!(who & (CBX_WHO_ATTACKER | CBX_WHO_DEFENDER))) return fail(CBX_ERR_INVALID, "reset: nobody selected");
  CUDA_TRY(cudaSetDevice(b->device));
  b->p.reset_mask = mask_or_null;
  b->p.att_actions = nullptr; b->p.def_actions = nullptr; b->p.scan_u = nullptr; b->p.detect_u = nullptr;
  int op = CBX_OP_RESET | ((who & CBX_WHO_ATTACKER) ? CBX_OP_ATTACKER : 0) | ((who & CBX_WHO_DEFENDER) ? CBX_OP_DEFENDER : 0);
  return timed_launch(b, op, (cudaStream_t)cuda_stream);
}

int cbx_batch_notify_reset(cbx_batch* b, const uint8_t* mask_or_null, int who, double last_reward, void* cuda_stream) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  if (!(who & (CBX_WHO_ATTACKER | CBX_WHO_DEFENDER))) return fail(CBX_ERR_INVALID, "notify: nobody selected");
  CUDA_TRY(cudaSetDevice(b->device));
  b->p.reset_mask = mask_or_null;
  b->p.notify_last_reward = (float)last_reward;
  b->p.att_actions = nullptr; b->p.def_actions = nullptr; b->p.scan_u = nullptr; b->p.detect_u = nullptr;
  int op = CBX_OP_RESET | CBX_OP_NOTIFY | ((who & CBX_WHO_ATTACKER) ? CBX_OP_ATTACKER : 0) | ((who & CBX_WHO_DEFENDER) ? CBX_OP_DEFENDER : 0);
  return timed_launch(b, op, (cudaStream_t)cuda_stream);
}

int cbx_batch_reset(cbx_batch* b, const uint8_t* mask_or_null, void* cuda_stream) {
  return cbx_batch_reset_ex(b, mask_or_null, CBX_WHO_ATTACKER | CBX_WHO_DEFENDER, cuda_stream);
}

int cbx_batch_step(cbx_batch* b, const int32_t* att, const int32_t* def, const cbx_tape* tape, void* cuda_stream) {
  return cbx_batch_step_ex(b, att, def, tape, CBX_WHO_ATTACKER | CBX_WHO_DEFENDER, cuda_stream);
}

int cbx_batch_step_i16(cbx_batch* b, const int16_t* att, const int16_t* def, void* cuda_stream) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  b->p.act_i16 = 1;
  const int rc = cbx_batch_step_ex(b, (const int32_t*)att, (const int32_t*)def, nullptr, CBX_WHO_ATTACKER | CBX_WHO_DEFENDER, cuda_stream);
  b->p.act_i16 = 0;  // the launch took its copy of the parameters
  return rc;
}

int cbx_batch_step_ex(cbx_batch* b, const int32_t* att, const int32_t* def, const cbx_tape* tape, int who, void* cuda_stream) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  const cbx_config& c = b->p.cfg;
  if (c.mode != CBX_MODE_MARLON || !c.def_enabled) who &= ~CBX_WHO_DEFENDER;
  if (c.mode != CBX_MODE_MARLON) who |= CBX_WHO_ATTACKER;
  if (!(who & (CBX_WHO_ATTACKER | CBX_WHO_DEFENDER))) return fail(CBX_ERR_INVALID, "step: nobody selected");
  if ((who & CBX_WHO_ATTACKER) && !att) return fail(CBX_ERR_INVALID, "attacker actions required");
  if ((who & CBX_WHO_DEFENDER) && !def) return fail(CBX_ERR_INVALID, "defender actions required (def_enabled)");
  CUDA_TRY(cudaSetDevice(b->device));
  b->p.reset_mask = nullptr;
  b->p.att_actions = att; b->p.def_actions = def;
  b->p.scan_u = tape ? tape->scan_u : nullptr;
  b->p.detect_u = tape ? tape->detect_u : nullptr;
  if (tape && (!tape->scan_u || !tape->detect_u)) return fail(CBX_ERR_INVALID, "tape needs both scan_u and detect_u");
  int op = ((who & CBX_WHO_ATTACKER) ? CBX_OP_ATTACKER : 0) | ((who & CBX_WHO_DEFENDER) ? CBX_OP_DEFENDER : 0);
  return timed_launch(b, op, (cudaStream_t)cuda_stream);
}

static int step_host_impl(cbx_batch* b, const void* h_att, const void* h_def, size_t esz, void* host_out, size_t host_out_bytes,
                          int flags, void* cuda_stream);

int cbx_batch_host_prepare(cbx_batch* b) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  if (b->host_ready) return CBX_OK;
  CUDA_TRY(cudaSetDevice(b->device));
  const size_t n = (size_t)b->p.n_envs;
  cudaError_t e = cudaMallocHost((void**)&b->h_att, n * 10 * 4);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&b->h_def, n * 12 * 4);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&b->h_out, n * 12);
  if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_att, n * 10 * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_def, n * 12 * 4);
  if (e != cudaSuccess) {  // all or nothing: a later call starts over instead of running on half of the buffers
    host_release(b);
    cudaGetLastError();
    return fail(CBX_ERR_CUDA, "host staging buffers: %s", cudaGetErrorString(e));
  }
  b->host_ready = 1;
  return CBX_OK;
}

int cbx_batch_step_host(cbx_batch* b, const int32_t* h_att, const int32_t* h_def, void* host_out, size_t host_out_bytes, void* cuda_stream) {
  return step_host_impl(b, h_att, h_def, 4, host_out, host_out_bytes, 0, cuda_stream);
}

int cbx_batch_step_host_i16(cbx_batch* b, const int16_t* h_att, const int16_t* h_def, void* host_out, size_t host_out_bytes,
                            void* cuda_stream) {
  return step_host_impl(b, h_att, h_def, 2, host_out, host_out_bytes, 0, cuda_stream);
}

int cbx_batch_step_host_ex(cbx_batch* b, const void* h_att, const void* h_def, int elem_bytes, void* host_out, size_t host_out_bytes,
                           int flags, void* cuda_stream) {
  if (elem_bytes != 2 && elem_bytes != 4) return fail(CBX_ERR_INVALID, "elem_bytes must be 2 (int16) or 4 (int32)");
  return step_host_impl(b, h_att, h_def, (size_t)elem_bytes, host_out, host_out_bytes, flags, cuda_stream);
}

static int step_host_impl(cbx_batch* b, const void* h_att, const void* h_def, const size_t esz, void* host_out, size_t host_out_bytes,
                          const int flags, void* cuda_stream) {
  if (!b || !h_att || !host_out) return fail(CBX_ERR_INVALID, "null argument");
  const cbx_config& c = b->p.cfg;
  const int64_t n = b->p.n_envs;
  const int aw = c.mode == CBX_MODE_MARLON ? 10 : 5;
  const bool need_def = c.mode == CBX_MODE_MARLON && c.def_enabled;
  if (need_def && !h_def) return fail(CBX_ERR_INVALID, "defender actions required");
  const size_t out_bytes = (size_t)n * (4 + 4 + 4);
  if (host_out_bytes < out_bytes) return fail(CBX_ERR_INVALID, "host_out needs %zu bytes", out_bytes);
  CUDA_TRY(cudaSetDevice(b->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (!b->host_ready) { const int rc0 = cbx_batch_host_prepare(b); if (rc0) return rc0; }
  // caller buffers that are already page-locked go to the device directly; pageable ones through the pinned staging
  // one query per buffer: page-locked or not, and the address the device sees it at
  auto pinned_view = [](const void* ptr) -> void* {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
  };
  // Page-locked action buffers are read by the step kernel IN PLACE over PCIe (pinned allocations are mapped into the device
  // address space under unified addressing): the transfer of a tile's actions overlaps the other tiles' work instead of
  // preceding the launch.  Pageable buffers go through the library's pinned staging first.  CBX_HOST_ZEROCOPY=0 restores
  // the explicit H2D copies.
  static const bool zero_copy = [] { const char* e = getenv("CBX_HOST_ZEROCOPY"); return !(e && e[0] == '0'); }();
  const bool nosync = flags & CBX_HOST_NOSYNC;
  if (nosync && (!pinned_view(h_att) || (need_def && !pinned_view(h_def)) || !pinned_view(host_out)))
    return fail(CBX_ERR_INVALID, "CBX_HOST_NOSYNC needs page-locked action and result buffers (the call returns before they are used)");
  const void* src_att = h_att;
  const void* k_att = pinned_view(h_att);
  if (!k_att) { memcpy(b->h_att, h_att, (size_t)n * aw * esz); src_att = b->h_att; k_att = pinned_view(b->h_att); }
  const void *src_def = nullptr, *k_def = nullptr;
  if (need_def) {
    src_def = h_def;
    k_def = pinned_view(h_def);
    if (!k_def) { memcpy(b->h_def, h_def, (size_t)n * 12 * esz); src_def = b->h_def; k_def = pinned_view(b->h_def); }
  }
  if (!zero_copy) k_att = k_def = nullptr;
  if (!k_att || (need_def && !k_def)) {  // explicit copies
    CUDA_TRY(cudaMemcpyAsync(b->d_att, src_att, (size_t)n * aw * esz, cudaMemcpyHostToDevice, st));
    if (need_def) CUDA_TRY(cudaMemcpyAsync(b->d_def, src_def, (size_t)n * 12 * esz, cudaMemcpyHostToDevice, st));
    k_att = b->d_att;
    k_def = need_def ? b->d_def : nullptr;
  }
  // With both agents stepping, every env writes all six result arrays each step: the kernel then also writes them straight
  // into the caller's page-locked block (posted PCIe writes, visible after the stream sync) and no D2H copy follows.
  // CBX_HOST_RESULTS=0 restores the copy.
  static const bool mirror_ok = [] { const char* e = getenv("CBX_HOST_RESULTS"); return !(e && e[0] == '0'); }();
  void* out_view = pinned_view(host_out);
  const bool out_pinned = out_view != nullptr;
  uint8_t* mirror = (mirror_ok && zero_copy && need_def) ? (uint8_t*)out_view : nullptr;
  b->p.host_results = mirror;
  b->p.act_i16 = esz == 2;
  int rc = cbx_batch_step(b, (const int32_t*)k_att, (const int32_t*)k_def, nullptr, cuda_stream);
  b->p.host_results = nullptr;
  b->p.act_i16 = 0;
  if (rc) return rc;
  const cbx_views& v = b->p.v;
  uint8_t* o = out_pinned ? (uint8_t*)host_out : b->h_out;
  if (mirror) {
    // nothing to copy
  } else if (b->p.n_pad == n) {  // [att_reward | def_reward | att_terminated | att_truncated | def_terminated | def_truncated] is one block
    CUDA_TRY(cudaMemcpyAsync(o, v.att_reward, out_bytes, cudaMemcpyDeviceToHost, st));
  } else {
    CUDA_TRY(cudaMemcpyAsync(o, v.att_reward, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(o + (size_t)n * 4, v.def_reward, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(o + (size_t)n * 8, v.att_terminated, (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(o + (size_t)n * 9, v.att_truncated, (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(o + (size_t)n * 10, v.def_terminated, (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(o + (size_t)n * 11, v.def_truncated, (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  if (nosync) return CBX_OK;
  CUDA_TRY(cudaStreamSynchronize(st));
  if (!out_pinned) memcpy(host_out, o, out_bytes);
  return CBX_OK;
}

}  // extern "C"

// ---- observation arrays to host memory, packed (cbx_batch_fetch_host) ----
namespace {
struct fetch_item { const void* src; size_t row_bytes; };
void fetch_items(const cbx_batch* b, fetch_item* it) {
  const cbx_views& v = b->p.v;
  const cbx_layout& L = b->p.lay;
  it[0] = {v.scalars, 32};
  it[1] = {v.leaked_credentials, (size_t)16 * L.LEAK};
  it[2] = {v.credential_cache_matrix, (size_t)8 * L.C};
  it[3] = {v.discovered_nodes_properties, (size_t)4 * L.N * L.nprops};
  it[4] = {v.nodes_privilegelevel, (size_t)4 * L.N};
  it[5] = {v.owned_bits, (size_t)4 * L.OW};
  it[6] = {v.local_vulnerability, (size_t)L.sz_local};
  it[7] = {v.remote_vulnerability, (size_t)L.sz_remote};
  it[8] = {v.connect, (size_t)L.sz_connect};
  it[9] = {v.def_infected_nodes, (size_t)L.n};
  it[10] = {v.def_incoming_firewall, (size_t)6 * L.n};
  it[11] = {v.def_outgoing_firewall, (size_t)6 * L.n};
  it[12] = {v.def_services_status, (size_t)L.nservices};
  it[13] = {v.att_reward, 12};
}
}  // namespace

extern "C" int64_t cbx_batch_fetch_host_layout(const cbx_batch* b, uint32_t fields, int64_t* offsets) {
  if (!b) return -1;
  fetch_item it[CBX_F_COUNT];
  fetch_items(b, it);
  int64_t o = 0;
  for (int k = 0; k < CBX_F_COUNT; ++k) {
    const bool on = ((fields >> k) & 1u) && it[k].src && it[k].row_bytes;
    if (offsets) offsets[k] = on ? o : -1;
    if (on) o = (o + (int64_t)it[k].row_bytes * b->p.n_envs + 255) / 256 * 256;
  }
  return o;
}

extern "C" int cbx_batch_fetch_host(cbx_batch* b, uint32_t fields, void* host_out, size_t host_out_bytes, void* cuda_stream) {
  if (!b || !host_out) return fail(CBX_ERR_INVALID, "null argument");
  int64_t off[CBX_F_COUNT];
  const int64_t total = cbx_batch_fetch_host_layout(b, fields, off);
  if ((int64_t)host_out_bytes < total) return fail(CBX_ERR_INVALID, "host_out needs %lld bytes", (long long)total);
  CUDA_TRY(cudaSetDevice(b->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  fetch_item it[CBX_F_COUNT];
  fetch_items(b, it);
  const size_t n = (size_t)b->p.n_envs, np = (size_t)b->p.n_pad;
  uint8_t* o = (uint8_t*)host_out;
  for (int k = 0; k < CBX_F_COUNT; ++k) {
    if (off[k] < 0) continue;
    if (k == 13 && np != n) {  // rewards / flags: six arrays laid out for the padded env count
      const uint8_t* s = (const uint8_t*)it[k].src;
      uint8_t* d = o + off[k];
      CUDA_TRY(cudaMemcpyAsync(d, s, n * 4, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(d + n * 4, s + np * 4, n * 4, cudaMemcpyDeviceToHost, st));
      for (int q = 0; q < 4; ++q) CUDA_TRY(cudaMemcpyAsync(d + n * (8 + q), s + np * (8 + q), n, cudaMemcpyDeviceToHost, st));
    } else {
      CUDA_TRY(cudaMemcpyAsync(o + off[k], it[k].src, it[k].row_bytes * n, cudaMemcpyDeviceToHost, st));
    }
  }
  return CBX_OK;
}

extern "C" {
int cbx_batch_sample_actions(cbx_batch* b, int32_t* att, int32_t* def, uint64_t seed, void* cuda_stream) {
  if (!b || !att) return fail(CBX_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cbx_launch_sample(&b->p, att, def, seed, b->sample_step++, (cudaStream_t)cuda_stream));
  b->launches++;
  return CBX_OK;
}

int cbx_batch_views(cbx_batch* b, cbx_views* out) {
  if (!b || !out) return fail(CBX_ERR_INVALID, "null argument");
  *out = b->p.v;
  return CBX_OK;
}

int cbx_batch_stats_reset(cbx_batch* b, void* cuda_stream) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaMemsetAsync(b->p.v.episode_stats, 0, CBX_STAT_COUNT * sizeof(double), (cudaStream_t)cuda_stream));
  return CBX_OK;
}

int64_t cbx_export_words(const cbx_scenario* s, const cbx_config* cfg) {
  if (!s || !cfg) return -1;
  return CBX_X_HEADER_WORDS + 10 * (int64_t)s->n + cfg->maximum_total_credentials + (s->nsecrets + 31) / 32;
}

int64_t cbx_batch_export_words(const cbx_batch* b) { return b ? cbx_export_words(b->scn, &b->p.cfg) : -1; }

int cbx_batch_export_state(cbx_batch* b, int64_t begin, int64_t end, int32_t* out, void* cuda_stream) {
  if (!b || !out || begin < 0 || end > b->p.n_envs || begin >= end) return fail(CBX_ERR_INVALID, "bad export range");
  CUDA_TRY(cudaSetDevice(b->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const cbx_layout& L = b->p.lay;
  const int64_t m = end - begin;
  // state is a tiled structure of arrays: tile t holds words [S][CBX_TILE] contiguously
  const int64_t t0 = begin / CBX_TILE, t1 = (end + CBX_TILE - 1) / CBX_TILE;
  std::vector<uint32_t> raw((size_t)(t1 - t0) * L.S * CBX_TILE);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaMemcpy(raw.data(), b->p.state + t0 * L.S * CBX_TILE, raw.size() * 4, cudaMemcpyDeviceToHost));
  const int n = L.n;
  const int64_t W = cbx_export_words(b->scn, &b->p.cfg);
  const bool def = b->p.cfg.mode == CBX_MODE_MARLON && b->p.cfg.def_enabled;
  for (int64_t i = 0; i < m; ++i) {
    const int64_t ge = begin + i;
    auto w = [&](int off) { return raw[(size_t)((ge / CBX_TILE - t0) * L.S + off) * CBX_TILE + (size_t)(ge % CBX_TILE)]; };
    auto byte = [&](int off, int k) { return (w(off + k / 4) >> ((k & 3) * 8)) & 0xFFu; };
    auto bit = [&](int off, int k) { return (w(off + k / 32) >> (k & 31)) & 1u; };
    int32_t* x = out + i * W;
    memset(x, 0, (size_t)W * 4);
    const uint32_t hdr = w(L.o_hdr), av = w(L.o_avail);
    x[CBX_X_STEPCOUNT] = (int32_t)w(L.o_stepcount);
    x[CBX_X_DONE] = (hdr >> 24) & 1;
    x[CBX_X_N_DISCOVERED] = hdr & 0xFF;
    x[CBX_X_N_CACHED] = (hdr >> 8) & 0xFFFF;
    x[CBX_X_ATT_TIMESTEPS] = (int32_t)w(L.o_att_ts);
    x[CBX_X_DEF_TIMESTEPS] = (int32_t)w(L.o_def_ts);
    x[CBX_X_ATT_RESET_REQUEST] = (hdr >> 25) & 1;
    x[CBX_X_DEF_RESET_REQUEST] = def ? (hdr >> 26) & 1 : 0;
    x[CBX_X_HAS_BREACHED_SLA] = (hdr >> 27) & 1;
    x[CBX_X_ATT_VALID] = (int32_t)w(L.o_att_valid); x[CBX_X_ATT_INVALID] = (int32_t)w(L.o_att_invalid);
    x[CBX_X_DEF_VALID] = (int32_t)w(L.o_def_valid); x[CBX_X_DEF_INVALID] = (int32_t)w(L.o_def_invalid);
    int live = 0;
    for (int k = 0; k < n; ++k) live += bit(L.o_notrunning, k);
    x[CBX_X_LIVE_IMAGING_COUNT] = live;
    x[CBX_X_SHADOW_IMAGING_COUNT] = (av >> 8) & 0xFF;
    x[CBX_X_PREV_SHADOW_IMAGING_COUNT] = (av >> 16) & 0xFF;
    int32_t* p = x + CBX_X_HEADER_WORDS;
    const int nd = hdr & 0xFF, nc = (hdr >> 8) & 0xFFFF;
    for (int k = 0; k < n; ++k) p[k] = k < nd ? (int32_t)byte(L.o_disc_order, k) : -1;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = bit(L.o_installed, k);
    p += n;
    for (int k = 0; k < n; ++k) p[k] = (w(L.o_priv + k / 16) >> ((k % 16) * 2)) & 3;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = byte(L.o_cd_live, k);
    p += n;
    for (int k = 0; k < n; ++k) p[k] = byte(L.o_fw >= 0 ? L.o_cd_live : L.o_cd_shadow, k);  // the actuator the defender's wrapper is bound to
    p += n;
    for (int k = 0; k < n; ++k) p[k] = bit(L.o_everowned, k);
    p += n;
    for (int k = 0; k < n; ++k) p[k] = (int32_t)w(L.o_props + k * L.PW);
    p += n;
    for (int k = 0; k < n; ++k) p[k] = L.PW > 1 ? (int32_t)w(L.o_props + k * L.PW + 1) : 0;
    p += n;
    for (int k = 0; k < n; ++k) p[k] = (int32_t)w(L.o_attacked + k * L.AW);
    p += n;
    for (int k = 0; k < n; ++k) p[k] = L.o_tags < 0 ? 0 : (int32_t)((w(L.o_tags + k / 8) >> ((k % 8) * 4)) & 15u);
    p += n;
    for (int k = 0; k < L.C; ++k) p[k] = k < nc ? (int32_t)((w(L.o_cache + k / 2) >> ((k & 1) * 16)) & 0xFFFFu) : -1;
    p += L.C;
    for (int k = 0; k < (L.nsecrets + 31) / 32; ++k) p[k] = (int32_t)w(L.o_gathered + k);
  }
  return CBX_OK;
}

int cbx_gae(const float* rewards, const float* values, const uint8_t* episode_starts, const float* last_values, const uint8_t* last_dones,
            double gamma, double gae_lambda, int n_steps, int64_t n_envs, float* advantages, float* returns, void* cuda_stream) {
  if (!rewards || !values || !episode_starts || !last_values || !last_dones || !advantages || !returns) return fail(CBX_ERR_INVALID, "null argument");
  if (n_steps <= 0 || n_envs <= 0) return fail(CBX_ERR_INVALID, "empty rollout");
  CUDA_TRY(cbx_launch_gae(rewards, values, episode_starts, last_values, last_dones, (float)gamma, (float)gae_lambda, n_steps, n_envs,
                          advantages, returns, (cudaStream_t)cuda_stream));
  return CBX_OK;
}

int64_t cbx_batch_launch_count(const cbx_batch* b) { return b ? b->launches : -1; }

int cbx_batch_phase_cycles(cbx_batch* b, int enable, uint64_t* out16) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  CUDA_TRY(cudaSetDevice(b->device));
  if (enable && !b->p.prof) {
    void* pp = nullptr;
    CUDA_TRY(cudaMalloc(&pp, 16 * sizeof(unsigned long long)));
    b->allocs.push_back(pp);
    CUDA_TRY(cudaMemset(pp, 0, 16 * sizeof(unsigned long long)));
    b->p.prof = (unsigned long long*)pp;
  }
  if (out16 && b->p.prof) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out16, b->p.prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemset(b->p.prof, 0, 16 * sizeof(unsigned long long)));
  }
  if (!enable) b->p.prof = nullptr;
  return CBX_OK;
}

int cbx_batch_tile_counter(cbx_batch* b, int32_t* out2) {
  if (!b || !out2) return fail(CBX_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaDeviceSynchronize());
  if (b->p.pipe.enabled) {  // the last launch's own counter: every logic warp drew tickets until one was past the end
    int32_t drawn = 0;
    CUDA_TRY(cudaMemcpy(&drawn, b->ticket_ring + (b->p.seq % kTicketRing), sizeof(int32_t), cudaMemcpyDeviceToHost));
    out2[0] = drawn - (b->p.pipe.dynamic ? b->p.n_tiles : 0);  // exactly n_tiles draws per launch (0 with the static order)
    out2[1] = 0;
  } else {
    CUDA_TRY(cudaMemcpy(out2, b->p.tile_counter, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  return CBX_OK;
}

int cbx_batch_kernel_info(const cbx_batch* b, int32_t* out8) {
  if (!b || !out8) return fail(CBX_ERR_INVALID, "null argument");
  const cbx_pipe_plan& Q = b->p.pipe;
  const bool wide = b->p.wide.enabled;
  out8[0] = Q.enabled ? 1 : wide ? 2 : 0;
  out8[1] = Q.enabled ? b->pipe_grid : wide ? b->wide_grid : b->grid;
  out8[2] = Q.enabled ? (Q.wl + Q.we + (b->p.overlap ? 1 : 0)) * 32 : wide ? b->p.wide.nwarps * 32 : CBX_THREADS;
  out8[3] = Q.enabled ? Q.total_bytes : wide ? b->p.wide.total_bytes : b->smem_bytes;
  out8[4] = Q.enabled ? Q.wl : 0;
  out8[5] = Q.enabled ? Q.we : 0;
  out8[6] = b->p.enc.warp_env;
  out8[7] = b->use_tma | ((Q.enabled ? Q.dynamic : wide ? b->p.wide.dynamic : 0) ? 2 : 0) | ((Q.enabled && b->p.overlap) ? 4 : 0) |
            ((wide ? 0 : b->p.l2_hints & 15) << 4);
  return CBX_OK;
}

int cbx_batch_enable_timing(cbx_batch* b, int enabled) {
  if (!b) return fail(CBX_ERR_INVALID, "null batch");
  if (!enabled && b->region_open == 1) {  // overlapped launches: the bracket ends HERE (an event on the launch stream, no sync)
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaEventRecord(b->ev[1], b->region_stream));
    b->region_open = 2;
  }
  b->timing = enabled;
  return CBX_OK;
}

int cbx_batch_step_kernel_ms(cbx_batch* b, double* mean_ms, int64_t* launches) {
  if (!b || !mean_ms || !launches) return fail(CBX_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(b->device));
  if (b->region_open) {  // overlapped launches: one bracket around the whole run, averaged over its launches
    if (b->region_open == 1) CUDA_TRY(cudaEventRecord(b->ev[1], b->region_stream));
    CUDA_TRY(cudaEventSynchronize(b->ev[1]));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]));
    b->ms_sum += ms;
    b->ms_count += b->region_count;
    b->region_open = 0; b->region_count = 0;
  }
  for (size_t k = 0; k + 1 < b->ev_used; k += 2) {
    CUDA_TRY(cudaEventSynchronize(b->ev[k + 1]));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, b->ev[k], b->ev[k + 1]));
    b->ms_sum += ms;
    b->ms_count++;
  }
  b->ev_used = 0;
  *mean_ms = b->ms_count ? b->ms_sum / (double)b->ms_count : 0.0;
  *launches = b->ms_count;
  b->ms_sum = 0; b->ms_count = 0;
  return CBX_OK;
}

// sizes of the ABI structs, so the Python mirror can be checked without a GPU
size_t cbx_abi_sizeof(int which) { return which == 0 ? sizeof(cbx_config) : which == 1 ? sizeof(cbx_views) : sizeof(cbx_tape); }

}  // extern "C"
