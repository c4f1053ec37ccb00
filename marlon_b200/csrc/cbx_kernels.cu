// cbx_kernels.cu -- the step kernels for sm_100a and what they share.
//
// This file: the FUSED step kernel (general path: any bounds, dense masks of any size), the on-device valid-action sampler,
// the GAE kernel and their launch helpers.  The encoder building blocks, the MARLon wrapper steps and the two game-logic
// phases of one env live in cbx_shared.cuh, shared with
//   cbx_pipe.cu / cbx_pipe.cuh -- pipelined, warp-specialised kernel (small state, dense masks: the bench workload): logic
//                                 warps ahead of encoder warps that hand mask rows to the TMA engine;
//   cbx_wide.cu / cbx_wide.cuh -- warp-per-tile kernel for large per-env state with factored masks (Chain-100, generated
//                                 networks).
// All three produce identical bytes; cbx_batch_create (cbx_api.cu) picks one per batch.
//
// Fused kernel: one persistent CTA per resident slot loops over tiles of 32 envs:
//   (0) TMA bulk copies stage the scenario tables (once per CTA) and the tile's S x 32 state words into shared memory;
//   (1) warp 0 plays the step, one thread per env, on the shared-memory tile (cbx_device.cuh);
//   (2) all 128 threads encode the tile's observations and action masks straight into the output tensors;
//   (3) the state tile is streamed back with a TMA bulk store.
// Algorithmic bytes per env-step and the rooflines are stated in DESIGN.md.
#include "cbx_shared.cuh"

namespace cbx {

// ---- the kernel ----------------------------------------------------------------------------------------------------------
template <bool USE_TMA, int ENC>
__global__ void __launch_bounds__(CBX_THREADS, CBX_MIN_CTAS) cbx_step_kernel(const __grid_constant__ cbx_params p, const int op) {
  extern __shared__ __align__(128) uint32_t smem[];
  const cbx_layout& L = p.lay;
  const cbx_config& cfg = p.cfg;
  uint32_t* s_tb = smem + p.plan.tables;
  uint32_t* s_st = smem + p.plan.state;
  uint32_t* s_sg = smem + p.plan.stage;
  uint32_t* s_desc = smem + p.plan.desc;
  int32_t* s_act = reinterpret_cast<int32_t*>(smem + p.plan.acts);
  uint2* s_lut = reinterpret_cast<uint2*>(smem + p.plan.lut);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.plan.bars);
  __shared__ uint32_t s_masks[4][kGroups];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const bool reset_only = op & CBX_OP_RESET, who_att = op & CBX_OP_ATTACKER, who_def = op & CBX_OP_DEFENDER;
  const bool marlon = cfg.mode == CBX_MODE_MARLON;
  const bool def_on = marlon && cfg.def_enabled && who_def;
  const int DW = p.enc.desc_words;
  const int AW = marlon ? 10 : 5;  // attacker action words per env
  const uint32_t* s_init = s_tb + p.table_words;  // initial per-env state follows the scenario blob
  const uint32_t table_bytes = (uint32_t)(p.table_words + ((L.S + 3) & ~3) + p.fwx_words) * 4u;
  const uint32_t* s_fx = p.fwx_words ? s_init + ((L.S + 3) & ~3) : nullptr;  // live defender binding: firewall extension tables

  // bits -> bytes expansion table: 8 mask bits to 8 bytes of 0/1
  for (int k = tid; k < 256; k += CBX_THREADS) {
    uint32_t lo = ((k & 0xF) * 0x00204081u) & 0x01010101u, hi = (((k >> 4) & 0xF) * 0x00204081u) & 0x01010101u;
    s_lut[k] = make_uint2(lo, hi);
  }
  // tiles of this CTA: strided for a single scenario (balance); one contiguous range for a multi-scenario batch, so that the
  // staged scenario tables change rarely (envs are grouped by scenario, cbx_batch_create_multi)
  const bool multi = p.n_scenarios > 1;
  const int tile_begin = multi ? (int)((int64_t)blockIdx.x * p.n_tiles / gridDim.x) : (int)blockIdx.x;
  const int tile_end = multi ? (int)((int64_t)(blockIdx.x + 1) * p.n_tiles / gridDim.x) : p.n_tiles;
  const int tile_step = multi ? 1 : (int)gridDim.x;
  int cur_scn = (multi && tile_begin < tile_end) ? p.tile_scn[tile_begin] : 0;
  uint32_t tb_phase = 0;
  auto stage_tables = [&](int scn) {  // all threads
    const uint32_t* src = p.tables + (size_t)scn * p.table_stride;
    if (USE_TMA) {
      if (tid == 0) {
        mbar_expect_tx(&bars[0], table_bytes);
        tma_load_1d(s_tb, src, table_bytes, &bars[0]);
      }
      mbar_wait(&bars[0], tb_phase);
      tb_phase ^= 1;
    } else {
      for (uint32_t k = tid; k < table_bytes / 4; k += CBX_THREADS) s_tb[k] = src[k];
      __syncthreads();
    }
  };
  if (USE_TMA) {
    if (tid == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  stage_tables(cur_scn);

  Acc acc;
#pragma unroll
  for (int k = 0; k < CBX_STAT_COUNT; ++k) acc.v[k] = 0.0;
  long long prof_t = p.prof ? clock64() : 0;
#define CBX_PROF(slot)                                                                         \
  if (p.prof && tid == 0) {                                                                    \
    long long _now = clock64();                                                                \
    atomicAdd(p.prof + (slot), (unsigned long long)(_now - prof_t));                           \
    prof_t = _now;                                                                             \
  }
  CBX_PROF(0)  // prologue: LUT + tables
  uint32_t st_phase = 0;
  int slice_of_kind[3] = {p.slice_of_kind[0], p.slice_of_kind[1], p.slice_of_kind[2]};
  constexpr uint32_t kRowBytes = CBX_TILE * 4u;

  for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
    const int64_t e0 = (int64_t)tile * CBX_TILE;
    const int n_valid = (int)min((int64_t)CBX_TILE, p.n_envs - e0);
    if (multi && p.tile_scn[tile] != cur_scn) {  // next scenario: restage its tables (every thread is past the last tile)
      cur_scn = p.tile_scn[tile];
      __syncthreads();
      stage_tables(cur_scn);
    }
    // ---- (0) stage the state tile (S rows of CBX_TILE words) and the tile's actions ----
    if (USE_TMA) {
      if (tid < 32) {
        if (tid == 0) {  // the tile's state is one contiguous block (tiled structure of arrays): a single bulk copy
          mbar_expect_tx(&bars[1], (uint32_t)L.S * kRowBytes);
          tma_load_1d_pol(s_st, p.state + (int64_t)tile * L.S * CBX_TILE, (uint32_t)L.S * kRowBytes, &bars[1], (p.l2_hints & 1) != 0,
                          l2_policy_evict_last());
        }
      }
    } else {
      for (int k = tid; k < L.S * CBX_TILE; k += CBX_THREADS)
        s_st[k] = p.state[(int64_t)tile * L.S * CBX_TILE + k];
    }
    if (!reset_only) {  // coalesced: the tile's actions are contiguous in the [n, width] action arrays
      if (p.att_actions && (who_att || !marlon))
        for (int k = tid; k < n_valid * AW; k += CBX_THREADS) s_act[k] = load_act(p.att_actions, e0 * AW + k, p.act_i16);
      if (def_on)
        for (int k = tid; k < n_valid * 12; k += CBX_THREADS) s_act[CBX_TILE * 10 + k] = load_act(p.def_actions, e0 * 12 + k, p.act_i16);
    }
    if (USE_TMA) {
      mbar_wait(&bars[1], st_phase);
      st_phase ^= 1;
    }
    __syncthreads();

    CBX_PROF(1)  // state tile + action load
    const bool active = tid < n_valid;
    Ctx c;
    c.st = s_st + tid; c.sg = s_sg + tid; c.tb = s_tb; c.L = &L; c.cfg = &cfg; c.env = e0 + tid; c.fx = s_fx;

    // ---- (1) game logic, phase 1: attacker move (or reset); one thread per env ----
    if (tid < CBX_TILE) {
      uint32_t att_done = 0, keep = 1;
      if (active) {
        logic_phase1(c, p, op, s_act + tid * AW, s_init, slice_of_kind, acc);
        att_done = c.g(STG_ATT_DONE);
        keep = c.g(STG_OBS_KIND) == OBS_KEEP;
      }
      uint32_t done_mask = __ballot_sync(0xFFFFFFFFu, att_done != 0);
      uint32_t keep_mask = __ballot_sync(0xFFFFFFFFu, keep != 0);
      if ((tid & 31) == 0) { s_masks[0][warp] = done_mask; s_masks[1][warp] = keep_mask; }
    }
    __syncthreads();
    CBX_PROF(2)  // logic phase 1 (attacker move)
    EnvMask att_done_mask, keep1;
#pragma unroll
    for (int k = 0; k < kGroups; ++k) { att_done_mask.w[k] = s_masks[0][k]; keep1.w[k] = s_masks[1][k]; }
    Tile t;
    t.L = &L; t.tb = s_tb; t.st = s_st; t.sg = s_sg; t.desc = s_desc; t.lut = s_lut; t.K = &p.enc; t.DW = DW;
    t.fx = s_fx; t.init = s_init;

    // ---- (1b) terminal observations of the envs that finished (rare): encode BEFORE the auto-reset ----
    if (att_done_mask.any() && cfg.auto_reset && cfg.emit_terminal_obs && !reset_only) {
      if (tid < CBX_TILE && active && att_done_mask.test(tid)) build_desc(c, s_desc + tid * DW, DW, nullptr);
      __syncthreads();
      Target tt = make_target(p.v, L, e0, true);
      encode_attacker<ENC>(t, tt, n_valid, mask_and_not(att_done_mask, keep1));
      const EnvMask cp = mask_and(att_done_mask, keep1);
      if (cp.any()) {
        Target tm = make_target(p.v, L, e0, false);
        copy_rows(tt.scalars, tm.scalars, 32, n_valid, cp);
        copy_rows(tt.leaked, tm.leaked, 16 * L.LEAK, n_valid, cp);
        copy_rows(tt.cachem, tm.cachem, 8 * L.C, n_valid, cp);
        copy_rows(tt.props, tm.props, 4 * L.N * L.nprops, n_valid, cp);
        copy_rows(tt.priv, tm.priv, 4 * L.N, n_valid, cp);
        copy_rows(tt.local, tm.local, L.sz_local, n_valid, cp);
        copy_rows(tt.remote, tm.remote, L.sz_remote, n_valid, cp);
        copy_rows(tt.connect, tm.connect, L.sz_connect, n_valid, cp);
      }
      __syncthreads();
    }

    CBX_PROF(3)  // terminal observations
    // ---- (2) game logic, phase 2: attacker auto-reset, defender move, encoder descriptors ----
    if (tid < CBX_TILE) {
      uint32_t def_done = 0, keep = 1;
      if (active) {
        def_done = logic_phase2(c, p, op, s_act + CBX_TILE * 10 + tid * 12, s_init, s_desc + tid * DW, acc);
        keep = c.g(STG_OBS_KIND) == OBS_KEEP;
      }
      uint32_t dmask = __ballot_sync(0xFFFFFFFFu, def_done != 0);
      uint32_t kmask = __ballot_sync(0xFFFFFFFFu, keep != 0);
      if ((tid & 31) == 0) { s_masks[2][warp] = dmask; s_masks[3][warp] = kmask; }
    }
    __syncthreads();
    CBX_PROF(4)  // logic phase 2 (auto-reset, defender move, descriptors)
    EnvMask def_done_mask, keep2;
#pragma unroll
    for (int k = 0; k < kGroups; ++k) { def_done_mask.w[k] = s_masks[2][k]; keep2.w[k] = s_masks[3][k]; }
    const EnvMask enc_mask = mask_not(keep2);

    // ---- (3) encode ----
    {
      Target tm = make_target(p.v, L, e0, false);
      encode_attacker<ENC>(t, tm, n_valid, enc_mask);
      if (marlon && cfg.def_enabled && who_def && !(op & CBX_OP_NOTIFY)) {
        encode_defender<ENC>(t, tm, n_valid, mask_all(), true);
        if (def_done_mask.any() && cfg.emit_terminal_obs) {
          // terminal defender observation = infected nodes seen by the step that ended the episode
          __syncthreads();
          if (tid < CBX_TILE && active && def_done_mask.test(tid))
            for (int k = 0; k < L.Wn; ++k) s_desc[tid * DW + D_OWNED + L.OW + k] = c.g(STG_DEF_TERM_INST + k);
          __syncthreads();
          Target tt = make_target(p.v, L, e0, true);
          encode_defender<ENC>(t, tt, n_valid, def_done_mask, false);
        }
      }
    }
    __syncthreads();

    CBX_PROF(5)  // encode
    // ---- (4) deferred defender auto-reset (DummyVecEnv resets after the step; the attacker's observation of this
    //          step was taken before it), then stream the state tile back ----
    if (tid < CBX_TILE && active && def_done_mask.test(tid)) c.defender_reset(s_init);
    if (USE_TMA) {
      fence_async_smem();
      __syncthreads();
      if (tid < 32) {
        if (tid == 0)
          tma_store_1d_pol(p.state + (int64_t)tile * L.S * CBX_TILE, s_st, (uint32_t)L.S * kRowBytes, (p.l2_hints & 1) != 0, l2_policy_evict_last());
        tma_store_commit();
        tma_store_wait_read();  // the tile buffer is reused by the next iteration
      }
      __syncthreads();
    } else {
      __syncthreads();
      for (int k = tid; k < L.S * CBX_TILE; k += CBX_THREADS)
        p.state[(int64_t)tile * L.S * CBX_TILE + k] = s_st[k];
      __syncthreads();
    }
    CBX_PROF(6)  // state write-back
  }
  if (USE_TMA && tid < 32) tma_store_wait_all();

  // ---- episode statistics: warp shuffle reduce, one atomic per slot per warp ----
  if (tid < CBX_TILE) {
#pragma unroll
    for (int k = 0; k < CBX_STAT_COUNT; ++k) {
      double x = acc.v[k];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, o);
      if ((tid & 31) == 0 && x != 0.0) atomicAdd(p.v.episode_stats + k, x);
    }
  }
}

}  // namespace cbx

namespace cbx {

// ---- valid actions with CyberBattleEnv.sample_valid_action's distribution (ENV:959-1047) ------------------------------------
// The reference draws whole proposals until the action mask admits one (ENV:1041-1047): kind uniform over [0, 1, 2] (no 2 while
// the credential cache is empty, ENV:972-976); quirk B.9: kind 1 builds a LOCAL action, kind 0 a REMOTE one; sources uniform
// over the nodes with privilege >= LocalUser (ENV:832-838), targets over the discovered nodes, vulnerability / port ids over the
// whole id range, credentials over the cache.  The mask wants agent_installed on the source and, for a local action, the
// vulnerability on that node (ENV:643-677): rejected proposals are redrawn kind and all, so local actions come out rarer than
// 1 / kinds.  Draws: Philox4x32-10 keyed (seed; env, step, 0x5A170000 + attempt); the oracle's orc_sample_actions does the
// same arithmetic (tests compare the two bit for bit and the oracle's frequencies with the live reference's).
// One thread per env reads its state words straight from HBM (column access is coalesced across the warp).
__global__ void cbx_sample_kernel(const __grid_constant__ cbx_params p, int32_t* att, int32_t* def, uint64_t seed, uint32_t step) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= p.n_envs) return;
  const cbx_layout& L = p.lay;
  const uint32_t* tables = p.tables + (size_t)(p.tile_scn ? p.tile_scn[env / CBX_TILE] : 0) * p.table_stride;
  auto W = [&](int off) { return p.state[((env / CBX_TILE) * L.S + off) * CBX_TILE + (env % CBX_TILE)]; };
  const uint32_t hdr = W(L.o_hdr);
  const int nd = hdr & 0xFF, nc = (hdr >> 8) & 0xFFFF;
  auto node_at = [&](int s) { return (int)((W(L.o_disc_order + (s >> 2)) >> ((s & 3) * 8)) & 0xFFu); };
  // discovery indices of the owned nodes (privilege >= LocalUser), in discovery order
  int owned[256];
  int n_owned = 0;
  for (int s = 0; s < nd; ++s) {
    const int node = node_at(s);
    if ((W(L.o_priv + (node >> 4)) >> ((node & 15) * 2)) & 3u) owned[n_owned++] = s;
  }
  int kind = CBX_KIND_REMOTE;
  int a[4] = {0, 0, 0, 0};
  for (uint32_t it = 0; it < 64u; ++it) {
    uint32_t r[4];
    philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), step, 0x5A170000u + it, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const uint32_t k = __umulhi(r[0], nc > 0 ? 3u : 2u);
    const int src = n_owned ? owned[__umulhi(r[1], (uint32_t)n_owned)] : 0;
    const int node = nd ? node_at(src) : 0;
    bool valid = n_owned && ((W(L.o_installed + (node >> 5)) >> (node & 31)) & 1u);
    a[0] = src; a[1] = a[2] = a[3] = 0;
    if (k == 1u) {
      kind = CBX_KIND_LOCAL;
      a[1] = (int)__umulhi(r[2], (uint32_t)L.L);
      valid = valid && (tables[tables[CBX_H_OFF_VULN] + (node * (L.L + L.R) + a[1]) * CBX_VULN_WORDS] & 1u);
    } else if (k == 0u) {
      kind = CBX_KIND_REMOTE;
      a[1] = (int)__umulhi(r[2], (uint32_t)max(nd, 1));
      a[2] = (int)__umulhi(r[3], (uint32_t)L.R);
    } else {
      uint32_t q[4];
      philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), step, 0x5A180000u + it, (uint32_t)seed, (uint32_t)(seed >> 32), q);
      kind = CBX_KIND_CONNECT;
      a[1] = (int)__umulhi(r[2], (uint32_t)max(nd, 1));
      a[2] = (int)__umulhi(r[3], (uint32_t)L.P);
      a[3] = (int)__umulhi(q[0], (uint32_t)nc);
    }
    if (valid) break;
  }
  uint32_t r2[4] = {0, 0, 0, 0};
  if (def) philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), step, 0x5A190000u, (uint32_t)seed, (uint32_t)(seed >> 32), r2);
  if (p.cfg.mode == CBX_MODE_MARLON) {
    int32_t* o = att + env * 10;
    for (int k = 0; k < 10; ++k) o[k] = 0;
    int idx = 0;
    for (int k = 0; k < 3; ++k)
      if (p.cfg.kind_of_index[k] == kind) idx = k;
    o[0] = idx;
    const int width = kind == CBX_KIND_LOCAL ? 2 : kind == CBX_KIND_REMOTE ? 3 : 4;
    for (int k = 0; k < width; ++k) o[p.slice_of_kind[kind] + k] = a[k];
    if (def) {
      int32_t* d = def + env * 12;
      const uint32_t n = tables[CBX_H_N_NODES];  // the scenario's own node count
      d[0] = (int)__umulhi(r2[1], 5u);
      uint32_t x = r2[2], y = r2[3];
      d[1] = (int)__umulhi(x, n); x = x * 1664525u + 1013904223u;
      d[2] = (int)__umulhi(x, n); x = x * 1664525u + 1013904223u;
      d[3] = (int)__umulhi(x, 6u); x = x * 1664525u + 1013904223u;
      d[4] = (int)__umulhi(x, 2u); x = x * 1664525u + 1013904223u;
      d[5] = (int)__umulhi(x, n); x = x * 1664525u + 1013904223u;
      d[6] = (int)__umulhi(y, 6u); y = y * 1664525u + 1013904223u;
      d[7] = (int)__umulhi(y, 2u); y = y * 1664525u + 1013904223u;
      d[8] = (int)__umulhi(y, n); y = y * 1664525u + 1013904223u;
      d[9] = (int)__umulhi(y, 3u); y = y * 1664525u + 1013904223u;
      d[10] = (int)__umulhi(y, n); y = y * 1664525u + 1013904223u;
      d[11] = (int)__umulhi(y, 3u);
    }
  } else {
    int32_t* o = att + env * 5;
    o[0] = kind; o[1] = a[0]; o[2] = a[1]; o[3] = a[2]; o[4] = a[3];
  }
}

}  // namespace cbx

namespace cbx {
// ---- generalised advantage estimation over a device-resident rollout --------------------------------
// stable-baselines3 RolloutBuffer.compute_returns_and_advantage, the routine MARLon's on_rollout_end reaches through
// baseline_marlon_agent.py:276-284: a backward scan per env over [T, n] arrays (one thread per env, coalesced rows).
__global__ void cbx_gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ episode_starts,
                               const float* __restrict__ last_values, const uint8_t* __restrict__ last_dones, const float gamma,
                               const float lam, const int T, const int64_t n, float* __restrict__ advantages, float* __restrict__ returns) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float next_value = last_values[e];
  float next_non_terminal = 1.0f - (float)(last_dones[e] != 0);
  float gae = 0.0f;
  for (int t = T - 1; t >= 0; --t) {
    const int64_t i = (int64_t)t * n + e;
    const float v = values[i];
    const float delta = rewards[i] + gamma * next_value * next_non_terminal - v;
    gae = delta + gamma * lam * next_non_terminal * gae;
    advantages[i] = gae;
    returns[i] = gae + v;
    next_value = v;
    next_non_terminal = 1.0f - (float)(episode_starts[i] != 0);
  }
}
}  // namespace cbx

// ---- launch helpers used by cbx_api.cu ------------------------------------------------------------------------------------
template <bool A, int B>
static cudaError_t attrs_of(int smem_bytes, int* blocks_per_sm) {
  cudaError_t e = cudaFuncSetAttribute(cbx::cbx_step_kernel<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, cbx::cbx_step_kernel<A, B>, CBX_THREADS, smem_bytes);
}
extern "C" {
// enc: encoder variant (cbx_enc_consts.warp_env): 0 generic, 1 warp-per-env, 2 ToyCtf(12,10) static, 3 Chain-10(12,12) static.
// The statically specialised variants exist with TMA staging only.
cudaError_t cbx_launch_step(const cbx_params* p, int op, int grid, int smem_bytes, int use_tma, cudaStream_t stream) {
  const int enc = p->enc.warp_env;
  if (use_tma) {
    switch (enc) {
      case 3: cbx::cbx_step_kernel<true, 3><<<grid, CBX_THREADS, smem_bytes, stream>>>(*p, op); break;
      case 2: cbx::cbx_step_kernel<true, 2><<<grid, CBX_THREADS, smem_bytes, stream>>>(*p, op); break;
      case 1: cbx::cbx_step_kernel<true, 1><<<grid, CBX_THREADS, smem_bytes, stream>>>(*p, op); break;
      default: cbx::cbx_step_kernel<true, 0><<<grid, CBX_THREADS, smem_bytes, stream>>>(*p, op); break;
    }
  } else {
#ifdef CBX_EXPERIMENTS  // plain-copy staging (CBX_NO_TMA=1) is a debugging aid of experiment builds only
    if (enc >= 1) cbx::cbx_step_kernel<false, 1><<<grid, CBX_THREADS, smem_bytes, stream>>>(*p, op);
    else cbx::cbx_step_kernel<false, 0><<<grid, CBX_THREADS, smem_bytes, stream>>>(*p, op);
#else
    return cudaErrorNotSupported;
#endif
  }
  return cudaGetLastError();
}
cudaError_t cbx_launch_gae(const float* rewards, const float* values, const uint8_t* episode_starts, const float* last_values,
                           const uint8_t* last_dones, float gamma, float lam, int T, int64_t n, float* advantages, float* returns,
                           cudaStream_t stream) {
  const int threads = 128;
  cbx::cbx_gae_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(rewards, values, episode_starts, last_values, last_dones,
                                                                                     gamma, lam, T, n, advantages, returns);
  return cudaGetLastError();
}
cudaError_t cbx_launch_sample(const cbx_params* p, int32_t* att, int32_t* def, uint64_t seed, uint32_t step, cudaStream_t stream) {
  const int threads = 128;
  const int grid = (int)((p->n_envs + threads - 1) / threads);
  cbx::cbx_sample_kernel<<<grid, threads, 0, stream>>>(*p, att, def, seed, step);
  return cudaGetLastError();
}
cudaError_t cbx_kernel_attrs(int smem_bytes, int use_tma, int enc, int* blocks_per_sm) {
  if (use_tma) {
    switch (enc) {
      case 3: return attrs_of<true, 3>(smem_bytes, blocks_per_sm);
      case 2: return attrs_of<true, 2>(smem_bytes, blocks_per_sm);
      case 1: return attrs_of<true, 1>(smem_bytes, blocks_per_sm);
      default: return attrs_of<true, 0>(smem_bytes, blocks_per_sm);
    }
  }
#ifdef CBX_EXPERIMENTS
  if (enc >= 1) return attrs_of<false, 1>(smem_bytes, blocks_per_sm);
  return attrs_of<false, 0>(smem_bytes, blocks_per_sm);
#else
  return cudaErrorNotSupported;
#endif
}
}
