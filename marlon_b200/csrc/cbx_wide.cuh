// cbx_wide.cuh -- warp-per-tile step kernel for LARGE per-env state (Chain-100, generated networks); included by cbx_kernels.cu.
//
// Why: with 400 state words per env a staged 32-env state tile is 51 KB of shared memory, so the fused kernel fits two CTAs
// (two game-logic warps) per SM, and the pipelined kernel's per-tile field images (225 KB at Chain-100) do not fit at all:
// config 4 ran at 26 % of the HBM roofline.  The game logic touches only a few dozen of those 400 words per step.
//
// Here nothing big is staged:
//   * the per-env state stays in HBM / L2 and the one-thread-per-env game logic works on it IN PLACE: word w of the 32 envs
//     of a tile is one 128-byte line (tiled structure of arrays), so every access of a warp is one coalesced transaction;
//   * the scenario tables are read through L1 from global memory (30 KB at Chain-100, hot), which also makes a batch over
//     several scenarios (cbx_batch_create_multi) free: a tile just points at its scenario's tables;
//   * each warp owns a tile end to end -- no CTA-wide barrier in the loop, 12 tiles in flight per SM instead of 2;
//   * the int32 observation fields are produced by the env's own thread, 32 words at a time, into a padded 32 x 33 shared-
//     memory square and leave transposed: one fully coalesced 128-byte store per env and chunk (0.1 instructions per word
//     where the element-wise encoder needed ~20).
// Factored masks only (what these configurations use: a dense Chain-100 connect mask is 8.5 MB per env).
namespace cbx {

constexpr int kImgStride = 33;  // padded row of the transpose square: conflict-free both ways

// ---- thread-per-env word generators (same values as encode_attacker / build_field_images) -------------------------------
struct GenStaging {  // k-th staging word from `base`
  const Ctx& c; int base;
  __device__ __forceinline__ uint32_t operator()(int k) { return c.g(base + k); }
};
struct GenCache {  // credential_cache_matrix [C][2]: (target discovery index, port) of every cached credential (ENV:920-922)
  const Ctx& c; int nc; bool blank; uint32_t a, b;
  __device__ __forceinline__ uint32_t operator()(int k) {
    if (!(k & 1)) {
      a = b = 0;
      const int i = k >> 1;
      if (!blank && i < nc) {
        const uint32_t* rec = c.triple((int)c.half(c.L->o_cache, i));
        a = c.byte(c.L->o_disc_idx, (int)rec[0]);
        b = rec[1];
      }
      return a;
    }
    return b;
  }
};
struct GenProps {  // discovered_nodes_properties [N][props] (ENV:811-830; 2 = unknown only in blank observations, ENV:765)
  const Ctx& c; int nd, nprops; bool blank; int kk, pi; uint32_t lo, hi;
  __device__ __forceinline__ uint32_t operator()(int) {
    if (pi == 0) {
      lo = hi = 0;
      if (!blank && kk < nd) {
        const int node = (int)c.byte(c.L->o_disc_order, kk);
        lo = c.w(c.L->o_props + node * c.L->PW);
        if (c.L->PW > 1) hi = c.w(c.L->o_props + node * c.L->PW + 1);
      }
    }
    const uint32_t v = blank ? 2u : (((pi < 32 ? lo : hi) >> (pi & 31)) & 1u);
    if (++pi == nprops) { pi = 0; ++kk; }
    return v;
  }
};
struct GenPriv {  // nodes_privilegelevel [N] in discovery order (ENV:840-857), as the observation saw it (staging snapshot)
  const Ctx& c; int nd; bool blank;
  __device__ __forceinline__ uint32_t operator()(int k) {
    if (blank || k >= nd) return 0u;
    const uint32_t node = c.byte(c.L->o_disc_order, k);
    return (c.g(c.L->g_priv + (node >> 4)) >> ((node & 15) * 2)) & 3u;
  }
};

// One int32 field of a tile: every thread generates ITS env's words 32 at a time into the padded square, then the warp
// stores the square transposed -- env e's 32 words are 128 contiguous bytes of dst.
template <class Gen>
__device__ __forceinline__ void emit_field_rows(int32_t* dst, const int wpe, const int n_valid, const uint32_t mask, uint32_t* img,
                                                const int lane, const bool mine, Gen gen) {
  if (!dst) return;
  for (int k0 = 0; k0 < wpe; k0 += 32) {
    const int m = min(32, wpe - k0);
    if (mine)
      for (int j = 0; j < m; ++j) img[lane * kImgStride + j] = gen(k0 + j);
    __syncwarp();
    if (lane < m) {
      int32_t* d = dst + k0 + lane;
#pragma unroll 4
      for (int e = 0; e < n_valid; ++e)
        if ((mask >> e) & 1u) d[(size_t)e * wpe] = (int32_t)img[e * kImgStride + lane];
    }
    __syncwarp();
  }
}

// Static rows of the defender observation for ONE env of the scenario `tb`: [6n incoming | 6n outgoing | nsvc services],
// followed by the first 4 bytes again (word reads wrap around the end of a row).  n / nsvc: layout sizes (zero padded).
__device__ __forceinline__ void build_defender_rows(const uint32_t* tb, uint8_t* rows, const int n, const int nsvc, const int lane) {
  const int n_own = (int)tb[CBX_H_N_NODES], nsvc_own = (int)tb[CBX_H_N_SERVICES];
  const int n6 = 6 * n;
  uint8_t* rin = rows;
  uint8_t* rout = rows + n6 + 4;
  uint8_t* rsvc = rout + n6 + 4;
  for (int i = lane; i < n6 + 4; i += 32) {
    const int q = i < n6 ? i : i - n6;
    const int node = q / 6, r = q - node * 6;
    const uint32_t dob = node < n_own ? tb[tb[CBX_H_OFF_NODE] + node * CBX_NODE_WORDS + CBX_N_DEFOBS] : 0u;
    rin[i] = (uint8_t)((dob >> r) & 1u);
    rout[i] = (uint8_t)((dob >> (8 + r)) & 1u);
  }
  for (int i = lane; i < nsvc + 4; i += 32) rsvc[i] = (uint8_t)((nsvc > 0 ? (i < nsvc ? i : i - nsvc) : 0) < nsvc_own);
}

// a tile's worth (32 envs) of a byte field whose rows all equal `row` (period `len` bytes): coalesced word stores
__device__ __forceinline__ void emit_periodic_bytes(int8_t* dst, const uint8_t* row, const int len, const int lane) {
  if (!dst || len == 0) return;
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);  // 32 * len bytes from a 128-byte aligned tile base: whole words
  const int words = CBX_TILE * len / 4;
  int i = (lane * 4) % len;              // byte offset within the row of this lane's first word
  const int step = 128 % len;            // advance per iteration (32 lanes x 4 bytes)
  for (int w = lane; w < words; w += 32) {
    const uint32_t v = (uint32_t)row[i] | ((uint32_t)row[i + 1 < len ? i + 1 : i + 1 - len] << 8) |
                       ((uint32_t)row[i + 2 < len ? i + 2 : i + 2 - len] << 16) | ((uint32_t)row[i + 3 < len ? i + 3 : i + 3 - len] << 24);
    d[w] = v;
    i += step;
    if (i >= len) i -= len;
  }
}

template <int ENC>
__global__ void __launch_bounds__(CBX_WIDE_WARPS * 32, 1) cbx_wide_kernel(const __grid_constant__ cbx_params p, const int op) {
  extern __shared__ __align__(128) uint32_t smem[];
  const cbx_layout& L = p.lay;
  const cbx_config& cfg = p.cfg;
  const cbx_wide_plan& Q = p.wide;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool reset_only = op & CBX_OP_RESET, who_def = op & CBX_OP_DEFENDER, who_att = op & CBX_OP_ATTACKER;
  const bool marlon = cfg.mode == CBX_MODE_MARLON;
  const bool def_on = marlon && cfg.def_enabled && who_def;
  const bool def_encode = def_on && !(op & CBX_OP_NOTIFY);
  const int DW = p.enc.desc_words;
  const int AW = marlon ? 10 : 5;
  uint2* s_lut = reinterpret_cast<uint2*>(smem + Q.lut);
  uint32_t* wb = smem + Q.warps + warp * Q.warp_words;  // this warp's private area
  uint32_t* sg = wb + Q.w_stage;
  uint32_t* desc = wb + Q.w_desc;
  int32_t* act = reinterpret_cast<int32_t*>(wb + Q.w_acts);
  uint32_t* img = wb + Q.w_img;
  uint8_t* drows = reinterpret_cast<uint8_t*>(wb + Q.w_drows);
  for (int k = tid; k < 256; k += (int)blockDim.x) {
    uint32_t lo = ((k & 0xF) * 0x00204081u) & 0x01010101u, hi = (((k >> 4) & 0xF) * 0x00204081u) & 0x01010101u;
    s_lut[k] = make_uint2(lo, hi);
  }
  __syncthreads();  // the only CTA-wide barrier

  Acc acc;
#pragma unroll
  for (int k = 0; k < CBX_STAT_COUNT; ++k) acc.v[k] = 0.0;
  int slice_of_kind[3] = {p.slice_of_kind[0], p.slice_of_kind[1], p.slice_of_kind[2]};
  int rows_scn = -1;  // scenario the defender rows in `drows` were built for
  const int n6 = 6 * L.n;
  const int nw = (int)blockDim.x >> 5;  // warps per CTA (<= CBX_WIDE_WARPS)
  const int gw = (int)gridDim.x * nw;
  for (int tile = (int)blockIdx.x * nw + warp; tile < p.n_tiles; tile += gw) {
    const int64_t e0 = (int64_t)tile * CBX_TILE;
    const int n_valid = (int)min((int64_t)CBX_TILE, p.n_envs - e0);
    const int scn = p.tile_scn ? p.tile_scn[tile] : 0;
    const uint32_t* tb = p.tables + (size_t)scn * p.table_stride;  // global memory, L1-resident
    const uint32_t* s_init = tb + p.table_words;
    uint32_t* gst = p.state + (int64_t)tile * L.S * CBX_TILE;        // the tile's state, in place
    if (!reset_only) {
      if (p.att_actions && (who_att || !marlon))
        for (int q = lane; q < n_valid * AW; q += 32) act[q] = p.att_actions[e0 * AW + q];
      if (def_on)
        for (int q = lane; q < n_valid * 12; q += 32) act[CBX_TILE * 10 + q] = p.def_actions[e0 * 12 + q];
    }
    __syncwarp();
    const bool active = lane < n_valid;
    Ctx c;
    c.st = gst + lane; c.sg = sg + lane; c.tb = tb; c.L = &L; c.cfg = &cfg; c.env = e0 + lane;
    uint32_t att_done = 0, keep = 1;
    if (active) {
      logic_phase1(c, p, op, act + lane * AW, s_init, slice_of_kind, acc);
      att_done = c.g(STG_ATT_DONE);
      keep = c.g(STG_OBS_KIND) == OBS_KEEP;
    }
    const uint32_t att_done_mask = __ballot_sync(0xFFFFFFFFu, att_done != 0);
    const uint32_t keep1 = __ballot_sync(0xFFFFFFFFu, keep != 0);
    Tile t;
    t.L = &L; t.tb = tb; t.st = gst; t.sg = sg; t.desc = desc; t.lut = s_lut; t.K = &p.enc; t.DW = DW;
    // terminal observations of the envs that finished: BEFORE the auto-reset (rare; element-wise encoder)
    if (att_done_mask && cfg.auto_reset && cfg.emit_terminal_obs && !reset_only) {
      if (active && ((att_done_mask >> lane) & 1u)) build_desc(c, desc + lane * DW, DW, nullptr);
      __syncwarp();
      Target tt = make_target(p.v, L, e0, true);
      EnvMask m_enc, m_cp;
#pragma unroll
      for (int q = 0; q < kGroups; ++q) { m_enc.w[q] = 0; m_cp.w[q] = 0; }
      m_enc.w[0] = att_done_mask & ~keep1;
      m_cp.w[0] = att_done_mask & keep1;
      encode_attacker<ENC>(t, tt, n_valid, m_enc, 0, 1);
      if (m_cp.any()) {
        Target tm = make_target(p.v, L, e0, false);
        copy_rows(tt.scalars, tm.scalars, 32, n_valid, m_cp, lane, 32);
        copy_rows(tt.leaked, tm.leaked, 16 * L.LEAK, n_valid, m_cp, lane, 32);
        copy_rows(tt.cachem, tm.cachem, 8 * L.C, n_valid, m_cp, lane, 32);
        copy_rows(tt.props, tm.props, 4 * L.N * L.nprops, n_valid, m_cp, lane, 32);
        copy_rows(tt.priv, tm.priv, 4 * L.N, n_valid, m_cp, lane, 32);
      }
      __syncwarp();
    }
    uint32_t def_done = 0;
    keep = 1;
    if (active) {
      def_done = logic_phase2(c, p, op, act + CBX_TILE * 10 + lane * 12, s_init, desc + lane * DW, acc);
      keep = c.g(STG_OBS_KIND) == OBS_KEEP;
      if (def_done && cfg.emit_terminal_obs && p.v.term_def_infected_nodes) {
        int8_t* ti = p.v.term_def_infected_nodes + c.env * L.n;
        for (int i = 0; i < L.n; ++i) ti[i] = (int8_t)((c.g(STG_DEF_TERM_INST + (i >> 5)) >> (i & 31)) & 1u);
      }
    }
    const uint32_t enc_mask = ~__ballot_sync(0xFFFFFFFFu, keep != 0);
    __syncwarp();
    // ---- the attacker's observation fields, transposed through the padded square ----
    {
      const Target tm = make_target(p.v, L, e0, false);
      const uint32_t* de = desc + lane * DW;
      const bool mine = active && !keep;
      const int nd = mine ? (int)de[D_ND] : 0, nc = mine ? (int)de[D_NC] : 0;
      const bool blank = mine && de[D_KIND] == OBS_BLANK;
      emit_field_rows(tm.scalars, 8, n_valid, enc_mask, img, lane, mine, GenStaging{c, STG_SCALARS});
      emit_field_rows(tm.leaked, 4 * L.LEAK, n_valid, enc_mask, img, lane, mine, GenStaging{c, L.g_leaked});
      emit_field_rows(tm.cachem, 2 * L.C, n_valid, enc_mask, img, lane, mine, GenCache{c, nc, blank, 0u, 0u});
      emit_field_rows(tm.props, L.N * L.nprops, n_valid, enc_mask, img, lane, mine, GenProps{c, nd, L.nprops, blank, 0, 0, 0u, 0u});
      emit_field_rows(tm.priv, L.N, n_valid, enc_mask, img, lane, mine, GenPriv{c, nd, blank});
      // ---- the defender's observation of the tile ----
      if (def_encode) {
        if (n_valid == CBX_TILE) {
          if (rows_scn != scn) { build_defender_rows(tb, drows, L.n, L.nservices, lane); rows_scn = scn; __syncwarp(); }
          // infected_nodes [32][n]: bit i of env e's installed bits (descriptor), 4 bytes per lane and iteration
          const int words = CBX_TILE * L.n / 4;
          uint32_t* di32 = reinterpret_cast<uint32_t*>(tm.infected);
          const FastDiv dn(p.enc.d_n);
          for (int w = lane; w < words; w += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
              const uint32_t b = (uint32_t)w * 4u + bb, e = dn.div(b), i = b - e * L.n;
              v |= ((desc[e * DW + D_OWNED + L.OW + (i >> 5)] >> (i & 31)) & 1u) << (8 * bb);
            }
            di32[w] = v;
          }
          emit_periodic_bytes(tm.fw_in, drows, n6, lane);
          emit_periodic_bytes(tm.fw_out, drows + n6 + 4, n6, lane);
          emit_periodic_bytes(tm.services, drows + 2 * (n6 + 4), L.nservices, lane);
        } else {
          encode_defender_by_warp<DimsDyn>(t, tm, n_valid, mask_all(), true, 0, 1);  // ragged last tile
        }
      }
    }
    // deferred defender auto-reset (DummyVecEnv resets after the step; the observations above were taken before it)
    if (active && def_done) c.defender_reset(s_init);
    __syncwarp();
  }
  // episode statistics: warp shuffle reduce, one atomic per slot per warp
#pragma unroll
  for (int k = 0; k < CBX_STAT_COUNT; ++k) {
    double x = acc.v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, o);
    if (lane == 0 && x != 0.0) atomicAdd(p.v.episode_stats + k, x);
  }
}

}  // namespace cbx
