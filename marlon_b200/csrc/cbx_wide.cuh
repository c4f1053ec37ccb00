// cbx_wide.cuh -- warp-per-tile step kernel for LARGE per-env state (Chain-100, generated networks); included by cbx_wide.cu.
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
//   * each warp owns a tile end to end -- no CTA-wide barrier in the loop, up to 14 tiles in flight per SM instead of 2;
//   * the int32 observation fields are produced by the env's own thread, 32 words at a time, into a padded 32 x 36 shared-
//     memory square and leave transposed with 16-byte stores (4 envs x 128 contiguous bytes per instruction); the property
//     matrix is a bit stream expanded four words at a time through a 16-entry uint4 table.
// Factored masks only (what these configurations use: a dense Chain-100 connect mask is 8.5 MB per env).
namespace cbx {

constexpr int kImgStride = 36;  // padded row of the transpose square: 16-byte aligned, conflict-free for 128-bit accesses both ways

// Streaming writer of one int32 field of a tile: every thread `put`s ITS env's words in order; after every 32 words the warp
// stores the padded square transposed -- env e's 32 words are 128 contiguous bytes of the output tensor.  All lanes call
// put() the same number of times (the field has the same length for every env), so the flushes are warp-uniform.
struct FieldWriter {
  int32_t* dst;     // the tile's first row of the output tensor (nullptr: field not materialised)
  uint32_t* img;    // this warp's 32 x kImgStride square
  const uint4* lut4;  // 17 entries: the 4 bits of a nibble as 4 int32 words; entry 16 = {2,2,2,2} ("unknown", blank observations)
  int wpe, n_valid, lane, j, k0;
  uint32_t mask;    // envs whose observation is (re)written
  bool mine;        // this lane's env is one of them
  __device__ __forceinline__ void flush(const int m) {  // the square holds words [k0, k0 + m) of every env
    __syncwarp();
    if ((wpe & 3) == 0) {  // env rows are 16-byte aligned: 4 envs x 128 bytes per store instruction
      const int sub = lane >> 3, c4 = (lane & 7) * 4;
      if (c4 < m) {
#pragma unroll
        for (int it = 0; it < CBX_TILE / 4; ++it) {
          const int e = it * 4 + sub;
          if (e < n_valid && ((mask >> e) & 1u))
            *reinterpret_cast<uint4*>(dst + (size_t)e * wpe + k0 + c4) = *reinterpret_cast<const uint4*>(img + e * kImgStride + c4);
        }
      }
    } else if (lane < m) {
      int32_t* d = dst + k0 + lane;
      const uint32_t* s = img + lane;
      for (int e = 0; e < n_valid; ++e)
        if ((mask >> e) & 1u) d[(size_t)e * wpe] = (int32_t)s[e * kImgStride];
    }
    __syncwarp();
    k0 += m;
    j = 0;
  }
  __device__ __forceinline__ void put(const uint32_t v) {
    if (mine) img[lane * kImgStride + j] = v;
    if (++j == 32) flush(32);
  }
  __device__ __forceinline__ void put4(const uint4 v) {  // j is a multiple of 4
    if (mine) *reinterpret_cast<uint4*>(img + lane * kImgStride + j) = v;
    j += 4;
    if (j == 32) flush(32);
  }
  __device__ __forceinline__ void finish() {
    if (j) flush(j);
  }
};

// Bits -> int32 words, four at a time: the property matrix of an env is one long bit stream (props bits per discovered node).
struct BitStream {
  FieldWriter& fw;
  uint64_t acc;
  int nb;
  bool blank;  // this env's observation is blank: every word is 2 (same control flow as the other lanes: flushes are collective)
  __device__ __forceinline__ void append(const uint32_t lo, const uint32_t hi, const int nbits) {  // nbits <= 64 - 3
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    acc |= v << nb;
    nb += nbits;
    while (nb >= 4) {
      fw.put4(fw.lut4[blank ? 16u : ((uint32_t)acc & 15u)]);
      acc >>= 4;
      nb -= 4;
    }
  }
  __device__ __forceinline__ void finish() {  // trailing bits (field length not a multiple of 4): word by word
    for (int k = 0; k < nb; ++k) fw.put(blank ? 2u : ((uint32_t)(acc >> k) & 1u));
    nb = 0;
  }
};

// Static rows of the defender observation for ONE env of the scenario `tb`: incoming [6n] | outgoing [6n] | services [nsvc],
// each followed by its first 4 bytes again (reads wrap around the end of a row) and padded to a multiple of 4 bytes.
__device__ __forceinline__ int defender_row_stride(const int len) { return (len + 4 + 3) & ~3; }
__device__ __forceinline__ void build_defender_rows(const uint32_t* tb, uint8_t* rows, const int n, const int nsvc, const int lane) {
  const int n_own = (int)tb[CBX_H_N_NODES], nsvc_own = (int)tb[CBX_H_N_SERVICES];
  const int n6 = 6 * n;
  uint8_t* rin = rows;
  uint8_t* rout = rows + defender_row_stride(n6);
  uint8_t* rsvc = rout + defender_row_stride(n6);
  for (int i = lane; i < n6 + 4; i += 32) {
    const int q = i < n6 ? i : i - n6;
    const int node = q / 6, r = q - node * 6;
    const uint32_t dob = node < n_own ? tb[tb[CBX_H_OFF_NODE] + node * CBX_NODE_WORDS + CBX_N_DEFOBS] : 0u;
    rin[i] = (uint8_t)((dob >> r) & 1u);
    rout[i] = (uint8_t)((dob >> (8 + r)) & 1u);
  }
  for (int i = lane; i < nsvc + 4; i += 32) rsvc[i] = (uint8_t)((nsvc > 0 ? i % nsvc : 0) < nsvc_own);
}

// a tile's worth (32 envs) of a byte field whose rows all equal `row` (period `len` >= 4 bytes, wrap copy behind it):
// coalesced word stores, each word = 4 bytes of the periodic string = two aligned shared-memory words funnel-shifted
__device__ __forceinline__ void emit_periodic_bytes(int8_t* dst, const uint8_t* row, const int len, const int lane) {
  if (!dst || len == 0) return;
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);  // 32 * len bytes from a 128-byte aligned tile base: whole words
  const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row);
  const int words = CBX_TILE * len / 4;
  int i = (lane * 4) % len;              // byte offset within the row of this lane's first word
  const int step = 128 % len;            // advance per iteration (32 lanes x 4 bytes)
  for (int w = lane; w < words; w += 32) {
    const uint32_t lo = r32[i >> 2], hi = r32[(i >> 2) + 1];
    d[w] = __funnelshift_r(lo, hi, (i & 3) * 8);
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
  if (tid < 17) {
    uint4* l4 = reinterpret_cast<uint4*>(smem + Q.lut4);
    l4[tid] = tid == 16 ? make_uint4(2u, 2u, 2u, 2u) : make_uint4(tid & 1u, (tid >> 1) & 1u, (tid >> 2) & 1u, (tid >> 3) & 1u);
  }
  __syncthreads();  // the only CTA-wide barrier

  Acc acc;
#pragma unroll
  for (int k = 0; k < CBX_STAT_COUNT; ++k) acc.v[k] = 0.0;
  int slice_of_kind[3] = {p.slice_of_kind[0], p.slice_of_kind[1], p.slice_of_kind[2]};
  int rows_scn = -1;  // scenario the defender rows in `drows` were built for
  long long prof_t = p.prof ? clock64() : 0;
  long long pp[6] = {0, 0, 0, 0, 0, 0};  // per-warp phase cycles (instrumentation slots 1..6), flushed once at the end
#define CBX_WPROF(slot)                    \
  if (p.prof) {                            \
    long long _now = clock64();            \
    pp[(slot) - 1] += _now - prof_t;       \
    prof_t = _now;                         \
  }
  const int n6 = 6 * L.n;
  const int nw = (int)blockDim.x >> 5;  // warps per CTA (<= CBX_WIDE_WARPS)
  const int gw = (int)gridDim.x * nw;
  // tile order: static stride, or (Q.dynamic) every tile after a warp's first is the next ticket of a global counter, so that
  // warps on faster SMs -- and, in a multi-scenario batch, warps that drew cheap tiles -- take more of them
  auto next_tile = [&](int t) {
    if (!Q.dynamic) return t + gw;
    int x = 0;
    if (lane == 0) x = gw + atomicAdd(p.tile_counter, 1);
    return __shfl_sync(0xFFFFFFFFu, x, 0);
  };
  for (int tile = (int)blockIdx.x * nw + warp; tile < p.n_tiles; tile = next_tile(tile)) {
    const int64_t e0 = (int64_t)tile * CBX_TILE;
    const int n_valid = (int)min((int64_t)CBX_TILE, p.n_envs - e0);
    const int scn = p.tile_scn ? p.tile_scn[tile] : 0;
    const uint32_t* tb = p.tables + (size_t)scn * p.table_stride;  // global memory, L1-resident
    const uint32_t* s_init = tb + p.table_words;
    uint32_t* gst = p.state + (int64_t)tile * L.S * CBX_TILE;        // the tile's state, in place
    // pull the tile's S lines towards L2 now, all at once: the game logic's dependent accesses then pay L2 latency, not HBM's
    for (int r = lane; r < L.S; r += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(gst + (size_t)r * CBX_TILE));
    if (!reset_only) {
      if (p.att_actions && (who_att || !marlon))
        for (int q = lane; q < n_valid * AW; q += 32) act[q] = load_act(p.att_actions, e0 * AW + q, p.act_i16);
      if (def_on)
        for (int q = lane; q < n_valid * 12; q += 32) act[CBX_TILE * 10 + q] = load_act(p.def_actions, e0 * 12 + q, p.act_i16);
    }
    __syncwarp();
    CBX_WPROF(1)  // actions in
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
    CBX_WPROF(2)  // attacker logic
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
    CBX_WPROF(4)  // auto-reset, defender logic, descriptors
    // ---- the attacker's observation fields, transposed through the padded square ----
    {
      const Target tm = make_target(p.v, L, e0, false);
      const uint32_t* de = desc + lane * DW;
      const bool mine = active && !keep;
      const int nd = mine ? (int)de[D_ND] : 0, nc = mine ? (int)de[D_NC] : 0;
      const bool blank = mine && de[D_KIND] == OBS_BLANK;
      FieldWriter fw;
      fw.img = img; fw.n_valid = n_valid; fw.lane = lane; fw.mask = enc_mask; fw.mine = mine;
      fw.lut4 = reinterpret_cast<const uint4*>(smem + Q.lut4);
      auto begin = [&](int32_t* dst, int wpe) { fw.dst = dst; fw.wpe = wpe; fw.j = 0; fw.k0 = 0; };
      begin(tm.scalars, 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) fw.put(mine ? c.g(STG_SCALARS + k) : 0u);
      fw.finish();
      begin(tm.leaked, 4 * L.LEAK);
      for (int k = 0; k < 4 * L.LEAK; ++k) fw.put((mine && k < 4 * L.LEAKS) ? c.g(L.g_leaked + k) : 0u);
      fw.finish();
      begin(tm.cachem, 2 * L.C);  // credential_cache_matrix [C][2]: (target discovery index, port) per cached credential
      for (int i = 0; i < L.C; ++i) {
        uint32_t a = 0, b = 0;
        if (!blank && i < nc) {
          const uint32_t* rec = c.triple((int)c.half(L.o_cache, i));
          a = c.byte(L.o_disc_idx, (int)rec[0]);
          b = rec[1];
        }
        fw.put(a);
        fw.put(b);
      }
      fw.finish();
      begin(tm.props, L.N * L.nprops);  // discovered_nodes_properties [N][props]; 2 = unknown only in blank observations
      if (L.nprops > 60) {  // word by word (warp-uniform choice: the flushes inside are collective)
        for (int kk = 0; kk < L.N; ++kk) {
          uint32_t lo = 0, hi = 0;
          if (!blank && kk < nd) {
            const int node = (int)c.byte(L.o_disc_order, kk);
            lo = c.w(L.o_props + node * L.PW);
            if (L.PW > 1) hi = c.w(L.o_props + node * L.PW + 1);
          }
          for (int pi = 0; pi < L.nprops; ++pi) fw.put(blank ? 2u : (((pi < 32 ? lo : hi) >> (pi & 31)) & 1u));
        }
      } else {
        BitStream bs{fw, 0ull, 0, blank};
        const uint32_t keep_lo = L.nprops >= 32 ? 0xFFFFFFFFu : ((1u << L.nprops) - 1u);
        const uint32_t keep_hi = L.nprops > 32 ? ((1u << (L.nprops - 32)) - 1u) : 0u;
        for (int kk = 0; kk < L.N; ++kk) {
          uint32_t lo = 0, hi = 0;
          if (kk < nd) {
            const int node = (int)c.byte(L.o_disc_order, kk);
            lo = c.w(L.o_props + node * L.PW) & keep_lo;
            if (L.PW > 1) hi = c.w(L.o_props + node * L.PW + 1) & keep_hi;
          }
          bs.append(lo, hi, L.nprops);
        }
        bs.finish();
      }
      fw.finish();
      begin(tm.priv, L.N);  // nodes_privilegelevel [N] in discovery order, as the observation saw it (staging snapshot)
      for (int k = 0; k < L.N; ++k) {
        uint32_t val = 0;
        if (!blank && k < nd) {
          const uint32_t node = c.byte(L.o_disc_order, k);
          val = (c.g(L.g_priv + (node >> 4)) >> ((node & 15) * 2)) & 3u;
        }
        fw.put(val);
      }
      fw.finish();
      CBX_WPROF(5)  // attacker observation fields
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
          emit_periodic_bytes(tm.fw_out, drows + defender_row_stride(n6), n6, lane);
          emit_periodic_bytes(tm.services, drows + 2 * defender_row_stride(n6), L.nservices, lane);
        } else {
          encode_defender_by_warp<DimsDyn>(t, tm, n_valid, mask_all(), true, 0, 1);  // ragged last tile
        }
      }
    }
    // deferred defender auto-reset (DummyVecEnv resets after the step; the observations above were taken before it)
    if (active && def_done) c.defender_reset(s_init);
    __syncwarp();
    CBX_WPROF(6)  // defender observation + deferred reset
  }
  if (p.prof && lane == 0)
    for (int k = 0; k < 6; ++k)
      if (pp[k]) atomicAdd(p.prof + 1 + k, (unsigned long long)pp[k]);
#undef CBX_WPROF
  // episode statistics: warp shuffle reduce, one atomic per slot per warp
#pragma unroll
  for (int k = 0; k < CBX_STAT_COUNT; ++k) {
    double x = acc.v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, o);
    if (lane == 0 && x != 0.0) atomicAdd(p.v.episode_stats + k, x);
  }
  if (Q.dynamic) {  // the last CTA to finish leaves the ticket counter at zero for the next launch
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(p.tile_counter + 1, 1) == (int)gridDim.x - 1) {
        p.tile_counter[0] = 0;
        p.tile_counter[1] = 0;
        __threadfence();
      }
    }
  }
}

}  // namespace cbx
