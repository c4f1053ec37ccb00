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
//   * the int32 observation fields leave in two phases.  Phase 1, one thread per env: GATHER what the field is made of from
//     the env's own state column into a compact image in shared memory -- the property matrix as a bit stream (props bits per
//     discovered node), privilege levels as 2-bit codes, the credential cache as (target index, port) byte pairs.  Phase 2,
//     the whole warp on one env at a time: EXPAND the image to int32 words in registers and store them where they belong --
//     every store instruction writes 512 contiguous bytes of one env's row.  (Round 1 produced the words thread-per-env into
//     a 32 x 36 square and stored it transposed: 4 scattered 128-byte segments per instruction and ~2.5 x the instructions;
//     profiles/r01_ncu_lines_wide_kernel.txt.)
//   * the static rows of the defender's observation (firewall / service status: the LearningDefender acts on a stale copy,
//     SURVEY.md B.1, so they never change) are written by RESET launches only; a step writes the infected-nodes row.
// Factored masks only (what these configurations use: a dense Chain-100 connect mask is 8.5 MB per env).
namespace cbx {

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
  uint32_t* wb = smem + Q.warps + warp * Q.warp_words;  // this warp's private area
  uint32_t* sg = wb + Q.w_stage;
  uint32_t* desc = wb + Q.w_desc;
  uint32_t* img = wb + Q.w_img;                          // packed field images, IS words per env; the tile's actions before that
  int32_t* act = reinterpret_cast<int32_t*>(img);
  const int IS = Q.img_stride;                           // odd: thread-per-env writes are bank-conflict free
  constexpr uint32_t kFull = 0xFFFFFFFFu;

  Acc acc;
#pragma unroll
  for (int k = 0; k < CBX_STAT_COUNT; ++k) acc.v[k] = 0.0;
  int slice_of_kind[3] = {p.slice_of_kind[0], p.slice_of_kind[1], p.slice_of_kind[2]};
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
  const int NPROPS = L.nprops, PW = L.PW;
  const uint32_t keep_lo = NPROPS >= 32 ? 0xFFFFFFFFu : ((1u << NPROPS) - 1u);
  const uint32_t keep_hi = NPROPS > 32 ? (NPROPS >= 64 ? 0xFFFFFFFFu : ((1u << (NPROPS - 32)) - 1u)) : 0u;
  const int wpe_leak = 4 * L.LEAK, wpe_cachem = 2 * L.C, wpe_props = L.N * NPROPS, wpe_priv = L.N;
  // tile order: static stride, or (Q.dynamic) every tile after a warp's first is the next ticket of a global counter, so that
  // warps on faster SMs -- and, in a multi-scenario batch, warps that drew cheap tiles -- take more of them
  auto next_tile = [&](int t) {
    if (!Q.dynamic) return t + gw;
    int x = 0;
    if (lane == 0) x = gw + atomicAdd(p.tile_counter, 1);
    return __shfl_sync(kFull, x, 0);
  };
  __shared__ __align__(8) uint64_t s_abar[CBX_WIDE_WARPS];  // one mbarrier per warp: the tile's action block
  uint32_t aphase = 0;
  if (lane == 0) {
    mbar_init(&s_abar[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  for (int tile = (int)blockIdx.x * nw + warp; tile < p.n_tiles; tile = next_tile(tile)) {
    const int64_t e0 = (int64_t)tile * CBX_TILE;
    const int n_valid = (int)min((int64_t)CBX_TILE, p.n_envs - e0);
    const int scn = p.tile_scn ? p.tile_scn[tile] : 0;
    const uint32_t* tb = p.tables + (size_t)scn * p.table_stride;  // global memory, L1-resident
    const uint32_t* s_init = tb + p.table_words;
    uint32_t* gst = p.state + (int64_t)tile * L.S * CBX_TILE;        // the tile's state, in place
    // pull the tile's S lines towards L2 now, all at once: the game logic's dependent accesses then pay L2 latency, not HBM's
    for (int r = lane; r < L.S; r += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(gst + (size_t)r * CBX_TILE));
    auto load_actions = [&]() {
      if (reset_only) return;
      const bool la = p.att_actions && (who_att || !marlon);
      // a full tile's int32 actions are one contiguous 16-byte aligned block per agent: two bulk copies on the warp's mbarrier
      // (22 dependent load -> shared-store pairs took 9 us per tile; the image area was last used through the generic proxy)
      const bool bulk = !p.act_i16 && n_valid == CBX_TILE && (la || def_on) &&
                        ((((uintptr_t)p.att_actions) | ((uintptr_t)p.def_actions)) & 15u) == 0;
      if (bulk) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint32_t att_bytes = la ? (uint32_t)(CBX_TILE * AW * 4) : 0u, def_bytes = def_on ? (uint32_t)(CBX_TILE * 12 * 4) : 0u;
          mbar_expect_tx(&s_abar[warp], att_bytes + def_bytes);
          if (att_bytes) tma_load_1d(act, p.att_actions + e0 * AW, att_bytes, &s_abar[warp]);
          if (def_bytes) tma_load_1d(act + CBX_TILE * 10, p.def_actions + e0 * 12, def_bytes, &s_abar[warp]);
        }
        mbar_wait(&s_abar[warp], aphase);
        aphase ^= 1u;
        return;
      }
      if (la)
        for (int q = lane; q < n_valid * AW; q += 32) act[q] = load_act(p.att_actions, e0 * AW + q, p.act_i16);
      if (def_on)
        for (int q = lane; q < n_valid * 12; q += 32) act[CBX_TILE * 10 + q] = load_act(p.def_actions, e0 * 12 + q, p.act_i16);
    };
    load_actions();
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
    const uint32_t att_done_mask = __ballot_sync(kFull, att_done != 0);
    const uint32_t keep1 = __ballot_sync(kFull, keep != 0);
    CBX_WPROF(2)  // attacker logic
    const bool need_term = att_done_mask && cfg.auto_reset && cfg.emit_terminal_obs && !reset_only;
    uint32_t def_done = 0;
    // pass 0 (rare): terminal observations of the envs that finished, BEFORE the auto-reset; pass 1: this step's observations
    for (int pass = need_term ? 0 : 1; pass < 2; ++pass) {
      uint32_t emit_mask;
      if (pass == 0) {
        if (active && ((att_done_mask >> lane) & 1u)) build_desc(c, desc + lane * DW, DW, nullptr);
        emit_mask = att_done_mask & ~keep1;
        const uint32_t cp = att_done_mask & keep1;
        if (cp) {  // intercepted-and-truncated step: the terminal observation is the previous one
          EnvMask m_cp;
#pragma unroll
          for (int q = 0; q < kGroups; ++q) m_cp.w[q] = 0;
          m_cp.w[0] = cp;
          const Target tt = make_target(p.v, L, e0, true), tm = make_target(p.v, L, e0, false);
          copy_rows(tt.scalars, tm.scalars, 32, n_valid, m_cp, lane, 32);
          copy_rows(tt.leaked, tm.leaked, 16 * L.LEAK, n_valid, m_cp, lane, 32);
          copy_rows(tt.cachem, tm.cachem, 8 * L.C, n_valid, m_cp, lane, 32);
          copy_rows(tt.props, tm.props, 4 * L.N * L.nprops, n_valid, m_cp, lane, 32);
          copy_rows(tt.priv, tm.priv, 4 * L.N, n_valid, m_cp, lane, 32);
        }
      } else {
        if (need_term) {  // the terminal pass used the image area: the tile's actions again
          __syncwarp();
          load_actions();
          __syncwarp();
        }
        keep = 1;
        if (active) {
          def_done = logic_phase2(c, p, op, act + CBX_TILE * 10 + lane * 12, s_init, desc + lane * DW, acc);
          keep = c.g(STG_OBS_KIND) == OBS_KEEP;
          if (def_done && cfg.emit_terminal_obs && p.v.term_def_infected_nodes) {
            int8_t* ti = p.v.term_def_infected_nodes + c.env * L.n;
            for (int i = 0; i < L.n; ++i) ti[i] = (int8_t)((c.g(STG_DEF_TERM_INST + (i >> 5)) >> (i & 31)) & 1u);
          }
        }
        emit_mask = ~__ballot_sync(kFull, keep != 0);
        CBX_WPROF(4)  // auto-reset, defender logic, descriptors
      }
      __syncwarp();
      emit_mask &= n_valid == CBX_TILE ? kFull : ((1u << n_valid) - 1u);
      if (!emit_mask) continue;
      // ---- the attacker's observation fields of the envs in emit_mask ----
      const Target tm = make_target(p.v, L, e0, pass == 0);
      const uint32_t* de = desc + lane * DW;
      const bool mine = (emit_mask >> lane) & 1u;
      const int nd = mine ? (int)de[D_ND] : 0, nc = mine ? (int)de[D_NC] : 0;
      const bool blank = mine && de[D_KIND] == OBS_BLANK;
      const int nd_eff = blank ? 0 : nd, nc_eff = blank ? 0 : nc;  // a blank observation shows nothing but its counts
      int nleak = 0;
      if (mine) {
        // scalars: this thread's own 8 words, two 16-byte stores (a tile's rows are contiguous: 1 KB per warp)
        uint4* so = reinterpret_cast<uint4*>(tm.scalars + (size_t)lane * 8);
        so[0] = make_uint4(c.g(STG_SCALARS + 0), c.g(STG_SCALARS + 1), c.g(STG_SCALARS + 2), c.g(STG_SCALARS + 3));
        so[1] = make_uint4(c.g(STG_SCALARS + 4), c.g(STG_SCALARS + 5), c.g(STG_SCALARS + 6), c.g(STG_SCALARS + 7));
        for (int k = 0; k < L.LEAKS; ++k) nleak += c.g(L.g_leaked + k) != 0;  // slots fill in order (ENV:890-907)
      }
      // leaked_credentials [LEAK][4]: almost always all zero; a used slot is read from the staging words as it is
      for (int e = 0; e < n_valid; ++e) {
        if (!((emit_mask >> e) & 1u)) continue;
        const int nl = __shfl_sync(kFull, nleak, e);
        emit_row(tm.leaked + (size_t)e * wpe_leak, 0, wpe_leak, wpe_leak, lane, [&](int w0) {
          const int slot = w0 >> 2;
          if (slot >= nl) return make_uint4(0u, 0u, 0u, 0u);
          const uint32_t q = sg[(L.g_leaked + slot) * CBX_TILE + e];
          return make_uint4(leak_field(q, 0), leak_field(q, 1), leak_field(q, 2), leak_field(q, 3));
        });
      }
      // credential_cache_matrix [C][2] = (target discovery index, port) per cached credential: 16 bits per entry
      {
        const int ncmax = __reduce_max_sync(kFull, nc_eff);
        const int EC = 2 * Q.img_words;  // entries per chunk
        for (int c0 = 0; c0 < L.C; c0 += EC) {
          const int c1 = min(L.C, c0 + EC), chi = min(c1, ncmax);
          for (int i = c0; i < chi; i += 2) {  // phase 1: every thread packs ITS env's entries i, i + 1
            uint32_t pk = 0;
            if (i < nc_eff) {
              const uint32_t cw = c.w(L.o_cache + (i >> 1));
              const uint32_t* r0 = c.triple((int)(cw & 0xFFFFu));
              pk = c.byte(L.o_disc_idx, (int)r0[0]) | (r0[1] << 8);
              if (i + 1 < nc_eff) {
                const uint32_t* r1 = c.triple((int)(cw >> 16));
                pk |= (c.byte(L.o_disc_idx, (int)r1[0]) | (r1[1] << 8)) << 16;
              }
            }
            img[lane * IS + ((i - c0) >> 1)] = pk;
          }
          __syncwarp();
          for (int e = 0; e < n_valid; ++e) {  // phase 2
            if (!((emit_mask >> e) & 1u)) continue;
            const int nce = __shfl_sync(kFull, nc_eff, e);
            const uint32_t* im = img + e * IS;
            emit_row(tm.cachem + (size_t)e * wpe_cachem, 2 * c0, 2 * c1, wpe_cachem, lane, [&](int w0) {
              const int i = w0 >> 1;  // entries i, i + 1
              if (i >= nce) return make_uint4(0u, 0u, 0u, 0u);
              const uint32_t pk = im[(i - c0) >> 1];
              return make_uint4(pk & 0xFFu, (pk >> 8) & 0xFFu, (pk >> 16) & 0xFFu, pk >> 24);
            });
          }
          __syncwarp();
        }
      }
      // discovered_nodes_properties [N][props] as a bit stream (props bits per discovered node, discovery order) and
      // nodes_privilegelevel [N] as 2-bit codes behind it; a chunk holds NPC nodes (all of them at Chain-100)
      {
        const int ndmax = __reduce_max_sync(kFull, nd_eff);
        const bool any_blank = __any_sync(kFull, blank);
        const int NPC = Q.nodes_per_chunk;  // a multiple of 4: a chunk starts on a 16-byte boundary of the row
        for (int k0 = 0; k0 < L.N; k0 += NPC) {
          const int k1 = min(L.N, k0 + NPC), khi = min(k1, ndmax);
          const int priv_at = Q.img_words - ((NPC + 15) >> 4);  // 2-bit codes of the chunk's nodes, behind the bit stream
          {  // phase 1
            uint64_t bits = 0;
            int nb = 0, wi = 0;
            uint32_t dw = 0, pv = 0;
            for (int k = k0; k < khi; ++k) {
              if ((k & 3) == 0 || k == k0) dw = c.w(L.o_disc_order + (k >> 2));
              uint32_t lo = 0, hi = 0, lvl = 0;
              if (k < nd_eff) {
                const uint32_t node = (dw >> ((k & 3) * 8)) & 0xFFu;
                lo = c.w(L.o_props + node * PW) & keep_lo;
                if (PW > 1) hi = c.w(L.o_props + node * PW + 1) & keep_hi;
                lvl = (c.g(L.g_priv + (node >> 4)) >> ((node & 15) * 2)) & 3u;
              }
              bits |= (uint64_t)lo << nb;
              nb += NPROPS < 32 ? NPROPS : 32;
              if (nb >= 32) { img[lane * IS + wi++] = (uint32_t)bits; bits >>= 32; nb -= 32; }
              if (NPROPS > 32) {
                bits |= (uint64_t)hi << nb;
                nb += NPROPS - 32;
                if (nb >= 32) { img[lane * IS + wi++] = (uint32_t)bits; bits >>= 32; nb -= 32; }
              }
              const int j = (k - k0) & 15;
              pv |= lvl << (2 * j);
              if (j == 15) { img[lane * IS + priv_at + ((k - k0) >> 4)] = pv; pv = 0; }
            }
            if (nb > 0) img[lane * IS + wi] = (uint32_t)bits;
            if (khi > k0 && ((khi - k0) & 15)) img[lane * IS + priv_at + ((khi - 1 - k0) >> 4)] = pv;
          }
          __syncwarp();
          for (int e = 0; e < n_valid; ++e) {  // phase 2
            if (!((emit_mask >> e) & 1u)) continue;
            const int nde = __shfl_sync(kFull, nd_eff, e);
            const bool bl = any_blank && __shfl_sync(kFull, (int)blank, e);
            const uint32_t* im = img + e * IS;
            const int lim = (min(nde, k1) - k0) * NPROPS;  // bits of this chunk that the env's own thread wrote
            emit_row(tm.props + (size_t)e * wpe_props, k0 * NPROPS, k1 * NPROPS, wpe_props, lane, [&](int w0) {
              if (bl) return make_uint4(2u, 2u, 2u, 2u);  // "unknown": blank observations only (ENV:765)
              const int b = w0 - k0 * NPROPS;
              if (b >= lim) return make_uint4(0u, 0u, 0u, 0u);
              const uint32_t nib = im[b >> 5] >> (b & 31);  // b is a multiple of 4: a nibble never straddles a word
              return make_uint4(nib & 1u, (nib >> 1) & 1u, (nib >> 2) & 1u, (nib >> 3) & 1u);
            });
            emit_row(tm.priv + (size_t)e * wpe_priv, k0, k1, wpe_priv, lane, [&](int w0) {
              if (w0 >= nde) return make_uint4(0u, 0u, 0u, 0u);
              const int b = 2 * (w0 - k0);
              const uint32_t q = im[priv_at + (b >> 5)] >> (b & 31);
              return make_uint4(q & 3u, (q >> 2) & 3u, (q >> 4) & 3u, (q >> 6) & 3u);
            });
          }
          __syncwarp();
        }
      }
      if (pass == 1) CBX_WPROF(5)  // attacker observation fields
    }
    // ---- the defender's observation of the tile ----
    if (def_encode) {
      const Target tm = make_target(p.v, L, e0, false);
      if (n_valid == CBX_TILE) {
        // infected_nodes [32][n] bytes: every thread lays out ITS env's row in the (unpadded) tile image, then the warp copies
        // the 32 * n contiguous bytes out with 16-byte stores
        uint8_t* ib = reinterpret_cast<uint8_t*>(img);
        const uint32_t* di = desc + lane * DW + D_OWNED + L.OW;
        if ((L.n & 1) == 0) {
          uint16_t* row = reinterpret_cast<uint16_t*>(ib + lane * L.n);
          for (int i = 0; i < L.n; i += 2) {
            const uint32_t two = (di[i >> 5] >> (i & 31)) & 3u;
            row[i >> 1] = (uint16_t)((two & 1u) | ((two & 2u) << 7));
          }
        } else {
          for (int i = 0; i < L.n; ++i) ib[lane * L.n + i] = (uint8_t)((di[i >> 5] >> (i & 31)) & 1u);
        }
        __syncwarp();
        uint4* dst = reinterpret_cast<uint4*>(tm.infected);
        const uint4* src = reinterpret_cast<const uint4*>(img);
        for (int q = lane; q < CBX_TILE * L.n / 16; q += 32) dst[q] = src[q];
        __syncwarp();
        if (reset_only) {  // the static rows: written when an env is (re)created, never by a step (they cannot change)
          uint8_t* drows = reinterpret_cast<uint8_t*>(img);
          build_defender_rows(tb, drows, L.n, L.nservices, lane);
          __syncwarp();
          emit_periodic_bytes(tm.fw_in, drows, n6, lane);
          emit_periodic_bytes(tm.fw_out, drows + defender_row_stride(n6), n6, lane);
          emit_periodic_bytes(tm.services, drows + 2 * defender_row_stride(n6), L.nservices, lane);
          __syncwarp();
        }
      } else {
        Tile t;
        t.L = &L; t.tb = tb; t.st = gst; t.sg = sg; t.desc = desc; t.lut = nullptr; t.K = &p.enc; t.DW = DW;
        encode_defender_by_warp<DimsDyn>(t, tm, n_valid, mask_all(), reset_only, 0, 1);  // ragged last tile
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
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(kFull, x, o);
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
