// cbx_pipe.cuh -- pipelined (warp-specialised) step kernel for sm_100a; included by cbx_pipe.cu.
//
// Why: in the fused kernel every CTA alternates "warp 0 plays 32 envs" (latency bound, ~13 us, store path idle) and "all warps
// encode" (bandwidth bound); the CTAs of an SM start together and stay in phase, so HBM idles while the SMs think
// (profiles/r01_section_skip_experiment.txt: 0.065 ms of a 0.2 ms step is logic with nothing under it).  And the store
// path itself is the second loss: 8-byte st.global rows reach 4.2 TB/s where TMA bulk stores of the same bytes reach
// 6.0 TB/s (profiles/r01_store_path_microbench.txt).
//
// One persistent CTA per SM, two kinds of warps:
//   * `wl` LOGIC warps, one thread per env, each on its own tile of 32 envs: TMA-load the state tile, play both agents
//     (logic_phase1/2), then every thread lays out ITS env's small observation fields (scalars, leaked credentials,
//     credential cache, property matrix, privilege levels, local-vulnerability mask) env-major in shared memory -- a
//     tile's rows are contiguous in every output tensor, so each field leaves as ONE bulk copy per tile -- applies the
//     deferred defender auto-reset, TMA-stores the state tile.  The tile's encoder descriptors go to a slot.
//   * `we` ENCODER warps, one env at a time: build the template row of the env's remote / connect action masks in the
//     warp's shared-memory buffer and hand it to the TMA engine: one cp.async.bulk per connect-mask row (or row pair when
//     rows are 8 mod 16 bytes) whose source is the template or a shared zero row -- every row of those masks is either
//     zero or the env's template.  One of them also emits the defender's observation of the tile.
// mbarriers: ready[slot] (1 arrival: the logic warp), empty[slot] (`we` arrivals: the encoder warps).
namespace cbx {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- overlapped launches (cbx_params.overlap): per-tile completion counters in HBM ----
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Completions are reported in two hops so that no WORKING warp ever executes a gpu-scope fence (measured: with every warp
// fencing and adding to the HBM counter itself, membar stalls were 30 % of all stall samples of the kernel,
// profiles/r02_ncu_lines_pipe_kernel_overlap_v1.txt): a warp whose copies of a tile have completed adds one to the tile's
// entry of a small ring in SHARED memory (cta-scope release); a dedicated publisher warp polls the ring and, once all parts
// of a tile are in, does the cross-proxy + gpu-scope fences and the release store of the launch's sequence number to HBM.
constexpr int kDoneRing = 32;              // tiles of one CTA whose completion is being tracked (ordinals j, j + 1, ...)
constexpr uint32_t kStopTile = 0xFFFFFFFFu;
// one more completed part of the CTA's tile with ordinal j: everything this warp wrote to the tile (bulk copies that have
// completed, plain stores of all its lanes before the preceding __syncwarp) happens-before the publisher's release store
__device__ __forceinline__ void tile_part_done(uint32_t* s_parts, const int j) {
  __threadfence_block();
  atomicAdd(s_parts + (j & (kDoneRing - 1)), 1u);
}
// cp.async.bulk.wait_group takes an immediate: all but the n most recent bulk groups of this thread have completed
__device__ __forceinline__ void tma_store_wait_all_but(const int n) {
  switch (n) {
    case 0: asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.bulk.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.bulk.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.bulk.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.bulk.wait_group 6;" ::: "memory"); break;
    case 7: asm volatile("cp.async.bulk.wait_group 7;" ::: "memory"); break;
    case 8: asm volatile("cp.async.bulk.wait_group 8;" ::: "memory"); break;
    default: asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); break;  // stricter than asked for
  }
}

template <int ENC> struct DimsOf { typedef DimsDyn T; };
template <> struct DimsOf<2> { typedef DimsToyCtf T; };
template <> struct DimsOf<3> { typedef DimsChain10 T; };

struct FieldImages {  // env-major images of a tile's small observation fields (shared memory)
  int32_t *scal, *leak, *cachem, *props, *priv;
  uint32_t* local;
};

// One thread lays out the small observation fields of ITS env (ENV:859-933 encodings; same values as encode_attacker).
// `de` = the env's encoder descriptor (owned-by-discovery-index bits), `e` = env index within the tile.
template <class D>
__device__ __forceinline__ void build_field_images(const Ctx& c, const uint32_t* de, const FieldImages& im, const int e, const bool dense) {
  const cbx_layout* L = c.L;
  const int N = CBX_DIM(D, N, L->N), NL = CBX_DIM(D, L, L->L);
  const int NC = CBX_DIM(D, C, L->C), NPROPS = CBX_DIM(D, NPROPS, L->nprops), LEAK = CBX_DIM(D, LEAK, L->LEAK);
  const int PW = D::kStatic ? (D::NPROPS + 31) / 32 : L->PW;
  const int nd = (int)de[D_ND], nc = (int)de[D_NC];
  const bool blank = de[D_KIND] == OBS_BLANK;
  int32_t* o;
  o = im.scal + e * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = (int32_t)c.g(STG_SCALARS + k);
  o = im.leak + e * 4 * LEAK;
#pragma unroll 4
  for (int k = 0; k < LEAK; ++k) {  // one packed staging word per slot -> the row's four words
    const uint32_t q = k < L->LEAKS ? c.g(L->g_leaked + k) : 0u;
    o[4 * k + 0] = (int32_t)leak_field(q, 0); o[4 * k + 1] = (int32_t)leak_field(q, 1);
    o[4 * k + 2] = (int32_t)leak_field(q, 2); o[4 * k + 3] = (int32_t)leak_field(q, 3);
  }
  o = im.cachem + e * 2 * NC;
#pragma unroll 2
  for (int k = 0; k < NC; ++k) {
    uint32_t a = 0, b = 0;
    if (!blank && k < nc) {
      const uint32_t* rec = c.triple((int)c.half(L->o_cache, k));
      a = c.byte(L->o_disc_idx, (int)rec[0]);
      b = rec[1];
    }
    o[2 * k] = (int32_t)a;
    o[2 * k + 1] = (int32_t)b;
  }
  o = im.props + e * N * NPROPS;
  for (int k = 0; k < N; ++k) {
    uint32_t lo = 0, hi = 0;
    if (!blank && k < nd) {
      const int node = (int)c.byte(L->o_disc_order, k);
      lo = c.w(L->o_props + node * PW);
      if (PW > 1) hi = c.w(L->o_props + node * PW + 1);
    }
#pragma unroll 2
    for (int pi = 0; pi < NPROPS; ++pi) {
      const uint32_t bit = ((pi < 32 ? lo : hi) >> (pi & 31)) & 1u;
      o[k * NPROPS + pi] = blank ? 2 : (int32_t)bit;
    }
  }
  o = im.priv + e * N;
#pragma unroll 2
  for (int k = 0; k < N; ++k) {
    uint32_t val = 0;
    if (!blank && k < nd) {
      const uint32_t node = c.byte(L->o_disc_order, k);
      val = (c.g(L->g_priv + (node >> 4)) >> ((node & 15) * 2)) & 3u;
    }
    o[k] = (int32_t)val;
  }
  if (dense) {  // local-vulnerability mask [N][L] bytes, 4 per word (N * L is a multiple of 4 on this path)
    uint32_t* lo = im.local + e * (N * NL / 4);
    uint32_t word = 0;
    int i = 0;
    for (int s = 0; s < N; ++s) {
      const bool own = (de[D_OWNED + (s >> 5)] >> (s & 31)) & 1u;
      const uint32_t* rec = c.vuln_rec(own ? (int)c.byte(L->o_disc_order, s) : 0, 0);
      for (int v = 0; v < NL; ++v, ++i) {
        if (own) word |= (rec[v * CBX_VULN_WORDS] & 1u) << (8 * (i & 3));
        if ((i & 3) == 3) { lo[i >> 2] = word; word = 0; }
      }
    }
  }
}

// The remote / connect action masks of one env, by one warp: template rows into the warp buffer, then bulk copies.
template <class D>
__device__ __forceinline__ void encode_masks_pipe(const uint32_t* de, const uint2* lut, const cbx_layout* L, const cbx_enc_consts& K,
                                                  int8_t* remote_env, int8_t* connect_env, uint8_t* wb, const cbx_pipe_plan& Q,
                                                  const uint8_t* zero, const int lane, const bool prof, long long (&pacc)[3], const bool hint,
                                                  const uint64_t pol) {
  long long pt = prof ? clock64() : 0;
#define CBX_EPROF(slot)                                                  \
  if (prof) {                                                            \
    long long _now = clock64();                                          \
    pacc[(slot) - 13] += _now - pt;                                      \
    pt = _now;                                                           \
  }
  const int N = CBX_DIM(D, N, L->N), NR = CBX_DIM(D, R, L->R), NP = CBX_DIM(D, P, L->P), NC = CBX_DIM(D, C, L->C);
  const int ROWR = N * NR, ROWC = N * NP * NC;
  const uint32_t nc = de[D_NC];
  const int skip = CBX_SKIP(K);  // -DCBX_EXPERIMENTS builds only (CBX_DEBUG_SKIP): 128 no connect copies, 256 no copies at all
  // the bulk copies that read this buffer for the warp's previous env must have finished reading it
  tma_store_wait_read();
  __syncwarp();
  CBX_EPROF(13)  // waiting for the TMA engine to finish reading the previous env's rows
  // remote mask image: 8-byte granules never straddle a row (ROWR % 8 == 0)
  const int limr = (int)de[D_LIMR];
  uint2* rimg = reinterpret_cast<uint2*>(wb + Q.b_remote);
#pragma unroll
  for (int g = lane; g < N * ROWR / 8; g += 32) {
    const uint32_t b = (uint32_t)g * 8u;
    const uint32_t s = D::kStatic ? b / (uint32_t)(D::kStatic ? D::N * D::R : 1) : FastDiv(K.d_rowr).div(b);
    const int w = (int)(b - s * ROWR);
    const uint32_t m = ((de[D_OWNED + (s >> 5)] >> (s & 31)) & 1u) ? lowmask(min(max(limr - w, 0), 8)) : 0u;
    rimg[g] = lut[m];
  }
  // connect template row (8-byte granules), written to every place a bulk copy reads it from:
  //   gs == 1: [T]            gs == 2: [0 T | T 0 | T T] (the zero halves were cleared once at kernel start)
  const int limc = (int)de[D_LIMC];
  const uint64_t base = ((uint64_t)de[D_BHI] << 32) | de[D_BLO];
  uint8_t* cb = wb + Q.b_conn;
#pragma unroll
  for (int g = lane; g < ROWC / 8; g += 32) {
    const int w = g * 8;
    uint32_t m = lowmask(min(max(limc - w, 0), 8));
    const int ph = D::kStatic ? w % (D::kStatic ? D::C : 1) : (int)((uint32_t)w - FastDiv(K.d_C).div((uint32_t)w) * NC);
    if (NC <= 48) m &= (uint32_t)(base >> ph);
    else m &= lowmask(min(max((int)nc - ph, 0), 8)) | (lowmask(min(max(NC - ph + (int)nc, 0), 8)) & ~lowmask(min(max(NC - ph, 0), 8)));
    const uint2 v = lut[m & 0xFFu];
    if (Q.gs == 1) {
      *reinterpret_cast<uint2*>(cb + 8 * g) = v;
    } else {
      *reinterpret_cast<uint2*>(cb + ROWC + 8 * g) = v;
      *reinterpret_cast<uint2*>(cb + 2 * ROWC + 8 * g) = v;
      *reinterpret_cast<uint2*>(cb + 4 * ROWC + 8 * g) = v;
      *reinterpret_cast<uint2*>(cb + 5 * ROWC + 8 * g) = v;
    }
  }
  fence_async_smem();  // generic-proxy writes above -> visible to the bulk copies below
  __syncwarp();
  CBX_EPROF(14)  // building the rows in shared memory
  // ---- hand the rows to the TMA engine (lanes issue in parallel; every lane closes its own bulk group) ----
  if (!(skip & 256)) {
    if (lane == 31) tma_store_1d_pol(remote_env, wb + Q.b_remote, (uint32_t)(N * ROWR), hint, pol);
    if (skip & 128) {
    } else if (Q.gs == 1) {
      for (int s = lane; s < N; s += 32) {
        const bool own = (de[D_OWNED + (s >> 5)] >> (s & 31)) & 1u;
        tma_store_1d_pol(connect_env + (size_t)s * ROWC, own ? cb : zero, (uint32_t)ROWC, hint, pol);
      }
    } else {
      for (int j = lane; j < N / 2; j += 32) {
        const uint32_t pr = (de[D_OWNED + ((2 * j) >> 5)] >> ((2 * j) & 31)) & 3u;  // bit 0: row 2j owned, bit 1: row 2j+1
        const uint8_t* src = pr == 3u ? cb + 4 * ROWC : pr == 1u ? cb + 2 * ROWC : pr == 2u ? cb : zero;
        tma_store_1d_pol(connect_env + (size_t)j * 2 * ROWC, src, (uint32_t)(2 * ROWC), hint, pol);
      }
    }
  }
  tma_store_commit();
  CBX_EPROF(15)  // issuing the bulk copies
#undef CBX_EPROF
}

// The defender's observation of a whole tile (MultiBinary arrays, DWR:492-534): the firewall / service parts are static per
// scenario (the LearningDefender acts on a stale copy) and leave from a CTA-wide image built once.
template <class D>
__device__ __forceinline__ void encode_defender_tile(const Tile& t, const Target& o, const int n_valid, uint8_t* wb, const cbx_pipe_plan& Q,
                                                     const uint8_t* def_static, const int lane, const bool hint, const uint64_t pol) {
  const cbx_layout* L = t.L;
  const int n = CBX_DIM(D, NN, L->n), nsvc = CBX_DIM(D, NSVC, L->nservices);
  const int OW = D::kStatic ? (D::N + 31) / 32 : L->OW;
  if (n_valid != CBX_TILE) {  // ragged last tile: plain stores (live binding: the logic warp writes the firewall rows)
    encode_defender_by_warp<D>(t, o, n_valid, mask_all(), !t.fx, 0, 1);
    if (t.fx) {
      for (int idx = lane; idx < n_valid * nsvc; idx += 32) o.services[idx] = 1;
    }
    return;
  }
  tma_store_wait_read();
  __syncwarp();
  uint8_t* img = wb + Q.b_inf;
  for (int idx = lane; idx < CBX_TILE * n; idx += 32) {
    const uint32_t e = D::kStatic ? (uint32_t)idx / (uint32_t)(D::kStatic ? D::NN : 1) : FastDiv(t.K->d_n).div((uint32_t)idx);
    const int i = idx - (int)e * n;
    const uint32_t* di = t.desc + e * t.DW + D_OWNED + OW;
    img[idx] = (uint8_t)((di[i >> 5] >> (i & 31)) & 1u);
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) tma_store_1d_pol(o.infected, img, (uint32_t)(CBX_TILE * n), hint, pol);
  if (lane == 1 && !t.fx) tma_store_1d_pol(o.fw_in, def_static, (uint32_t)(CBX_TILE * 6 * n), hint, pol);
  if (lane == 2 && !t.fx) tma_store_1d_pol(o.fw_out, def_static + CBX_TILE * 6 * n, (uint32_t)(CBX_TILE * 6 * n), hint, pol);
  if (lane == 3 && nsvc > 0) tma_store_1d_pol(o.services, def_static + Q.def_svc, (uint32_t)(CBX_TILE * nsvc), hint, pol);
  tma_store_commit();
}

// field image -> output tensor for the selected envs of a tile, plain coalesced stores (ragged or partially kept tiles)
__device__ __forceinline__ void copy_field_rows(int32_t* dst, const int32_t* img, const int wpe, const int n_valid, const uint32_t mask,
                                                const int lane) {
  for (int e = 0; e < n_valid; ++e) {
    if (!((mask >> e) & 1u)) continue;
    for (int w = lane; w < wpe; w += 32) dst[e * wpe + w] = img[e * wpe + w];
  }
}

// LIVE: the live defender binding (per-env firewall rule lists).  A compile-time switch, not a test of cbx_params.fwx_words:
// with Ctx::fx a constant nullptr the stale-binding instantiation carries none of the live branches of the game logic (as a
// run-time pointer they cost the logic warps 12 % at 1 048 576 envs per launch, where the state does not stay in L2).
template <int ENC, bool LIVE>
__global__ void __launch_bounds__(512, 1) cbx_pipe_kernel(const __grid_constant__ cbx_params p, const int op) {
  typedef typename DimsOf<ENC>::T D;
  extern __shared__ __align__(128) uint32_t smem[];
  const cbx_layout& L = p.lay;
  const cbx_config& cfg = p.cfg;
  const cbx_pipe_plan& Q = p.pipe;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool reset_only = op & CBX_OP_RESET, who_def = op & CBX_OP_DEFENDER, who_att = op & CBX_OP_ATTACKER;
  const bool marlon = cfg.mode == CBX_MODE_MARLON;
  const bool def_on = marlon && cfg.def_enabled && who_def;
  const bool def_encode = def_on && !(op & CBX_OP_NOTIFY);
  const bool dense = p.v.connect != nullptr;
  const int DW = p.enc.desc_words;
  const int AW = marlon ? 10 : 5;
  uint32_t* s_tb = smem + Q.tables;
  uint2* s_lut = reinterpret_cast<uint2*>(smem + Q.lut);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q.bars);
  uint64_t* bar_load = bars + 1;             // [wl]    state-tile loads, one per logic warp
  uint64_t* bar_ready = bar_load + Q.wl;     // [nslot] logic -> encoders
  uint64_t* bar_empty = bar_ready + Q.nslot; // [nslot] encoders -> logic
  uint64_t* bar_rdef = bar_empty + Q.nslot;  // [nslot] logic -> the encoder of the tile's defender observation
  uint32_t* s_parts = smem + Q.done_ring;          // [kDoneRing] completed parts per tracked tile
  uint32_t* s_ptile = s_parts + kDoneRing;         // [kDoneRing] its tile index (kStopTile: the logic warp has no more tiles)
  volatile uint32_t* s_pub = s_ptile + kDoneRing;  // [1] tiles the publisher warp has dealt with
  uint8_t* s_zero = reinterpret_cast<uint8_t*>(smem + Q.zero);
  uint8_t* s_defst = reinterpret_cast<uint8_t*>(smem + Q.def_static);
  const uint32_t* s_init = s_tb + p.table_words;
  const uint32_t* s_fx = LIVE ? s_init + ((L.S + 3) & ~3) : nullptr;  // live defender binding: firewall extension tables
  const uint32_t table_bytes = (uint32_t)(p.table_words + ((L.S + 3) & ~3) + p.fwx_words) * 4u;
  const int nthreads = (int)blockDim.x;  // logic + encoder warps (+ the publisher warp when launches overlap)
  constexpr uint32_t kRowBytes = CBX_TILE * 4u;
  // Overlapped launches: the next launch of the stream may start its CTAs on every SM this launch's CTA has left (its CTAs
  // then wait tile by tile on tile_done, never on this grid as a whole)
  const bool overlap = p.overlap != 0;
  if (overlap) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t seq_need = p.seq - 1u;  // tile_done[t] holds the sequence number of the last launch that completed tile t
  // dynamic tile order: this launch's OWN ticket counter (a fresh slot of a ring the host re-zeroes in halves: any number of
  // launches may be in flight at once when the grid is smaller than the machine, so no counter is shared or reset in-kernel)
  int* const tickets = p.tickets;
  const bool fw_rows = LIVE && def_encode && Q.i_fwin >= 0;  // live binding: per-env firewall rows, laid out by the logic warps
  const bool keep_state = p.l2_hints & 1, stream_masks = p.l2_hints & 2, stream_rest = p.l2_hints & 4, keep_tables = p.l2_hints & 8;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();

  // the scenario tables are on their way (one bulk copy, ~2 us from L2 / HBM) while the CTA fills its constant buffers
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    for (int w = 0; w < Q.wl; ++w) mbar_init(&bar_load[w], 1);
    for (int s = 0; s < Q.nslot; ++s) { mbar_init(&bar_ready[s], 1); mbar_init(&bar_empty[s], Q.we); mbar_init(&bar_rdef[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bars[0], table_bytes);
    tma_load_1d_pol(s_tb, p.tables, table_bytes, &bars[0], keep_tables, pol_keep);
  }
  for (int k = tid; k < 256; k += nthreads) {
    uint32_t lo = ((k & 0xF) * 0x00204081u) & 0x01010101u, hi = (((k >> 4) & 0xF) * 0x00204081u) & 0x01010101u;
    s_lut[k] = make_uint2(lo, hi);
  }
  // zero row + encoder buffers (the zero halves of the row-pair templates stay zero for the whole kernel)
  for (int k = Q.zero + tid; k < Q.def_static; k += nthreads) smem[k] = 0;
  for (int k = Q.wbufs + tid; k < Q.wbufs + Q.we * Q.wbuf_words; k += nthreads) smem[k] = 0;
  for (int k = tid; k < 2 * kDoneRing + 1; k += nthreads) s_parts[k] = 0;
  __syncthreads();
  mbar_wait(&bars[0], 0);
  if (def_encode) {  // static parts of the defender observation for a tile of 32 envs: [32][6n] in, [32][6n] out, [32][nsvc]
    const int n6 = 6 * L.n;
    if (!s_fx)
      for (int idx = tid; idx < CBX_TILE * n6; idx += nthreads) {
        const int i = idx % n6, node = i / 6, r = i - node * 6;
        const uint32_t dob = s_tb[s_tb[CBX_H_OFF_NODE] + node * CBX_NODE_WORDS + CBX_N_DEFOBS];
        s_defst[idx] = (uint8_t)((dob >> r) & 1u);
        s_defst[CBX_TILE * n6 + idx] = (uint8_t)((dob >> (8 + r)) & 1u);
      }
    for (int idx = tid; idx < CBX_TILE * L.nservices; idx += nthreads) s_defst[Q.def_svc + idx] = 1;
  }
  fence_async_smem();
  __syncthreads();

  const int my_tiles = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int spw = Q.nslot / Q.wl;  // slots per logic warp
  long long prof_t = p.prof ? clock64() : 0;
  long long pp[5] = {0, 0, 0, 0, 0};  // per-warp phase cycles (slots 8..12), flushed once at the end
#define CBX_PPROF(slot)                                                        \
  if (p.prof) {                                                                \
    long long _now = clock64();                                                \
    pp[(slot) - 8] += _now - prof_t;                                           \
    prof_t = _now;                                                             \
  }

  if (warp < Q.wl) {
    // =============================== game-logic warp ===============================
    Acc acc;
#pragma unroll
    for (int k = 0; k < CBX_STAT_COUNT; ++k) acc.v[k] = 0.0;
    int slice_of_kind[3] = {p.slice_of_kind[0], p.slice_of_kind[1], p.slice_of_kind[2]};
    uint32_t load_phase = 0;
    uint32_t* lb = smem + Q.lbufs + warp * Q.lbuf_words;  // state tile | staging | actions | field images
    uint32_t* sg = lb + Q.l_stage;
    int32_t* act = reinterpret_cast<int32_t*>(lb + Q.l_acts);
    FieldImages im;
    im.scal = reinterpret_cast<int32_t*>(lb + Q.i_scal); im.leak = reinterpret_cast<int32_t*>(lb + Q.i_leak);
    im.cachem = reinterpret_cast<int32_t*>(lb + Q.i_cachem); im.props = reinterpret_cast<int32_t*>(lb + Q.i_props);
    im.priv = reinterpret_cast<int32_t*>(lb + Q.i_priv); im.local = lb + Q.i_local;
    const int wpe_leak = 4 * L.LEAK, wpe_cachem = 2 * L.C, wpe_props = L.N * L.nprops, wpe_priv = L.N, wpe_local = L.sz_local / 4;
    int u = 0;  // my u-th tile
    // Tile order.  Static: tile j of this CTA is blockIdx.x + j * gridDim.x, warp w takes j = w, w + wl, ...  Dynamic: a warp's
    // first tile is the static one (no ticket latency in the pipeline fill), every later one is the next ticket of a global
    // counter -- CTAs that run ahead (the SMs do not all see the same memory latency) take more tiles and the launch ends
    // when the work does, not when the slowest CTA has finished a fixed share.  A warp that draws a ticket past the end
    // publishes a stop marker in its next slot.  Measured: +2 % at 28 tiles per CTA, +9..11 % from 55 on,
    // -2.5 % at 14 (most of a short launch is assigned before any CTA has shown its speed), hence the threshold in the plan.
    int next_tile = (int)blockIdx.x + warp * (int)gridDim.x;
    int done_j = -1;  // overlapped launches: ordinal of the tile whose stores were issued last, not yet reported complete
    for (int j = warp;; j += Q.wl, ++u) {
      const int slot = warp + Q.wl * (u % spw), use = u / spw;
      uint32_t* desc = smem + Q.slots + slot * Q.slot_words;
      uint32_t* hdr = desc + Q.s_hdr;
      int tile;
      if (Q.dynamic) {
        if (u > 0 && lane == 0) next_tile = (int)gridDim.x * Q.wl + atomicAdd(tickets, 1);
        tile = __shfl_sync(0xFFFFFFFFu, next_tile, 0);
        if (tile >= p.n_tiles) {
          if (overlap && lane == 0) {  // tell the publisher warp that this logic warp is done
            while (j - (int)*s_pub >= kDoneRing) __nanosleep(32);
            s_ptile[j & (kDoneRing - 1)] = kStopTile;
            __threadfence_block();
            atomicAdd(s_parts + (j & (kDoneRing - 1)), (uint32_t)(1 + Q.we));
          }
          if (use > 0) mbar_wait(&bar_empty[slot], (uint32_t)(use - 1) & 1u);
          if (lane == 0) hdr[CBX_SH_TILE] = 0xFFFFFFFFu;
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_ready[slot]);
          break;
        }
      } else {
        if (j >= my_tiles) break;
        tile = (int)blockIdx.x + j * (int)gridDim.x;
      }
      const int64_t e0 = (int64_t)tile * CBX_TILE;
      const int n_valid = (int)min((int64_t)CBX_TILE, p.n_envs - e0);
      const uint4* gstate = reinterpret_cast<const uint4*>(p.state + (int64_t)tile * L.S * CBX_TILE);
      const int state_q = L.S * CBX_TILE / 4;  // 16-byte words of a state tile
      const bool need_att = !reset_only && p.att_actions && (who_att || !marlon);
      const bool need_def = !reset_only && def_on;
      // a full tile's actions are one contiguous 16-byte aligned block per agent: they ride the same mbarrier as the state
      // tile (bulk copies also read page-locked HOST memory efficiently: cbx_batch_step_host hands host pointers over)
      const bool bulk_acts = Q.logic_tma && n_valid == CBX_TILE && ((((uintptr_t)p.att_actions) | ((uintptr_t)p.def_actions)) & 15u) == 0;
      if (overlap) {  // every write of the earlier launches to this tile (state, observations, results) must have landed
        if (lane == 0) {
          while ((int32_t)(ld_acquire_gpu(p.tile_done + tile) - seq_need) < 0) __nanosleep(64);
          asm volatile("fence.proxy.async;" ::: "memory");
          while (j - (int)*s_pub >= kDoneRing) __nanosleep(32);  // the ring entry of ordinal j - kDoneRing has been retired
          s_ptile[j & (kDoneRing - 1)] = (uint32_t)tile;
        }
        __syncwarp();
      }
      if (Q.logic_tma) {
        // the previous tile's state store and field copies (issued by lanes 0..6) must be done reading this warp's buffer
        tma_store_wait_read();
        __syncwarp();
        if (lane == 0) {
          // int16 actions land in the upper half of their int32 area and are widened in place after the wait
          const uint32_t asz = p.act_i16 ? 2u : 4u, half = p.act_i16 ? 1u : 0u;
          const uint32_t att_bytes = (bulk_acts && need_att) ? (uint32_t)(CBX_TILE * AW) * asz : 0u;
          const uint32_t def_bytes = (bulk_acts && need_def) ? (uint32_t)(CBX_TILE * 12) * asz : 0u;
          mbar_expect_tx(&bar_load[warp], (uint32_t)L.S * kRowBytes + att_bytes + def_bytes);
          tma_load_1d_pol(lb, gstate, (uint32_t)L.S * kRowBytes, &bar_load[warp], keep_state, pol_keep);
          if (att_bytes)
            tma_load_1d_pol(reinterpret_cast<char*>(act) + half * att_bytes, reinterpret_cast<const char*>(p.att_actions) + e0 * AW * asz,
                            att_bytes, &bar_load[warp], stream_rest, pol_stream);
          if (def_bytes)
            tma_load_1d_pol(reinterpret_cast<char*>(act + CBX_TILE * 10) + half * def_bytes,
                            reinterpret_cast<const char*>(p.def_actions) + e0 * 12 * asz, def_bytes, &bar_load[warp], stream_rest, pol_stream);
        }
      } else {
        uint4* d = reinterpret_cast<uint4*>(lb);
#pragma unroll 7
        for (int q = lane; q < state_q; q += 32) d[q] = gstate[q];
      }
      if (!bulk_acts) {
        if (need_att)
          for (int q = lane; q < n_valid * AW; q += 32) act[q] = load_act(p.att_actions, e0 * AW + q, p.act_i16);
        if (need_def)
          for (int q = lane; q < n_valid * 12; q += 32) act[CBX_TILE * 10 + q] = load_act(p.def_actions, e0 * 12 + q, p.act_i16);
      }
      if (Q.logic_tma) {
        mbar_wait(&bar_load[warp], load_phase);
        load_phase ^= 1;
      }
      if (bulk_acts && p.act_i16) {  // widen this lane's rows: every lane reads before any lane writes (the areas overlap)
        int32_t va[10], vd[12];
        const int16_t* a16 = reinterpret_cast<const int16_t*>(act) + CBX_TILE * AW + lane * AW;
        const int16_t* d16 = reinterpret_cast<const int16_t*>(act + CBX_TILE * 10) + CBX_TILE * 12 + lane * 12;
#pragma unroll
        for (int k = 0; k < 10; ++k) va[k] = (need_att && k < AW) ? (int32_t)a16[k] : 0;
#pragma unroll
        for (int k = 0; k < 12; ++k) vd[k] = need_def ? (int32_t)d16[k] : 0;
        __syncwarp();
        if (need_att) {
#pragma unroll
          for (int k = 0; k < 10; ++k)
            if (k < AW) act[lane * AW + k] = va[k];
        }
        if (need_def) {
#pragma unroll
          for (int k = 0; k < 12; ++k) act[CBX_TILE * 10 + lane * 12 + k] = vd[k];
        }
      }
      __syncwarp();
      CBX_PPROF(9)  // state tile + actions in
      const bool active = lane < n_valid;
      Ctx c;
      c.st = lb + lane; c.sg = sg + lane; c.tb = s_tb; c.L = &L; c.cfg = &cfg; c.env = e0 + lane; c.fx = s_fx;
      uint32_t att_done = 0, keep = 1;
      if (active) {
        logic_phase1(c, p, op, act + lane * AW, s_init, slice_of_kind, acc);
        att_done = c.g(STG_ATT_DONE);
        keep = c.g(STG_OBS_KIND) == OBS_KEEP;
      }
      const uint32_t att_done_mask = __ballot_sync(0xFFFFFFFFu, att_done != 0);
      const uint32_t keep1 = __ballot_sync(0xFFFFFFFFu, keep != 0);
      CBX_PPROF(10)  // game logic
      if (use > 0) mbar_wait(&bar_empty[slot], (uint32_t)(use - 1) & 1u);  // the encoders are done with the slot's last tile
      CBX_PPROF(8)  // waiting for a free descriptor slot
      Tile t;
      t.L = &L; t.tb = s_tb; t.st = lb; t.sg = sg; t.desc = desc; t.lut = s_lut; t.K = &p.enc; t.DW = DW;
      t.fx = s_fx; t.init = s_init;
      // terminal observations of the envs that finished: encoded by this warp BEFORE the auto-reset (plain stores)
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
        if (m_cp.any()) {  // intercepted-and-truncated step: the terminal observation is the previous one
          Target tm = make_target(p.v, L, e0, false);
          copy_rows(tt.scalars, tm.scalars, 32, n_valid, m_cp, lane, 32);
          copy_rows(tt.leaked, tm.leaked, 16 * L.LEAK, n_valid, m_cp, lane, 32);
          copy_rows(tt.cachem, tm.cachem, 8 * L.C, n_valid, m_cp, lane, 32);
          copy_rows(tt.props, tm.props, 4 * L.N * L.nprops, n_valid, m_cp, lane, 32);
          copy_rows(tt.priv, tm.priv, 4 * L.N, n_valid, m_cp, lane, 32);
          copy_rows(tt.local, tm.local, L.sz_local, n_valid, m_cp, lane, 32);
          copy_rows(tt.remote, tm.remote, L.sz_remote, n_valid, m_cp, lane, 32);
          copy_rows(tt.connect, tm.connect, L.sz_connect, n_valid, m_cp, lane, 32);
        }
        __syncwarp();
      }
      // ---- phase 2a: the attacker's auto-reset and the encoder descriptors.  The action masks depend on nothing the defender
      //      does (it acts on its stale copy), so the encoder warps are released BEFORE the defender moves.
      keep = 1;
      if (active) {
        if (!reset_only && c.g(STG_ATT_DONE) && cfg.auto_reset) {
          if (marlon) c.attacker_reset(s_init);
          else { c.cyber_reset(s_init); c.setf32(L.o_att_return, 0.f); }
          c.stage_reset_obs();
        }
        keep = c.g(STG_OBS_KIND) == OBS_KEEP;
        build_desc(c, desc + lane * DW, DW, nullptr);
        if (!keep) {
          uint32_t* ob = p.v.owned_bits + c.env * L.OW;
          for (int k = 0; k < L.OW; ++k) ob[k] = desc[lane * DW + D_OWNED + k];
        }
      }
      const uint32_t enc_mask = ~__ballot_sync(0xFFFFFFFFu, keep != 0);
      if (lane == 0) { hdr[CBX_SH_ENC_MASK] = enc_mask; hdr[CBX_SH_TILE] = (uint32_t)tile; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ready[slot]);
      // ---- phase 2b: the defender's move; its observation of the tile is released separately ----
      uint32_t def_done = 0;
      if (active) {
        if (!reset_only && def_on) defender_wrapper_step(c, p, act + CBX_TILE * 10 + lane * 12, acc);
        def_done = c.g(STG_DEF_DONE) && cfg.auto_reset;
        if (def_done) {
          // main defender observation after an auto-reset shows the fresh environment (DWR:477)
          for (int k = 0; k < L.Wn; ++k) desc[lane * DW + D_OWNED + L.OW + k] = s_init[L.o_installed + k];
          if (cfg.emit_terminal_obs && p.v.term_def_infected_nodes) {
            // terminal defender observation = infected nodes seen by the step that ended the episode
            int8_t* ti = p.v.term_def_infected_nodes + c.env * L.n;
            for (int i = 0; i < L.n; ++i) ti[i] = (int8_t)((c.g(STG_DEF_TERM_INST + (i >> 5)) >> (i & 31)) & 1u);
          }
        }
        if (fw_rows) {
          // live binding: the defender has just acted on THIS environment -- its observation shows the installed bits after
          // the move (a re-image removes the agent) and the rule lists as they are now, or the fresh environment's after its
          // auto-reset (DWR:477)
          uint32_t* de = desc + lane * DW;
          if (!def_done)
            for (int k = 0; k < L.Wn; ++k) de[D_OWNED + L.OW + k] = c.w(L.o_installed + k);
          // ... and the tile's firewall rows [32][6n] bytes, this env's 6n bytes each way: six rule bytes per node from the six
          // presence bits of the node's list (the byte LUT of the mask encoder), as three 16-bit stores
          uint8_t* rin = reinterpret_cast<uint8_t*>(lb + Q.i_fwin) + lane * 6 * L.n;
          uint8_t* rout = reinterpret_cast<uint8_t*>(lb + Q.i_fwout) + lane * 6 * L.n;
          const uint32_t* grp = s_fx + CBX_FX_WORDS + s_tb[CBX_H_N_PORTS];
          for (int node = 0; node < L.n; ++node) {
            const uint32_t gw = grp[node];
            const int gi = (int)(gw & 0xFFFFu), go = (int)(gw >> 16);
            const uint2 bi = s_lut[(def_done ? s_init[L.o_fw + 2 * gi] : c.w(L.o_fw + 2 * gi)) & 63u];
            const uint2 bo = s_lut[(def_done ? s_init[L.o_fw + 2 * go] : c.w(L.o_fw + 2 * go)) & 63u];
            uint16_t* pi = reinterpret_cast<uint16_t*>(rin + 6 * node);
            uint16_t* po = reinterpret_cast<uint16_t*>(rout + 6 * node);
            pi[0] = (uint16_t)bi.x; pi[1] = (uint16_t)(bi.x >> 16); pi[2] = (uint16_t)bi.y;
            po[0] = (uint16_t)bo.x; po[1] = (uint16_t)(bo.x >> 16); po[2] = (uint16_t)bo.y;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_rdef[slot]);
      CBX_PPROF(10)
      // ---- the small observation fields of the tile: one thread per env, env-major images, one bulk copy per field ----
      if (active && !keep) build_field_images<D>(c, desc + lane * DW, im, lane, dense);
      // deferred defender auto-reset (DummyVecEnv resets after the step; the observations above were taken before it)
      if (active && def_done) c.defender_reset(s_init);
      if (Q.logic_tma) fence_async_smem();
      __syncwarp();
      {
        const Target tm = make_target(p.v, L, e0, false);
        const bool full = n_valid == CBX_TILE && enc_mask == 0xFFFFFFFFu;
        if (Q.logic_tma) {
          if (lane == 0) tma_store_1d_pol(p.state + (int64_t)tile * L.S * CBX_TILE, lb, (uint32_t)L.S * kRowBytes, keep_state, pol_keep);
          if (full) {
            if (lane == 1) tma_store_1d_pol(tm.scalars, im.scal, CBX_TILE * 32u, stream_rest, pol_stream);
            if (lane == 2) tma_store_1d_pol(tm.leaked, im.leak, (uint32_t)(CBX_TILE * 4 * wpe_leak), stream_rest, pol_stream);
            if (lane == 3) tma_store_1d_pol(tm.cachem, im.cachem, (uint32_t)(CBX_TILE * 4 * wpe_cachem), stream_rest, pol_stream);
            if (lane == 4) tma_store_1d_pol(tm.props, im.props, (uint32_t)(CBX_TILE * 4 * wpe_props), stream_rest, pol_stream);
            if (lane == 5) tma_store_1d_pol(tm.priv, im.priv, (uint32_t)(CBX_TILE * 4 * wpe_priv), stream_rest, pol_stream);
            if (lane == 6 && dense) tma_store_1d_pol(tm.local, im.local, (uint32_t)(CBX_TILE * 4 * wpe_local), stream_rest, pol_stream);
          }
          if (fw_rows && n_valid == CBX_TILE) {  // the defender's observation is written for every env, kept attacker rows or not
            if (lane == 7) tma_store_1d_pol(tm.fw_in, lb + Q.i_fwin, (uint32_t)(CBX_TILE * 6 * L.n), stream_rest, pol_stream);
            if (lane == 8) tma_store_1d_pol(tm.fw_out, lb + Q.i_fwout, (uint32_t)(CBX_TILE * 6 * L.n), stream_rest, pol_stream);
          }
          if (lane < 9) tma_store_commit();
          if (overlap) {  // the tile before this one: its bulk copies (one group per lane) have completed -> report it
            if (lane < 9) tma_store_wait_all_but(1);
            __syncwarp();
            if (lane == 0 && done_j >= 0) tile_part_done(s_parts, done_j);
            done_j = j;
          }
        } else {
          // plain 16-byte copies: a tile's rows are contiguous in every tensor (fully coalesced 512-byte warp stores)
          auto copy16 = [&](void* dst, const void* src, int n16) {
            uint4* dd = reinterpret_cast<uint4*>(dst);
            const uint4* ss = reinterpret_cast<const uint4*>(src);
#pragma unroll 4
            for (int q = lane; q < n16; q += 32) dd[q] = ss[q];
          };
          copy16(p.state + (int64_t)tile * L.S * CBX_TILE, lb, state_q);
          if (full) {
            copy16(tm.scalars, im.scal, CBX_TILE * 8 / 4);
            copy16(tm.leaked, im.leak, CBX_TILE * wpe_leak / 4);
            copy16(tm.cachem, im.cachem, CBX_TILE * wpe_cachem / 4);
            copy16(tm.props, im.props, CBX_TILE * wpe_props / 4);
            copy16(tm.priv, im.priv, CBX_TILE * wpe_priv / 4);
            if (dense) copy16(tm.local, im.local, CBX_TILE * wpe_local / 4);
          }
          if (fw_rows && n_valid == CBX_TILE) {
            copy16(tm.fw_in, lb + Q.i_fwin, CBX_TILE * 6 * L.n / 16);
            copy16(tm.fw_out, lb + Q.i_fwout, CBX_TILE * 6 * L.n / 16);
          }
          if (overlap) {
            __syncwarp();
            if (lane == 0 && done_j >= 0) tile_part_done(s_parts, done_j);
            done_j = j;
          }
        }
        if (!full) {
          copy_field_rows(tm.scalars, im.scal, 8, n_valid, enc_mask, lane);
          copy_field_rows(tm.leaked, im.leak, wpe_leak, n_valid, enc_mask, lane);
          copy_field_rows(tm.cachem, im.cachem, wpe_cachem, n_valid, enc_mask, lane);
          copy_field_rows(tm.props, im.props, wpe_props, n_valid, enc_mask, lane);
          copy_field_rows(tm.priv, im.priv, wpe_priv, n_valid, enc_mask, lane);
          if (dense) copy_field_rows(reinterpret_cast<int32_t*>(tm.local), reinterpret_cast<const int32_t*>(im.local), wpe_local, n_valid, enc_mask, lane);
        }
        if (fw_rows && n_valid != CBX_TILE) {  // ragged last tile: the valid envs' rows, byte by byte
          const uint8_t* si = reinterpret_cast<const uint8_t*>(lb + Q.i_fwin);
          const uint8_t* so = reinterpret_cast<const uint8_t*>(lb + Q.i_fwout);
          for (int q = lane; q < n_valid * 6 * L.n; q += 32) { tm.fw_in[q] = (int8_t)si[q]; tm.fw_out[q] = (int8_t)so[q]; }
        }
        __syncwarp();
      }
      CBX_PPROF(11)  // field images + state write-back
    }
    tma_store_wait_all();  // every lane that issued bulk copies waits for its own
    if (overlap) {
      __syncwarp();
      if (lane == 0 && done_j >= 0) tile_part_done(s_parts, done_j);
    }
    // episode statistics: warp shuffle reduce, one atomic per slot per warp
#pragma unroll
    for (int k = 0; k < CBX_STAT_COUNT; ++k) {
      double x = acc.v[k];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, o);
      if (lane == 0 && x != 0.0) atomicAdd(p.v.episode_stats + k, x);
    }
  } else if (warp < Q.wl + Q.we) {
    // =============================== encoder warp ===============================
    const int wid = warp - Q.wl;
    uint8_t* wb = reinterpret_cast<uint8_t*>(smem + Q.wbufs + wid * Q.wbuf_words);
    long long pacc[3] = {0, 0, 0};
    uint32_t stopped = 0;  // dynamic order: logic warps that have published their stop marker
    const uint32_t all_stopped = (1u << Q.wl) - 1u;
    int done_j = -1;  // overlapped launches: ordinal of the tile this warp issued last, not yet reported complete
    for (int j = 0;; ++j) {
      const int lw = j % Q.wl, u = j / Q.wl;
      if (Q.dynamic) {
        if (stopped == all_stopped) break;
        if ((stopped >> lw) & 1u) continue;
      } else if (j >= my_tiles) {
        break;
      }
      const int slot = lw + Q.wl * (u % spw), use = u / spw;
      const uint32_t* desc = smem + Q.slots + slot * Q.slot_words;
      mbar_wait(&bar_ready[slot], (uint32_t)use & 1u);
      CBX_PPROF(12)  // encoders waiting for a tile
      const int tile = Q.dynamic ? (int)desc[Q.s_hdr + CBX_SH_TILE] : (int)blockIdx.x + j * (int)gridDim.x;
      if (tile < 0) {  // stop marker
        stopped |= 1u << lw;
        continue;
      }
      const uint32_t enc_mask = desc[Q.s_hdr + CBX_SH_ENC_MASK];
      const int64_t e0 = (int64_t)tile * CBX_TILE;
      const int n_valid = (int)min((int64_t)CBX_TILE, p.n_envs - e0);
      int groups = 0;  // bulk groups every lane of this warp commits for this tile
      if (dense) {
        for (int e = wid; e < n_valid; e += Q.we)
          if ((enc_mask >> e) & 1u) {
            encode_masks_pipe<D>(desc + e * DW, s_lut, &L, p.enc, p.v.remote_vulnerability + (e0 + e) * L.sz_remote,
                                 p.v.connect + (e0 + e) * (int64_t)L.sz_connect, wb, Q, s_zero, lane, p.prof != nullptr, pacc, stream_masks, pol_stream);
            ++groups;
          }
      }
      if (def_encode && wid == j % Q.we) {  // the defender has moved by now (it is released after the action masks)
        mbar_wait(&bar_rdef[slot], (uint32_t)use & 1u);
        Tile t;
        t.L = &L; t.tb = s_tb; t.st = nullptr; t.sg = nullptr; t.desc = desc; t.lut = s_lut; t.K = &p.enc; t.DW = DW;
        t.fx = s_fx; t.init = s_init;
        const Target tm = make_target(p.v, L, e0, false);
        encode_defender_tile<D>(t, tm, n_valid, wb, Q, s_defst, lane, stream_rest, pol_stream);
        if (n_valid == CBX_TILE) ++groups;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[slot]);
      if (overlap) {  // this warp's copies of the tile BEFORE this one have completed (groups retire in order): report it
        tma_store_wait_all_but(groups);
        __syncwarp();
        if (lane == 0 && done_j >= 0) tile_part_done(s_parts, done_j);
        done_j = j;
      }
      prof_t = p.prof ? clock64() : 0;
    }
    tma_store_wait_all();
    if (overlap) {
      __syncwarp();
      if (lane == 0 && done_j >= 0) tile_part_done(s_parts, done_j);
    }
    if (p.prof && lane == 0)
      for (int k = 0; k < 3; ++k) atomicAdd(p.prof + 13 + k, (unsigned long long)pacc[k]);
  } else if (overlap && lane == 0) {
    // =============================== publisher warp (overlapped launches) ===============================
    // tiles of this CTA in ordinal order: once all 1 + we parts of a tile have been reported, everything the CTA wrote to
    // it is complete -> release the launch's sequence number into the tile's counter in HBM
    const uint32_t parts = (uint32_t)(1 + Q.we);
    uint32_t stopped = 0;
    const uint32_t all_stopped = (1u << Q.wl) - 1u;
    for (int j = 0;; ++j) {
      const int lw = j % Q.wl;
      if (Q.dynamic) {
        if (stopped == all_stopped) break;
        if ((stopped >> lw) & 1u) continue;
      } else if (j >= my_tiles) {
        break;
      }
      volatile uint32_t* cnt = s_parts + (j & (kDoneRing - 1));
      while (*cnt < parts) __nanosleep(128);
      __threadfence_block();
      const uint32_t tile = s_ptile[j & (kDoneRing - 1)];
      *cnt = 0;
      __threadfence_block();
      *s_pub = (uint32_t)(j + 1);
      if (tile == kStopTile) {
        stopped |= 1u << lw;
        continue;
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.tile_done + tile), "r"(p.seq) : "memory");
    }
  }
  if (p.prof && lane == 0)
    for (int k = 0; k < 5; ++k)
      if (pp[k]) atomicAdd(p.prof + 8 + k, (unsigned long long)pp[k]);
#undef CBX_PPROF
  // stream order for whatever follows the NEXT launch: a grid that ends implies the grid before it has ended (it has, long
  // ago -- its tiles were consumed above -- so this never waits in practice)
  if (overlap) asm volatile("griddepcontrol.wait;" ::: "memory");
}


// launch helpers, one set per binding (cbx_pipe.cu: stale, cbx_pipe_live.cu: live -- two translation units compile in parallel)
template <bool LIVE>
static cudaError_t pipe_attrs_t(int enc, int smem_bytes) {
  switch (enc) {
    case 3: return cudaFuncSetAttribute(cbx_pipe_kernel<3, LIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    case 2: return cudaFuncSetAttribute(cbx_pipe_kernel<2, LIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    default: return cudaFuncSetAttribute(cbx_pipe_kernel<1, LIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  }
}
template <bool LIVE>
static cudaError_t launch_pipe_t(const cbx_params* p, int op, int grid, cudaStream_t stream) {
  // overlapped launches: one more warp, the publisher of the per-tile completion counters
  const int threads = (p->pipe.wl + p->pipe.we + (p->overlap ? 1 : 0)) * 32;
  if (p->overlap) {
    // programmatic dependent launch: this grid's CTAs may start while the previous launch of the stream is still draining;
    // the kernel orders its accesses tile by tile through cbx_params.tile_done
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)p->pipe.total_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    switch (p->enc.warp_env) {
      case 3: return cudaLaunchKernelEx(&cfg, cbx_pipe_kernel<3, LIVE>, *p, op);
      case 2: return cudaLaunchKernelEx(&cfg, cbx_pipe_kernel<2, LIVE>, *p, op);
      default: return cudaLaunchKernelEx(&cfg, cbx_pipe_kernel<1, LIVE>, *p, op);
    }
  }
  switch (p->enc.warp_env) {
    case 3: cbx_pipe_kernel<3, LIVE><<<grid, threads, p->pipe.total_bytes, stream>>>(*p, op); break;
    case 2: cbx_pipe_kernel<2, LIVE><<<grid, threads, p->pipe.total_bytes, stream>>>(*p, op); break;
    default: cbx_pipe_kernel<1, LIVE><<<grid, threads, p->pipe.total_bytes, stream>>>(*p, op); break;
  }
  return cudaGetLastError();
}

}  // namespace cbx
