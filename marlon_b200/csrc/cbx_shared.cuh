// cbx_shared.cuh -- what the three step kernels share: TMA / mbarrier helpers, the observation encoder building blocks,
// the MARLon wrapper steps and the two game-logic phases of one env (logic_phase1 / 2, over the rules in cbx_device.cuh).
// Included by cbx_kernels.cu (fused kernel, sampler, GAE), cbx_pipe.cu (pipelined kernel) and cbx_wide.cu (warp-per-tile
// kernel); the three translation units compile in parallel (marlon_b200/build.py).
#ifndef CBX_SHARED_CUH_
#define CBX_SHARED_CUH_
#include <cuda_runtime.h>
#include <stdint.h>

#include "cbx_device.cuh"

namespace cbx {

struct FastDiv {  // x / d for x < 2^31 (d fixed per batch; cbx_fastdiv computed on the host)
  uint32_t m, s;
  __device__ __forceinline__ FastDiv(uint32_t m_, uint32_t s_) : m(m_), s(s_) {}
  __device__ __forceinline__ FastDiv(const cbx_fastdiv& f) : m(f.m), s(f.s) {}
  __device__ __forceinline__ uint32_t div(uint32_t x) const { return (m ? __umulhi(x, m) : x) >> s; }
};

// desc (env-major): 0 nd | 1 nc | 2 obs kind | 3 lim_remote | 4 lim_connect | 5 base_lo | 6 base_hi | 7 enc flags |
//                   [8, 8+OW) owned-by-discovery-index bits | [8+OW, 8+OW+Wn) installed bits for the defender observation
enum { D_ND = 0, D_NC, D_KIND, D_LIMR, D_LIMC, D_BLO, D_BHI, D_FLAGS, D_OWNED };

// ---- TMA / mbarrier helpers (PTX) -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
// L2 cache policies for bulk copies: the per-env state is re-read by the next launch (evict_last keeps its lines while the
// observation stream, 50x its size, passes through L2 -- evict_first), cbx_params.l2_hints selects which are applied
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_1d_hint(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_1d_pol(void* dst_gmem, const void* src_smem, uint32_t bytes, bool hint, uint64_t pol) {
  if (hint) tma_store_1d_hint(dst_gmem, src_smem, bytes, pol);
  else tma_store_1d(dst_gmem, src_smem, bytes);
}
__device__ __forceinline__ void tma_load_1d_pol(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, bool hint, uint64_t pol) {
  if (hint) tma_load_1d_hint(dst_smem, src_gmem, bytes, bar, pol);
  else tma_load_1d(dst_smem, src_gmem, bytes, bar);
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Output stores.  CBX_STREAMING_STORES=1 selects st.global.cs (evict-first); default is a plain write-back store.
#ifndef CBX_STREAMING_STORES
#define CBX_STREAMING_STORES 0
#endif
template <class T>
__device__ __forceinline__ void st_out(T* p, T v) {
#if CBX_STREAMING_STORES
  __stcs(p, v);
#else
  *p = v;
#endif
}
__device__ __forceinline__ void st_stream(void* p, uint4 v) { st_out(reinterpret_cast<uint4*>(p), v); }

// Per-tile env selection mask (bit e = env e of the tile); a tile has CBX_TILE/32 groups of 32 envs.
constexpr int kGroups = CBX_TILE / 32;
struct EnvMask {
  uint32_t w[kGroups];
  __device__ __forceinline__ bool test(int e) const { return (w[e >> 5] >> (e & 31)) & 1u; }
  __device__ __forceinline__ bool any() const {
    uint32_t a = 0;
#pragma unroll
    for (int k = 0; k < kGroups; ++k) a |= w[k];
    return a != 0;
  }
};
__device__ __forceinline__ EnvMask mask_and_not(const EnvMask& a, const EnvMask& b) {
  EnvMask r;
#pragma unroll
  for (int k = 0; k < kGroups; ++k) r.w[k] = a.w[k] & ~b.w[k];
  return r;
}
__device__ __forceinline__ EnvMask mask_and(const EnvMask& a, const EnvMask& b) {
  EnvMask r;
#pragma unroll
  for (int k = 0; k < kGroups; ++k) r.w[k] = a.w[k] & b.w[k];
  return r;
}
__device__ __forceinline__ EnvMask mask_not(const EnvMask& a) {
  EnvMask r;
#pragma unroll
  for (int k = 0; k < kGroups; ++k) r.w[k] = ~a.w[k];
  return r;
}
__device__ __forceinline__ EnvMask mask_all() {
  EnvMask r;
#pragma unroll
  for (int k = 0; k < kGroups; ++k) r.w[k] = 0xFFFFFFFFu;
  return r;
}

// Section-skip timing experiments exist only in -DCBX_EXPERIMENTS builds (marlon_b200/build.py build_variant); in the
// release library the constant 0 removes every such branch at compile time.
#ifdef CBX_EXPERIMENTS
#define CBX_SKIP(K) ((K).debug_skip)
#else
#define CBX_SKIP(K) 0
#endif

// ---- encoder -----------------------------------------------------------------------------------------------------------
struct Target {  // output pointers already offset to the tile's first env
  int32_t *scalars, *leaked, *cachem, *props, *priv;
  int8_t *local, *remote, *connect;
  int8_t *infected, *fw_in, *fw_out, *services;
  uint32_t* owned_bits;
};

struct Tile {
  const cbx_layout* L;
  const uint32_t* tb;
  const uint32_t* st;    // state tile
  const uint32_t* sg;    // staging
  const uint32_t* desc;  // env-major
  const uint2* lut;
  const cbx_enc_consts* K;
  int DW;
  const uint32_t* fx = nullptr;    // live defender binding: firewall extension tables (cbx.h CBX_FX_*), else nullptr
  const uint32_t* init = nullptr;  //   and the initial per-env state (what a defender observation shows right after its auto-reset)
  __device__ __forceinline__ uint32_t w(int e, int off) const { return st[off * CBX_TILE + e]; }
  __device__ __forceinline__ uint32_t g(int e, int off) const { return sg[off * CBX_TILE + e]; }
  __device__ __forceinline__ uint32_t d(int e, int k) const { return desc[e * DW + k]; }
  __device__ __forceinline__ uint32_t byte(int e, int off, int i) const { return (w(e, off + (i >> 2)) >> ((i & 3) * 8)) & 0xFFu; }
  __device__ __forceinline__ bool owned(int e, int s) const { return (d(e, D_OWNED + (s >> 5)) >> (s & 31)) & 1u; }
  // DWR:506-517: does (node, direction) have a rule named `r` (one of the defender's six)?  Static per scenario under the
  // reference's stale binding; under the live binding the env's own rule lists -- the initial ones when the defender's episode
  // has just been auto-reset (D_FLAGS bit 0: the observation shows the fresh environment, DWR:477)
  __device__ __forceinline__ uint32_t fw_rule_bit(int e, int node, int r, bool outgoing, int n_own) const {
    if (node >= n_own) return 0u;
    if (!fx) return (tb[tb[CBX_H_OFF_NODE] + node * CBX_NODE_WORDS + CBX_N_DEFOBS] >> ((outgoing ? 8 : 0) + r)) & 1u;
    const uint32_t gw = fx[CBX_FX_WORDS + tb[CBX_H_N_PORTS] + node];
    const int g = (int)(outgoing ? (gw >> 16) : (gw & 0xFFFFu));
    const uint32_t present = (d(e, D_FLAGS) & 1u) ? init[L->o_fw + 2 * g] : w(e, L->o_fw + 2 * g);
    return (present >> r) & 1u;
  }
};

// generic writer for int32 fields: `wpe` words per env, f(e, wi) -> value
template <class F>
__device__ __forceinline__ void write_i32(int32_t* dst, int wpe, FastDiv dv, int n_valid, const EnvMask& enc_mask, F f) {
  if (!dst || wpe == 0) return;
  const uint32_t total = (uint32_t)wpe * n_valid;
  for (uint32_t v = threadIdx.x * 4; v < total; v += CBX_THREADS * 4) {
    uint32_t vals[4];
    uint32_t keep = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t idx = v + q;
      vals[q] = 0;
      if (idx < total) {
        uint32_t e = dv.div(idx);
        if (enc_mask.test((int)e)) { vals[q] = (uint32_t)f((int)e, (int)(idx - e * wpe)); keep |= 1u << q; }
      }
    }
    if (keep == 15u) st_stream(dst + v, make_uint4(vals[0], vals[1], vals[2], vals[3]));
    else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((keep >> q) & 1u) dst[v + q] = (int32_t)vals[q];
    }
  }
}

// generic (slow) writer for int8 fields: f(e, i) -> 0/1
template <class F>
__device__ __forceinline__ void write_i8(int8_t* dst, int bpe, FastDiv dv, int n_valid, const EnvMask& enc_mask, F f) {
  if (!dst || bpe == 0) return;
  const uint32_t total = (uint32_t)bpe * n_valid;
  for (uint32_t v = threadIdx.x * 16; v < total; v += CBX_THREADS * 16) {
    uint32_t words[4] = {0, 0, 0, 0};
    uint32_t keep = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      uint32_t idx = v + q;
      if (idx < total) {
        uint32_t e = dv.div(idx);
        if (enc_mask.test((int)e)) {
          words[q >> 2] |= (uint32_t)(f((int)e, (int)(idx - e * bpe)) & 0xFF) << ((q & 3) * 8);
          keep |= 1u << q;
        }
      }
    }
    if (keep == 0xFFFFu) st_stream(dst + v, make_uint4(words[0], words[1], words[2], words[3]));
    else {
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if ((keep >> q) & 1u) dst[v + q] = (int8_t)((words[q >> 2] >> ((q & 3) * 8)) & 0xFF);
    }
  }
}

__device__ __forceinline__ uint32_t lowmask(int k) { return k <= 0 ? 0u : (k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u)); }

// 16 pattern bits of the connect mask starting at credential phase k: bit j <=> ((k + j) mod C) < nc
__device__ __forceinline__ uint32_t pat16(const Tile& t, int e, int k, int C) {
  if (C <= 48) {
    uint64_t base = ((uint64_t)t.d(e, D_BHI) << 32) | t.d(e, D_BLO);
    return (uint32_t)(base >> k) & 0xFFFFu;
  }
  const int nc = (int)t.d(e, D_NC);
  uint32_t m = lowmask(min(max(nc - k, 0), 16));
  m |= lowmask(min(max(C - k + nc, 0), 16)) & ~lowmask(min(max(C - k, 0), 16));
  return m;
}

// Row-structured dense masks: byte (s, w) of an env = owned[s] & (w < lim) & pattern(w mod C).
//   remote  [N][N*R]:   lim = n_discovered * R,     no pattern
//   connect [N][N*P*C]: lim = n_discovered * P * C, pattern = (credential index < n_cached)
template <bool CONNECT>
__device__ __forceinline__ uint32_t rowmask_byte(const Tile& t, int e, uint32_t i, int row_len, FastDiv drow, int C, FastDiv dC) {
  uint32_t s = drow.div(i), w = i - s * row_len;
  uint32_t lim = CONNECT ? t.d(e, D_LIMC) : t.d(e, D_LIMR);
  bool on = t.owned(e, (int)s) && w < lim;
  if (CONNECT && on) on = (w - dC.div(w) * C) < t.d(e, D_NC);
  return on ? 1u : 0u;
}

template <bool CONNECT>
__device__ __forceinline__ void write_rowmask(int8_t* dst, int bpe, FastDiv denv, int row_len, FastDiv drow, int C, FastDiv dC,
                                              int n_valid, const EnvMask& enc_mask, const Tile& t) {
  if (!dst || bpe == 0) return;
  const uint32_t total = (uint32_t)bpe * n_valid;
  const bool fast = (bpe % 16 == 0) && row_len >= 16;
  for (uint32_t v = threadIdx.x * 16; v < total; v += CBX_THREADS * 16) {
    if (fast) {
      uint32_t e = denv.div(v);
      if (!enc_mask.test((int)e)) continue;
      uint32_t i = v - e * bpe;
      uint32_t s = drow.div(i), w0 = i - s * row_len;
      int lim = (int)(CONNECT ? t.d(e, D_LIMC) : t.d(e, D_LIMR));
      int a = min(16, row_len - (int)w0);
      uint32_t m = t.owned(e, (int)s) ? lowmask(min(max(lim - (int)w0, 0), a)) : 0u;
      if (a < 16 && t.owned(e, (int)s + 1)) m |= lowmask(min(lim, 16 - a)) << a;
      if (CONNECT) m &= pat16(t, (int)e, (int)(i - dC.div(i) * C), C);
      uint2 lo = t.lut[m & 0xFFu], hi = t.lut[(m >> 8) & 0xFFu];
      st_stream(dst + v, make_uint4(lo.x, lo.y, hi.x, hi.y));
    } else {
      uint32_t words[4] = {0, 0, 0, 0};
      uint32_t keep = 0;
      for (int q = 0; q < 16; ++q) {
        uint32_t idx = v + q;
        if (idx < total) {
          uint32_t e = denv.div(idx);
          if (enc_mask.test((int)e)) {
            words[q >> 2] |= rowmask_byte<CONNECT>(t, (int)e, idx - e * bpe, row_len, drow, C, dC) << ((q & 3) * 8);
            keep |= 1u << q;
          }
        }
      }
      if (keep == 0xFFFFu) st_stream(dst + v, make_uint4(words[0], words[1], words[2], words[3]));
      else
        for (int q = 0; q < 16; ++q)
          if ((keep >> q) & 1u) dst[v + q] = (int8_t)((words[q >> 2] >> ((q & 3) * 8)) & 0xFF);
    }
  }
}


// ---- warp-per-env encoder (fast path) ------------------------------------------------------------------------------------
// One warp writes one env's whole observation: the env index is warp-uniform, so the per-env quantities (counts, owned
// bits) live in registers and no lane does index arithmetic across envs.  Every row of the remote / connect masks is
// either zero or the env's one template row: each lane computes its (at most CBX_MAXG) granules of that
// template once per env, in registers, and then only issues stores, one row after the other.
#define CBX_MAXG 4

template <int U> struct Gran;
template <> struct Gran<16> {
  typedef uint4 T;
  static __device__ __forceinline__ T zero() { return make_uint4(0, 0, 0, 0); }
  static __device__ __forceinline__ T expand(uint32_t m, const uint2* lut) {
    uint2 lo = lut[m & 0xFFu], hi = lut[(m >> 8) & 0xFFu];
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
};
template <> struct Gran<8> {
  typedef uint2 T;
  static __device__ __forceinline__ T zero() { return make_uint2(0, 0); }
  static __device__ __forceinline__ T expand(uint32_t m, const uint2* lut) { return lut[m & 0xFFu]; }
};
template <> struct Gran<4> {
  typedef uint32_t T;
  static __device__ __forceinline__ T zero() { return 0u; }
  static __device__ __forceinline__ T expand(uint32_t m, const uint2*) { return ((m & 0xFu) * 0x00204081u) & 0x01010101u; }
};

// granule `g` (U bytes at row offset g*U) of an env's template row
template <int U, bool CONNECT>
__device__ __forceinline__ typename Gran<U>::T template_granule(int g, int lim, int nc, uint64_t base, int C, FastDiv dC, const uint2* lut) {
  const int w = g * U;
  uint32_t m = lowmask(min(max(lim - w, 0), U));
  if (CONNECT) {
    const int ph = (int)((uint32_t)w - dC.div((uint32_t)w) * C);
    uint32_t pm;
    if (C <= 48) pm = (uint32_t)(base >> ph) & 0xFFFFu;
    else pm = lowmask(min(max(nc - ph, 0), 16)) | (lowmask(min(max(C - ph + nc, 0), 16)) & ~lowmask(min(max(C - ph, 0), 16)));
    m &= pm;
  }
  return Gran<U>::expand(m, lut);
}

// Long rows (more than 16 granules): one row after the other, NG granules per lane.
template <int U, int NG, bool CONNECT>
__device__ __noinline__ void warp_rowmask(int8_t* envdst, int row_len, int N, int C, cbx_fastdiv dCv, const uint32_t* desc_e,
                                          const uint2* lut, int lane) {
  typedef typename Gran<U>::T G;
  const int gpr = row_len / U;
  const int lim = (int)(CONNECT ? desc_e[D_LIMC] : desc_e[D_LIMR]);
  const int nc = (int)desc_e[D_NC];
  const uint64_t base = ((uint64_t)desc_e[D_BHI] << 32) | desc_e[D_BLO];
  G tm[NG];
  bool valid[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int g = lane + 32 * k;
    valid[k] = g < gpr;
    tm[k] = template_granule<U, CONNECT>(g, lim, nc, base, C, FastDiv(dCv), lut);
  }
  G* p = reinterpret_cast<G*>(envdst) + lane;
  uint32_t ow = desc_e[D_OWNED];
  for (int s = 0; s < N; ++s, p += gpr) {
    if (s && (s & 31) == 0) ow = desc_e[D_OWNED + (s >> 5)];
    if ((ow >> (s & 31)) & 1u) {  // warp-uniform
#pragma unroll
      for (int k = 0; k < NG; ++k)
        if (valid[k]) st_out(p + 32 * k, tm[k]);
    } else {
#pragma unroll
      for (int k = 0; k < NG; ++k)
        if (valid[k]) st_out(p + 32 * k, Gran<U>::zero());
    }
  }
}

// Short rows (at most 16 granules): 32 / gpr rows per store instruction.
template <int U, bool CONNECT>
__device__ __noinline__ void warp_rowmask_packed(int8_t* envdst, int row_len, int N, int C, cbx_fastdiv dCv, const uint32_t* desc_e,
                                                 const uint2* lut, int lane) {
  typedef typename Gran<U>::T G;
  const int gpr = row_len / U;
  const int rpi = 32 / gpr;  // rows per iteration
  const int r = lane / gpr, g = lane - r * gpr;
  const bool active = r < rpi;
  const int lim = (int)(CONNECT ? desc_e[D_LIMC] : desc_e[D_LIMR]);
  const uint64_t base = ((uint64_t)desc_e[D_BHI] << 32) | desc_e[D_BLO];
  const G tm = template_granule<U, CONNECT>(g, lim, (int)desc_e[D_NC], base, C, FastDiv(dCv), lut);
  G* p = reinterpret_cast<G*>(envdst) + lane;  // row r, granule g of the first group == granule index lane
  for (int s = r; s < N + r; s += rpi, p += rpi * gpr) {
    if (active && s < N) {
      const bool own = (desc_e[D_OWNED + (s >> 5)] >> (s & 31)) & 1u;
      st_out(p, own ? tm : Gran<U>::zero());
    }
  }
}

template <int U, bool CONNECT>
__device__ __forceinline__ void warp_rowmask_u(int8_t* envdst, int row_len, int N, int C, cbx_fastdiv dC, const uint32_t* de,
                                               const uint2* lut, int lane) {
  const int gpr = row_len / U;
  if (gpr <= 16) warp_rowmask_packed<U, CONNECT>(envdst, row_len, N, C, dC, de, lut, lane);
  else if (gpr <= 32) warp_rowmask<U, 1, CONNECT>(envdst, row_len, N, C, dC, de, lut, lane);
  else if (gpr <= 64) warp_rowmask<U, 2, CONNECT>(envdst, row_len, N, C, dC, de, lut, lane);
  else warp_rowmask<U, 4, CONNECT>(envdst, row_len, N, C, dC, de, lut, lane);
}

template <bool CONNECT>
__device__ __forceinline__ void warp_rowmask_dispatch(int unit, int8_t* envdst, int row_len, int N, int C, cbx_fastdiv dC,
                                                      const uint32_t* de, const uint2* lut, int lane) {
  if (unit == 16) warp_rowmask_u<16, CONNECT>(envdst, row_len, N, C, dC, de, lut, lane);
  else if (unit == 8) warp_rowmask_u<8, CONNECT>(envdst, row_len, N, C, dC, de, lut, lane);
  else warp_rowmask_u<4, CONNECT>(envdst, row_len, N, C, dC, de, lut, lane);
}

// ---- compile-time dimension sets ---------------------------------------------------------------------------------------
// The encoder's loop bounds are the scenario / bounds dimensions.  For MARLon's canonical configurations they are baked in
// at compile time (every row loop unrolls to bare stores with immediate offsets); DimsDyn reads them from the layout.
struct DimsDyn {
  static constexpr bool kStatic = false;
  static constexpr int N = 0, L = 0, R = 0, P = 0, C = 0, NPROPS = 0, LEAK = 0, NN = 0, NSVC = 0;
};
template <int N_, int L_, int R_, int P_, int C_, int NPROPS_, int LEAK_, int NN_, int NSVC_>
struct DimsStatic {
  static constexpr bool kStatic = true;
  static constexpr int N = N_, L = L_, R = R_, P = P_, C = C_, NPROPS = NPROPS_, LEAK = LEAK_, NN = NN_, NSVC = NSVC_;
};
// ToyCtf with MARLon's bounds (ppo/train_marl.py:12-14: 12 nodes, 10 credentials): 7 ports, 3 local / 8 remote ids, 10 props
typedef DimsStatic<12, 3, 8, 7, 10, 10, 5, 10, 13> DimsToyCtf;
// Chain size 10 at (12, 12): 8 ports, 5 local / 2 remote ids, 14 properties, 12 nodes, 22 services
typedef DimsStatic<12, 5, 2, 8, 12, 14, 5, 12, 22> DimsChain10;
#define CBX_DIM(D, name, dyn) (D::kStatic ? (int)D::name : (int)(dyn))

// Static-dimension row masks: everything but the env's counts and owned bits is a compile-time constant.
template <class D, bool CONNECT>
__device__ __forceinline__ void rowmask_static(int8_t* envdst, const uint32_t* desc_e, const uint2* lut, int lane) {
  constexpr int ROW = CONNECT ? D::N * D::P * D::C : D::N * D::R;
  constexpr int U = (ROW % 16 == 0) ? 16 : (ROW % 8 == 0) ? 8 : 4;
  constexpr int GPR = ROW / U;
  constexpr int C = CONNECT ? D::C : 1;
  typedef typename Gran<U>::T G;
  const int lim = (int)(CONNECT ? desc_e[D_LIMC] : desc_e[D_LIMR]);
  const int nc = (int)desc_e[D_NC];
  const uint64_t base = ((uint64_t)desc_e[D_BHI] << 32) | desc_e[D_BLO];
  const uint32_t ow = desc_e[D_OWNED];  // D::N <= 32
  auto granule = [&](int g) -> G {
    const int w = g * U;
    uint32_t m = lowmask(min(max(lim - w, 0), U));
    if (CONNECT) {
      const int ph = w % C;
      uint32_t pm;
      if (C <= 48) pm = (uint32_t)(base >> ph) & 0xFFFFu;
      else pm = lowmask(min(max(nc - ph, 0), 16)) | (lowmask(min(max(C - ph + nc, 0), 16)) & ~lowmask(min(max(C - ph, 0), 16)));
      m &= pm;
    }
    return Gran<U>::expand(m, lut);
  };
  if constexpr (GPR <= 16) {
    constexpr int RPI = 32 / GPR;
    const int r = lane / GPR, g = lane - r * GPR;
    const G tm = granule(g);
    G* p = reinterpret_cast<G*>(envdst) + lane;
    if (r < RPI) {
#pragma unroll
      for (int s0 = 0; s0 < D::N; s0 += RPI) {
        const int s = s0 + r;
        if (s < D::N) st_out(p + s0 * GPR, ((ow >> s) & 1u) ? tm : Gran<U>::zero());
      }
    }
  } else {
    constexpr int NG = (GPR + 31) / 32;
    G tm[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) tm[k] = granule(lane + 32 * k);
    G* p = reinterpret_cast<G*>(envdst) + lane;
#pragma unroll
    for (int s = 0; s < D::N; ++s) {
      if ((ow >> s) & 1u) {  // warp-uniform
#pragma unroll
        for (int k = 0; k < NG; ++k)
          if (32 * k + 32 <= GPR || lane + 32 * k < GPR) st_out(p + s * GPR + 32 * k, tm[k]);
      } else {
#pragma unroll
        for (int k = 0; k < NG; ++k)
          if (32 * k + 32 <= GPR || lane + 32 * k < GPR) st_out(p + s * GPR + 32 * k, Gran<U>::zero());
      }
    }
  }
}

// (wid, nw): this warp's index within the group of warps that share the tile's encoding, and the size of that group
template <class D>
__device__ __forceinline__ void encode_attacker_by_warp(const Tile& t, const Target& o, int n_valid, const EnvMask& enc_mask,
                                                        int wid, int nw) {
  const cbx_layout* L = t.L;
  const cbx_enc_consts& K = *t.K;
  const int lane = threadIdx.x & 31;
  const int N = CBX_DIM(D, N, L->N), NL = CBX_DIM(D, L, L->L), NR = CBX_DIM(D, R, L->R), NP = CBX_DIM(D, P, L->P);
  const int NC = CBX_DIM(D, C, L->C), NPROPS = CBX_DIM(D, NPROPS, L->nprops), LEAK = CBX_DIM(D, LEAK, L->LEAK);
  const int PW = D::kStatic ? (D::NPROPS + 31) / 32 : L->PW;
  for (int e = wid; e < n_valid; e += nw) {
    if (!enc_mask.test((int)e)) continue;
    const uint32_t* de = t.desc + e * t.DW;
    const uint32_t nd = de[D_ND], nc = de[D_NC];
    const bool blank = de[D_KIND] == OBS_BLANK;
    const int skip = CBX_SKIP(K);
    if (!(skip & 1)) {
    if (lane < 8) st_out(o.scalars + e * 8 + lane, (int32_t)t.g(e, STG_SCALARS + lane));
#pragma unroll
    for (int w = lane; w < 4 * LEAK; w += 32) st_out(o.leaked + e * 4 * LEAK + w, w < 4 * L->LEAKS ? (int32_t)leak_field(t.g(e, L->g_leaked + (w >> 2)), w & 3) : 0);
#pragma unroll
    for (int w = lane; w < 2 * NC; w += 32) {
      const int c = w >> 1;
      uint32_t val = 0;
      if (!blank && c < (int)nc) {
        uint32_t tr = (t.w(e, L->o_cache + (c >> 1)) >> ((c & 1) * 16)) & 0xFFFFu;
        const uint32_t* rec = t.tb + t.tb[CBX_H_OFF_TRIPLE] + 3 * tr;
        val = (w & 1) ? rec[1] : t.byte(e, L->o_disc_idx, (int)rec[0]);
      }
      st_out(o.cachem + e * 2 * NC + w, (int32_t)val);
    }
    const int npw = N * NPROPS;
#pragma unroll
    for (int w = lane; w < npw; w += 32) {
      uint32_t val = 2u;
      if (!blank) {
        uint32_t k = D::kStatic ? (uint32_t)w / (uint32_t)(D::kStatic ? D::NPROPS : 1) : FastDiv(K.d_nprops).div((uint32_t)w);
        uint32_t pi = w - k * NPROPS;
        val = 0u;
        if (k < nd) {
          uint32_t node = t.byte(e, L->o_disc_order, (int)k);
          val = (t.w(e, L->o_props + node * PW + (pi >> 5)) >> (pi & 31)) & 1u;
        }
      }
      st_out(o.props + e * npw + w, (int32_t)val);
    }
#pragma unroll
    for (int w = lane; w < N; w += 32) {
      uint32_t val = 0;
      if (!blank && w < (int)nd) {
        uint32_t node = t.byte(e, L->o_disc_order, w);
        val = (t.g(e, L->g_priv + (node >> 4)) >> ((node & 15) * 2)) & 3u;
      }
      st_out(o.priv + e * N + w, (int32_t)val);
    }
    }
    if (o.local) {
      const int sz_local = N * NL;
      if (!(skip & 2)) {
      int8_t* dst = o.local + (size_t)e * sz_local;
#pragma unroll
      for (int q = lane; q < sz_local / 4; q += 32) {  // fast path guarantees sz_local % 4 == 0
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint32_t i = q * 4 + b;
          uint32_t s = D::kStatic ? i / (uint32_t)(D::kStatic ? D::L : 1) : FastDiv(K.d_L).div(i);
          uint32_t v = i - s * NL;
          if ((de[D_OWNED + (s >> 5)] >> (s & 31)) & 1u) {
            uint32_t node = t.byte(e, L->o_disc_order, (int)s);
            word |= (t.tb[t.tb[CBX_H_OFF_VULN] + (node * (NL + NR) + v) * CBX_VULN_WORDS] & 1u) << (8 * b);
          }
        }
        st_out(reinterpret_cast<uint32_t*>(dst) + q, word);
      }
      }
      if constexpr (D::kStatic) {
        if (!(skip & 4)) rowmask_static<D, false>(o.remote + (size_t)e * (D::N * D::N * D::R), de, t.lut, lane);
        if (!(skip & 8)) rowmask_static<D, true>(o.connect + (size_t)e * (D::N * D::N * D::P * D::C), de, t.lut, lane);
      } else {
        warp_rowmask_dispatch<false>(K.tmpl_unit_r, o.remote + (size_t)e * L->sz_remote, N * NR, N, 1, cbx_fastdiv{0u, 0u}, de, t.lut, lane);
        warp_rowmask_dispatch<true>(K.tmpl_unit_c, o.connect + (size_t)e * L->sz_connect, N * NP * NC, N, NC, K.d_C, de, t.lut, lane);
      }
    }
  }
}

template <class D>
__device__ __forceinline__ void encode_defender_by_warp(const Tile& t, const Target& o, int n_valid, const EnvMask& enc_mask,
                                                        bool static_too, int wid, int nw) {
  const cbx_layout* L = t.L;
  const int lane = threadIdx.x & 31;
  const int n = CBX_DIM(D, NN, L->n), nsvc = CBX_DIM(D, NSVC, L->nservices);
  // the scenario's own node / service counts (smaller than the layout's in a padded multi-scenario batch: zero fill)
  const int n_own = D::kStatic ? n : (int)t.tb[CBX_H_N_NODES], nsvc_own = D::kStatic ? nsvc : (int)t.tb[CBX_H_N_SERVICES];
  const int OW = D::kStatic ? (D::N + 31) / 32 : L->OW;
  for (int e = wid; e < n_valid; e += nw) {
    if (!enc_mask.test((int)e)) continue;
    const uint32_t* di = t.desc + e * t.DW + D_OWNED + OW;
    if (CBX_SKIP(*t.K) & 16) continue;
#pragma unroll
    for (int i = lane; i < n; i += 32) o.infected[(size_t)e * n + i] = (int8_t)((di[i >> 5] >> (i & 31)) & 1u);
    if (!static_too) continue;
#pragma unroll
    for (int i = lane; i < 6 * n; i += 32) {
      const int node = i / 6, r = i - node * 6;
      o.fw_in[(size_t)e * 6 * n + i] = (int8_t)t.fw_rule_bit(e, node, r, false, n_own);
      o.fw_out[(size_t)e * 6 * n + i] = (int8_t)t.fw_rule_bit(e, node, r, true, n_own);
    }
#pragma unroll
    for (int i = lane; i < nsvc; i += 32) o.services[(size_t)e * nsvc + i] = (int8_t)(i < nsvc_own);
  }
}

// Encode the attacker observation of the envs selected by enc_mask (bit e = env e of the tile).
// ENC: 0 generic flat encoder | 1 warp-per-env, runtime dimensions | 2 warp-per-env ToyCtf(12,10) | 3 warp-per-env Chain-10(12,12)
template <int ENC>
__device__ __forceinline__ void encode_attacker(const Tile& t, const Target& o, int n_valid, const EnvMask& enc_mask,
                                                int wid = threadIdx.x >> 5, int nw = CBX_THREADS / 32) {
  const cbx_layout* L = t.L;
  const cbx_enc_consts& K = *t.K;
  if (ENC == 1) { encode_attacker_by_warp<DimsDyn>(t, o, n_valid, enc_mask, wid, nw); return; }
  if (ENC == 2) { encode_attacker_by_warp<DimsToyCtf>(t, o, n_valid, enc_mask, wid, nw); return; }
  if (ENC == 3) { encode_attacker_by_warp<DimsChain10>(t, o, n_valid, enc_mask, wid, nw); return; }
  write_i32(o.scalars, 8, FastDiv(0u, 3u), n_valid, enc_mask, [&](int e, int wi) { return t.g(e, STG_SCALARS + wi); });
  write_i32(o.leaked, 4 * L->LEAK, K.d_leaked, n_valid, enc_mask, [&](int e, int wi) { return wi < 4 * L->LEAKS ? leak_field(t.g(e, L->g_leaked + (wi >> 2)), wi & 3) : 0u; });
  write_i32(o.cachem, 2 * L->C, K.d_cachem, n_valid, enc_mask, [&](int e, int wi) -> uint32_t {
    int c = wi >> 1;
    if (t.d(e, D_KIND) == OBS_BLANK || c >= (int)t.d(e, D_NC)) return 0u;
    uint32_t tr = (t.w(e, L->o_cache + (c >> 1)) >> ((c & 1) * 16)) & 0xFFFFu;
    const uint32_t* rec = t.tb + t.tb[CBX_H_OFF_TRIPLE] + 3 * tr;
    return (wi & 1) ? rec[1] : t.byte(e, L->o_disc_idx, (int)rec[0]);
  });
  write_i32(o.props, L->N * L->nprops, K.d_props, n_valid, enc_mask, [&](int e, int wi) -> uint32_t {
    if (t.d(e, D_KIND) == OBS_BLANK) return 2u;
    uint32_t k = FastDiv(K.d_nprops).div((uint32_t)wi), p = wi - k * L->nprops;
    if (k >= t.d(e, D_ND)) return 0u;
    uint32_t node = t.byte(e, L->o_disc_order, (int)k);
    return (t.w(e, L->o_props + node * L->PW + (p >> 5)) >> (p & 31)) & 1u;
  });
  write_i32(o.priv, L->N, K.d_priv, n_valid, enc_mask, [&](int e, int wi) -> uint32_t {
    if (t.d(e, D_KIND) == OBS_BLANK || wi >= (int)t.d(e, D_ND)) return 0u;
    uint32_t node = t.byte(e, L->o_disc_order, wi);
    return (t.g(e, L->g_priv + (node >> 4)) >> ((node & 15) * 2)) & 3u;
  });
  write_i8(o.local, L->sz_local, K.d_local, n_valid, enc_mask, [&](int e, int i) -> uint32_t {
    uint32_t s = FastDiv(K.d_L).div((uint32_t)i), v = i - s * L->L;
    if (!t.owned(e, (int)s)) return 0u;
    uint32_t node = t.byte(e, L->o_disc_order, (int)s);
    return t.tb[t.tb[CBX_H_OFF_VULN] + (node * (L->L + L->R) + v) * CBX_VULN_WORDS] & 1u;
  });
  write_rowmask<false>(o.remote, L->sz_remote, K.d_remote, L->N * L->R, K.d_rowr, 1, FastDiv(0u, 0u), n_valid, enc_mask, t);
  write_rowmask<true>(o.connect, L->sz_connect, K.d_connect, L->N * L->P * L->C, K.d_rowc, L->C, K.d_C, n_valid, enc_mask, t);
}

template <int ENC>
__device__ __forceinline__ void encode_defender(const Tile& t, const Target& o, int n_valid, const EnvMask& enc_mask, bool static_too,
                                                int wid = threadIdx.x >> 5, int nw = CBX_THREADS / 32) {
  const cbx_layout* L = t.L;
  const cbx_enc_consts& K = *t.K;
  if (ENC == 1) { encode_defender_by_warp<DimsDyn>(t, o, n_valid, enc_mask, static_too, wid, nw); return; }
  if (ENC == 2) { encode_defender_by_warp<DimsToyCtf>(t, o, n_valid, enc_mask, static_too, wid, nw); return; }
  if (ENC == 3) { encode_defender_by_warp<DimsChain10>(t, o, n_valid, enc_mask, static_too, wid, nw); return; }
  write_i8(o.infected, L->n, K.d_n, n_valid, enc_mask,
           [&](int e, int i) -> uint32_t { return (t.d(e, D_OWNED + L->OW + (i >> 5)) >> (i & 31)) & 1u; });
  if (!static_too) return;
  const int n_own = (int)t.tb[CBX_H_N_NODES], nsvc_own = (int)t.tb[CBX_H_N_SERVICES];
  write_i8(o.fw_in, 6 * L->n, K.d_6n, n_valid, enc_mask, [&](int e, int i) -> uint32_t {
    int node = i / 6, r = i - node * 6;
    return t.fw_rule_bit(e, node, r, false, n_own);
  });
  write_i8(o.fw_out, 6 * L->n, K.d_6n, n_valid, enc_mask, [&](int e, int i) -> uint32_t {
    int node = i / 6, r = i - node * 6;
    return t.fw_rule_bit(e, node, r, true, n_own);
  });
  write_i8(o.services, L->nservices, K.d_svc, n_valid, enc_mask, [&](int, int i) -> uint32_t { return i < nsvc_own ? 1u : 0u; });
}

// ---- the warp writes one env's row of an int32 field (cbx_wide.cuh, cbx_pipe.cuh: gather thread-per-env, expand warp-per-env) ----
// `f(w0)` returns words [w0, w0 + 4) of the row (w0 a multiple of 4); words at or beyond wpe are not stored.  The row's
// alignment decides the store width: 16 bytes when wpe is a multiple of 4 words, 8 when even, else 4.
template <class F>
__device__ __forceinline__ void emit_row(int32_t* row, const int w_begin, const int w_end, const int wpe, const int lane, F f) {
  for (int w0 = w_begin + 4 * lane; w0 < w_end; w0 += 128) {
    const uint4 v = f(w0);
    if ((wpe & 3) == 0) {
      *reinterpret_cast<uint4*>(row + w0) = v;
    } else if ((wpe & 1) == 0) {
      *reinterpret_cast<uint2*>(row + w0) = make_uint2(v.x, v.y);
      if (w0 + 2 < wpe) *reinterpret_cast<uint2*>(row + w0 + 2) = make_uint2(v.z, v.w);
    } else {
      row[w0] = (int32_t)v.x;
      if (w0 + 1 < wpe) row[w0 + 1] = (int32_t)v.y;
      if (w0 + 2 < wpe) row[w0 + 2] = (int32_t)v.z;
      if (w0 + 3 < wpe) row[w0 + 3] = (int32_t)v.w;
    }
  }
}

// copy env rows main -> terminal buffers (terminal observation of an intercepted-and-truncated step is the previous one)
static __device__ void copy_rows(void* dst, const void* src, int bpe, int n_valid, const EnvMask& mask, int tidx = threadIdx.x,
                          int nthreads = CBX_THREADS) {
  if (!dst || !src || bpe == 0) return;
  for (int e = 0; e < n_valid; ++e) {
    if (!mask.test(e)) continue;
    const uint8_t* s = (const uint8_t*)src + (size_t)e * bpe;
    uint8_t* d = (uint8_t*)dst + (size_t)e * bpe;
    for (int k = tidx; k < bpe; k += nthreads) d[k] = s[k];
  }
}

static __device__ Target make_target(const cbx_views& v, const cbx_layout& L, int64_t e0, bool term) {
  Target o;
  const int64_t N = L.N;
  if (!term) {
    o.scalars = v.scalars + e0 * 8;
    o.leaked = v.leaked_credentials + e0 * 4 * L.LEAK;
    o.cachem = v.credential_cache_matrix + e0 * 2 * L.C;
    o.props = v.discovered_nodes_properties + e0 * N * L.nprops;
    o.priv = v.nodes_privilegelevel + e0 * N;
    o.local = v.local_vulnerability ? v.local_vulnerability + e0 * L.sz_local : nullptr;
    o.remote = v.remote_vulnerability ? v.remote_vulnerability + e0 * L.sz_remote : nullptr;
    o.connect = v.connect ? v.connect + e0 * (int64_t)L.sz_connect : nullptr;
    o.infected = v.def_infected_nodes ? v.def_infected_nodes + e0 * L.n : nullptr;
    o.fw_in = v.def_incoming_firewall ? v.def_incoming_firewall + e0 * 6 * L.n : nullptr;
    o.fw_out = v.def_outgoing_firewall ? v.def_outgoing_firewall + e0 * 6 * L.n : nullptr;
    o.services = v.def_services_status ? v.def_services_status + e0 * L.nservices : nullptr;
    o.owned_bits = v.owned_bits + e0 * L.OW;
  } else {
    o.scalars = v.term_scalars ? v.term_scalars + e0 * 8 : nullptr;
    o.leaked = v.term_leaked_credentials ? v.term_leaked_credentials + e0 * 4 * L.LEAK : nullptr;
    o.cachem = v.term_credential_cache_matrix ? v.term_credential_cache_matrix + e0 * 2 * L.C : nullptr;
    o.props = v.term_discovered_nodes_properties ? v.term_discovered_nodes_properties + e0 * N * L.nprops : nullptr;
    o.priv = v.term_nodes_privilegelevel ? v.term_nodes_privilegelevel + e0 * N : nullptr;
    o.local = v.term_local_vulnerability ? v.term_local_vulnerability + e0 * L.sz_local : nullptr;
    o.remote = v.term_remote_vulnerability ? v.term_remote_vulnerability + e0 * L.sz_remote : nullptr;
    o.connect = v.term_connect ? v.term_connect + e0 * (int64_t)L.sz_connect : nullptr;
    o.infected = v.term_def_infected_nodes ? v.term_def_infected_nodes + e0 * L.n : nullptr;
    o.fw_in = o.fw_out = o.services = nullptr;
    o.owned_bits = nullptr;
  }
  return o;
}

// env-major descriptor for the encoder, built by the env's logic thread
static __device__ void build_desc(const Ctx& c, uint32_t* d, int DW, const uint32_t* def_inst_override) {
  const cbx_layout* L = c.L;
  const uint32_t kind = c.g(STG_OBS_KIND);
  const int nd = c.nd(), nc = c.nc();
  d[D_ND] = (uint32_t)nd;
  d[D_NC] = (uint32_t)nc;
  d[D_KIND] = kind;
  d[D_LIMR] = (uint32_t)(nd * L->R);
  d[D_LIMC] = (uint32_t)(nd * L->P * L->C);
  uint64_t base = 0;
  if (L->C <= 48) {  // bit b set iff (b mod C) < nc: replicate the first period by doubling
    base = nc >= 64 ? ~0ull : ((1ull << nc) - 1ull);
    for (int sh = L->C; sh < 64; sh <<= 1) base |= base << sh;
  }
  d[D_BLO] = (uint32_t)base;
  d[D_BHI] = (uint32_t)(base >> 32);
  d[D_FLAGS] = 0;
  for (int k = 0; k < L->OW; ++k) d[D_OWNED + k] = 0;
  if (kind != OBS_BLANK) {
    for (int s = 0; s < nd; ++s) {
      uint32_t node = c.byte(L->o_disc_order, s);
      if ((c.g(L->g_inst + (node >> 5)) >> (node & 31)) & 1u) d[D_OWNED + (s >> 5)] |= 1u << (s & 31);
    }
  }
  for (int k = 0; k < L->Wn; ++k) d[D_OWNED + L->OW + k] = def_inst_override ? def_inst_override[k] : c.w(L->o_installed + k);
  if (def_inst_override) d[D_FLAGS] |= 1u;  // the defender's observation shows the fresh environment (firewall rows too, live binding)
  (void)DW;
}

// one action element: the arrays are int32, or int16 for the compact host format (cbx_params.act_i16)
__device__ __forceinline__ int32_t load_act(const int32_t* base, const int64_t idx, const int i16) {
  return i16 ? (int32_t)reinterpret_cast<const int16_t*>(base)[idx] : base[idx];
}

struct Acc {  // per-thread episode statistics, reduced once per CTA
  double v[CBX_STAT_COUNT];
};

// AttackerEnvWrapper.step (ATT:255-398) for one env; the VecEnv auto-reset is applied later by the caller
static __device__ void attacker_wrapper_step(const Ctx& c, const cbx_params& p, const int32_t* aa, int slice_of_kind[3], Acc& acc) {
  const cbx_layout* L = c.L;
  const cbx_config* cfg = c.cfg;
  int32_t info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // action[0] outside 0..2 or a negative coordinate cannot come out of the MultiDiscrete space; the reference would raise for
  // that one env.  Here such an action is the wrapper-level invalid action (_action_in_discovered_range false), never an index.
  const bool kind_ok = aa[0] >= 0 && aa[0] <= 2;
  int kind = cfg->kind_of_index[kind_ok ? aa[0] : 0];
  const int32_t* coords = aa + slice_of_kind[kind];
  const int ndisc = c.nd();
  bool in_range = kind_ok && coords[0] >= 0 && coords[0] < ndisc &&
                  (kind == CBX_KIND_LOCAL || (coords[1] >= 0 && coords[1] < ndisc));  // ATT:233-253
  double reward_modifier = 0.0, reward = 0.0, cyber_reward = 0.0;
  int terminated = 0, truncated = 0;
  if (!in_range) {
    c.w(L->o_att_invalid) += 1;
    reward_modifier += cfg->att_invalid_action_reward_modifier;
    info[5] = 1;
    c.g(STG_OBS_KIND) = OBS_KEEP;
  } else {
    c.w(L->o_att_valid) += 1;
    StepOut so = c.cyber_step(kind, coords, p.scan_u, p.detect_u);
    reward = cyber_reward = so.reward;
    terminated = so.terminated;
    info[0] = (int32_t)__float_as_uint((float)so.reward);
    info[1] = (int32_t)__float_as_uint((float)so.raw);
    info[2] = so.outcome;
    info[3] = so.error;
  }
  c.setflag(HDR_HAS_CYBER, true);
  c.setf32(L->o_last_cyber, (float)reward);
  c.w(L->o_att_ts) += 1;
  if (c.flag(HDR_ATT_RR)) truncated = 1;
  if ((int)c.w(L->o_att_ts) >= cfg->att_max_timesteps) truncated = 1;
  reward = reward + reward_modifier;
  c.setflag(HDR_HAS_REWARD, true);
  c.setf32(L->o_last_reward, (float)reward);
  c.setf32(L->o_att_return, c.f32(L->o_att_return) + (float)reward);
  info[4] = (int32_t)c.w(L->o_stepcount);
  p.v.att_reward[c.env] = (float)reward;
  p.v.att_terminated[c.env] = (uint8_t)terminated;
  p.v.att_truncated[c.env] = (uint8_t)truncated;
  if (p.host_results) {  // cbx_batch_step_host: the caller's page-locked result block, written in place over PCIe
    uint8_t* h = p.host_results;
    reinterpret_cast<float*>(h)[c.env] = (float)reward;
    h[8 * p.n_envs + c.env] = (uint8_t)terminated;
    h[9 * p.n_envs + c.env] = (uint8_t)truncated;
  }
  p.v.network_availability[c.env] = c.live_availability();
  acc.v[CBX_STAT_ENV_STEPS] += 1;
  const int done = terminated || truncated;
  c.g(STG_ATT_DONE) = (uint32_t)done;
  if (done) {
    info[6] = (int32_t)c.w(L->o_att_ts);
    double ret = (double)c.f32(L->o_att_return);
    int len = (int)c.w(L->o_att_ts);
    acc.v[CBX_STAT_EPISODES] += 1;
    acc.v[CBX_STAT_ATT_RETURN] += ret;
    acc.v[CBX_STAT_ATT_RETURN_SQ] += ret * ret;
    acc.v[CBX_STAT_EP_LEN] += len;
    acc.v[CBX_STAT_EP_LEN_SQ] += (double)len * len;
    acc.v[CBX_STAT_ATT_VALID] += c.w(L->o_att_valid);
    acc.v[CBX_STAT_ATT_INVALID] += c.w(L->o_att_invalid);
    if (terminated && cyber_reward == cfg->winning_reward) acc.v[CBX_STAT_ATT_WINS] += 1;
    if (!terminated && (int)c.w(L->o_att_ts) >= cfg->att_max_timesteps) acc.v[CBX_STAT_TIMEOUTS] += 1;
  }
  int4* ip = reinterpret_cast<int4*>(p.v.att_info + c.env * 8);
  ip[0] = make_int4(info[0], info[1], info[2], info[3]);
  ip[1] = make_int4(info[4], info[5], info[6], info[7]);
}

// DefenderEnvWrapper.step (DWR:197-327) with LearningDefender.executeAction on the stale copy (LDF:31-107)
static __device__ void defender_wrapper_step(const Ctx& c, const cbx_params& p, const int32_t* da, Acc& acc) {
  const cbx_layout* L = c.L;
  const cbx_config* cfg = c.cfg;
  double reward = 0.0;
  int term = 0, trunc = 0;
  const bool empty = da[0] < 0;
  const bool valid = empty ? true : c.defender_action_valid(da);
  if (!valid) { c.w(L->o_def_invalid) += 1; reward += cfg->def_invalid_action_reward; }
  else c.w(L->o_def_valid) += 1;
  int down;
  if (c.fx) {
    // LIVE binding: LearningDefender.executeAction on the environment the attacker plays in (LDF:31-107).  First
    // on_attacker_step_taken() of the live actuator -- the only place it ticks in MARLon, CyberBattleEnv.step calls it only for
    // a built-in defender agent (ENV:1156-1158) -- so availability is the one BEFORE this action
    down = c.tick(L->o_cd_live, L->o_notrunning);
    if (valid && !empty) {
      if (da[0] == 0) c.reimage_live(da[1]);                       // ACT:700-712 on the live env
      else if (da[0] == 1) c.fw_block(da[2], da[4] != 0, da[3]);  // LDF:50-58
      else if (da[0] == 2) c.fw_allow(da[5], da[7] != 0, da[6]);  // LDF:60-69 (stop / start service: no-ops)
    }
  } else {
    down = c.tick(L->o_cd_shadow, -1);  // on_attacker_step_taken() of the stale actuator: availability BEFORE the action
    if (valid && !empty && da[0] == 0) c.setbyte(L->o_cd_shadow, da[1], 16);
  }
  uint32_t a = c.w(L->o_avail);
  const int prev_down = (int)((a >> 16) & 0xFFu);
  const double cur = c.availability(down), prev = c.availability(prev_down);
  const double worsening = prev - cur;
  if (c.flag(HDR_HAS_CYBER)) reward += -1.0 * (double)c.f32(L->o_last_cyber);
  if (cur < cfg->maintain_sla) {
    if (!c.flag(HDR_BREACHED)) {
      reward += cfg->def_loss_reward;
      if (cfg->def_reset_on_constraint_broken) term = 1;
      c.setflag(HDR_BREACHED, true);
      acc.v[CBX_STAT_SLA_BREACHES] += 1;
    } else if (worsening > 0) reward += -cfg->def_sla_worsening_penalty_scale * worsening;
  } else c.setflag(HDR_BREACHED, false);
  // bits 0-7: the live env's count at its last tick (info["network_availability"]) -- under the live binding that is this tick
  c.w(L->o_avail) = (c.fx ? (uint32_t)down : (a & 0xFFu)) | ((uint32_t)down << 8) | ((uint32_t)down << 16);
  if (c.defender_goal_reached()) { reward = cfg->winning_reward; term = 1; }
  c.w(L->o_def_ts) += 1;
  if (c.flag(HDR_DEF_RR)) { trunc = 1; reward = -1.0 * (double)c.f32(L->o_last_att); }
  else if ((int)c.w(L->o_def_ts) >= cfg->def_max_timesteps) trunc = 1;
  c.setf32(L->o_def_return, c.f32(L->o_def_return) + (float)reward);
  p.v.def_reward[c.env] = (float)reward;
  p.v.def_terminated[c.env] = (uint8_t)term;
  p.v.def_truncated[c.env] = (uint8_t)trunc;
  if (p.host_results) {
    uint8_t* h = p.host_results;
    reinterpret_cast<float*>(h + 4 * p.n_envs)[c.env] = (float)reward;
    h[10 * p.n_envs + c.env] = (uint8_t)term;
    h[11 * p.n_envs + c.env] = (uint8_t)trunc;
  }
  const int done = term || trunc;
  c.g(STG_DEF_DONE) = (uint32_t)done;
  if (done) {
    double ret = (double)c.f32(L->o_def_return);
    acc.v[CBX_STAT_DEF_RETURN] += ret;
    acc.v[CBX_STAT_DEF_RETURN_SQ] += ret * ret;
    acc.v[CBX_STAT_DEF_VALID] += c.w(L->o_def_valid);
    acc.v[CBX_STAT_DEF_INVALID] += c.w(L->o_def_invalid);
    for (int k = 0; k < L->Wn; ++k) c.g(STG_DEF_TERM_INST + k) = c.w(L->o_installed + k);
  }
}

static __device__ void cyber_only_step(const Ctx& c, const cbx_params& p, const int32_t* a, Acc& acc) {
  const cbx_layout* L = c.L;
  StepOut so = c.cyber_step(a[0], a + 1, p.scan_u, p.detect_u);
  int4* ip = reinterpret_cast<int4*>(p.v.att_info + c.env * 8);
  int done = so.terminated && so.error != CBX_E_STEP_AFTER_DONE;
  ip[0] = make_int4((int)__float_as_uint((float)so.reward), (int)__float_as_uint((float)so.raw), so.outcome, so.error);
  ip[1] = make_int4((int)c.w(L->o_stepcount), 0, done ? (int)c.w(L->o_stepcount) : 0, 0);
  p.v.att_reward[c.env] = (float)so.reward;
  p.v.att_terminated[c.env] = (uint8_t)so.terminated;
  p.v.att_truncated[c.env] = 0;
  p.v.network_availability[c.env] = c.live_availability();
  c.g(STG_ATT_DONE) = (uint32_t)done;
  if (so.error == CBX_E_STEP_AFTER_DONE) return;
  acc.v[CBX_STAT_ENV_STEPS] += 1;
  c.setf32(L->o_att_return, c.f32(L->o_att_return) + (float)so.reward);
  if (done) {
    double ret = (double)c.f32(L->o_att_return);
    int len = (int)c.w(L->o_stepcount);
    acc.v[CBX_STAT_EPISODES] += 1;
    acc.v[CBX_STAT_ATT_RETURN] += ret;
    acc.v[CBX_STAT_ATT_RETURN_SQ] += ret * ret;
    acc.v[CBX_STAT_EP_LEN] += len;
    acc.v[CBX_STAT_EP_LEN_SQ] += (double)len * len;
    if (so.reward == c.cfg->winning_reward) acc.v[CBX_STAT_ATT_WINS] += 1;
  }
}

// ---- the two game-logic phases of one env (shared by the fused and the pipelined kernel) --------------------------------
// Phase 1: the attacker's move (or an explicit reset / notify_reset).  `aa` = this env's attacker action words.
__device__ __forceinline__ void logic_phase1(const Ctx& c, const cbx_params& p, const int op, const int32_t* aa, const uint32_t* s_init,
                                             int slice_of_kind[3], Acc& acc) {
  const cbx_layout& L = p.lay;
  const cbx_config& cfg = p.cfg;
  const bool reset_only = op & CBX_OP_RESET, who_att = op & CBX_OP_ATTACKER, who_def = op & CBX_OP_DEFENDER;
  const bool marlon = cfg.mode == CBX_MODE_MARLON;
  c.g(STG_ATT_DONE) = 0; c.g(STG_DEF_DONE) = 0;
  if (reset_only) {
    c.g(STG_OBS_KIND) = OBS_KEEP;
    if (op & CBX_OP_NOTIFY) {
      if (!p.reset_mask || p.reset_mask[c.env]) {
        if (who_att) c.setflag(HDR_ATT_RR, true);
        if (who_def) { c.setflag(HDR_DEF_RR, true); c.setf32(L.o_last_att, p.notify_last_reward); }
      }
    } else if (!p.reset_mask || p.reset_mask[c.env]) {
      if (marlon) {  // attacker.reset() then defender.reset(), either or both
        if (who_att) c.attacker_reset(s_init);
        if (who_def && cfg.def_enabled) c.defender_reset(s_init);
      } else {
        c.cyber_reset(s_init);
        c.setf32(L.o_att_return, 0.f);
      }
      if (who_att || !marlon) {
        c.stage_reset_obs();
        p.v.att_reward[c.env] = 0.f; p.v.att_terminated[c.env] = 0; p.v.att_truncated[c.env] = 0;
        int4* ip = reinterpret_cast<int4*>(p.v.att_info + c.env * 8);
        ip[0] = make_int4(0, 0, 0, 0); ip[1] = make_int4(0, 0, 0, 0);
      }
      if (who_def) { p.v.def_reward[c.env] = 0.f; p.v.def_terminated[c.env] = 0; p.v.def_truncated[c.env] = 0; }
      p.v.network_availability[c.env] = 1.0;
    }
  } else if (marlon) {
    if (who_att) attacker_wrapper_step(c, p, aa, slice_of_kind, acc);
    else c.g(STG_OBS_KIND) = OBS_KEEP;
  } else {
    cyber_only_step(c, p, aa, acc);
  }
}

// Phase 2: the attacker's auto-reset, the defender's move, the encoder descriptor.  `da` = this env's defender action words.
// Returns whether the defender finished an episode that the VecEnv protocol resets (deferred: its observation comes first).
__device__ __forceinline__ uint32_t logic_phase2(const Ctx& c, const cbx_params& p, const int op, const int32_t* da, const uint32_t* s_init,
                                                 uint32_t* desc_e, Acc& acc) {
  const cbx_layout& L = p.lay;
  const cbx_config& cfg = p.cfg;
  const bool reset_only = op & CBX_OP_RESET;
  const bool marlon = cfg.mode == CBX_MODE_MARLON;
  const bool def_on = marlon && cfg.def_enabled && (op & CBX_OP_DEFENDER);
  if (!reset_only) {
    if (c.g(STG_ATT_DONE) && cfg.auto_reset) {
      if (marlon) c.attacker_reset(s_init);
      else { c.cyber_reset(s_init); c.setf32(L.o_att_return, 0.f); }
      c.stage_reset_obs();
    }
    if (def_on) defender_wrapper_step(c, p, da, acc);
  }
  const uint32_t def_done = c.g(STG_DEF_DONE) && cfg.auto_reset;
  // main defender observation: after an auto-reset it shows the fresh environment (DWR:477)
  build_desc(c, desc_e, p.enc.desc_words, def_done ? s_init + L.o_installed : nullptr);
  if (c.g(STG_OBS_KIND) != OBS_KEEP) {
    uint32_t* ob = p.v.owned_bits + c.env * L.OW;
    for (int k = 0; k < L.OW; ++k) ob[k] = desc_e[D_OWNED + k];
  }
  return def_done;
}

}  // namespace cbx

#endif  // CBX_SHARED_CUH_
