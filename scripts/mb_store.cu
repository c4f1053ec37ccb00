// mb_store.cu -- store-path microbenchmark for the mask encoder (ToyCtf (12,10): connect [12][840] + remote [12][96] int8 per env).
// Which way of emitting "row s = owned[s] ? template : 0" streams closest to the HBM write peak?
//   V0 rows8   : warp per env, 8-byte stores row by row (rows are 840 B = 8 mod 16) -- what the step kernel does today
//   V1 pair16  : warp per env, 16-byte stores over ROW PAIRS (1680 B = 105 granules), all stores 16-byte aligned
//   V2 tma     : warp per env, template row pairs built in shared memory, one TMA bulk store per row pair
//   V3 flat    : pure 16-byte fill of the same byte count (upper bound of the st.global path)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/mb_store scripts/mb_store.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int N = 12, P = 7, C = 10, R = 8;
constexpr int ROWC = N * P * C;  // 840
constexpr int ROWR = N * R;      // 96
constexpr int SZC = N * ROWC, SZR = N * ROWR;

struct Desc { uint32_t nd, nc, owned, pad; };

__device__ __forceinline__ uint32_t lowmask(int k) { return k <= 0 ? 0u : (k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u)); }
__device__ __forceinline__ uint2 expand8(uint32_t m) {
  return make_uint2(((m & 0xF) * 0x00204081u) & 0x01010101u, (((m >> 4) & 0xF) * 0x00204081u) & 0x01010101u);
}
__device__ __forceinline__ uint64_t period_base(int nc) {
  uint64_t base = (1ull << nc) - 1ull;
  for (int sh = C; sh < 64; sh <<= 1) base |= base << sh;
  return base;
}
// 8 bytes of the connect template row at byte offset w
__device__ __forceinline__ uint2 tmpl8_c(int w, int lim, uint64_t base) {
  uint32_t m = lowmask(min(max(lim - w, 0), 8));
  m &= (uint32_t)(base >> (w % C)) & 0xFFu;
  return expand8(m);
}
__device__ __forceinline__ uint2 tmpl8_r(int w, int lim) { return expand8(lowmask(min(max(lim - w, 0), 8))); }

template <int MODE>  // 0 default, 1 __stcs, 2 st.global.L1::no_allocate
__device__ __forceinline__ void st16(uint4* p, uint4 v) {
  if (MODE == 1) __stcs(p, v);
  else if (MODE == 2) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else *p = v;
}
template <int MODE>
__device__ __forceinline__ void st8(uint2* p, uint2 v) {
  if (MODE == 1) __stcs(p, v);
  else *p = v;
}

// remote mask, shared by V0/V1: 96-byte rows = 6 granules of 16 B, 5 rows per warp store
template <int MODE>
__device__ __forceinline__ void remote_rows(int8_t* dst, const Desc& d, int lane) {
  const int r = lane / 6, g = lane - r * 6;
  const int lim = d.nd * R;
  uint2 a = tmpl8_r(g * 16, lim), b = tmpl8_r(g * 16 + 8, lim);
  const uint4 tm = make_uint4(a.x, a.y, b.x, b.y);
  uint4* p = reinterpret_cast<uint4*>(dst) + lane;
  if (r < 5) {
#pragma unroll
    for (int s0 = 0; s0 < N; s0 += 5) {
      const int s = s0 + r;
      if (s < N) st16<MODE>(p + s0 * 6, ((d.owned >> s) & 1u) ? tm : make_uint4(0, 0, 0, 0));
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_rows8(int8_t* connect, int8_t* remote, const Desc* desc, int n_envs) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int e = blockIdx.x * wpb + (threadIdx.x >> 5); e < n_envs; e += gridDim.x * wpb) {
    const Desc d = desc[e];
    const uint64_t base = period_base(d.nc);
    const int lim = d.nd * P * C;
    uint2 tm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tm[k] = tmpl8_c((lane + 32 * k) * 8, lim, base);
    uint2* p = reinterpret_cast<uint2*>(connect + (size_t)e * SZC) + lane;
#pragma unroll
    for (int s = 0; s < N; ++s) {
      const bool own = (d.owned >> s) & 1u;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (32 * k + 32 <= 105 || lane + 32 * k < 105) st8<MODE>(p + s * 105 + 32 * k, own ? tm[k] : make_uint2(0, 0));
    }
    remote_rows<MODE>(remote + (size_t)e * SZR, d, lane);
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_pair16(int8_t* connect, int8_t* remote, const Desc* desc, int n_envs) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int e = blockIdx.x * wpb + (threadIdx.x >> 5); e < n_envs; e += gridDim.x * wpb) {
    const Desc d = desc[e];
    const uint64_t base = period_base(d.nc);
    const int lim = d.nd * P * C;
    // granule q of a row pair covers bytes [16q, 16q+16): halves lo / hi belong to row A (offset < 840) or row B
    uint4 tm[4];
    uint32_t loA = 0, hiA = 0;  // bit k: the lo / hi half of granule lane+32k lies in row A
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int q = lane + 32 * k;
      const int o0 = q * 16, o1 = q * 16 + 8;
      uint2 a = tmpl8_c(o0 < ROWC ? o0 : o0 - ROWC, lim, base), b = tmpl8_c(o1 < ROWC ? o1 : o1 - ROWC, lim, base);
      tm[k] = make_uint4(a.x, a.y, b.x, b.y);
      if (o0 < ROWC) loA |= 1u << k;
      if (o1 < ROWC) hiA |= 1u << k;
    }
    uint4* p = reinterpret_cast<uint4*>(connect + (size_t)e * SZC) + lane;
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      const uint32_t ownA = (d.owned >> (2 * j)) & 1u, ownB = (d.owned >> (2 * j + 1)) & 1u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (32 * k + 32 <= 105 || lane + 32 * k < 105) {
          const bool lo = ((loA >> k) & 1u) ? ownA : ownB, hi = ((hiA >> k) & 1u) ? ownA : ownB;
          uint4 v = tm[k];
          if (!lo) { v.x = 0; v.y = 0; }
          if (!hi) { v.z = 0; v.w = 0; }
          st16<MODE>(p + j * 105 + 32 * k, v);
        }
      }
    }
    remote_rows<MODE>(remote + (size_t)e * SZR, d, lane);
  }
}

// ---- TMA variant ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(K) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// per warp and buffer: [ZT: 0(840) T(840)] [TZ: T(840) 0(840)] [TT: T T] [remote 1152]; zeros written once.  NBUF buffers per warp.
constexpr int BUF_BYTES = 3 * 1680 + SZR;  // 6192
// lane 0 issues all seven bulk stores (one bulk group per env)
template <int NBUF>
__global__ void __launch_bounds__(128) k_tma1(int8_t* connect, int8_t* remote, const Desc* desc, int n_envs) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  uint8_t* zrow = smem;
  uint8_t* mybuf = smem + 1792 + (size_t)warp * NBUF * BUF_BYTES;
  for (int k = threadIdx.x; k < 1792 / 4; k += blockDim.x) reinterpret_cast<uint32_t*>(zrow)[k] = 0;
  for (int k = lane; k < NBUF * BUF_BYTES / 4; k += 32) reinterpret_cast<uint32_t*>(mybuf)[k] = 0;
  __syncthreads();
  int it = 0;
  for (int e = blockIdx.x * wpb + warp; e < n_envs; e += gridDim.x * wpb, ++it) {
    const Desc d = desc[e];
    const uint64_t base = period_base(d.nc);
    const int lim = d.nd * P * C;
    uint8_t* buf = mybuf + (it % NBUF) * BUF_BYTES;
    if (lane == 0) tma_wait_read<NBUF - 1>();
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = lane + 32 * k;
      if (g < 105) {
        const uint2 v = tmpl8_c(g * 8, lim, base);
        *reinterpret_cast<uint2*>(buf + 840 + g * 8) = v;
        *reinterpret_cast<uint2*>(buf + 1680 + g * 8) = v;
        *reinterpret_cast<uint2*>(buf + 3360 + g * 8) = v;
        *reinterpret_cast<uint2*>(buf + 3360 + 840 + g * 8) = v;
      }
    }
    {
      const int limr = d.nd * R;
      for (int g = lane; g < SZR / 16; g += 32) {
        const int s = g / 6, w = (g - s * 6) * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((d.owned >> s) & 1u) { uint2 a = tmpl8_r(w, limr), b = tmpl8_r(w + 8, limr); v = make_uint4(a.x, a.y, b.x, b.y); }
        *reinterpret_cast<uint4*>(buf + 5040 + g * 16) = v;
      }
    }
    fence_async();
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint32_t ownA = (d.owned >> (2 * j)) & 1u, ownB = (d.owned >> (2 * j + 1)) & 1u;
        const uint8_t* src = (ownA && ownB) ? buf + 3360 : ownA ? buf + 1680 : ownB ? buf : zrow;
        tma_store_1d(connect + (size_t)e * SZC + j * 1680, src, 1680);
      }
      tma_store_1d(remote + (size_t)e * SZR, buf + 5040, SZR);
      tma_commit();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128) k_flat(uint4* dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) st16<MODE>(dst + i, make_uint4(1, 1, 1, 1));
}

// warp-contiguous flat fill: each warp writes whole 11 232-byte "env" chunks like the encoders do (same locality as V0/V1)
template <int MODE>
__global__ void __launch_bounds__(128) k_flat_env(uint4* dst, int n_envs) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  constexpr int G = (SZC + SZR) / 16;  // 702
  for (int e = blockIdx.x * wpb + (threadIdx.x >> 5); e < n_envs; e += gridDim.x * wpb) {
    uint4* p = dst + (size_t)e * G + lane;
#pragma unroll
    for (int k = 0; k < (G + 31) / 32; ++k)
      if (lane + 32 * k < G) st16<MODE>(p + 32 * k, make_uint4(1, 1, 1, 1));
  }
}

static uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// CPU check of the produced masks
static int check(const std::vector<int8_t>& hc, const std::vector<int8_t>& hr, const std::vector<Desc>& d, int n) {
  int bad = 0;
  for (int e = 0; e < n && bad < 5; ++e) {
    for (int s = 0; s < N; ++s) {
      const bool own = (d[e].owned >> s) & 1u;
      for (int w = 0; w < ROWC; ++w) {
        int exp = own && w < (int)d[e].nd * P * C && (w % C) < (int)d[e].nc;
        if (hc[(size_t)e * SZC + s * ROWC + w] != exp) { if (bad < 5) printf("  connect mismatch e=%d s=%d w=%d got %d exp %d\n", e, s, w, hc[(size_t)e * SZC + s * ROWC + w], exp); ++bad; break; }
      }
      for (int w = 0; w < ROWR; ++w) {
        int exp = own && w < (int)d[e].nd * R;
        if (hr[(size_t)e * SZR + s * ROWR + w] != exp) { if (bad < 5) printf("  remote mismatch e=%d s=%d w=%d\n", e, s, w); ++bad; break; }
      }
    }
  }
  return bad;
}

int main(int argc, char** argv) {
  int n_envs = argc > 1 ? atoi(argv[1]) : 65536;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  std::vector<Desc> hd(n_envs);
  for (int e = 0; e < n_envs; ++e) {
    uint32_t h = hash32(e * 2654435761u + 12345u);
    uint32_t nd = 1 + h % 10, nc = (h >> 8) % 11;
    uint32_t owned = hash32(h) & hash32(h + 1) & ((1u << nd) - 1u);
    owned |= 1u;
    hd[e] = Desc{nd, nc, owned, 0};
  }
  Desc* dd; int8_t *dc, *dr;
  CK(cudaMalloc(&dd, sizeof(Desc) * n_envs));
  CK(cudaMemcpy(dd, hd.data(), sizeof(Desc) * n_envs, cudaMemcpyHostToDevice));
  // one allocation so that the flat fill covers exactly the same bytes
  const size_t bytes = (size_t)n_envs * (SZC + SZR);
  CK(cudaMalloc(&dc, bytes));
  dr = dc + (size_t)n_envs * SZC;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<int8_t> hc((size_t)n_envs * SZC), hr((size_t)n_envs * SZR);
  printf("n_envs %d  bytes/launch %.1f MB  SMs %d\n", n_envs, bytes / 1e6, sms);

  auto run = [&](const char* name, auto launch, bool verify) {
    CK(cudaMemset(dc, 0xEE, bytes));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e9, sum = 0;
    const int reps = 10;
    for (int i = 0; i < reps; ++i) {
      CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best; sum += ms;
    }
    CK(cudaGetLastError());
    int bad = -1;
    if (verify) {
      CK(cudaMemcpy(hc.data(), dc, hc.size(), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hr.data(), dr, hr.size(), cudaMemcpyDeviceToHost));
      bad = check(hc, hr, hd, n_envs);
    }
    printf("%-28s best %.4f ms  %7.1f GB/s   mean %.4f ms  %7.1f GB/s  %s\n", name, best, bytes / best / 1e6, sum / reps, bytes / (sum / reps) / 1e6,
           bad < 0 ? "" : bad ? "MISMATCH" : "ok");
  };
  for (int cps : {4, 8, 16}) {
    const int grid = sms * cps;
    char nm[64];
    printf("-- %d CTAs/SM x 128 threads\n", cps);
    snprintf(nm, 64, "flat16 default"); run(nm, [&] { k_flat<0><<<grid, 128>>>((uint4*)dc, bytes / 16); }, false);
    snprintf(nm, 64, "flat16 stcs"); run(nm, [&] { k_flat<1><<<grid, 128>>>((uint4*)dc, bytes / 16); }, false);
    snprintf(nm, 64, "flat16 no_allocate"); run(nm, [&] { k_flat<2><<<grid, 128>>>((uint4*)dc, bytes / 16); }, false);
    snprintf(nm, 64, "flat16 env-chunks"); run(nm, [&] { k_flat_env<0><<<grid, 128>>>((uint4*)dc, n_envs); }, false);
    snprintf(nm, 64, "V0 rows8 default"); run(nm, [&] { k_rows8<0><<<grid, 128>>>(dc, dr, dd, n_envs); }, true);
    snprintf(nm, 64, "V0 rows8 stcs"); run(nm, [&] { k_rows8<1><<<grid, 128>>>(dc, dr, dd, n_envs); }, false);
    snprintf(nm, 64, "V1 pair16 default"); run(nm, [&] { k_pair16<0><<<grid, 128>>>(dc, dr, dd, n_envs); }, true);
    snprintf(nm, 64, "V1 pair16 stcs"); run(nm, [&] { k_pair16<1><<<grid, 128>>>(dc, dr, dd, n_envs); }, false);
    snprintf(nm, 64, "V1 pair16 no_allocate"); run(nm, [&] { k_pair16<2><<<grid, 128>>>(dc, dr, dd, n_envs); }, false);
  }
  for (int cps : {1, 2, 4, 8}) {
    const int grid = sms * cps;
    printf("-- TMA, %d CTAs/SM x 128 threads\n", cps);
    {
      const int smem = 1792 + 4 * 1 * BUF_BYTES;
      CK(cudaFuncSetAttribute(k_tma1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      run("V2 tma lane0 1 buf", [&] { k_tma1<1><<<grid, 128, smem>>>(dc, dr, dd, n_envs); }, true);
    }
    {
      const int smem = 1792 + 4 * 2 * BUF_BYTES;
      CK(cudaFuncSetAttribute(k_tma1<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      run("V2 tma lane0 2 buf", [&] { k_tma1<2><<<grid, 128, smem>>>(dc, dr, dd, n_envs); }, true);
    }
    if (cps <= 4) {
      const int smem = 1792 + 4 * 4 * BUF_BYTES;
      CK(cudaFuncSetAttribute(k_tma1<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      run("V2 tma lane0 4 buf", [&] { k_tma1<4><<<grid, 128, smem>>>(dc, dr, dd, n_envs); }, true);
    }
  }
  return 0;
}
