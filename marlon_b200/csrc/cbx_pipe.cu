// cbx_pipe.cu -- translation unit of the pipelined, warp-specialised step kernel (cbx_pipe.cuh) and its launch helpers.
#include "cbx_shared.cuh"
#include "cbx_pipe.cuh"

extern "C" {
// pipelined kernel: enc = 1 runtime dimensions, 2 ToyCtf(12,10), 3 Chain-10(12,12)
cudaError_t cbx_pipe_attrs(int enc, int smem_bytes) {
  switch (enc) {
    case 3: return cudaFuncSetAttribute(cbx::cbx_pipe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    case 2: return cudaFuncSetAttribute(cbx::cbx_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    default: return cudaFuncSetAttribute(cbx::cbx_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  }
}
cudaError_t cbx_launch_pipe(const cbx_params* p, int op, int grid, cudaStream_t stream) {
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
      case 3: return cudaLaunchKernelEx(&cfg, cbx::cbx_pipe_kernel<3>, *p, op);
      case 2: return cudaLaunchKernelEx(&cfg, cbx::cbx_pipe_kernel<2>, *p, op);
      default: return cudaLaunchKernelEx(&cfg, cbx::cbx_pipe_kernel<1>, *p, op);
    }
  }
  switch (p->enc.warp_env) {
    case 3: cbx::cbx_pipe_kernel<3><<<grid, threads, p->pipe.total_bytes, stream>>>(*p, op); break;
    case 2: cbx::cbx_pipe_kernel<2><<<grid, threads, p->pipe.total_bytes, stream>>>(*p, op); break;
    default: cbx::cbx_pipe_kernel<1><<<grid, threads, p->pipe.total_bytes, stream>>>(*p, op); break;
  }
  return cudaGetLastError();
}
}
