// cbx_pipe.cu -- translation unit of the pipelined, warp-specialised step kernel (cbx_pipe.cuh), stale defender binding, and
// the launch entry points (the live-binding instantiations are cbx_pipe_live.cu).
#include "cbx_shared.cuh"
#include "cbx_pipe.cuh"

extern "C" {
cudaError_t cbx_pipe_attrs_live(int enc, int smem_bytes);
cudaError_t cbx_launch_pipe_live(const cbx_params* p, int op, int grid, cudaStream_t stream);

// pipelined kernel: enc = 1 runtime dimensions, 2 ToyCtf(12,10), 3 Chain-10(12,12); live = the live defender binding
cudaError_t cbx_pipe_attrs(int enc, int smem_bytes, int live) {
  return live ? cbx_pipe_attrs_live(enc, smem_bytes) : cbx::pipe_attrs_t<false>(enc, smem_bytes);
}
cudaError_t cbx_launch_pipe(const cbx_params* p, int op, int grid, cudaStream_t stream) {
  return p->fwx_words ? cbx_launch_pipe_live(p, op, grid, stream) : cbx::launch_pipe_t<false>(p, op, grid, stream);
}
}
