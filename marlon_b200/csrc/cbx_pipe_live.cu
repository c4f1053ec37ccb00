// cbx_pipe_live.cu -- the pipelined step kernel (cbx_pipe.cuh) instantiated for the LIVE defender binding.
#include "cbx_shared.cuh"
#include "cbx_pipe.cuh"

extern "C" {
cudaError_t cbx_pipe_attrs_live(int enc, int smem_bytes) { return cbx::pipe_attrs_t<true>(enc, smem_bytes); }
cudaError_t cbx_launch_pipe_live(const cbx_params* p, int op, int grid, cudaStream_t stream) {
  return cbx::launch_pipe_t<true>(p, op, grid, stream);
}
}
