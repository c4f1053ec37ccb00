// cbx_wide.cu -- translation unit of the warp-per-tile step kernel for large per-env state (cbx_wide.cuh) and its launch helpers.
#define CBX_STATE_IN_PLACE 1  // cbx_device.cuh: loops over per-node state arrays keep several loads in flight
#include "cbx_shared.cuh"
#include "cbx_wide.cuh"

extern "C" {
// warp-per-tile kernel for large per-env state (factored masks)
cudaError_t cbx_wide_attrs(int smem_bytes) {
  return cudaFuncSetAttribute(cbx::cbx_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}
cudaError_t cbx_launch_wide(const cbx_params* p, int op, int grid, cudaStream_t stream) {
  cbx::cbx_wide_kernel<1><<<grid, p->wide.nwarps * 32, p->wide.total_bytes, stream>>>(*p, op);
  return cudaGetLastError();
}
}
