// peer.cu — multi-GPU result exchange WITHOUT a collective kernel.
// An orbit sweep / tiled frame ends with "rank 0 holds every rank's RGBA".  The cast kernel is a persistent
// launch that fills every SM, so an NCCL gather kernel hardly finds room beside it: gather and render serialise
// (measured: 8 GPUs, 1.20 ms per step instead of 0.98).  Instead every rank's SHADE kernel stores its RGBA
// straight into rank 0's HBM through NVLink peer memory (a CUDA-IPC mapping of one buffer rank 0 owns), and
// ranks hand frames over with two stream-ordered flag kernels:
//     peer r, frame k :  wait(released >= k-1)  ->  render (shade writes slot k&1 of rank 0)  ->  signal(arrived[r] = k+1)
//     rank 0, frame k :  ... same ...           ->  wait(arrived[0..N) >= k+1)  ->  consume  ->  signal(released = k+1)
// No SM is taken away from the traversal kernel, no data crosses NVLink twice, nothing is staged.
#include "common.cuh"

namespace {

constexpr unsigned long long WAIT_TIMEOUT_NS = 5ull * 1000 * 1000 * 1000;  // a lost peer must not hang the GPU

__global__ void signal_kernel(uint32_t* flag, uint32_t value) {
  __threadfence_system();  // everything this stream wrote before (the frame) is visible before the flag
  asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(flag), "r"(value) : "memory");
}

__global__ void wait_geq_kernel(const uint32_t* flags, uint32_t n, uint32_t value, uint32_t* timed_out) {
  const uint32_t i = threadIdx.x;
  if (i >= n) return;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
    if (v >= value) break;
    __nanosleep(500);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > WAIT_TIMEOUT_NS) { *timed_out = 1u; break; }
  }
  __threadfence_system();
}

}  // namespace

J3DG_API int j3dg_peer_alloc(j3dg_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char* handle_out) {
  if (!ctx || !dev_ptr || !bytes) return J3DG_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == J3DG_IPC_HANDLE_BYTES, "IPC handle size");
  cudaSetDevice(ctx->device);
  *dev_ptr = nullptr;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return j3dg_cuda_fail(ctx, e, "cudaMalloc", __FILE__, __LINE__);
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess && handle_out) {
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
  }
  if (e != cudaSuccess) { cudaFree(p); return j3dg_cuda_fail(ctx, e, "cudaIpcGetMemHandle", __FILE__, __LINE__); }
  *dev_ptr = p;
  return J3DG_OK;
}

J3DG_API int j3dg_peer_free(j3dg_ctx* ctx, void* dev_ptr) {
  if (!ctx) return J3DG_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  CU_CHECK(ctx, cudaFree(dev_ptr));
  return J3DG_OK;
}

J3DG_API int j3dg_peer_open(j3dg_ctx* ctx, const unsigned char* handle, void** dev_ptr) {
  if (!ctx || !handle || !dev_ptr) return J3DG_EINVAL;
  cudaSetDevice(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  *dev_ptr = nullptr;
  CU_CHECK(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return J3DG_OK;
}

J3DG_API int j3dg_peer_close(j3dg_ctx* ctx, void* dev_ptr) {
  if (!ctx) return J3DG_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  CU_CHECK(ctx, cudaIpcCloseMemHandle(dev_ptr));
  return J3DG_OK;
}

J3DG_API int j3dg_stream_signal(j3dg_ctx* ctx, uint32_t* flag, uint32_t value) {
  if (!ctx || !flag) return J3DG_EINVAL;
  cudaSetDevice(ctx->device);
  { const int st = j3dg_check_sticky(ctx); if (st != J3DG_OK) return st; }
  signal_kernel<<<1, 1, 0, ctx->stream>>>(flag, value);
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}

J3DG_API int j3dg_stream_wait_geq(j3dg_ctx* ctx, const uint32_t* flags, uint32_t n, uint32_t value) {
  if (!ctx || !flags || !n || n > 32) { j3dg_set_error(ctx, "j3dg_stream_wait_geq: 1..32 flags"); return J3DG_EINVAL; }
  cudaSetDevice(ctx->device);
  { const int st = j3dg_check_sticky(ctx); if (st != J3DG_OK) return st; }
  wait_geq_kernel<<<1, 32, 0, ctx->stream>>>(flags, n, value, ctx->d_status);
  KERNEL_CHECK(ctx);
  return J3DG_OK;
}

J3DG_API int j3dg_stream_wait_status(j3dg_ctx* ctx, int* timed_out) {
  if (!ctx || !timed_out) return J3DG_EINVAL;
  cudaSetDevice(ctx->device);
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  *timed_out = (int)ctx->h_status[0];  // sticky: cleared by j3dg_ctx_status(reset) only
  return J3DG_OK;
}
