// peak_kernels.cu — on-box measurement of the INT32 ALU issue rate, the roofline denominator
// for the alignment kernel ("lanes/SM/clk must be measured on-box with an
// IADD3 / IMNMX / VIADDMNMX microbenchmark").  Not part of the hot path.
#include <cuda_runtime.h>

#include "sdgpu_internal.cuh"

namespace {

constexpr int ILP = 8;
constexpr int ITERS = 4096;

// which: 0 = add (IADD3/VIADD), 1 = max (VIMNMX), 2 = fused add+max (VIADDMNMX), 3 = max3 (VIMNMX3),
//        4 = max (ALU pipe) interleaved 1:1 with a multiply-add (IMAD, FMA pipe), both counted,
//        5 = fused add+max (ALU pipe) interleaved 1:1 with IMAD, both counted: the instruction mix of the DP kernels
template <int WHICH>
__global__ void __launch_bounds__(256) int32PeakKernel(int* out, int a, int b, int m) {
  if (WHICH >= 4) {
    // each chain alternates the two instruction kinds (x feeds y, y feeds x), so neither the maxima nor the
    // multiply-adds of consecutive iterations can be merged; m is a run-time 1 the compiler cannot see
    int x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k, y[k] = threadIdx.x - k;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        if (WHICH == 4) x[k] = max(x[k], y[k]);
        if (WHICH == 5) x[k] = __viaddmax_s32(x[k], a, y[k]);
        y[k] = y[k] * m + x[k];
      }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k] ^ y[k];
    if (s == 0x7fffffff) out[0] = s;
    return;
  }
  int x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (WHICH == 0) x[k] = x[k] + a;
      if (WHICH == 1) x[k] = max(x[k], b - x[k]);
      if (WHICH == 2) x[k] = __viaddmax_s32(x[k], a, b);
      if (WHICH == 3) x[k] = __vimax3_s32(x[k], a - x[k], b);
    }
    a ^= it;  // keep the loop from being folded
  }
  int s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  if (s == 0x7fffffff) out[0] = s;
}

}  // namespace

extern "C" int sdgpu_measure_int32_peak(sdgpu_ctx* ctx, int which, double* instrPerSec) {
  if (!ctx || !instrPerSec || which < 0 || which > 5) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = sd_ensure_stage(ctx, 256);
  if (rc) return rc;
  const int grid = ctx->numSMs * 8, block = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    SD_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    int* o = reinterpret_cast<int*>(ctx->dStage);
    if (which == 0) int32PeakKernel<0><<<grid, block, 0, ctx->stream>>>(o, 3, 5, ctx->numSMs > 0 ? 1 : 0);
    if (which == 1) int32PeakKernel<1><<<grid, block, 0, ctx->stream>>>(o, 3, 5, ctx->numSMs > 0 ? 1 : 0);
    if (which == 2) int32PeakKernel<2><<<grid, block, 0, ctx->stream>>>(o, 3, 5, ctx->numSMs > 0 ? 1 : 0);
    if (which == 3) int32PeakKernel<3><<<grid, block, 0, ctx->stream>>>(o, 3, 5, ctx->numSMs > 0 ? 1 : 0);
    if (which == 4) int32PeakKernel<4><<<grid, block, 0, ctx->stream>>>(o, 3, 5, ctx->numSMs > 0 ? 1 : 0);
    if (which == 5) int32PeakKernel<5><<<grid, block, 0, ctx->stream>>>(o, 3, 5, ctx->numSMs > 0 ? 1 : 0);
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    SD_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    SD_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (rep > 0 && ms < best) best = ms;
  }
  // thread-level instructions of the measured kind (WHICH 1 and 3 also issue one subtract each, not counted;
  // WHICH 4 and 5 count both instructions of each pair)
  const double n = double(grid) * block * double(ITERS) * ILP * (which >= 4 ? 2.0 : 1.0);
  *instrPerSec = n / (double(best) * 1e-3);
  return SDGPU_OK;
}
