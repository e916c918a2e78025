// csrc/tpt_render_fast.cu -- FAST instantiation. Compiled with -use_fast_math (FMA contraction,
// approximate div/sqrt/sin/cos, flush-to-zero): same estimator, same Philox stream, fp32 only.
#define TPT_PAR false
#define TPT_SUFFIX fast
#include "tpt_kernels.cuh"

namespace tptd {
// FP32 issue-rate probe: the roofline denominator of the render kernels, measured on the device the
// bench runs on instead of derived from the data sheet. 16 independent FMA chains per thread (enough
// to cover the 4-cycle FMA latency at any occupancy), 1184 CTAs x 256 threads = 8 per SM.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float a, float b, int iters, float *sink) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = a + (float)(threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 12345.678f) sink[0] = s; // never true for the probe's arguments: keeps the chains alive
}

cudaError_t launch_fp32_peak_probe(int blocks, int iters, float *sink, cudaStream_t st) {
  fp32_peak_kernel<<<blocks, 256, 0, st>>>(0.999f, 0.001f, iters, sink);
  return cudaGetLastError();
}
} // namespace tptd
