// diag.cu -- device micro-measurements the benchmark reports beside its rooflines.
//
// The general-hexahedron fill is bound by the FP64 pipe, not by HBM; its ceiling is the DFMA
// rate of this part, which no vendor table states reliably for B200.  txasm_measure_fp64_peak times dependent-free
// DFMA chains (8 per thread, all SMs full) with CUDA events on the handle's stream.
#include "txasm_internal.hpp"

namespace txasm {

__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double seed, double *__restrict__ sink)
{
  double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double m = 1.0000001, c = 1e-9;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) sink[0] = s;       // never true: keeps the chains alive
}

}  // namespace txasm

using namespace txasm;

extern "C" int txasm_measure_fp64_peak(txasm_handle h, double *tflops)
{
  if (!h || !tflops) return TXASM_EINVAL;
  if (h->sticky) return TXASM_ECUDA;
  TX_CUDA(h, cudaSetDevice(h->device));
  double *sink = nullptr;
  TX_CUDA(h, cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 8192, grid = h->n_sm * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, h->stream);
    k_dfma_peak<<<grid, threads, 0, h->stream>>>(iters, 1.0 + rep, sink);
    cudaEventRecord(e1, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { cudaFree(sink); return cuda_fail(h, e, "k_dfma_peak", __FILE__, __LINE__); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * iters * (double)grid * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;     // first launch: warm-up
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = best;
  return TXASM_OK;
}
