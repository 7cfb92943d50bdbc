// peak.cu — the FP64 roofline denominator measured on the device in use: a DFMA microkernel (independent chains, no
// memory traffic), timed with CUDA events.  bench.py reports the FP64 fraction of the dominant kernel against this
// number ("of measured-in-run"); MEASURED_PEAKS.json only carries the HBM and bf16 tensor peaks.
#include <cuda_runtime.h>

#include "../../include/relxill_b200.h"

namespace {
constexpr int PK_CHAINS = 8, PK_ITERS = 4096, PK_NT = 256;
__global__ void __launch_bounds__(PK_NT) k_fp64_peak(double *out, double a, double b) {
  double x[PK_CHAINS];
#pragma unroll
  for (int c = 0; c < PK_CHAINS; c++) x[c] = (double) (threadIdx.x + c) * 1e-3;
#pragma unroll 4
  for (int it = 0; it < PK_ITERS; it++) {
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < PK_CHAINS; c++) s += x[c];
  if (s == 123.456) out[0] = s;   // never true for these inputs: keeps the chains alive
}
}  // namespace

extern "C" double relxill_b200_measure_fp64_peak(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1.0;
  double *d = nullptr;
  if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 16;
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    k_fp64_peak<<<grid, PK_NT>>>(d, 0.999999, 1e-7);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * PK_CHAINS * (double) PK_ITERS * PK_NT * grid;
    if (rep > 0 && ms > 0.f) best = flop / (ms * 1e-3) / 1e12 > best ? flop / (ms * 1e-3) / 1e12 : best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}
