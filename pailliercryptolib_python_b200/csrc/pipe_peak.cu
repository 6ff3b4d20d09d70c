// pipe_peak.cu -- live measurement of the integer-pipe roofline denominator (phe_int_pipe_peak).
// The Montgomery kernels are built from IMAD.WIDE.U32 (32x32+64 multiply-accumulate); this kernel issues
// nothing else (checked in SASS): 24 independent accumulator chains per thread, 8 warps per SM sub-partition.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/phe_b200.h"

namespace {
// Same instruction pattern as one row of the Montgomery product: acc[j] += a[j] * b with the multiplier b
// changing every row, so ptxas cannot fold the products (the r01 first version of this probe used a loop-invariant
// product, which ptxas turned into IADD3 pairs and which over-stated the peak by 2x).
constexpr int CH = 24, INNER = 16;
__global__ void __launch_bounds__(512) k_imad_wide_peak(uint32_t* out, uint32_t seed, int trips) {
  uint64_t acc[CH];
  uint32_t a[CH];
  uint32_t b = seed * 3u + blockIdx.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < CH; ++c) { acc[c] = (uint64_t)(threadIdx.x + c) << 13; a[c] = seed * (2 * c + 5) + threadIdx.x; }
  for (int t = 0; t < trips; ++t) {
#pragma unroll
    for (int i = 0; i < INNER; ++i) {
      b = __funnelshift_l(b, b, 7) ^ seed;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] += (uint64_t)a[c] * b;
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s ^= acc[c];
  if (s == 0x1234567ull) out[0] = (uint32_t)s;
}
}  // namespace

extern "C" int phe_int_pipe_peak(int reps, double* mac_per_s) {
  if (!mac_per_s) return 1;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint32_t* d = nullptr;
  if (cudaMalloc(&d, 4) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int trips = 1500, grid = sms * 2, block = 512;   // 8 warps per SM sub-partition
  k_imad_wide_peak<<<grid, block>>>(d, 1u, 50);   // warm-up
  double best = 0;
  for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
    cudaEventRecord(e0);
    k_imad_wide_peak<<<grid, block>>>(d, 7u + r, trips);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double macs = (double)grid * block * (double)trips * INNER * CH;
    const double rate = macs / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *mac_per_s = best;
  return 0;
}
