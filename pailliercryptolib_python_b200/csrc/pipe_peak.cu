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
// FP64 pipe: 20 independent DFMA.RZ chains per thread, one multiplicand changing every step.
constexpr int DCH = 20;
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, double seed, int trips) {
  double acc[DCH], a[DCH];
#pragma unroll
  for (int c = 0; c < DCH; ++c) { acc[c] = c + 0.5; a[c] = 1.0 + 1e-9 * (threadIdx.x * 32 + c); }
  double y = 1.0 + 1e-7 * seed;
  for (int t = 0; t < trips; ++t) {
#pragma unroll
    for (int i = 0; i < INNER; ++i) {
      y = y * 1.0000001;
#pragma unroll
      for (int c = 0; c < DCH; ++c) acc[c] = __fma_rz(a[c], y, acc[c]);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < DCH; ++c) s += acc[c];
  if (s == 0.1234567) out[0] = s;
}

// The exact instruction mix of one 52x52-bit product of csrc/mont52.cuh (2 DFMA.RZ + 1 DADD + one 3-input 64-bit
// integer add = IADD3 + IADD3.X), 20 independent products per step, nothing else: what the Montgomery kernels could
// reach with no quotient digits, shuffles, shared memory or carry handling (same kernel as tools/dfma_mix_probe.cu).
__global__ void __launch_bounds__(256, 1) k_product_mix_peak(unsigned long long* out, double seed, int trips) {
  constexpr double TWO104 = 20282409603651670423947251286016.0, TWO104P52 = 20282409603651674927546878656512.0;
  double a[DCH];
  unsigned long long acc[DCH];
#pragma unroll
  for (int c = 0; c < DCH; ++c) { a[c] = 4503599627370495.0 - (threadIdx.x * 64 + c) * 1048577.0; acc[c] = c; }
  double y = 4503599627370401.0 - seed - threadIdx.x;
  for (int t = 0; t < trips; ++t) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      y -= 1025.0;
      double ph[DCH], pl[DCH];
#pragma unroll
      for (int c = 0; c < DCH; ++c) ph[c] = __fma_rz(a[c], y, TWO104);
#pragma unroll
      for (int c = 0; c < DCH; ++c) pl[c] = TWO104P52 - ph[c];
#pragma unroll
      for (int c = 0; c < DCH; ++c) pl[c] = __fma_rz(a[c], y, pl[c]);
      unsigned long long hprev = (unsigned long long)__double_as_longlong(ph[DCH - 1]);
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        acc[c] += (unsigned long long)__double_as_longlong(pl[c]) + hprev;
        hprev = (unsigned long long)__double_as_longlong(ph[c]);
      }
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int c = 0; c < DCH; ++c) s ^= acc[c];
  if (s == 0x123456789ull) out[0] = s;
}

template <class F> int time_best(int reps, double work, double* rate_out, F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
    cudaEventRecord(e0);
    launch(r);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = work / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *rate_out = best;
  return 0;
}
}  // namespace

extern "C" int phe_fp64_pipe_peak(int reps, double* dfma_per_s) {
  if (!dfma_per_s) return 1;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return 1;
  const int trips = 2000, grid = sms * 2, block = 256;   // 4 warps per SM sub-partition
  k_dfma_peak<<<grid, block>>>(d, 1.0, 20);
  const int rc = time_best(reps, (double)grid * block * (double)trips * INNER * DCH, dfma_per_s,
                           [&](int r) { k_dfma_peak<<<grid, block>>>(d, 3.0 + r, trips); });
  cudaFree(d);
  return rc;
}

extern "C" int phe_product_mix_peak(int reps, double* products_per_s) {
  if (!products_per_s) return 1;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return 1;
  const int trips = 2000, grid = sms, block = 256;   // 2 warps per SM sub-partition, 254 registers
  k_product_mix_peak<<<grid, block>>>(d, 1.0, 20);
  const int rc = time_best(reps, (double)grid * block * (double)trips * 8 * DCH, products_per_s,
                           [&](int r) { k_product_mix_peak<<<grid, block>>>(d, 3.0 + r, trips); });
  cudaFree(d);
  return rc;
}

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
