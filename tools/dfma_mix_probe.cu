// dfma_mix_probe.cu -- what the inner loop of csrc/mont52.cuh can reach at best: the exact instruction mix of one
// 52x52-bit product (2 DFMA.RZ + 1 DADD + one 3-input 64-bit integer add = IADD3 + IADD3.X) with no quotient
// digit, no shuffles and no shared memory, CH independent chains per thread.  Output: one JSON object with the
// product rate per variant and warps per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dfma_mix_probe dfma_mix_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr double TWO104 = 20282409603651670423947251286016.0;
constexpr double TWO104P52 = 20282409603651674927546878656512.0;
constexpr int CH = 20, INNER = 8;

// MODE 0: full product mix (3-input adds).  MODE 1: FP64 part only (results xor-folded rarely).
// MODE 2: 2-input adds (4 IADD3 per product).  MODE 3: product mix but the hi halves are not accumulated (1 add of 2 inputs).
template <int MODE> __global__ void __launch_bounds__(256, 1) k(uint64_t* out, double seed, int trips) {
  double a[CH];
  uint64_t acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { a[c] = 4503599627370495.0 - (threadIdx.x * 64 + c) * 1048577.0; acc[c] = c; }
  double y = 4503599627370401.0 - seed - threadIdx.x;
  for (int t = 0; t < trips; ++t) {
#pragma unroll
    for (int i = 0; i < INNER; ++i) {
      y -= 1025.0;
      double ph[CH], pl[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) ph[c] = __fma_rz(a[c], y, TWO104);
#pragma unroll
      for (int c = 0; c < CH; ++c) pl[c] = TWO104P52 - ph[c];
#pragma unroll
      for (int c = 0; c < CH; ++c) pl[c] = __fma_rz(a[c], y, pl[c]);
      uint64_t hprev = (uint64_t)__double_as_longlong(ph[CH - 1]);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint64_t l = (uint64_t)__double_as_longlong(pl[c]), h = (uint64_t)__double_as_longlong(ph[c]);
        if (MODE == 0) acc[c] += l + hprev;
        if (MODE == 1) { if (c == 0) acc[0] ^= l ^ h; else acc[c] ^= (l & h) >> 63; }
        if (MODE == 2) { acc[c] += l; asm volatile("" : "+l"(acc[c])); acc[c] += hprev; }
        if (MODE == 3) acc[c] += l;
        hprev = h;
      }
      if (MODE == 3) acc[0] ^= hprev;
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s ^= acc[c];
  if (s == 0x123456789ull) out[0] = s;
}

template <int MODE> double run(uint64_t* d, int sms, int warps_per_smsp) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int block = 32 * 4 * warps_per_smsp, trips = 3000;
  k<MODE><<<sms, block>>>(d, 1.0, 10);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<MODE><<<sms, block>>>(d, 3.0 + r, trips); cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return (double)sms * block * trips * INNER * CH / (best * 1e-3) / 1e12;   // T products/s (lane products)
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  uint64_t* d; CK(cudaMalloc(&d, 64));
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"unit\":\"T lane-products/s (one 52x52 product = 2 DFMA + 1 DADD [+ adds])\"", p.name, sms);
  for (int w : {1, 2}) {
    printf(",\"full_mix_w%d\":%.3f", w, run<0>(d, sms, w));
    printf(",\"fp64_only_w%d\":%.3f", w, run<1>(d, sms, w));
    printf(",\"two_input_adds_w%d\":%.3f", w, run<2>(d, sms, w));
    printf(",\"lo_add_only_w%d\":%.3f", w, run<3>(d, sms, w));
  }
  printf("}\n");
  return 0;
}
