// Integer-pipe peak microbenchmark for sm_100a.
// Measures warp-instruction issue rates of the instructions the Montgomery
// kernels are built from, all SMs busy, many independent chains per thread.
// Output: one JSON object on stdout.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

constexpr int CH = 8;        // independent chains per thread
constexpr int INNER = 64;    // unrolled ops per chain per loop trip

// mode 0: IMAD.WIDE.U32 (64-bit accumulate, no carry)
__global__ void k_wide(uint32_t* out, uint32_t seed, int trips) {
  uint64_t acc[CH]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = (uint64_t)(threadIdx.x + c) << 13;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c];
  if (s == 0x1234567) out[0] = (uint32_t)s;
}

// mode 1: IMAD.WIDE.U32.X carry chains (mad.lo.cc + madc.hi.cc pairs), CH links per chain
__global__ void k_widex(uint32_t* out, uint32_t seed, int trips) {
  uint32_t lo[CH], hi[CH]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x, top = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) { lo[c] = threadIdx.x + c; hi[c] = c; }
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
      asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[0]) : "r"(a), "r"(b));
      asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[0]) : "r"(a), "r"(b));
#pragma unroll
      for (int c = 1; c < CH; c++) {
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(a), "r"(b));
      }
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(top));
    }
  }
  uint32_t s = top;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= lo[c] ^ hi[c];
  if (s == 0x1234567) out[0] = s;
}

// mode 2: IMAD (32-bit lo)
__global__ void k_imad(uint32_t* out, uint32_t seed, int trips) {
  uint32_t acc[CH]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = threadIdx.x + c;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++)
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c];
  if (s == 0x1234567) out[0] = s;
}

// mode 3: DFMA
__global__ void k_dfma(uint32_t* out, uint32_t seed, int trips) {
  double acc[CH]; double a = 1.0 + 1e-9 * (seed + threadIdx.x), b = 1e-12 * blockIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = threadIdx.x + c;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++)
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[c]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += acc[c];
  if (s == 0.1234567) out[0] = 1;
}

// mode 4: SHFL.IDX
__global__ void k_shfl(uint32_t* out, uint32_t seed, int trips) {
  uint32_t acc[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = threadIdx.x + c + seed;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++) acc[c] = __shfl_sync(0xffffffffu, acc[c], (c + i) & 31);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c];
  if (s == 0x1234567) out[0] = s;
}

// mode 5: IADD3 with carry chain (add.cc/addc.cc)
__global__ void k_iadd(uint32_t* out, uint32_t seed, int trips) {
  uint32_t acc[CH]; uint32_t a = seed + threadIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = threadIdx.x + c;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(acc[0]) : "r"(a));
#pragma unroll
      for (int c = 1; c < CH; c++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc[c]) : "r"(a));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c];
  if (s == 0x1234567) out[0] = s;
}

// mode 6: mixed: 2 IMAD.WIDE.X chains interleaved with SHFL + IADD (row-like mix: 16 wide + 2 shfl + 4 iadd)
__global__ void k_mix(uint32_t* out, uint32_t seed, int trips) {
  uint32_t lo[CH], hi[CH]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x, top = 0, sh = threadIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) { lo[c] = threadIdx.x + c; hi[c] = c; }
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER / 2; i++) {
#pragma unroll
      for (int rep = 0; rep < 2; rep++) {
        asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[0]) : "r"(a), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[0]) : "r"(a), "r"(b));
#pragma unroll
        for (int c = 1; c < CH; c++) {
          asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(a), "r"(b));
        }
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(top));
      }
      sh = __shfl_sync(0xffffffffu, sh, (i + 1) & 31);
      b = __shfl_sync(0xffffffffu, b ^ sh, i & 31);
      top += sh;
    }
  }
  uint32_t s = top;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= lo[c] ^ hi[c];
  if (s == 0x1234567) out[0] = s;
}

typedef void (*kern_t)(uint32_t*, uint32_t, int);

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  uint32_t* out; CK(cudaMalloc(&out, 64));
  struct { const char* name; kern_t k; double ops_per_thread_trip; } modes[] = {
    {"imad_wide_u32", k_wide, (double)CH * INNER},
    {"imad_wide_u32_x_chain", k_widex, (double)CH * INNER},
    {"imad_lo_u32", k_imad, (double)CH * INNER},
    {"dfma", k_dfma, (double)CH * INNER},
    {"shfl_idx", k_shfl, (double)CH * INNER},
    {"iadd3_x_chain", k_iadd, (double)CH * INNER},
    {"mix_wide_x_plus_shfl", k_mix, (double)CH * INNER},
  };
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz_attr\":%d", p.name, sms, clk_khz);
  const int threads = 256, blocks_per_sm = 4, trips = 200;
  for (auto& m : modes) {
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
      CK(cudaEventRecord(e0));
      m.k<<<sms * blocks_per_sm, threads>>>(out, 12345u + rep, trips);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double ops = (double)sms * blocks_per_sm * threads * trips * m.ops_per_thread_trip;
      double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best) best = rate;
    }
    printf(",\"%s_tops\":%.4f", m.name, best / 1e12);
  }
  // sustained: 2 s of IMAD.WIDE.X chains
  {
    int launches = 0; CK(cudaEventRecord(e0));
    float ms = 0;
    do { for (int i = 0; i < 20; i++) k_widex<<<sms * blocks_per_sm, threads>>>(out, 7u, trips); launches += 20;
         CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); } while (ms < 2000.f);
    double ops = (double)launches * sms * blocks_per_sm * threads * trips * CH * INNER;
    printf(",\"imad_wide_u32_x_chain_sustained_tops\":%.4f", ops / (ms * 1e-3) / 1e12);
  }
  printf("}\n");
  return 0;
}
