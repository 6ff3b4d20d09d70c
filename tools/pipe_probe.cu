// Probe: is IMAD.WIDE.U32.X half rate inherently or latency bound?  Plus ALU co-issue checks.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
constexpr int INNER = 32;

template<int CH> __global__ void k_widex(uint32_t* out, uint32_t seed, int trips) {
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
// carry-out only + addc consumer per MAC
template<int CH> __global__ void k_wide_cout(uint32_t* out, uint32_t seed, int trips) {
  uint32_t lo[CH], hi[CH], cy[CH]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) { lo[c] = threadIdx.x + c; hi[c] = c; cy[c] = 0; }
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(a), "r"(b));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(cy[c]));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= lo[c] ^ hi[c] ^ cy[c];
  if (s == 0x1234567) out[0] = s;
}
// IMAD.WIDE (no carry) + NALU alu ops per 2 MACs
template<int CH, int NALU> __global__ void k_wide_alu(uint32_t* out, uint32_t seed, int trips) {
  uint64_t acc[CH]; uint32_t x[CH]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) { acc[c] = threadIdx.x + c; x[c] = c + seed; }
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
        if ((c & 1) == 0) {
#pragma unroll
          for (int u = 0; u < NALU; u++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
        }
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c] ^ x[c];
  if (s == 0x1234567) out[0] = (uint32_t)s;
}
// mad.hi.u32
template<int CH> __global__ void k_madhi(uint32_t* out, uint32_t seed, int trips) {
  uint32_t acc[CH]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = threadIdx.x + c;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c];
  if (s == 0x1234567) out[0] = s;
}
template<int CH> __global__ void k_shfl(uint32_t* out, uint32_t seed, int trips) {
  uint32_t acc[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = threadIdx.x * 7 + c + seed;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        uint32_t r; asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(acc[c]), "r"(acc[c] & 31));
        acc[c] = r + 1;
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c];
  if (s == 0x1234567) out[0] = s;
}
// 64-bit shift-right + add + and : the radix-2^30 normalisation step cost (3 ops / column)
template<int CH> __global__ void k_norm(uint32_t* out, uint32_t seed, int trips) {
  uint64_t acc[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = ((uint64_t)(threadIdx.x + c) << 33) + seed;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < INNER; i++) {
#pragma unroll
      for (int c = 0; c + 1 < CH; c++) { acc[c + 1] += acc[c] >> 30; acc[c] &= 0x3fffffffu; }
      acc[0] += acc[CH - 1] * 3;
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= acc[c];
  if (s == 0x1234567) out[0] = (uint32_t)s;
}

typedef void (*kern_t)(uint32_t*, uint32_t, int);
int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  uint32_t* out; CK(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct M { const char* name; kern_t k; double ops; int threads; int bps; };
  M modes[] = {
    {"widex_ch8_w8", k_widex<8>, 8.0 * INNER, 256, 4},
    {"widex_ch8_w16", k_widex<8>, 8.0 * INNER, 512, 4},
    {"widex_ch8_w4", k_widex<8>, 8.0 * INNER, 128, 4},
    {"widex_ch8_w2", k_widex<8>, 8.0 * INNER, 64, 4},
    {"widex_ch8_w1", k_widex<8>, 8.0 * INNER, 128, 1},
    {"widex_ch2_w8", k_widex<2>, 2.0 * INNER, 256, 4},
    {"widex_ch16_w8", k_widex<16>, 16.0 * INNER, 256, 4},
    {"wide_cout_addc_w8", k_wide_cout<8>, 8.0 * INNER, 256, 4},
    {"wide_alu0_w8", k_wide_alu<8, 0>, 8.0 * INNER, 256, 4},
    {"wide_alu1per2_w8", k_wide_alu<8, 1>, 8.0 * INNER, 256, 4},
    {"wide_alu2per2_w8", k_wide_alu<8, 2>, 8.0 * INNER, 256, 4},
    {"wide_alu3per2_w8", k_wide_alu<8, 3>, 8.0 * INNER, 256, 4},
    {"wide_alu0_w2", k_wide_alu<8, 0>, 8.0 * INNER, 64, 4},
    {"wide_alu0_w1", k_wide_alu<8, 0>, 8.0 * INNER, 128, 1},
    {"madhi_w8", k_madhi<8>, 8.0 * INNER, 256, 4},
    {"shfl_w8", k_shfl<8>, 8.0 * INNER, 256, 4},
    {"norm3op_cols_w8", k_norm<8>, 7.0 * INNER, 256, 4},
  };
  printf("{");
  bool first = true;
  for (auto& m : modes) {
    double best = 0; const int trips = 400;
    for (int rep = 0; rep < 4; rep++) {
      CK(cudaEventRecord(e0));
      m.k<<<sms * m.bps, m.threads>>>(out, 12345u + rep, trips);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double rate = (double)sms * m.bps * m.threads * trips * m.ops / (ms * 1e-3);
      if (rep > 0 && rate > best) best = rate;
    }
    printf("%s\"%s\":%.3f", first ? "" : ",", m.name, best / 1e12); first = false;
  }
  printf("}\n");
  return 0;
}
