// p-adic pair engine instantiations: L = limbs of p, q (10: 1024-bit keys, 20: 2048-bit, 30: 3072-bit).
#include "phe_launch.cuh"
namespace phe {
extern const PairOps g_pair_10 = PairLaunch<10>::ops();
extern const PairOps g_pair_20 = PairLaunch<20>::ops();
extern const PairOps g_pair_30 = PairLaunch<30>::ops();

// ---- row gather / scatter of packed matrices (exponent alignment, matmul operand maps, rotate, slices) --------------
// dst[i] = src[idx[i]] (gather) or dst[idx[i]] = src[i] (scatter); rows of `words` u32 words, words % 4 == 0
__global__ void __launch_bounds__(256) k_rows_move(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                   const long long* __restrict__ idx, long long n, int words, int scatter) {
  const int q = words / 4;   // 16-byte pieces per row
  const long long total = n * q;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / q;
    const int k = (int)(t - i * q);
    const long long r = idx[i];
    const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)(scatter ? i : r) * words) + k;
    uint4* d = reinterpret_cast<uint4*>(dst + (size_t)(scatter ? r : i) * words) + k;
    *d = *s;
  }
}

cudaError_t rows_move(const uint32_t* src, uint32_t* dst, const long long* idx, long long n, int words, int scatter,
                      cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const long long total = n * (words / 4);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  { TimedLaunch tl_(KK_ROWS, s);
  k_rows_move<<<grid, 256, 0, s>>>(src, dst, idx, n, words, scatter);
  }
  return cudaGetLastError();
}
}
