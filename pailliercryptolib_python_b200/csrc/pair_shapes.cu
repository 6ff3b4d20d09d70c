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

// ---- first half of the fixed-point decode on the device (fixedpoint.py:97-115 of the reference: FixedPointNumber.decode) --
// One thread per decrypted plaintext row m (nw words, 0 <= m < n): class 0 if m < 2^63 (mantissa = m), class 1 if
// n - m < 2^63 (a negative value: mantissa = -(n - m)), class 2 otherwise (a mantissa beyond 63 bits or an overflow:
// the host decodes those rows from the words).  Saves the host the 64-word scan and the D2H of 256 B per element.
__global__ void __launch_bounds__(256) k_classify_plain(const uint32_t* __restrict__ m, const uint32_t* __restrict__ n, int nw,
                                                        long long count, long long* __restrict__ mant,
                                                        unsigned char* __restrict__ cls) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t* row = m + (size_t)i * nw;
    uint32_t hi_or = 0, d_or = 0, borrow = 0, d0 = 0, d1 = 0;
    for (int j = 0; j < nw; ++j) {
      const uint32_t v = row[j], nj = n[j];
      const uint64_t t = (uint64_t)nj - v - borrow;
      borrow = (uint32_t)(t >> 63);
      const uint32_t dj = (uint32_t)t;
      if (j == 0) d0 = dj; else if (j == 1) d1 = dj; else { hi_or |= v; d_or |= dj; }
    }
    const uint32_t w0 = row[0], w1 = nw > 1 ? row[1] : 0u;
    if (hi_or == 0 && w1 < 0x80000000u) { cls[i] = 0; mant[i] = (long long)(((uint64_t)w1 << 32) | w0); }
    else if (borrow == 0 && d_or == 0 && d1 < 0x80000000u) { cls[i] = 1; mant[i] = -(long long)(((uint64_t)d1 << 32) | d0); }
    else { cls[i] = 2; mant[i] = 0; }
  }
}
cudaError_t classify_plain(const uint32_t* m, const uint32_t* n, int nw, long long count, long long* mant, unsigned char* cls,
                           cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int grid = (int)std::min<long long>((count + 255) / 256, (long long)sm_count() * 8);
  { TimedLaunch tl_(KK_ROWS, s);
  k_classify_plain<<<grid, 256, 0, s>>>(m, n, nw, count, mant, cls);
  }
  return cudaGetLastError();
}
}
