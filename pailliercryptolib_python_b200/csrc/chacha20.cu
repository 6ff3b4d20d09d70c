// chacha20.cu -- device-side CSPRNG for the obfuscator exponents r of ipcl::PublicKey::applyObfuscator
// (/root/reference/src/ipcl_python/bindings/ipcl_bindings_classes.cpp:71-83: r is drawn inside the library, never
// passed in by the caller).  The keystream is ChaCha20 (RFC 8439 block function, 20 rounds) under a 256-bit key and a
// 96-bit nonce that the host takes from getrandom(2) for every call; one thread produces one 64-byte block.  Drawing
// 12.8 MB of r for 100 000 encryptions from getrandom itself costs 40-55 ms on the host plus the H2D copy, more than the
// encrypt kernel (24 ms).
#include <cuda_runtime.h>

#include <cstdint>

namespace phe {

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return __funnelshift_l(v, v, c); }

#define PHE_QR(a, b, c, d)                    \
  a += b; d ^= a; d = rotl32(d, 16);          \
  c += d; b ^= c; b = rotl32(b, 12);          \
  a += b; d ^= a; d = rotl32(d, 8);           \
  c += d; b ^= c; b = rotl32(b, 7);

struct ChaChaKey { uint32_t key[8]; uint32_t nonce[3]; uint32_t counter0; };

// out[0 .. words): keystream words; block b (16 words) uses counter counter0 + b.  mask_every > 0: word index
// (mask_every - 1) of every row of mask_every words is ANDed with top_mask (r uniform in [0, 2^randbits)).
__global__ void __launch_bounds__(256) k_chacha20_fill(uint32_t* __restrict__ out, size_t words, ChaChaKey k,
                                                       int mask_every, uint32_t top_mask) {
  const size_t nblocks = (words + 15) / 16;
  for (size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x; blk < nblocks; blk += (size_t)gridDim.x * blockDim.x) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      k.key[0], k.key[1], k.key[2], k.key[3], k.key[4], k.key[5], k.key[6], k.key[7],
                      k.counter0 + (uint32_t)blk, k.nonce[0], k.nonce[1], k.nonce[2]};
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = s[i];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      PHE_QR(x[0], x[4], x[8], x[12]) PHE_QR(x[1], x[5], x[9], x[13]) PHE_QR(x[2], x[6], x[10], x[14]) PHE_QR(x[3], x[7], x[11], x[15])
      PHE_QR(x[0], x[5], x[10], x[15]) PHE_QR(x[1], x[6], x[11], x[12]) PHE_QR(x[2], x[7], x[8], x[13]) PHE_QR(x[3], x[4], x[9], x[14])
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const size_t w = blk * 16 + i;
      if (w < words) {
        uint32_t v = x[i] + s[i];
        if (mask_every > 0 && (w % (size_t)mask_every) == (size_t)(mask_every - 1)) v &= top_mask;
        out[w] = v;
      }
    }
  }
}

cudaError_t chacha20_fill(uint32_t* d_out, size_t words, const uint32_t key[8], const uint32_t nonce[3], uint32_t counter0,
                          int mask_every, uint32_t top_mask, cudaStream_t s) {
  if (words == 0) return cudaSuccess;
  ChaChaKey k;
  for (int i = 0; i < 8; ++i) k.key[i] = key[i];
  for (int i = 0; i < 3; ++i) k.nonce[i] = nonce[i];
  k.counter0 = counter0;
  const size_t nblocks = (words + 15) / 16;
  const int grid = (int)((nblocks + 255) / 256 > 148 * 8 ? 148 * 8 : (nblocks + 255) / 256);
  k_chacha20_fill<<<grid, 256, 0, s>>>(d_out, words, k, mask_every, top_mask);
  return cudaGetLastError();
}

}  // namespace phe
