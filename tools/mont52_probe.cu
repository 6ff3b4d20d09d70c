// mont52_probe.cu -- correctness + throughput probe of the radix-2^52 FP64 Montgomery product (csrc/mont52.cuh).
// Each lane group computes x^(2^nsq) mod n by nsq Montgomery squarings; a few items are checked against hostbn.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o mont52_probe mont52_probe.cu
// Output: one JSON object per shape.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "../pailliercryptolib_python_b200/csrc/hostbn.hpp"
#include "../pailliercryptolib_python_b200/csrc/mont52.cuh"

using hbn::BN;
using namespace phe;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int L, int TPI, int NT, int MINB, bool NSM>
__global__ void __launch_bounds__(NT, MINB) k_sqr_chain(const uint32_t* __restrict__ x_w, uint32_t* __restrict__ out_w, int nwords,
                                                  int count, const double* __restrict__ entries, uint64_t n0inv, int nsq) {
  using Env = DevEnv<TPI>;
  using S = Shape<L, TPI>;
  constexpr int GPB = NT / TPI;
  extern __shared__ __align__(16) double smem[];
  for (int i = threadIdx.x; i < 3 * S::KP; i += NT) smem[i] = entries[i];
  __syncthreads();
  const int g = threadIdx.x / TPI;
  double* b0 = smem + (size_t)(3 + g) * S::KP;
  double x[L];
  for (int base = blockIdx.x * GPB; base < count; base += gridDim.x * GPB) {
    const int want = base + g;
    const int item = want < count ? want : count - 1;
    limbs_from_words<L, TPI, Env>(x, x_w + (size_t)item * nwords, nwords);
    montmul<L, TPI, Env>(x, x, smem + S::KP, smem, n0inv);   // to Montgomery form
#pragma unroll 1
    for (int s = 0; s < nsq; ++s) {
      Env::sync();
      limbs_to_mem<L, TPI, Env>(b0, x);
      Env::sync();
      montmul<L, TPI, Env>(x, x, b0, smem, n0inv);
    }
    montmul<L, TPI, Env>(x, x, smem + 2 * S::KP, smem, n0inv);   // leave Montgomery form
    uint64_t xi[L];
    canonical_ints<L, TPI, Env>(xi, x, smem);
    Env::sync();
    ints_to_mem<L, TPI, Env>(reinterpret_cast<uint64_t*>(b0), xi);
    Env::sync();
    if (want < count)
      for (int v = Env::lane(); v < nwords; v += TPI)
        out_w[(size_t)item * nwords + v] = word_from_ints<L, TPI>(reinterpret_cast<const uint64_t*>(b0), v);
    Env::sync();
  }
}

template <int L, int TPI> void entry_from_bn(double* dst, const BN& v) {
  using S = Shape<L, TPI>;
  std::vector<uint32_t> w(S::K * 52 / 32 + 4, 0);
  v.to_words(w.data(), w.size());
  for (int t = 0; t < TPI; ++t)
    for (int j = 0; j < S::LP; ++j) {
      uint64_t limb = 0;
      if (j < L) {
        const int bit = (t * L + j) * 52;
        for (int k = 0; k < 52; ++k) { const int bb = bit + k; if ((w[bb >> 5] >> (bb & 31)) & 1u) limb |= 1ull << k; }
      }
      dst[t * S::LP + j] = (double)limb;
    }
}

template <int L, int TPI, int NT, int MINB = 1, bool NSM = false> int run(int mod_bits, int nsq, int ctas_per_sm, int waves) {
  using S = Shape<L, TPI>;
  constexpr int GPB = NT / TPI;
  const int nwords = mod_bits / 32;
  std::mt19937_64 rng(1234 + L * 7 + TPI);
  std::vector<uint32_t> nw(nwords);
  for (auto& v : nw) v = (uint32_t)rng();
  nw[0] |= 1u; nw[nwords - 1] |= 0x80000000u;
  const BN N = BN::from_words(nw.data(), nwords);
  const BN R = hbn::shl(BN(1), 52 * S::K);
  std::vector<double> ent(3 * S::KP);
  entry_from_bn<L, TPI>(ent.data(), N);
  entry_from_bn<L, TPI>(ent.data() + S::KP, hbn::mod(hbn::mul(R, R), N));
  entry_from_bn<L, TPI>(ent.data() + 2 * S::KP, BN(1));
  uint64_t n0 = (uint64_t)nw[0] | ((uint64_t)nw[1] << 32), inv = 1;
  for (int i = 0; i < 6; ++i) inv *= 2 - n0 * inv;
  const uint64_t n0inv = (0 - inv) & M52;

  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t smem = (size_t)(3 + GPB) * S::KP * sizeof(double);
  auto kern = k_sqr_chain<L, TPI, NT, MINB, NSM>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
  if (ctas_per_sm > 0 && ctas_per_sm < occ) occ = ctas_per_sm;
  const int grid = sms * occ;
  const int count = grid * GPB * waves;
  std::vector<uint32_t> xs((size_t)count * nwords), out((size_t)count * nwords);
  for (auto& v : xs) v = (uint32_t)rng();
  for (int i = 0; i < count; ++i) xs[(size_t)i * nwords + nwords - 1] &= 0x3fffffffu;   // < n
  // edge items: 0, 1, n-1
  for (int v = 0; v < nwords; ++v) { xs[v] = 0; xs[nwords + v] = v == 0; xs[2 * (size_t)nwords + v] = nw[v] - (v == 0); }
  uint32_t *dx, *dout; double* dent;
  CK(cudaMalloc(&dx, xs.size() * 4)); CK(cudaMalloc(&dout, out.size() * 4)); CK(cudaMalloc(&dent, ent.size() * 8));
  CK(cudaMemcpy(dx, xs.data(), xs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dent, ent.data(), ent.size() * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<grid, NT, smem>>>(dx, dout, nwords, count, dent, n0inv, 8);   // warm-up + correctness run
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  const int check[] = {0, 1, 2, 3, 4, GPB - 1, GPB, count / 2, count - 1};
  for (int i : check) {
    const BN x = BN::from_words(&xs[(size_t)i * nwords], nwords);
    const BN want = hbn::modexp(x, BN(256), N);
    std::vector<uint32_t> ww(nwords); want.to_words(ww.data(), nwords);
    if (memcmp(ww.data(), &out[(size_t)i * nwords], nwords * 4) != 0) { ++bad; fprintf(stderr, "MISMATCH L=%d TPI=%d item %d\n", L, TPI, i); }
  }
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    kern<<<grid, NT, smem>>>(dx, dout, nwords, count, dent, n0inv, nsq);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  const double mm = (double)count * (nsq + 2);
  const int kw = mod_bits / 32;
  printf("{\"L\":%d,\"TPI\":%d,\"NT\":%d,\"MINB\":%d,\"NSM\":%d,\"mod_bits\":%d,\"regs\":%d,\"ctas_per_sm\":%d,\"grid\":%d,\"count\":%d,\"nsq\":%d,\"ms\":%.3f,"
         "\"montmul_per_s\":%.4e,\"tmac32_per_s\":%.3f,\"mismatches\":%d}\n",
         L, TPI, NT, MINB, (int)NSM, mod_bits, fa.numRegs, occ, grid, count, nsq, best, mm / (best * 1e-3),
         mm * (2.0 * kw * kw + kw) / (best * 1e-3) / 1e12, bad);
  cudaFree(dx); cudaFree(dout); cudaFree(dent);
  return bad;
}

int main(int argc, char** argv) {
  const int nsq = argc > 1 ? atoi(argv[1]) : 600;
  const int only = argc > 2 ? atoi(argv[2]) : -1;   // run a single shape (for ncu)
  int bad = 0;
  if (only == 0) return run<20, 2, 128, 2, true>(2048, nsq, 0, 2);
  if (only == 1) return run<20, 2, 128, 3, true>(2048, nsq, 0, 2);
  bad += run<20, 2, 128, 2, true>(2048, nsq, 0, 2);
  bad += run<20, 2, 128, 3, true>(2048, nsq, 0, 2);
  bad += run<10, 4, 128, 3, true>(2048, nsq, 0, 2);
  bad += run<10, 4, 128, 4, true>(2048, nsq, 0, 2);
  bad += run<10, 4, 128, 5, true>(2048, nsq, 0, 2);
  bad += run<10, 4, 128, 6, true>(2048, nsq, 0, 2);
  bad += run<10, 8, 128, 4, true>(4096, nsq / 2, 0, 2);
  bad += run<10, 8, 128, 5, true>(4096, nsq / 2, 0, 2);
  bad += run<20, 4, 128, 2, true>(4096, nsq / 2, 0, 2);
  bad += run<15, 4, 128, 3, true>(3072, nsq / 2, 0, 2);
  return bad ? 1 : 0;
}
