// phe_launch.cuh -- templated launchers; included by each shape_*.cu with PHE_SHAPE_L / PHE_SHAPE_TPI defined.
#pragma once
#include <algorithm>

#include "phe_kernels.cuh"
#include "npair_kernels.cuh"
#include "phe_shapes.hpp"

namespace phe {

inline int sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) dev = 0;
  if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev];
}

// grid.x for a persistent-style launch: a whole number of resident waves, never more CTAs than work
template <class K> int grid_for(K kernel, size_t smem, int count, int gpb, int ny, int threads = NT) {
  int occ = 0;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
  if (occ < 1) occ = 1;
  const int resident = std::max(1, sm_count() * occ / ny);
  const int need = (count + gpb - 1) / gpb;
  return std::max(1, std::min(resident, need));
}

template <int L, int TPI> struct Launch {
  using KS = KShape<L, TPI>;

  static cudaError_t modmul(const uint32_t* a, const uint32_t* b, size_t b_stride, uint32_t* out, int nwords,
                            int count, const MontCtxArgs& ctx, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_modmul<L, TPI>, smem, count, KS::GPB, 1);
    { TimedLaunch tl_(KK_MODMUL, s);
    k_modmul<L, TPI><<<grid, NT, smem, s>>>(a, b, b_stride, out, nwords, count, ctx);
    }
    return cudaGetLastError();
  }

  template <int WIN> static cudaError_t powm_w(const PowmArgs& p, int ny, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_powm<L, TPI, WIN>, smem, p.count, KS::GPB, ny);
    { TimedLaunch tl_(KK_POWM, s);
    k_powm<L, TPI, WIN><<<dim3(grid, ny), NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t powm(int win, const PowmArgs& p, int ny, cudaStream_t s) {
    switch (win) {
      case 1: return powm_w<1>(p, ny, s);
      case 3: return powm_w<3>(p, ny, s);
      case 5: return powm_w<5>(p, ny, s);
      default: return cudaErrorInvalidValue;
    }
  }
  // shared-exponent program path; scratch: same 32 entries per group as the 5-bit fixed window
  static cudaError_t powm_prog(const PowmArgs& p, int ny, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_powm_prog<L, TPI>, smem, p.count, KS::GPB, ny);
    { TimedLaunch tl_(KK_POWM, s);
    k_powm_prog<L, TPI><<<dim3(grid, ny), NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static size_t powm_prog_tbl_words(int ny, int count) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_powm_prog<L, TPI>, smem, count, KS::GPB, ny);
    return (size_t)grid * ny * KS::GPB * ((size_t)KS::KP << (PROG_WS - 1)) * 2;
  }
  template <int WIN> static size_t tbl_words_w(int ny, int count) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_powm<L, TPI, WIN>, smem, count, KS::GPB, ny);
    return (size_t)grid * ny * KS::GPB * ((size_t)KS::KP << WIN) * 2;   // doubles -> u32 words
  }
  static size_t powm_tbl_words(int win, int ny, int count) {   // in u32 words (DevBuf unit)
    switch (win) {
      case 1: return tbl_words_w<1>(ny, count);
      case 3: return tbl_words_w<3>(ny, count);
      case 5: return tbl_words_w<5>(ny, count);
      default: return 0;
    }
  }

  static cudaError_t dec_prep(const DecPrepArgs& p, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_dec_prep<L, TPI>, smem, p.count, KS::GPB, 2);
    { TimedLaunch tl_(KK_DEC_PREP, s);
    k_dec_prep<L, TPI><<<dim3(grid, 2), NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t dec_tail(const DecTailArgs& p, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(DT_COUNT);
    const int grid = grid_for(k_dec_tail<L, TPI>, smem, p.count, KS::GPB, 1);
    { TimedLaunch tl_(KK_DEC_TAIL, s);
    k_dec_tail<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t inv_block(const InvArgs& p, bool unwind, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int nblocks = p.count / p.block;
    if (unwind) {
      const int grid = grid_for(k_inv_block<L, TPI, true>, smem, nblocks, KS::GPB, 1);
      { TimedLaunch tl_(KK_MODMUL, s);
      k_inv_block<L, TPI, true><<<grid, NT, smem, s>>>(p);
      }
    } else {
      const int grid = grid_for(k_inv_block<L, TPI, false>, smem, nblocks, KS::GPB, 1);
      { TimedLaunch tl_(KK_MODMUL, s);
      k_inv_block<L, TPI, false><<<grid, NT, smem, s>>>(p);
      }
    }
    return cudaGetLastError();
  }
  static int resident_groups() {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    return grid_for(k_inv_block<L, TPI, false>, smem, 1 << 30, KS::GPB, 1) * KS::GPB;
  }
  static cudaError_t dec_crt(const DecCrtArgs& p, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(DT_COUNT);
    const int grid = grid_for(k_dec_crt<L, TPI>, smem, p.count, KS::GPB, 1);
    { TimedLaunch tl_(KK_DEC_CRT, s);
    k_dec_crt<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t encrypt_comb(const EncCombArgs& p, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_encrypt_comb<L, TPI>, smem, p.count, KS::GPB, 1);
    { TimedLaunch tl_(KK_ENC_COMB, s);
    k_encrypt_comb<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t encrypt_finish(const EncFinishArgs& p, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_encrypt_finish<L, TPI>, smem, p.count, KS::GPB, 1);
    { TimedLaunch tl_(KK_ENC_FINISH, s);
    k_encrypt_finish<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t comb_build(const CombArgs& p0, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    cudaFuncSetAttribute(k_comb_bases<L, TPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    { TimedLaunch tl_(KK_COMB_BUILD, s);
    k_comb_bases<L, TPI><<<1, NT, smem, s>>>(p0);
    }
    cudaError_t e = cudaGetLastError();
    for (int k = 1; k < p0.wb && e == cudaSuccess; ++k) {
      CombArgs p = p0;
      p.level = k;
      const int count = p.nwin * ((1 << k) - 1);
      const int grid = grid_for(k_comb_level<L, TPI>, smem, count, KS::GPB, 1);
      { TimedLaunch tl_(KK_COMB_BUILD, s);
      k_comb_level<L, TPI><<<grid, NT, smem, s>>>(p);
      }
      e = cudaGetLastError();
    }
    return e;
  }

  // ---- n-adic pair engine ----
  using NS = NKShape<L, TPI>;
  template <int WIN> static cudaError_t mul_npair_w(const MulNPairArgs& p, cudaStream_t s) {
    const size_t smem = NS::smem_bytes();
    const int grid = grid_for(k_mul_npair<L, TPI, WIN>, smem, p.count, NS::GPB, 1);
    { TimedLaunch tl_(KK_MUL_NPAIR, s);
    k_mul_npair<L, TPI, WIN><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t mul_npair(int win, const MulNPairArgs& p, cudaStream_t s) {
    switch (win) {
      case 1: return mul_npair_w<1>(p, s);
      case 3: return mul_npair_w<3>(p, s);
      case 5: return mul_npair_w<5>(p, s);
      default: return cudaErrorInvalidValue;
    }
  }
  template <int WIN> static size_t mul_npair_tbl_w(int count) {
    const int grid = grid_for(k_mul_npair<L, TPI, WIN>, NS::smem_bytes(), count, NS::GPB, 1);
    return (size_t)grid * NS::GPB * ((size_t)2 * NS::KP << WIN) * 2;   // doubles -> u32 words
  }
  static size_t mul_npair_tbl_words(int win, int count) {
    switch (win) {
      case 1: return mul_npair_tbl_w<1>(count);
      case 3: return mul_npair_tbl_w<3>(count);
      case 5: return mul_npair_tbl_w<5>(count);
      default: return 0;
    }
  }
  static cudaError_t encrypt_npair(const EncNPairArgs& p, cudaStream_t s) {
    const size_t smem = NS::smem_bytes();
    const int grid = grid_for(k_encrypt_npair<L, TPI>, smem, p.count, NS::GPB, 1);
    { TimedLaunch tl_(KK_ENC_NPAIR, s);
    k_encrypt_npair<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t powm_prog_npair(const ProgNPairArgs& p, cudaStream_t s) {
    const size_t smem = NS::smem_bytes();
    const int grid = grid_for(k_powm_prog_npair<L, TPI>, smem, p.count, NS::GPB, 1);
    { TimedLaunch tl_(KK_MUL_NPAIR, s);
    k_powm_prog_npair<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static size_t powm_prog_npair_tbl_words(int count) {
    const int grid = grid_for(k_powm_prog_npair<L, TPI>, NS::smem_bytes(), count, NS::GPB, 1);
    return (size_t)grid * NS::GPB * ((size_t)2 * NS::KP << (PROG_WS - 1)) * 2;
  }
  static cudaError_t modmul1(const Modmul1Args& p, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const int grid = grid_for(k_modmul1<L, TPI>, smem, p.count, KS::GPB, 1);
    { TimedLaunch tl_(KK_MODMUL, s);
    k_modmul1<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t tree_level(const TreeLevelArgs& p, cudaStream_t s) {
    const size_t smem = KS::smem_bytes(ME_COUNT);
    const long long items = (long long)p.groups * (p.w / 2 + (p.w & 1));
    const int grid = grid_for(k_tree_level<L, TPI>, smem, (int)std::min<long long>(items, 1 << 30), KS::GPB, 1);
    { TimedLaunch tl_(KK_MODMUL, s);
    k_tree_level<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t scale_npair(const ScaleNPairArgs& p, cudaStream_t s) {
    const size_t smem = NS::smem_bytes();
    const int grid = grid_for(k_scale_npair<L, TPI>, smem, p.count, NS::GPB, 1);
    { TimedLaunch tl_(KK_MUL_NPAIR, s);
    k_scale_npair<L, TPI><<<grid, NT, smem, s>>>(p);
    }
    return cudaGetLastError();
  }
  static cudaError_t comb_build_npair(const CombNPairArgs& p0, cudaStream_t s) {
    const size_t smem = NS::smem_bytes();
    cudaFuncSetAttribute(k_comb_bases_npair<L, TPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    { TimedLaunch tl_(KK_COMB_BUILD, s);
    k_comb_bases_npair<L, TPI><<<1, NT, smem, s>>>(p0);
    }
    cudaError_t e = cudaGetLastError();
    for (int k = 1; k < p0.wb && e == cudaSuccess; ++k) {
      CombNPairArgs p = p0;
      p.level = k;
      const int count = p.nwin * ((1 << k) - 1);
      const int grid = grid_for(k_comb_level_npair<L, TPI>, smem, count, NS::GPB, 1);
      { TimedLaunch tl_(KK_COMB_BUILD, s);
      k_comb_level_npair<L, TPI><<<grid, NT, smem, s>>>(p);
      }
      e = cudaGetLastError();
    }
    return e;
  }

  static constexpr ShapeOps ops() {
    return ShapeOps{L, TPI, KS::KP, KS::GPB, LW * L * TPI, &modmul, &powm, &powm_tbl_words, &powm_prog, &powm_prog_tbl_words, &dec_prep, &dec_tail, &dec_crt, &inv_block, &resident_groups,
                    &encrypt_comb, &encrypt_finish, &comb_build,
                    &mul_npair, &mul_npair_tbl_words, &encrypt_npair, &comb_build_npair, &powm_prog_npair,
                    &powm_prog_npair_tbl_words, &modmul1, &tree_level, &scale_npair};
  }
};

// Launcher of the one-bignum-per-lane pair engine (L = limbs of p, q).
template <int L> struct PairLaunch {
  static constexpr int NTP = PairShape<L>::NTP;
  static int units_of(int count) { return 2 * ((count + 31) / 32); }   // both moduli: warp units
  static int grid(int count) {   // NT / 32 warps per CTA
    return grid_for(k_dec_pair<L>, PairShape<L>::smem_bytes(), units_of(count), NTP / 32, 1, NTP);
  }
  // resident warps of a launch over `count` ciphertexts (the host picks the number of segments from it)
  static int warps(int count) { return grid(count) * (NTP / 32); }
  // mod_p, mod_q: L limbs each (doubles); p.sched: sched_ints(count) ints, zeroed here on the stream
  static cudaError_t dec_pair(const DecPairArgs& p, const double* mod_p, const double* mod_q, cudaStream_t s) {
    const size_t smem = PairShape<L>::smem_bytes();
    const int g = grid(p.count);
    ModLimbs<L> m;
    for (int i = 0; i < L; ++i) { m.v[0][i] = mod_p[i]; m.v[1][i] = mod_q[i]; }
    cudaError_t e = cudaMemsetAsync(p.sched, 0, sched_ints(p.count) * sizeof(int), s);
    if (e != cudaSuccess) return e;
    { TimedLaunch tl_(KK_DEC_PAIR, s);
    k_dec_pair<L><<<g, NTP, smem, s>>>(p, m);
    }
    return cudaGetLastError();
  }
  static size_t sched_ints(int count) { return (size_t)PAIR_SCHED_RING + (size_t)units_of(count) * (PAIR_MAX_SEG - 1); }
  static size_t tbl_words(int count, int slots) {   // u32 words: per unit `slots` table entries + the parked pair
    return (size_t)units_of(count) * (slots + 1) * 2 * L * 32 * 2;
  }
  static constexpr PairOps ops() { return PairOps{L, &dec_pair, &tbl_words, &warps, &sched_ints}; }
};

}  // namespace phe
