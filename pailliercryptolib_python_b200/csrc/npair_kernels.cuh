// npair_kernels.cuh -- __global__ wrappers of the n-adic pair engine (npair_items.cuh): HE mul, DJN encrypt and the
// comb-table construction on (X0, X1) pairs of n-sized numbers instead of n^2-sized ones.
// Shared memory (doubles): [NE_COUNT constant entries][per group: xs0, x1, y0, y1 (KP each), e (KP + 2)] -- 106 KB per
// CTA at the (20, 2) shape of 2048-bit keys, two CTAs per SM.
#pragma once
#include "npair_items.cuh"
#include "phe_kernels.cuh"

namespace phe {

// Resident CTAs per SM asked of ptxas (i.e. the register cap).  Shared memory allows two CTAs; at L = 20 two CTAs of
// 255 registers is also what ptxas is given.  At L = 15 (3072-bit keys) the same bound made ptxas schedule
// k_encrypt_npair one product chain at a time (ncu r02: wait 3.4 stalled warps per issue, issue slots 0.33 busy, every
// DFMA of a row writing the same register); under the 168-register cap of three CTAs it interleaves the chains.
#ifndef PHE_NPAIR_CTAS_L15
#define PHE_NPAIR_CTAS_L15 3
#endif
template <int L> struct NPairCtas { static constexpr int V = (L == 15) ? PHE_NPAIR_CTAS_L15 : 2; };

template <int L, int TPI> struct NKShape {
  static constexpr int KP = Shape<L, TPI>::KP;
  static constexpr int GPB = NT / TPI;
  static constexpr int PER_GROUP = 5 * KP + 2;
  static constexpr size_t smem_bytes() { return (size_t)(NE_COUNT * KP + GPB * PER_GROUP) * sizeof(double); }
};

template <int L, int TPI> __device__ __forceinline__ NPairSmem npair_group_smem(double* smem) {
  using NS = NKShape<L, TPI>;
  double* g = smem + NE_COUNT * NS::KP + (size_t)(threadIdx.x / TPI) * NS::PER_GROUP;
  NPairSmem sm;
  sm.xs0 = g; sm.x1 = g + NS::KP; sm.y0 = g + 2 * NS::KP; sm.y1 = g + 3 * NS::KP;
  sm.e = reinterpret_cast<uint64_t*>(g + 4 * NS::KP);
  return sm;
}

// ---- HE mul: out = c^e mod n^2 ------------------------------------------------------------------------------------
struct MulNPairArgs {
  const uint32_t* c_w;    // [count][2 * chunk_words]
  int chunk_words;        // n_words
  const uint32_t* e_w;    // [count][e_words], or one row when e_stride == 0
  int e_words;
  size_t e_stride;
  int ebits;
  uint32_t* out_w;        // [count][2 * chunk_words]
  int count;
  NPairCtxArgs ctx;
  double* tbl;            // [gridDim.x * GPB][1 << WIN][2 * KP]
};

template <int L, int TPI, int WIN> __global__ void __launch_bounds__(NT, NPairCtas<L>::V) k_mul_npair(MulNPairArgs p) {
  using Env = DevEnv<TPI>;
  using NS = NKShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<NS::KP>(smem, p.ctx.entries, NE_COUNT);
  const NPairSmem sm = npair_group_smem<L, TPI>(smem);
  const int g = threadIdx.x / TPI;
  double* tbl = p.tbl + ((size_t)blockIdx.x * NS::GPB + g) * ((size_t)2 * NS::KP << WIN);
  const int cw = 2 * p.chunk_words;
  for (int base = blockIdx.x * NS::GPB; base < p.count; base += gridDim.x * NS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    NPairPowmCtl<L, TPI, Env, WIN> ctl;
    ctl.c_w = p.c_w + (size_t)item * cw; ctl.chunk_words = p.chunk_words;
    ctl.e_w = p.e_w + (size_t)item * p.e_stride; ctl.e_words = p.e_words; ctl.ebits = p.ebits;
    ctl.out_w = want < p.count ? p.out_w + (size_t)item * cw : nullptr; ctl.out_words = cw;
    ctl.cst = smem; ctl.tbl = tbl; ctl.sm = sm;
    npair_run<L, TPI, Env>(ctl, smem, p.ctx.n0inv, p.ctx.d_top, sm);
  }
}

// ---- exponent alignment in place: ct[idx[i]] <- ct[idx[i]]^(2^delta[i]), rows sorted by delta (descending) ---------
struct ScaleNPairArgs {
  uint32_t* ct;              // [rows][2 * chunk_words]
  int chunk_words;
  const long long* idx;      // [count] distinct rows
  const int* delta;          // [count], non-increasing
  int count;
  NPairCtxArgs ctx;
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, NPairCtas<L>::V) k_scale_npair(ScaleNPairArgs p) {
  using Env = DevEnv<TPI>;
  using NS = NKShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<NS::KP>(smem, p.ctx.entries, NE_COUNT);
  const NPairSmem sm = npair_group_smem<L, TPI>(smem);
  const int g = threadIdx.x / TPI;
  const int cw = 2 * p.chunk_words;
  for (int base = blockIdx.x * NS::GPB; base < p.count; base += gridDim.x * NS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    uint32_t* row = p.ct + (size_t)p.idx[item] * cw;
    NPairScaleCtl<L, TPI, Env> ctl;
    ctl.c_w = row; ctl.chunk_words = p.chunk_words;
    ctl.delta = p.delta[item]; ctl.max_delta = p.delta[base];     // sorted: the first row of the CTA's batch has the most
    ctl.out_w = want < p.count ? row : nullptr; ctl.out_words = cw;
    ctl.cst = smem; ctl.sm = sm;
    npair_run<L, TPI, Env>(ctl, smem, p.ctx.n0inv, p.ctx.d_top, sm);
  }
}

// ---- shared-exponent sliding-window modexp mod n^2 (classic obfuscator r^n) -----------------------------------------
struct ProgNPairArgs {
  const uint32_t* c_w;    // [count][nchunks * chunk_words]
  int chunk_words, nchunks;
  const uint32_t* prog; int nprog;
  uint32_t* out_w;        // [count][out_words]
  int out_words;
  int count;
  NPairCtxArgs ctx;
  double* tbl;            // [gridDim.x * GPB][1 << (PROG_WS - 1)][2 * KP]
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, NPairCtas<L>::V) k_powm_prog_npair(ProgNPairArgs p) {
  using Env = DevEnv<TPI>;
  using NS = NKShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<NS::KP>(smem, p.ctx.entries, NE_COUNT);
  const NPairSmem sm = npair_group_smem<L, TPI>(smem);
  const int g = threadIdx.x / TPI;
  double* tbl = p.tbl + ((size_t)blockIdx.x * NS::GPB + g) * ((size_t)2 * NS::KP << (PROG_WS - 1));
  const int cw = p.nchunks * p.chunk_words;
  for (int base = blockIdx.x * NS::GPB; base < p.count; base += gridDim.x * NS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    NPairProgCtl<L, TPI, Env, PROG_WS> ctl;
    ctl.c_w = p.c_w + (size_t)item * cw; ctl.chunk_words = p.chunk_words; ctl.nchunks = p.nchunks;
    ctl.prog = p.prog; ctl.nprog = p.nprog;
    ctl.out_w = want < p.count ? p.out_w + (size_t)item * p.out_words : nullptr; ctl.out_words = p.out_words;
    ctl.cst = smem; ctl.tbl = tbl; ctl.sm = sm;
    npair_run<L, TPI, Env>(ctl, smem, p.ctx.n0inv, p.ctx.d_top, sm);
  }
}

// ---- DJN encrypt ------------------------------------------------------------------------------------------------
struct EncNPairArgs {
  const uint32_t* m_w; int m_words;
  const uint32_t* r_w; int r_words;
  int nwin, wb;
  uint32_t* out_w; int out_words;
  int count;
  NPairCtxArgs ctx;
  const double* comb;     // [nwin][1 << wb][2 * KP]
  uint32_t* peer_out[NPAIR_MAX_PEERS];   // n_peers more [count][out_words] destinations (peer-mapped gather buffers)
  int n_peers;
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, NPairCtas<L>::V) k_encrypt_npair(const __grid_constant__ EncNPairArgs p) {
  using Env = DevEnv<TPI>;
  using NS = NKShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<NS::KP>(smem, p.ctx.entries, NE_COUNT);
  const NPairSmem sm = npair_group_smem<L, TPI>(smem);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * NS::GPB; base < p.count; base += gridDim.x * NS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    NPairEncCtl<L, TPI, Env> ctl;
    ctl.m_w = p.m_w + (size_t)item * p.m_words; ctl.m_words = p.m_words;
    ctl.r_w = p.r_w ? p.r_w + (size_t)item * p.r_words : nullptr; ctl.r_words = p.r_words;
    ctl.nwin = p.nwin; ctl.wb = p.wb;
    ctl.out_w = want < p.count ? p.out_w + (size_t)item * p.out_words : nullptr; ctl.out_words = p.out_words;
    ctl.cst = smem; ctl.comb = p.comb; ctl.sm = sm;
    ctl.peers = p.peer_out; ctl.n_peers = p.n_peers; ctl.peer_off = (size_t)item * p.out_words;
    npair_run<L, TPI, Env>(ctl, smem, p.ctx.n0inv, p.ctx.d_top, sm);
  }
}

// ---- comb table construction ------------------------------------------------------------------------------------
struct CombNPairArgs {
  const uint32_t* hs_w;   // hs canonical words (2 * chunk_words)
  int chunk_words;
  int nwin, wb, level;
  double* comb;
  NPairCtxArgs ctx;
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, NPairCtas<L>::V) k_comb_bases_npair(CombNPairArgs p) {
  using Env = DevEnv<TPI>;
  using NS = NKShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<NS::KP>(smem, p.ctx.entries, NE_COUNT);
  const NPairSmem sm = npair_group_smem<L, TPI>(smem);
  NPairCombBasesCtl<L, TPI, Env> ctl;
  ctl.hs_w = p.hs_w; ctl.chunk_words = p.chunk_words; ctl.nwin = p.nwin; ctl.wb = p.wb; ctl.comb = p.comb;
  ctl.writer = (threadIdx.x / TPI) == 0;   // every group of the single CTA runs the same chain; group 0 stores
  ctl.cst = smem; ctl.sm = sm;
  npair_run<L, TPI, Env>(ctl, smem, p.ctx.n0inv, p.ctx.d_top, sm);
}

template <int L, int TPI> __global__ void __launch_bounds__(NT, NPairCtas<L>::V) k_comb_level_npair(CombNPairArgs p) {
  using Env = DevEnv<TPI>;
  using NS = NKShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<NS::KP>(smem, p.ctx.entries, NE_COUNT);
  const NPairSmem sm = npair_group_smem<L, TPI>(smem);
  const int g = threadIdx.x / TPI;
  const int per_win = (1 << p.level) - 1;
  const int count = p.nwin * per_win;
  for (int base = blockIdx.x * NS::GPB; base < count; base += gridDim.x * NS::GPB) {
    const int want = base + g;
    const int item = want < count ? want : count - 1;   // past the end: redo the last product, skip the store
    const int j = item / per_win;
    NPairCombLevelCtl<L, TPI, Env> ctl;
    ctl.row = p.comb + (((size_t)j) << p.wb) * 2 * NS::KP; ctl.level = p.level; ctl.e = 1 + item % per_win;
    ctl.store = want < count; ctl.sm = sm;
    npair_run<L, TPI, Env>(ctl, smem, p.ctx.n0inv, p.ctx.d_top, sm);
  }
}

}  // namespace phe
