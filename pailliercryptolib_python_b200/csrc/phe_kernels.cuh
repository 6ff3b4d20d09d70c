// phe_kernels.cuh -- __global__ wrappers around the per-item bodies (paillier_items.cuh) for sm_100a.
//
// Launch shape: persistent-style grid (a multiple of the SM count), CTA = NT threads, a lane group of TPI
// adjacent lanes per bignum, item loops are block-uniform (so every shuffle is convergent and ptxas does not
// clone a WARPSYNC'd slow path).  Lanes whose item index is past the end recompute the last item and skip
// the store.
//
// Shared memory (doubles): [CTA-shared constant entries][per group: b0, b1].  The modulus is read from its
// shared-memory entry inside montmul, so no kernel keeps it in registers.
// __launch_bounds__(NT, 3): a 168-register cap, 3 CTAs (12 warps) per SM.  Left to itself (cap 255) ptxas flips,
// kernel by kernel, between a schedule that interleaves ~20 product chains (250 registers, ~7 stall cycles per
// DFMA in the SASS control words) and one that runs a single chain at a time (180-200 registers, ~17 stall cycles
// per DFMA: k_powm_prog<20,2> came out 47 % slower than k_powm<20,2,5> on the same arithmetic).  Under the explicit
// cap every kernel gets the interleaved schedule in 160-168 registers without spills, and the third CTA adds the
// warps that hide what latency is left (tools/mont52_probe.cu: 1.17 G vs 1.18 G 2048-bit products/s).
#pragma once
#include <cuda_runtime.h>

#include "paillier_items.cuh"

namespace phe {

constexpr int NT = 128;  // threads per CTA
template <int L> struct MinCtas { static constexpr int V = 3; };

template <int L, int TPI> struct KShape {
  static constexpr int KP = Shape<L, TPI>::KP;
  static constexpr int GPB = NT / TPI;  // groups per CTA
  static constexpr size_t smem_bytes(int n_shared_entries) {
    return (size_t)(n_shared_entries + 2 * GPB) * KP * sizeof(double);
  }
};

template <int L, int TPI>
__device__ __forceinline__ GroupSmem group_smem(double* smem, int n_shared_entries) {
  constexpr int KP = Shape<L, TPI>::KP;
  const int g = threadIdx.x / TPI;
  GroupSmem sm;
  sm.b0 = smem + (size_t)(n_shared_entries + 2 * g) * KP;
  sm.b1 = sm.b0 + KP;
  return sm;
}

template <int KP>
__device__ __forceinline__ void stage_entries(double* smem, const double* src, int n_entries) {
  const double2* s = reinterpret_cast<const double2*>(src);
  double2* d = reinterpret_cast<double2*>(smem);
  for (int i = threadIdx.x; i < n_entries * KP / 2; i += NT) d[i] = s[i];
  __syncthreads();
}

// Montgomery context entries as laid out by the host (phe_api.cu): indices into a [.. ][KP] block
enum MontEntry { ME_N = 0, ME_R2 = 1, ME_ONEM = 2, ME_ONE = 3, ME_X0 = 4, ME_COUNT = 5 };
// ME_X0: context-specific extra (n*R mod n^2 for the n^2 context; 2^(32 hw) R^2 mod x^2 for x^2 contexts)

struct MontCtxArgs {
  const double* entries;  // ME_COUNT entries of KP doubles
  uint64_t n0inv;         // -n^-1 mod 2^52
};

// ---- HE add ------------------------------------------------------------------------------------
// b_stride = 0 broadcasts a single b (ipcl CipherText::operator+ with other.size == 1)
template <int L, int TPI>
__global__ void __launch_bounds__(NT, MinCtas<L>::V) k_modmul(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                               size_t b_stride, uint32_t* __restrict__ out, int nwords, int count,
                                               MontCtxArgs ctx) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * KS::GPB; base < count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < count ? want : count - 1;
    uint32_t* dst = out + (size_t)item * nwords;
    item_modmul<L, TPI, Env>(a + (size_t)item * nwords, b + (size_t)item * b_stride, want < count ? dst : nullptr,
                             nwords, smem + ME_N * KS::KP, ctx.n0inv, smem + ME_R2 * KS::KP, sm);
  }
}

// ---- one-product HE add: broadcast second operand, constant second operand -----------------------------------------
// out[i] = a[i] * b * R^-1 mod N with b = b_w[i * b_stride] (words; b_stride = 0: one row for all) or the constant
// entry b_entry (global memory, KP doubles).  See item_modmul1.
struct Modmul1Args {
  const uint32_t* a; const uint32_t* b_w; size_t b_stride; const double* b_entry;
  uint32_t* out; int nwords, count;
  MontCtxArgs ctx;
};
template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_modmul1(Modmul1Args p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    item_modmul1<L, TPI, Env>(p.a + (size_t)item * p.nwords, p.b_w ? p.b_w + (size_t)item * p.b_stride : nullptr, p.b_entry,
                              want < p.count ? p.out + (size_t)item * p.nwords : nullptr, p.nwords, smem + ME_N * KS::KP,
                              p.ctx.n0inv, sm);
  }
}

// ---- one level of the add tree of sum / dot / matmul (ipcl_python.py:746-930; __padded_ct :810-827) ---------------
// src: [groups][w] rows, dst: [groups][wout] rows, wout = half + (w & 1), half = w / 2:
//   dst[g][j] = src[g][j] * src[g][j + half] * R^-1   (j < half);   dst[g][half] = src[g][2 half]   (w odd: carried over
// as a product by the Montgomery one, so that every lane group of a warp runs the same product)
struct TreeLevelArgs {
  const uint32_t* src; uint32_t* dst;
  int nwords, groups, w;
  MontCtxArgs ctx;
};
template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_tree_level(TreeLevelArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  const int half = p.w / 2, wout = half + (p.w & 1);
  const long long count = (long long)p.groups * wout;
  for (long long base = (long long)blockIdx.x * KS::GPB; base < count; base += (long long)gridDim.x * KS::GPB) {
    const long long want = base + g;
    const long long item = want < count ? want : count - 1;
    const long long grp = item / wout;
    const int j = (int)(item - grp * wout);
    const uint32_t* row = p.src + (size_t)grp * p.w * p.nwords;
    const bool pass = (j == half);     // the odd element of this level
    item_modmul1<L, TPI, Env>(row + (size_t)(pass ? 2 * half : j) * p.nwords, pass ? nullptr : row + (size_t)(j + half) * p.nwords,
                              smem + ME_ONEM * KS::KP, want < count ? p.dst + (size_t)item * p.nwords : nullptr, p.nwords,
                              smem + ME_N * KS::KP, p.ctx.n0inv, sm);
  }
}

// ---- batched modular inverse (Montgomery's trick): one block of `block` consecutive elements per lane group --------
struct InvArgs {
  const uint32_t* c_w;     // [count][nwords]
  int nwords, count, block;   // count % block == 0
  double* P;               // [count][KP] prefix products (Montgomery form)
  uint32_t* total_w;       // k_inv_prefix: [nblocks][nwords] block totals (canonical)
  const uint32_t* tinv_w;  // k_inv_unwind: [nblocks][nwords] inverses of the block totals
  uint32_t* out_w;         // k_inv_unwind: [count][nwords]
  MontCtxArgs ctx;
};

template <int L, int TPI, bool UNWIND>
__global__ void __launch_bounds__(NT, MinCtas<L>::V) k_inv_block(InvArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  const int nblocks = p.count / p.block;   // count is a multiple of block (the host pads with ones): every group of a
  const int cnt = p.block;                // warp runs the same number of products, so the shuffles stay convergent
  for (int base = blockIdx.x * KS::GPB; base < nblocks; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int b = want < nblocks ? want : nblocks - 1;      // past the end: redo the last block (same values stored)
    const int first = b * p.block;
    if (!UNWIND)
      item_inv_prefix<L, TPI, Env>(p.c_w + (size_t)first * p.nwords, p.nwords, cnt, p.P + (size_t)first * KS::KP,
                                   p.total_w + (size_t)b * p.nwords, smem + ME_N * KS::KP, p.ctx.n0inv,
                                   smem + ME_R2 * KS::KP, smem + ME_ONEM * KS::KP, smem + ME_ONE * KS::KP, sm);
    else
      item_inv_unwind<L, TPI, Env>(p.c_w + (size_t)first * p.nwords, p.nwords, cnt, p.P + (size_t)first * KS::KP,
                                   p.tinv_w + (size_t)b * p.nwords, want < nblocks ? p.out_w + (size_t)first * p.nwords : nullptr,
                                   smem + ME_N * KS::KP, p.ctx.n0inv, smem + ME_R2 * KS::KP, smem + ME_ONEM * KS::KP,
                                   smem + ME_ONE * KS::KP, sm);
  }
}

// ---- generic fixed-window modexp ----------------------------------------------------------------
// gridDim.y selects one of up to two modulus contexts (decrypt: y=0 -> p^2, y=1 -> q^2); each has its own
// exponent (shared by all items when e_stride == 0) and output array.
struct PowmArgs {
  const uint32_t* base_w;      // [count][base_words] or null
  int base_words;
  const double* base_mont[2];  // [count][KP] Montgomery-form entries or null
  const uint32_t* e_w[2];      // exponent words
  int e_words;
  size_t e_stride;             // 0: shared exponent
  int ebits[2];
  uint32_t* out_w[2];
  int out_words;
  int count;
  MontCtxArgs ctx[2];
  double* tbl;                 // [gridDim.y * gridDim.x * GPB][1<<WIN][KP] scratch
  const uint32_t* prog[2];     // k_powm_prog only: sliding-window program of the shared exponent (paillier_items.cuh)
  int nprog[2];
};

template <int L, int TPI, int WIN>
__global__ void __launch_bounds__(NT, MinCtas<L>::V) k_powm(PowmArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  const int y = blockIdx.y;
  stage_entries<KS::KP>(smem, p.ctx[y].entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  double* tbl = p.tbl + ((size_t)(y * gridDim.x + blockIdx.x) * KS::GPB + g) * ((size_t)KS::KP << WIN);
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    item_powm<L, TPI, Env, WIN>(p.base_w ? p.base_w + (size_t)item * p.base_words : nullptr, p.base_words,
                                p.base_mont[y] ? p.base_mont[y] + (size_t)item * KS::KP : nullptr,
                                p.e_w[y] + (size_t)item * p.e_stride, p.e_words, p.ebits[y],
                                want < p.count ? p.out_w[y] + (size_t)item * p.out_words : nullptr, p.out_words,
                                smem + ME_N * KS::KP, p.ctx[y].n0inv, smem + ME_R2 * KS::KP, smem + ME_ONEM * KS::KP,
                                smem + ME_ONE * KS::KP, tbl, sm);
  }
}

// ---- shared-exponent sliding-window modexp (decrypt, classic obfuscator) ---------------------------------
constexpr int PROG_WS = 6;   // window width of the programs built by the host: 32 odd powers per table
template <int L, int TPI>
__global__ void __launch_bounds__(NT, MinCtas<L>::V) k_powm_prog(PowmArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  const int y = blockIdx.y;
  stage_entries<KS::KP>(smem, p.ctx[y].entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  double* tbl = p.tbl + ((size_t)(y * gridDim.x + blockIdx.x) * KS::GPB + g) * ((size_t)KS::KP << (PROG_WS - 1));
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    item_powm_prog<L, TPI, Env, PROG_WS>(p.base_w ? p.base_w + (size_t)item * p.base_words : nullptr, p.base_words,
                                         p.base_mont[y] ? p.base_mont[y] + (size_t)item * KS::KP : nullptr,
                                         p.prog[y], p.nprog[y],
                                         want < p.count ? p.out_w[y] + (size_t)item * p.out_words : nullptr,
                                         p.out_words, smem + ME_N * KS::KP, p.ctx[y].n0inv, smem + ME_R2 * KS::KP,
                                         smem + ME_ONEM * KS::KP, smem + ME_ONE * KS::KP, tbl, sm);
  }
}

// ---- decrypt pre-reduction ------------------------------------------------------------------------
struct DecPrepArgs {
  const uint32_t* c_w;  // [count][2*hw]
  int hw;
  double* out[2];       // [count][KP]
  int count;
  MontCtxArgs ctx[2];
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_dec_prep(DecPrepArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  const int y = blockIdx.y;
  stage_entries<KS::KP>(smem, p.ctx[y].entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    // lanes past the end redo the last item and write identical data: benign
    item_dec_prep<L, TPI, Env>(p.c_w + (size_t)item * 2 * p.hw, p.hw, p.out[y] + (size_t)item * KS::KP,
                               smem + ME_N * KS::KP, p.ctx[y].n0inv, smem + ME_R2 * KS::KP, smem + ME_X0 * KS::KP, sm);
  }
}

// ---- decrypt tail -----------------------------------------------------------------------------------
struct DecTailArgs {
  const uint32_t* up_w;
  const uint32_t* uq_w;
  int u_words;
  uint32_t* m_w;
  int m_words;
  int count;
  const double* cst;       // DT_COUNT entries
  uint64_t n0invs[3];
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_dec_tail(DecTailArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.cst, DT_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, DT_COUNT);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    item_dec_tail<L, TPI, Env>(p.up_w + (size_t)item * p.u_words, p.uq_w + (size_t)item * p.u_words, p.u_words,
                               want < p.count ? p.m_w + (size_t)item * p.m_words : nullptr, p.m_words, smem,
                               p.n0invs, sm);
  }
}

// ---- decrypt on the p-adic pair engine: one (ciphertext, modulus) per lane ---------------------------------------
// Work unit = one warp's worth (32 ciphertexts) of ONE modulus: unit u -> modulus u & 1 (0: p, 1: q), ciphertexts
// [32 (u >> 1), +32).  There are no shuffles in this kernel, so warps run independently and pull work from a global
// counter.  A unit is one sequential 1024-bit exponentiation (~21 ms at 2048-bit keys) and 100 000 ciphertexts are
// 6250 units for 1184 resident warps = 5.28 per warp: dealt whole, the last round keeps a fraction of the warps busy
// for a full unit time (r01: 2.64 waves cost 3; r01/r02 with the remainder reserved for one CTA per SM: 5.74 unit
// times for 5.28 of work, tools/tail_probe.py).  So a unit is TIME-SLICED: the host cuts the program of the exponent
// into nseg segments of equal cost (phe_api.cu: split_pair_program), every segment ends by parking the running pair
// (X0, X1) in an extra slot of the unit's window table and the next one starts by reloading it.  Any warp continues any
// unit: the window table and the parked pair live in global memory indexed by UNIT (not by resident warp).  Pieces are
// dealt from a READY QUEUE: pop number t < units is segment 0 of unit t; a warp that finishes segment s of unit u
// parks the pair and pushes (s + 1, u) (fence, ticket from `tail`, st.release of the slot); pop number t >= units
// takes ring slot t - units, spinning (ld.acquire) only while that slot is still empty, i.e. while there is no ready
// work at all.  That is greedy list scheduling: no warp ever waits for a piece that is merely behind in some fixed
// order (a static segment-major order with a per-unit progress counter -- the first version -- stalled whole rounds
// below ~2.3 units per warp: profiles/r02_dec_sweep_static_order.json), pushes = units (nseg - 1) = ring slots, and
// the slot a waiting pop looks at is filled unless every final segment is already done.  The end of a launch is then
// ragged by one SEGMENT instead of one unit, whatever the batch size.
// The limbs of both moduli travel inside the kernel parameters and are selected into registers once per piece.
constexpr int PAIR_MAX_SEG = 16;
struct DecPairArgs {
  const uint32_t* c_w;       // [count][c_words]
  int c_words, chunk_words;
  const uint32_t* prog[2];   // pair-engine programs for x = p, q (paillier_items.cuh: PairOp), all segments back to back
  int seg_off[2][PAIR_MAX_SEG];   // start of segment s inside prog[y]
  int nseg;                  // segments per unit (1: the whole program in one piece); the same for both moduli: the host
                             // pads the shorter program with empty segments ([PO_END])
  uint32_t* out_w[2];        // m_p, m_q: [count][out_words]
  int out_words;
  int count;
  uint64_t n0inv[2];
  const double* cst[2];      // [PC_COUNT][2][L] constant pairs
  double* tbl;               // [units][slots + 1][2][L][32]: window table + the parked pair of every unit
  int slots;
  int* sched;                // zeroed before the launch: [PAIR_SCHED_HEAD] pops, [PAIR_SCHED_TAIL] pushes, [PAIR_SCHED_RING + k]
};                           // ring slot k: 1 + segment * units + unit of the k-th piece that became ready after segment 0
constexpr int PAIR_SCHED_HEAD = 0, PAIR_SCHED_TAIL = 32, PAIR_SCHED_RING = 64;
template <int L> struct ModLimbs { double v[2][L]; };

template <int L> struct PairShape {
  static constexpr int LE = (L + 2) & ~1;                       // entries of E per lane (L used), padded
  static constexpr int PER_LANE = 4 * L + LE;                   // doubles of shared memory per lane
  // threads per CTA.  L = 30 needs 1216 B of shared memory per lane: one 128-thread CTA per SM is all that fits; a
  // 160-thread CTA (a fifth warp, 195 KB) was measured and is no faster (518 vs 513 ms per 100 000 at 3072-bit keys)
  static constexpr int NTP = NT;
  static constexpr int CTAS = (L > 20) ? 1 : 2;
  // L > 20: the 2 L modulus limbs do not fit the register file next to the accumulators (k_dec_pair<30> spilled
  // 256 bytes with them in registers): they are staged into shared memory instead ([2][LE] doubles)
  static constexpr bool MOD_IN_SMEM = L > 20;
  static constexpr int MOD_DOUBLES = MOD_IN_SMEM ? 2 * LE : 0;
  static constexpr size_t smem_bytes() { return (size_t)(MOD_DOUBLES + PER_LANE * NTP) * sizeof(double); }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int L> __global__ void __launch_bounds__(PairShape<L>::NTP, PairShape<L>::CTAS) k_dec_pair(const DecPairArgs p, const ModLimbs<L> mod) {
  using PE = DevPairEnv;
  using PS = PairShape<L>;
  extern __shared__ __align__(16) double smem[];
  if (PS::MOD_IN_SMEM)
    for (int i = threadIdx.x; i < 2 * L; i += PS::NTP) smem[(i / L) * PS::LE + i % L] = mod.v[i / L][i % L];
  __syncthreads();
  const int warp = threadIdx.x >> 5, col = threadIdx.x & 31;
  double* wbase = smem + PS::MOD_DOUBLES + (size_t)warp * PS::PER_LANE * 32 + col;
  PairSmem<PE> sm;
  sm.xs0 = wbase;
  sm.x1 = wbase + L * 32;
  sm.y0 = wbase + 2 * L * 32;
  sm.y1 = wbase + 3 * L * 32;
  sm.e = reinterpret_cast<int64_t*>(wbase + 4 * L * 32);

  const int blocks = (p.count + 31) / 32;
  const int units = 2 * blocks;
  const int total = units * p.nseg;
  int* const ring = p.sched + PAIR_SCHED_RING;
#pragma unroll 1
  for (;;) {
    int t = 0;
    if (col == 0) {
      t = atomicAdd(p.sched + PAIR_SCHED_HEAD, 1);
      if (t >= units && t < total) {   // a piece some warp has made ready (or is about to)
        int v;
        while ((v = ld_acquire_gpu(ring + (t - units))) == 0) __nanosleep(128);
        t = v - 1;                     // segment * units + unit
      }
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total) break;
    __syncwarp();                      // the acquire of lane 0 orders the other lanes' loads of the parked pair / table
    const int seg = t / units, u = t - seg * units;
    const int y = u & 1;
    const int want = (u >> 1) * 32 + col;
    const int item = want < p.count ? want : p.count - 1;
    double* tbl = p.tbl + (size_t)u * ((size_t)(p.slots + 1) * 2 * L * 32) + col;
    const uint32_t* prog = p.prog[y] + p.seg_off[y][seg];
    if (PS::MOD_IN_SMEM) {
      item_dec_pair<L, PE>(p.c_w + (size_t)item * p.c_words, p.chunk_words, prog,
                           want < p.count ? p.out_w[y] + (size_t)item * p.out_words : nullptr, p.out_words,
                           smem + y * PS::LE, p.n0inv[y], p.cst[y], tbl, sm);
    } else {
      double n[L];
#pragma unroll
      for (int j = 0; j < L; ++j) n[j] = y ? mod.v[1][j] : mod.v[0][j];
      item_dec_pair<L, PE>(p.c_w + (size_t)item * p.c_words, p.chunk_words, prog,
                           want < p.count ? p.out_w[y] + (size_t)item * p.out_words : nullptr, p.out_words, n,
                           p.n0inv[y], p.cst[y], tbl, sm);
    }
    // NOTE on code generation: this kernel sits at 252-255 registers and ptxas' row loop (mont52.cuh: pair_pass) flips
    // between two forms with any change to what is live across item_dec_pair -- e.g. a per-modulus segment count cost
    // 3 % (119.3 vs 115.6 ms): the loop re-derived the shared-memory address of D inside every row (S2UR / ULEA + three
    // extra branches per iteration).  After touching this function check `cuobjdump -sass build/pair_shapes.o`: the hot loop
    // of k_dec_pair<20> is 918 instructions with 3 BRA.
    // (k_dec_pair<30> spills 448 B here.  Keeping only the piece number across the exponentiation and deriving (segment,
    // unit) again afterwards brings that to 376 B with a cleaner-looking row loop -- and 481 instead of 458 ms per 100 000
    // at 3072-bit keys; the same spelling at L = 20 makes ptxas re-derive shared-memory addresses inside the row loop.
    // Measured r02, not kept.)
    if (seg + 1 < p.nseg) {   // publish: every lane's stores, then the ring slot
      __syncwarp();
      if (col == 0) {
        __threadfence();
        const int k = atomicAdd(p.sched + PAIR_SCHED_TAIL, 1);
        st_release_gpu(ring + k, 1 + (seg + 1) * units + u);
      }
    }
  }
}

struct DecCrtArgs {
  const uint32_t* mp_w;
  const uint32_t* mq_w;
  int half_words;
  uint32_t* m_w;
  int m_words;
  int count;
  const double* cst;       // DT_COUNT entries
  uint64_t n0invs[3];
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_dec_crt(DecCrtArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.cst, DT_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, DT_COUNT);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    item_dec_crt<L, TPI, Env>(p.mp_w + (size_t)item * p.half_words, p.mq_w + (size_t)item * p.half_words, p.half_words,
                              want < p.count ? p.m_w + (size_t)item * p.m_words : nullptr, p.m_words, smem, p.n0invs, sm);
  }
}

// ---- DJN encrypt (fixed-base comb) -----------------------------------------------------------------
struct EncCombArgs {
  const uint32_t* m_w;   // [count][m_words]
  int m_words;
  const uint32_t* r_w;   // [count][r_words] or null (make_secure = false)
  int r_words;
  int nwin;
  int wb;                // comb digit width in bits (<= 16)
  uint32_t* out_w;       // [count][out_words]
  int out_words;
  int count;
  MontCtxArgs ctx;       // n^2 context; ME_X0 = n*R mod n^2
  const double* comb;    // [nwin][1 << wb][KP]
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_encrypt_comb(EncCombArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    item_encrypt_comb<L, TPI, Env>(p.m_w + (size_t)item * p.m_words, p.m_words,
                                      p.r_w ? p.r_w + (size_t)item * p.r_words : nullptr, p.r_words, p.nwin, p.wb,
                                      want < p.count ? p.out_w + (size_t)item * p.out_words : nullptr, p.out_words,
                                      smem + ME_N * KS::KP, p.ctx.n0inv, smem + ME_X0 * KS::KP, p.comb, sm);
  }
}

// ---- ct = (1 + m n) * obf (classic path) / ct * obf ---------------------------------------------------
struct EncFinishArgs {
  const uint32_t* m_w;
  int m_words;
  const uint32_t* obf_w;
  uint32_t* out_w;
  int out_words;
  int count;
  MontCtxArgs ctx;
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_encrypt_finish(EncFinishArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  const int g = threadIdx.x / TPI;
  for (int base = blockIdx.x * KS::GPB; base < p.count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < p.count ? want : p.count - 1;
    item_encrypt_finish<L, TPI, Env>(p.m_w + (size_t)item * p.m_words, p.m_words,
                                     p.obf_w + (size_t)item * p.out_words,
                                     want < p.count ? p.out_w + (size_t)item * p.out_words : nullptr, p.out_words,
                                     smem + ME_N * KS::KP, p.ctx.n0inv, smem + ME_X0 * KS::KP, smem + ME_R2 * KS::KP, sm);
  }
}

// ---- comb table construction (once per DJN key, on the first obfuscated encrypt) ------------------------------
// T[j][d] = hs^(d 2^(wb j)) R mod n^2, j < nwin, d < 2^wb.
// Phase A (k_comb_bases, one group): the chain hs^(2^i) R, i = 0 .. wb*nwin - 1, by repeated squaring;
//   element i = wb j + k is T[j][2^k].  T[j][0] = R mod N is written alongside.
// Phase B (k_comb_level, one launch per k = 1 .. wb-1): T[j][2^k + e] = T[j][e] * T[j][2^k] for 0 < e < 2^k --
//   nwin (2^k - 1) independent products per level, 4.2 M in total at wb = 16 (~15 ms).
struct CombArgs {
  const uint32_t* hs_w;   // hs canonical words
  int hs_words;
  int nwin;
  int wb;
  int level;              // k_comb_level: k
  double* comb;           // [nwin][1 << wb][KP]
  MontCtxArgs ctx;
};

template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_comb_bases(CombArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  double x[L];
  // every group of the single CTA computes the same chain; only group 0 stores
  const bool writer = (threadIdx.x / TPI) == 0;
  limbs_from_words<L, TPI, Env>(x, p.hs_w, p.hs_words);
  const double* bp = smem + ME_R2 * KS::KP;
  const int total = p.wb * p.nwin;
#pragma unroll 1
  for (int i = 0; i < total; ++i) {   // product i yields hs^(2^i) R
    montmul<L, TPI, Env>(x, x, bp, smem + ME_N * KS::KP, p.ctx.n0inv);
    const int j = i / p.wb, k = i - j * p.wb;
    double* row = p.comb + ((size_t)j << p.wb) * KS::KP;
    if (writer) {
      limbs_to_mem<L, TPI, Env>(row + ((size_t)1 << k) * KS::KP, x);
      if (k == 0) copy_entry<L, TPI, Env>(row, smem + ME_ONEM * KS::KP);
    }
    Env::sync();
    limbs_to_mem<L, TPI, Env>(sm.b0, x);
    Env::sync();
    bp = sm.b0;
  }
}

template <int L, int TPI> __global__ void __launch_bounds__(NT, MinCtas<L>::V) k_comb_level(CombArgs p) {
  using Env = DevEnv<TPI>;
  using KS = KShape<L, TPI>;
  extern __shared__ __align__(16) double smem[];
  stage_entries<KS::KP>(smem, p.ctx.entries, ME_COUNT);
  GroupSmem sm = group_smem<L, TPI>(smem, ME_COUNT);
  double x[L];
  const int g = threadIdx.x / TPI;
  const int per_win = (1 << p.level) - 1;
  const int count = p.nwin * per_win;
  for (int base = blockIdx.x * KS::GPB; base < count; base += gridDim.x * KS::GPB) {
    const int want = base + g;
    const int item = want < count ? want : count - 1;   // past the end: redo the last product (same value stored)
    const int j = item / per_win, e = 1 + item % per_win;
    double* row = p.comb + ((size_t)j << p.wb) * KS::KP;
    Env::sync();
    copy_entry<L, TPI, Env>(sm.b0, row + ((size_t)1 << p.level) * KS::KP);
    Env::sync();
    load_entry<L, TPI, Env>(x, row + (size_t)e * KS::KP);
    montmul<L, TPI, Env>(x, x, sm.b0, smem + ME_N * KS::KP, p.ctx.n0inv);
    limbs_to_mem<L, TPI, Env>(row + (((size_t)1 << p.level) + e) * KS::KP, x);
  }
}

}  // namespace phe
