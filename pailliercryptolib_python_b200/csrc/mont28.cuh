// mont28.cuh -- radix-2^28 cooperative Montgomery arithmetic for sm_100a.
//
// Replaces the reference's ipcl::modExp -> mbx_exp_mb8 inner loop (reached from
// /root/reference/src/ipcl_python/bindings/ipcl_bindings_classes.cpp:53-60 (encrypt), :127-133 (decrypt),
// :318-325 (CipherText + and *)).
//
// Why radix 2^28 and not MADC chains: measured on B200 (profiles/r01_pipe_probe.json) IMAD.WIDE.U32
// issues at 64 lanes/clk/SM (18.4 T/s) but the carry-chained IMAD.WIDE.U32.X form that mad.lo.cc/madc.hi.cc
// compiles to runs at half that (9.1 T/s), and IADD3/LOP3/SHF share the issue budget.  With 28-bit limbs a
// 64-bit column accumulator absorbs 2 products per row for up to 127 rows with no carry handling at all;
// columns are re-split when they pass column 0 of their owner thread, which doubles as the normalisation.
//
// Layout: a K = L*TPI limb number is spread over TPI adjacent lanes of a warp ("group"); lane t holds limbs
// [t*L, (t+1)*L) in registers.  The multiplier operand b is read row by row from shared memory.
//
// The same source compiles for the host (emulation used by the CPU tests: tests/emu) -- no inline PTX.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PHE_HD __host__ __device__ __forceinline__
#define PHE_D __device__ __forceinline__
#else
#define PHE_HD inline
#define PHE_D inline
#endif

namespace phe {

constexpr int LW = 28;                       // limb width in bits
constexpr uint32_t LMASK = (1u << LW) - 1u;  // limb mask

// Row stride of the shared-memory b operand: each lane-block of L limbs is padded to a multiple of 4 words
// so the row loop can fetch 4 rows with one LDS.128.
template <int L> struct Pad { static constexpr int LP = (L + 3) & ~3; };

PHE_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  sh &= 31u;
  return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
#endif
}

struct alignas(16) U4 { uint32_t x, y, z, w; };

#if defined(__CUDACC__)
// Device environment: a group is TPI adjacent lanes.
template <int TPI> struct DevEnv {
  static PHE_D int lane() { return (int)(threadIdx.x & (TPI - 1)); }
  static PHE_D uint32_t bcast(uint32_t v, int src) {
    if (TPI == 1) return v;
    return __shfl_sync(0xffffffffu, v, src, TPI);
  }
  // value held by lane+1 (top lane: unspecified, caller masks)
  static PHE_D uint32_t from_above(uint32_t v) {
    if (TPI == 1) return 0u;
    return __shfl_down_sync(0xffffffffu, v, 1, TPI);
  }
  // value held by lane-1 (lane 0: unspecified, caller masks)
  static PHE_D uint32_t from_below(uint32_t v) {
    if (TPI == 1) return 0u;
    return __shfl_up_sync(0xffffffffu, v, 1, TPI);
  }
  static PHE_D void sync() { __syncwarp(); }
};
#endif

// ------------------------------------------------------------------------------------------------
// Montgomery product, R = 2^(28*L*TPI).
//   r = a * b * R^-1 mod n, as "almost normalised" limbs (< 2^28 + 2^9), value < 2n provided
//   a*b < 2^44 * n * n (always true here: R > 2^48 n for every supported modulus size).
//   a: registers (limbs may be up to 2^29 + 2^10), b: shared memory, K limbs in [TPI][LP] padded layout,
//   limbs < 2^28 + 2^9.   n: registers, exact limbs.  n0inv = -n^-1 mod 2^28.
//   CAPQ: also return the Montgomery quotient digits q_i (lane t gets digits [t*L, (t+1)*L)).
// Overflow bound: a column lives for L rows and gains < 2^58.2 per row -> needs L <= 50.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env, bool CAPQ = false>
PHE_HD void montmul(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* b, const uint32_t (&n)[L],
                    uint32_t n0inv, uint32_t* qcap = nullptr) {
  static_assert(L >= 2 && L <= 50, "limbs per lane out of range");
  constexpr int LP = Pad<L>::LP;
  const int lane = Env::lane();
  const uint32_t topmask = (lane == TPI - 1) ? 0u : 0xffffffffu;
  uint64_t acc[L];
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] = 0;

#pragma unroll 1
  for (int blk = 0; blk < TPI; ++blk) {
    const U4* brow = reinterpret_cast<const U4*>(b + blk * LP);
#pragma unroll
    for (int c4 = 0; c4 < LP / 4; ++c4) {
      const U4 bv = brow[c4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        constexpr int dummy = 0; (void)dummy;
        const int r0 = c4 * 4 + e;   // row within the block; column c lives in acc[(c + r0) % L]
        if (r0 < L) {
          const uint32_t bi = (e == 0) ? bv.x : (e == 1) ? bv.y : (e == 2) ? bv.z : bv.w;
#pragma unroll
          for (int c = 0; c < L; ++c) acc[(c + r0) % L] += (uint64_t)a[c] * bi;
          uint32_t q = ((uint32_t)acc[r0 % L] * n0inv) & LMASK;
          q = Env::bcast(q, 0);
          if (CAPQ) { if (lane == blk) qcap[r0] = q; }
#pragma unroll
          for (int c = 0; c < L; ++c) acc[(c + r0) % L] += (uint64_t)n[c] * q;
          const uint64_t low = acc[r0 % L];
          acc[(r0 + 1) % L] += low >> LW;
          const uint32_t in = Env::from_above((uint32_t)low & LMASK) & topmask;
          acc[r0 % L] = in;  // new top column
        }
      }
    }
  }

  // Three-stage parallel normalisation (no ripple): limbs end < 2^28 + 2^9.
  uint32_t s1[L];
  uint64_t c1[L];
#pragma unroll
  for (int j = 0; j < L; ++j) { s1[j] = (uint32_t)acc[j] & LMASK; c1[j] = acc[j] >> LW; }
  uint32_t c1lo = Env::from_below((uint32_t)c1[L - 1]);
  uint32_t c1hi = Env::from_below((uint32_t)(c1[L - 1] >> 32));
  if (lane == 0) { c1lo = 0; c1hi = 0; }
  const uint64_t c1in = ((uint64_t)c1hi << 32) | c1lo;
  uint32_t c2prev;
  {
    // column L-1 first so its c2 can be shuffled while the rest is computed
    const uint64_t t = (uint64_t)s1[L - 1] + c1[L - 2];
    c2prev = Env::from_below((uint32_t)(t >> LW));
    if (lane == 0) c2prev = 0;
  }
#pragma unroll
  for (int j = 0; j < L; ++j) {
    const uint64_t t = (uint64_t)s1[j] + (j == 0 ? c1in : c1[j - 1]);
    r[j] = ((uint32_t)t & LMASK) + c2prev;
    c2prev = (uint32_t)(t >> LW);
  }
}

// ------------------------------------------------------------------------------------------------
// Exact helpers (used once per result, not in the exponentiation loop)
// ------------------------------------------------------------------------------------------------

// Local ripple: x limbs (any < 2^31) += cin; returns carry out; limbs exact afterwards.
template <int L> PHE_HD uint32_t ripple_add(uint32_t (&x)[L], uint32_t cin) {
  uint32_t c = cin;
#pragma unroll
  for (int j = 0; j < L; ++j) { const uint32_t v = x[j] + c; x[j] = v & LMASK; c = v >> LW; }
  return c;
}

// Make limbs exact (< 2^28) across the whole group.  Value must fit in K limbs.
template <int L, int TPI, class Env> PHE_HD void normalize_exact(uint32_t (&x)[L]) {
  const int lane = Env::lane();
  uint32_t cout = ripple_add<L>(x, 0u);
#pragma unroll 1
  for (int s = 1; s < TPI; ++s) {
    uint32_t cin = Env::from_below(cout);
    if (lane == 0) cin = 0;
    cout = ripple_add<L>(x, cin);
  }
}

// x (exact limbs) -= y (exact limbs) across the group; returns 1 (in every lane) if the result went negative
// (then x holds the result mod 2^(28K)).
template <int L, int TPI, class Env> PHE_HD uint32_t sub_exact(uint32_t (&x)[L], const uint32_t (&y)[L]) {
  const int lane = Env::lane();
  uint32_t bw = 0;
#pragma unroll
  for (int j = 0; j < L; ++j) { const uint32_t v = x[j] - y[j] - bw; x[j] = v & LMASK; bw = v >> 31; }
  uint32_t any = bw;   // a lane borrows out at most once over all rounds; the top lane's is the sign
#pragma unroll 1
  for (int s = 1; s < TPI; ++s) {
    uint32_t b2 = Env::from_below(bw);
    if (lane == 0) b2 = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) { const uint32_t v = x[j] - b2; x[j] = v & LMASK; b2 = v >> 31; }
    bw = b2;
    any |= bw;
  }
  return Env::bcast(any, TPI - 1);
}

// x (exact limbs) += y (exact limbs) across the group (carry out of the top is dropped).
template <int L, int TPI, class Env> PHE_HD void add_exact(uint32_t (&x)[L], const uint32_t (&y)[L]) {
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] += y[j];
  normalize_exact<L, TPI, Env>(x);
}

// if x >= n: x -= n   (x exact limbs)
template <int L, int TPI, class Env> PHE_HD void cond_sub(uint32_t (&x)[L], const uint32_t (&n)[L]) {
  uint32_t d[L];
#pragma unroll
  for (int j = 0; j < L; ++j) d[j] = x[j];
  const uint32_t neg = sub_exact<L, TPI, Env>(d, n);
  if (!neg) {
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = d[j];
  }
}

// Almost-normalised value < 2n  ->  canonical value in [0, n), exact limbs.
template <int L, int TPI, class Env> PHE_HD void canonicalize(uint32_t (&x)[L], const uint32_t (&n)[L]) {
  normalize_exact<L, TPI, Env>(x);
  cond_sub<L, TPI, Env>(x, n);
}

// ------------------------------------------------------------------------------------------------
// Layout conversion through a group-private shared-memory scratch area.
// ------------------------------------------------------------------------------------------------

// Limbs [lane*L, lane*L+L) of the little-endian u32 word array w[0..nwords) (words beyond are zero).
template <int L, int TPI, class Env>
PHE_HD void limbs_from_words(uint32_t (&x)[L], const uint32_t* w, int nwords) {
  const int lane = Env::lane();
#pragma unroll
  for (int j = 0; j < L; ++j) {
    const int bit = (lane * L + j) * LW;
    const int wi = bit >> 5;
    const uint32_t lo = (wi < nwords) ? w[wi] : 0u;
    const uint32_t hi = (wi + 1 < nwords) ? w[wi + 1] : 0u;
    x[j] = funnel_r(lo, hi, (uint32_t)(bit & 31)) & LMASK;
  }
}

// Lane-block padded limb array in shared memory ([TPI][LP]) <- registers.  Caller syncs.
template <int L, int TPI, class Env> PHE_HD void limbs_to_smem(uint32_t* dst, const uint32_t (&x)[L]) {
  constexpr int LP = Pad<L>::LP;
  uint32_t* d = dst + Env::lane() * LP;
#pragma unroll
  for (int j = 0; j < L; ++j) d[j] = x[j];
#pragma unroll
  for (int j = L; j < LP; ++j) d[j] = 0u;
}

template <int L, int TPI, class Env> PHE_HD void limbs_from_smem(uint32_t (&x)[L], const uint32_t* src) {
  constexpr int LP = Pad<L>::LP;
  const uint32_t* s = src + Env::lane() * LP;
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = s[j];
}

// Word v of the number whose exact limbs sit in the padded smem limb array.
template <int L, int TPI> PHE_HD uint32_t word_from_smem_limbs(const uint32_t* limbs, int v) {
  constexpr int LP = Pad<L>::LP;
  constexpr int K = L * TPI;
  const int bit = v * 32;
  const int g = bit / LW;
  const int o = bit - g * LW;
  uint64_t u = 0;
  if (g < K) u = limbs[(g / L) * LP + (g % L)];
  if (g + 1 < K) u |= (uint64_t)limbs[((g + 1) / L) * LP + ((g + 1) % L)] << LW;
  uint32_t res = (uint32_t)(u >> o);
  if (o > 24 && g + 2 < K) res |= limbs[((g + 2) / L) * LP + ((g + 2) % L)] << (56 - o);
  return res;
}

}  // namespace phe
