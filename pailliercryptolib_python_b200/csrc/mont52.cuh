// mont52.cuh -- radix-2^52 cooperative Montgomery arithmetic on the FP64 pipe of sm_100a.
//
// Replaces the reference's ipcl::modExp -> mbx_exp_mb8 inner loop (reached from
// /root/reference/src/ipcl_python/bindings/ipcl_bindings_classes.cpp:53-60 (encrypt), :127-133 (decrypt),
// :318-325 (CipherText + and *)).  mbx_exp_mb8 itself keeps radix-2^52 digits in 64-bit AVX-512 lanes and
// multiplies them with vpmadd52{l,h}uq; the B200 has no 52-bit integer multiplier, but its FP64 pipe issues
// DFMA at twice the rate of IMAD.WIDE (profiles/r01_pipe_probe2.json: 17.2 vs 8.2 T lane-ops/s) and one DFMA
// covers a 52x52-bit partial product half, i.e. 2.64 32x32 MACs' worth of bits in two instructions.
//
// Exact double-word product of two 52-bit integers held in doubles (Emmart/Zheng/Weems, ARITH 2018):
//     ph = fma_rz(a, b, 2^104)            = 2^104 + hi * 2^52        (hi = floor(a b / 2^52) sits in the mantissa)
//     pl = fma_rz(a, b, (2^104 + 2^52) - ph) = 2^52 + lo              (lo = a b mod 2^52 sits in the mantissa)
// The raw IEEE bit patterns of ph and pl are summed into 64-bit integer column accumulators (IADD3 on the ALU
// pipe, which is otherwise idle); the exponent-field constants 0x467<<52 and 0x433<<52 only touch the top 12
// bits, so everything done modulo 2^52 (the Montgomery quotient digit) can ignore them, and each column is
// pre-loaded with minus the total bias it will collect over its lifetime, so it is a true integer when retired.
//
// Layout: a K = L*TPI limb number is spread over TPI adjacent lanes ("group"); lane t holds limbs
// [t*L, (t+1)*L) as doubles in registers.  The multiplier operand b is read row by row from shared memory
// (doubles, [TPI][LP] padded layout).  All limbs entering and leaving montmul are exact integers < 2^52
// (required: a product must stay below 2^104).
//
// The same source compiles for the host (lock-step lane emulator of the CPU tests: tests/emu); the host build
// must run with the rounding mode set to FE_TOWARDZERO and be compiled with -frounding-math.
#pragma once
#include <cstdint>
#include <cstring>
#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

#if defined(__CUDACC__)
#define PHE_HD __host__ __device__ __forceinline__
#define PHE_D __device__ __forceinline__
#else
#define PHE_HD inline
#define PHE_D inline
#endif

namespace phe {

constexpr int LW = 52;
constexpr uint64_t M52 = (1ull << 52) - 1ull;
constexpr uint64_t BIAS_LO = 0x433ull << 52;   // bit pattern of 2^52
constexpr uint64_t BIAS_HI = 0x467ull << 52;   // bit pattern of 2^104
constexpr double TWO52 = 4503599627370496.0;                      // 2^52
constexpr double TWO104 = 20282409603651670423947251286016.0;     // 2^104
constexpr double TWO104P52 = 20282409603651674927546878656512.0;  // 2^104 + 2^52

// limbs per lane padded to an even count: every lane block starts 16-byte aligned
// ... and never a multiple of 16 doubles: lane t reads limb j of ITS block (modulus, operands) at t * LP + j, and
// with LP = 16 (L = 15) every lane of a group hits the same shared-memory bank -- ncu on k_encrypt_npair<15,4>:
// short_scoreboard 3.1 stalled warps per issue, issue_active 0.36.  LP = 18 spreads LDS.128 of 4 / 8 lanes over
// 16 / 32 distinct banks.
template <int L> struct Pad {
  static constexpr int LP0 = (L + 1) & ~1;
  static constexpr int LP = (LP0 % 16 == 0) ? LP0 + 2 : LP0;
};

template <int L, int TPI> struct Shape {
  static constexpr int LP = Pad<L>::LP;
  static constexpr int K = L * TPI;
  static constexpr int KP = LP * TPI;      // doubles per entry
  static constexpr int BITS = K * LW;
};

PHE_HD double u2d(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)v);
#else
  double d; std::memcpy(&d, &v, 8); return d;
#endif
}
PHE_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t v; std::memcpy(&v, &d, 8); return v;
#endif
}
PHE_HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rz(a, b, c);
#else
  return std::fma(a, b, c);   // host: thread rounding mode is FE_TOWARDZERO (tests/emu)
#endif
}
// integer < 2^52 <-> limb (double holding that integer)
PHE_HD double limb_of(uint64_t x) { return u2d(x | BIAS_LO) - TWO52; }
PHE_HD uint64_t int_of(double d) { return d2u(d + TWO52) & M52; }

#if defined(__CUDACC__)
template <int TPI> struct DevEnv {
  static PHE_D int lane() { return (int)(threadIdx.x & (TPI - 1)); }
  static PHE_D uint32_t bcast(uint32_t v, int src) {
    if (TPI == 1) return v;
    return __shfl_sync(0xffffffffu, v, src, TPI);
  }
  static PHE_D uint32_t from_above(uint32_t v) {
    if (TPI == 1) return 0u;
    return __shfl_down_sync(0xffffffffu, v, 1, TPI);
  }
  static PHE_D uint32_t from_below(uint32_t v) {
    if (TPI == 1) return 0u;
    return __shfl_up_sync(0xffffffffu, v, 1, TPI);
  }
  // true if the predicate holds in any lane of (a superset of) the group; callers are warp-convergent
  static PHE_D bool any(bool p) {
    if (TPI == 1) return p;
    return __any_sync(0xffffffffu, p) != 0;
  }
  static PHE_D void sync() { __syncwarp(); }
  // 16-byte asynchronous global -> shared copy (LDGSTS) and the matching wait; used to prefetch the next comb-table
  // entry of the DJN encrypt under the current Montgomery product
  static PHE_D void cp_async16(void* dst_shared, const void* src_global) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_shared);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_global) : "memory");
  }
  static PHE_D void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
  // bulk prefetch of `bytes` (a multiple of 16, 16-byte aligned) from global memory into the L2 by the TMA unit
  // (cp.async.bulk.prefetch.L2: one instruction per entry, no registers, no shared memory): the comb-table entry of the
  // next window of the DJN encrypt is pulled out of HBM while the current product runs (npair_items.cuh: NPairEncCtl)
  static PHE_D void prefetch_l2(const void* src_global, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_global), "r"(bytes) : "memory");
  }
};
#endif

#if defined(__CUDACC__)
// One bignum per lane (the p-adic pair engine): a "group" is a single lane, per-lane operands sit in shared memory as
// columns of a [limb][32] matrix.
struct DevPairEnv {
  static constexpr int STRIDE = 32;
  static PHE_D int lane() { return 0; }
  static PHE_D int column() { return (int)(threadIdx.x & 31); }
  static PHE_D uint32_t bcast(uint32_t v, int) { return v; }
  static PHE_D uint32_t from_above(uint32_t) { return 0u; }
  static PHE_D uint32_t from_below(uint32_t) { return 0u; }
  static PHE_D bool any(bool p) { return p; }
  static PHE_D void sync() { __syncwarp(); }
  // 8-byte asynchronous global -> shared copy (LDGSTS, no staging registers) and the matching wait: the table entry of
  // the next multiplication is fetched under the squarings that precede it (paillier_items.cuh: item_dec_pair)
  static PHE_D void cp_async8(void* dst_shared, const void* src_global) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_shared);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src_global) : "memory");
  }
  static PHE_D void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
};
#endif

template <class Env> PHE_HD uint64_t bcast64(uint64_t v, int src) {
  const uint32_t lo = Env::bcast((uint32_t)v, src), hi = Env::bcast((uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
template <class Env> PHE_HD uint64_t from_above64(uint64_t v) {
  const uint32_t lo = Env::from_above((uint32_t)v), hi = Env::from_above((uint32_t)(v >> 32));
  return ((uint64_t)hi << 32) | lo;
}

// bias a column must be pre-charged with (negated) when it will still receive `nlo` low halves and `nhi` high halves
PHE_HD constexpr uint64_t bias_of(int nlo, int nhi) {
  return (uint64_t)nlo * BIAS_LO + (uint64_t)nhi * BIAS_HI;   // mod 2^64
}

// The row loops below add the multiplicand part of row i + 1 inside the iteration of row i.  After the LAST row that part
// used to be skipped by a branch -- a conditional region in the last row of every unrolled chunk, inside which ptxas cannot
// interleave the quotient-digit chain (LDS -> DFMA -> DADD -> DFMA -> IADD3 -> 4 IMAD) with other products: ncu r02 put
// 6.4 % of the stall samples of k_dec_pair<20> on those 11 instructions.  With PHE52_GHOST the part always runs, after the
// last row with b = 0: L "ghost" products per pass that only add their exponent-field biases (one low half to every
// result column, one high half to every column but the first), which final_bias takes off again at compile time.
// k_dec_pair<20> 112.7 -> 108.6 ms, k_dec_pair<30> 451 -> 444 ms.  The multi-lane products (montmul, montmul_e: K = L TPI
// rows per pass, so the ghost products are a smaller share, but so is the gain) lose: k_encrypt_npair<20,2> 15.62 -> 15.82
// ms, HE mul 5.83 -> 5.72 M/s, HE add +0.3 %; they keep the branch (PHE52_GHOST_MULTI = 0).
#ifndef PHE52_GHOST
#define PHE52_GHOST 1
#endif
#ifndef PHE52_GHOST_MULTI
#define PHE52_GHOST_MULTI 0
#endif
PHE_HD constexpr uint64_t final_bias(int j, bool ghost) {
  return ghost ? bias_of(2 * j + 1, j ? 2 * j - 1 : 0) : bias_of(2 * (j + 1), 2 * j);
}

template <int L> PHE_HD uint32_t ripple(uint64_t (&x)[L], uint32_t cin) {
  uint64_t c = cin;
#pragma unroll
  for (int j = 0; j < L; ++j) { const uint64_t v = x[j] + c; x[j] = v & M52; c = v >> LW; }
  return (uint32_t)c;
}

// Columns (true integers, each < 2^63) -> exact limbs < 2^52 across the whole group.  Value must fit in K limbs.
template <int L, int TPI, class Env> PHE_HD void normalize_exact(uint64_t (&x)[L]) {
  const int lane = Env::lane();
  uint32_t cout = ripple<L>(x, 0u);
#pragma unroll 1
  for (int s = 1; s < TPI; ++s) {
    uint32_t cin = Env::from_below(cout);
    if (lane == 0) cin = 0;
    x[0] += cin;
    cout = 0;
    if (Env::any(x[0] > M52)) cout = ripple<L>(x, 0u);   // a second-order carry: probability ~2^-40 per product
  }
}

// ------------------------------------------------------------------------------------------------
// Product spans.  mac_first: column += lo(x*y), hprev = hi(x*y).  mac_span: for c in [C0, C1):
// acc[(c + rot) % L] += lo(x[c]*y) + hprev; hprev = hi(x[c]*y)  -- written in batches of G independent
// ph / sub / pl chains so that the FP64 pipe always has several products in flight.  G covers the whole span:
// with smaller batches (5, and for some kernels 10) ptxas re-serialises the chains one product at a time
// (~17 stall cycles per DFMA instead of ~7 in the SASS control words, 180 instead of 250 registers).
// x may be a register array or a pointer (modulus limbs in shared memory: identical for every item of a CTA,
// so they cost broadcast LDS.128 instead of 2L live registers per lane).
// ------------------------------------------------------------------------------------------------
#ifndef PHE52_G
#define PHE52_G 20
#endif
PHE_HD void mac_first(uint64_t& col, double x, double y, uint64_t& hprev) {
  const double ph = fma_rz(x, y, TWO104);
  const double pl = fma_rz(x, y, TWO104P52 - ph);
  col += d2u(pl);
  hprev = d2u(ph);
}

template <int L, int C0, int C1, class X, int G = PHE52_G>
PHE_HD void mac_span(uint64_t (&acc)[L], const X& x, double y, uint64_t& hprev, int rot) {
#pragma unroll
  for (int c = C0; c < C1; c += G) {
    double xv[G], ph[G], pl[G];
#pragma unroll
    for (int g = 0; g < G; ++g) if (c + g < C1) xv[g] = x[c + g];
#pragma unroll
    for (int g = 0; g < G; ++g) if (c + g < C1) ph[g] = fma_rz(xv[g], y, TWO104);
#pragma unroll
    for (int g = 0; g < G; ++g) if (c + g < C1) pl[g] = TWO104P52 - ph[g];
#pragma unroll
    for (int g = 0; g < G; ++g) if (c + g < C1) pl[g] = fma_rz(xv[g], y, pl[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) if (c + g < C1) { acc[(c + g + rot) % L] += d2u(pl[g]) + hprev; hprev = d2u(ph[g]); }
  }
}

// The row loop is unrolled U rows at a time (U divides L): inside a chunk column c of row u lives in
// acc[(c + u) % L]; after the chunk the accumulators are rotated back by U places.  A fully unrolled L-row body
// (77 KB at L = 20) misses the 32 KB L1.5 instruction cache on every pass -- ncu: no_instruction = 2.3 stalled
// warps per issue -- so the body is kept under ~20 KB.  U = 4 at L = 20: r02, k_dec_pair<20> 119.9 ms against 120.6 at
// U = 5 (profiles/r02_bench_u4_variant.json).
// A symmetric (triangular) pass 1 of the pair square -- L (L + 1) / 2 instead of L^2 multiplicand products, bit-exact in
// the emulator and on the GPU -- was built and measured in r02 and is NOT in here: its rows need per-block code next to
// the general row loop (27 KB loop body, uniform branches between 4-product blocks) and k_dec_pair<20> went from 120.6
// to 174.9 ms (ncu: no_instruction 1.2 and wait 1.15 stalled warps per issue, 3 % MORE executed instructions).
// profiles/r02_tri_square_experiment.patch, r02_ncu_k_dec_pair_tri_square_summary.txt.  A lighter form -- X0 = lo + hi B^(L/2),
// X0^2 = lo (lo + 2 hi B^(L/2)) + hi^2 B^L: full rows for the low half, half rows for the high half, 3/4 of the products,
// only 4 KB more code -- lost too (139.6 ms; 563 against 459 ms at L = 30): one uniform branch per row is enough to make
// ptxas reconcile the two paths with register moves (IMAD.MOV 0.33 -> 0.89 per product in the loop) and to cut the
// scheduling window.  profiles/r02_half_square_experiment.patch.  With the branch hoisted out -- two row loops, each with a
// straight-line body (14.3 + 12.1 KB, clean instruction mix) -- it is still 134.5 ms (505 at L = 30): the two loop bodies
// together no longer stay in the instruction cache while 8 warps are in different passes.
// profiles/r02_half_square_two_loops_experiment.patch.  The same two loops at U = 2 and U = 1, where both bodies DO fit
// (7.5 + 6.4 KB, 3.9 + 3.3 KB): 118.4 / 129.2 ms against 119.6 / 131.7 for the plain square at the same U and 115.5 at
// U = 4 (profiles/r02_unroll_halfsquare_ab.json) -- every loop back-edge drains the product chains (~50 cycles: U = 2
// costs 3.5 %, U = 1 14 %), and the half rows then give back 1-2 %, not the 5 % of their product count.
// ONE branch-free row loop of ~15 KB is what this kernel can afford.
#ifndef PHE52_U
#define PHE52_U 4
#endif
// L = 30 (one-lane pair engine of 3072-bit keys): 5 rows are 300 products = 27 KB of code, 3 rows 16 KB, 2 rows 11 KB.
// With the branch-free row body (PHE52_GHOST) a back-edge is cheap and the smallest body wins at this size:
// k_dec_pair<30> U = 1 415.9, U = 2 437.7, U = 3 444.2, U = 5 459.1 ms
#ifndef PHE52_U_WIDE
#define PHE52_U_WIDE 1
#endif
template <int L> struct Unroll { static constexpr int U = (L > 24 && L % PHE52_U_WIDE == 0) ? PHE52_U_WIDE : (L % PHE52_U == 0) ? PHE52_U : (L % 5 == 0) ? 5 : (L % 4 == 0) ? 4 : (L % 3 == 0) ? 3 : (L % 2 == 0) ? 2 : 1; };

// index of limb `row` in the padded [TPI][LP] layout
template <int L> PHE_HD int padded_index(int row) {
  constexpr int LP = Pad<L>::LP;
  if (LP == L) return row;
  return row + (row / L) * (LP - L);
}

// ------------------------------------------------------------------------------------------------
// Montgomery product, R = 2^(52*L*TPI):  r = a * b * R^-1 mod n, exact limbs, value < a*b/R + n, i.e. < 2n for
// a < 4n, b < 2n (R >= 8n: every shape is chosen with at least 8 bits of headroom).
//   a: registers; b: shared memory doubles [TPI][LP]; n_entry: modulus, padded entry in shared memory;
//   all exact limbs.  n0inv = -n^-1 mod 2^52.
//   CAPQ: also store the Montgomery quotient digits q_i to qcap (memory, padded [TPI][LP] layout, lane 0 writes).
// Row i adds a*b_i (A-part) and n*q_i (N-part) and retires column 0.  The loop is software-pipelined so the
// quotient digit never sits on the critical path: iteration i does N-part(i) for columns 0,1, retires column 0,
// adds a[0]*b_(i+1), derives q_(i+1) and starts its broadcast, and only then runs the remaining 2(L-1) products
// of N-part(i) and A-part(i+1), which hide the IMAD/SHFL latency of the digit.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env, bool CAPQ = false>
PHE_HD void montmul(double (&r)[L], const double (&a)[L], const double* b, const double* n_entry, uint64_t n0inv,
                    uint64_t* qcap = nullptr) {
  static_assert(L >= 2 && L <= 64, "limbs per lane out of range");
  constexpr int K = L * TPI;
  constexpr int U = Unroll<L>::U;
  constexpr bool GHOST = PHE52_GHOST_MULTI != 0;
  const int lane = Env::lane();
  const double* n = n_entry + lane * Pad<L>::LP;
  const uint64_t topmask = (lane == TPI - 1) ? 0ull : ~0ull;
  constexpr uint64_t INIT = 0ull - bias_of(2 * L, 2 * L);
  uint64_t acc[L];
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] = 0ull - bias_of(2 * (j + 1), 2 * j);

  uint64_t topA, q;
  double qd;
  {  // prologue: A-part of row 0 and its quotient digit
    const double b0 = b[0];
    uint64_t h;
    mac_first(acc[0], a[0], b0, h);
    q = bcast64<Env>((acc[0] * n0inv) & M52, 0);
    mac_span<L, 1, L>(acc, a, b0, h, 0);
    topA = h;
    qd = limb_of(q);
  }
#pragma unroll 1
  for (int row0 = 0; row0 < K; row0 += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {   // row i = row0 + u; column c of row i lives in acc[(c + u) % L]
      const int row = row0 + u;
      const bool last = (u == U - 1) && (row == K - 1);
      if (CAPQ) { if (lane == 0) qcap[padded_index<L>(row)] = q; }
      uint64_t hN, hA = 0;
      const double n0 = n[0], n1 = n[1];
      mac_first(acc[u % L], n0, qd, hN);
      {
        const double ph = fma_rz(n1, qd, TWO104);
        const double pl = fma_rz(n1, qd, TWO104P52 - ph);
        const uint64_t low = acc[u % L];                  // column 0 is complete: retire it
        acc[(u + 1) % L] += d2u(pl) + hN + (low >> LW);
        hN = d2u(ph);
        acc[u % L] = from_above64<Env>(low & M52) & topmask;   // becomes the new top column (column L of row i)
      }
      double bn = 0.0;
      if (GHOST) {   // branch-free: after the last row the multiplicand part runs with b = 0 (see pair_pass)
        bn = b[padded_index<L>(last ? row : row + 1)];
        if (last) bn = 0.0;
        mac_first(acc[(u + 1) % L], a[0], bn, hA);
        q = bcast64<Env>((acc[(u + 1) % L] * n0inv) & M52, 0);
      } else if (!last) {
        bn = b[padded_index<L>(row + 1)];
        mac_first(acc[(u + 1) % L], a[0], bn, hA);
        q = bcast64<Env>((acc[(u + 1) % L] * n0inv) & M52, 0);
      }
      mac_span<L, 2, L>(acc, n, qd, hN, u);
      acc[u % L] += topA + hN + INIT;
      if (GHOST || !last) {
        mac_span<L, 1, L>(acc, a, bn, hA, u + 1);
        topA = hA;
        qd = limb_of(q);
      }
    }
    if (U != L) {   // rotate back: column c returns to acc[c]
      uint64_t t[L];
#pragma unroll
      for (int j = 0; j < L; ++j) t[j] = acc[(j + U) % L];
#pragma unroll
      for (int j = 0; j < L; ++j) acc[j] = t[j];
    }
  }
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] += final_bias(j, GHOST);
  normalize_exact<L, TPI, Env>(acc);
#pragma unroll
  for (int j = 0; j < L; ++j) r[j] = limb_of(acc[j]);
}

// ------------------------------------------------------------------------------------------------
// montmul_e: montmul with the run-time extras the multi-lane pair engine (npair_items.cuh) needs; ONE body serves
// every pass of a pair product (and the final plain product), so a kernel has a single copy of the row loop:
//   ein   (or null): addend E, K + 1 exact limbs as integers in the padded [TPI][LP] layout, top limb at [TPI * LP]:
//                    r = (a b + E + m n) / R.  Lane 0 adds digit i to column 0 of row i.
//   cap   (or null): plain == false: the quotient digits q_i (integers, padded layout; lane 0 writes)
//                    plain == true : the low half of the product (limb i = retired column i)
//   plain          : q is forced to 0: r = floor((a b + E) / R), i.e. an ordinary 2K-limb product with cap as low half.
// cap may alias ein (digit i + 1 of ein is read in the iteration that writes digit i of cap).
// (Adding E whole before the first row, every lane its own L digits -- what pays in the one-lane engine -- loses here:
// k_encrypt_npair<20,2> 15.62 -> 15.96 ms.  So do the branch-free last row, 15.82, and other unrolls, U = 5 15.94, U = 2
// 16.69.  r02 A/B builds; the multi-lane product stays as it was at the start of the round.)
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void montmul_e(double (&r)[L], const double (&a)[L], const double* b, const double* n_entry, uint64_t n0inv,
                      const uint64_t* ein, uint64_t* cap, bool plain) {
  static_assert(L >= 2 && L <= 64, "limbs per lane out of range");
  constexpr int K = L * TPI;
  constexpr int U = Unroll<L>::U;
  constexpr bool GHOST = PHE52_GHOST_MULTI != 0;
  const int lane = Env::lane();
  const double* n = n_entry + lane * Pad<L>::LP;
  const uint64_t topmask = (lane == TPI - 1) ? 0ull : ~0ull;
  const uint64_t qmask = plain ? 0ull : M52;
  const bool lead = (lane == 0);
  constexpr uint64_t INIT = 0ull - bias_of(2 * L, 2 * L);
  uint64_t acc[L];
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] = 0ull - bias_of(2 * (j + 1), 2 * j);

  uint64_t topA, q;
  double qd;
  {  // prologue: A-part of row 0 and its quotient digit
    const double b0 = b[0];
    uint64_t h;
    mac_first(acc[0], a[0], b0, h);
    if (ein && lead) acc[0] += ein[0];
    q = bcast64<Env>((acc[0] * n0inv) & qmask, 0);
    mac_span<L, 1, L>(acc, a, b0, h, 0);
    topA = h;
    qd = limb_of(q);
  }
#pragma unroll 1
  for (int row0 = 0; row0 < K; row0 += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {   // row i = row0 + u; column c of row i lives in acc[(c + u) % L]
      const int row = row0 + u;
      const bool last = (u == U - 1) && (row == K - 1);
      uint64_t hN, hA = 0;
      const double n0 = n[0], n1 = n[1];
      mac_first(acc[u % L], n0, qd, hN);
      {
        const double ph = fma_rz(n1, qd, TWO104);
        const double pl = fma_rz(n1, qd, TWO104P52 - ph);
        const uint64_t low = acc[u % L];                  // column 0 is complete: retire it
        if (cap && lead) cap[padded_index<L>(row)] = plain ? (low & M52) : q;
        acc[(u + 1) % L] += d2u(pl) + hN + (low >> LW);
        hN = d2u(ph);
        acc[u % L] = from_above64<Env>(low & M52) & topmask;   // becomes the new top column (column L of row i)
      }
      double bn = 0.0;
      if (GHOST) {   // branch-free: after the last row the multiplicand part runs with b = 0 (see pair_pass)
        bn = b[padded_index<L>(last ? row : row + 1)];
        if (last) bn = 0.0;
        mac_first(acc[(u + 1) % L], a[0], bn, hA);
        if (ein && lead && !last) acc[(u + 1) % L] += ein[padded_index<L>(row + 1)];
        q = bcast64<Env>((acc[(u + 1) % L] * n0inv) & qmask, 0);
      } else if (!last) {
        bn = b[padded_index<L>(row + 1)];
        mac_first(acc[(u + 1) % L], a[0], bn, hA);
        if (ein && lead) acc[(u + 1) % L] += ein[padded_index<L>(row + 1)];
        q = bcast64<Env>((acc[(u + 1) % L] * n0inv) & qmask, 0);
      }
      mac_span<L, 2, L>(acc, n, qd, hN, u);
      acc[u % L] += topA + hN + INIT;
      if (GHOST || !last) {
        mac_span<L, 1, L>(acc, a, bn, hA, u + 1);
        topA = hA;
        qd = limb_of(q);
      }
    }
    if (U != L) {   // rotate back: column c returns to acc[c]
      uint64_t t[L];
#pragma unroll
      for (int j = 0; j < L; ++j) t[j] = acc[(j + U) % L];
#pragma unroll
      for (int j = 0; j < L; ++j) acc[j] = t[j];
    }
  }
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] += final_bias(j, GHOST);
  if (ein && lead) acc[0] += ein[TPI * Pad<L>::LP];        // top limb of E: column K
  normalize_exact<L, TPI, Env>(acc);
#pragma unroll
  for (int j = 0; j < L; ++j) r[j] = limb_of(acc[j]);
}

// ------------------------------------------------------------------------------------------------
// p-adic pair arithmetic for the CRT half of decrypt:  numbers mod x^2 (x = p or q) are kept as pairs (X0, X1),
// X0, X1 < 2x, meaning (X0 + X1 x) R^-1 mod x^2 with R = 2^(52 L) >= 2^8 x.  With  X0 Y0 = u R - m x  (u, m the
// result and the quotient of an ordinary Montgomery reduction mod x) one gets
//     (X0 + X1 x)(Y0 + Y1 x) R^-1  =  u + (X0 Y1 + X1 Y0 - m) R^-1 x   (mod x^2)
// so a product mod x^2 costs 6 L^2 limb products (3 multiplications + 3 reductions of L limbs) and a square 4 L^2,
// against 8 L^2 for one Montgomery product of 2L limbs -- and everything fits ONE lane (L = 20 at 2048-bit keys):
// no shuffles, 32 ciphertexts per warp.  The L function comes for free: c^(x-1) = 1 + L x.
//
// pair_pass is one reduction:  r = (a b - M + m x) / R  with M = sum_i e_in[i] 2^(52 i) subtracted column by column.
// Pass 1 of a product (a = X0, b = Y0) records its quotient digits, e_out[i] = q_i; pass 2 (a = X0, b = Y1) takes them as
// e_in: that is the "- m" term.  No multiple of x has to be added to keep things non-negative: a b - M + m' x is a multiple
// of R and, with a b >= 0 and M < R, greater than -R, hence >= 0 -- only single columns go negative on the way, and the
// column arithmetic is signed (arithmetic carry shifts, signed final ripple).  (r01 / early r02 fed D_i - q_i with
// D = ceil(R / x) x instead: one load, a masked subtraction and two moves more per row, 2.6 % of the row loop.)
// A third pass (a = X1, b = Y0) gives the other cross term and the two
// are added (a square needs only pass 2 with a = 2 X0, b = X1).  One code body serves every pass: a fused
// two-multiplicand pass saves one reduction per multiplication but needs a second, 26 KB loop body that evicts the hot
// one from the 32 KB instruction cache (measured: no_instruction 0.64 stalled warps per issue).  b is read from, and
// the result written to, (shared) memory with a stride of PE::STRIDE doubles between limbs (one column per lane); n
// is shared by all lanes.  Result: exact limbs, value < a b / R + x.  r_out may alias b.
// ------------------------------------------------------------------------------------------------
// rows per chunk of the one-lane engine: with the branch-free row body a loop back-edge costs little (U = 2: +0.5 %) and
// five rows of 40 products (17.7 KB) still sit in the instruction cache: k_dec_pair<20> U = 5 107.7, U = 4 108.5, U = 2 109.1 ms
// (Offset carries -- every column pre-charged with 2^63 - 2^11 more, so that the carry it hands on is the non-negative
// (value >> 52) and needs no sign word: one SHF less per row, 25 instructions less per pass, bit-exact -- measure the
// same: 107.2 vs 107.0 ms.  r02, not kept.)
// (Fetching the multiplier limb of a chunk's first row one chunk ahead -- ncu puts 3.5 % of the stall samples on the first
// DFMA behind that LDS -- costs two registers across the back-edge and LOSES: 110.4 vs 108.1 ms.  r02, not kept.)
#ifndef PHE52_U_PAIR
#define PHE52_U_PAIR 5
#endif
template <int L> struct PairUnroll { static constexpr int U = (L <= 20 && L % PHE52_U_PAIR == 0) ? PHE52_U_PAIR : Unroll<L>::U; };

template <int L, class PE>
PHE_HD void pair_pass(double* r_out, const double (&a)[L], const double* b, const int64_t* e_in, int64_t* e_out,
                      const double* n, uint64_t n0inv) {
  constexpr int U = PairUnroll<L>::U;
  constexpr int ST = PE::STRIDE;
  // product chains in flight per batch: a whole row (L) at every shape.  At L = 30 that is 90 registers of temporaries next
  // to 60 + 60 for a and the accumulators and ptxas spills a little more outside the row loop, but r02 measured
  // k_dec_pair<30> at 100 000 x 3072 bits: G = 30 451 ms, 20 465, 15 469, 10 457 (tools/ab_wide.sh).
#ifndef PHE52_G_WIDE
#define PHE52_G_WIDE L
#endif
  constexpr int G = (L > 20) ? PHE52_G_WIDE : PHE52_G;
  constexpr uint64_t INIT = 0ull - bias_of(2 * L, 2 * L);
  uint64_t acc[L];
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] = 0ull - bias_of(2 * (j + 1), 2 * j);
  // M is subtracted whole around the first row (digit j sits in window column j of row 0) instead of digit by digit inside
  // the rows: the same 3 instructions per digit, but only in the passes that have an M (r02: k_dec_pair<20> 114.0 -> 113.0
  // ms; k_dec_pair<30> lost with it at 3 rows per chunk, 450.6 -> 462.5 ms, and wins at 1 row per chunk, 416.6 -> 404.4 ms)
#ifndef PHE_PAIR_MUP_MAXL
#define PHE_PAIR_MUP_MAXL 64
#endif
  constexpr bool M_UPFRONT = (L <= PHE_PAIR_MUP_MAXL);
  constexpr bool GHOST = PHE52_GHOST != 0;

  uint64_t topA, q;
  double qd;
  {  // prologue: A-part of row 0 and the quotient digit
    const double b0 = b[0];
    uint64_t h;
    mac_first(acc[0], a[0], b0, h);
    if (e_in) acc[0] -= (uint64_t)e_in[0];
    q = (acc[0] * n0inv) & M52;
    mac_span<L, 1, L, double[L], G>(acc, a, b0, h, 0);
    topA = h;
    qd = limb_of(q);
  }
  // (after the first row, not before it: the pre-charge constants of the columns then fold into the first additions as
  // immediates instead of being moved into 2 L registers first)
  if (M_UPFRONT && e_in) {
#pragma unroll
    for (int j = 1; j < L; ++j) acc[j] -= (uint64_t)e_in[j * ST];
  }
#pragma unroll 1
  for (int row0 = 0; row0 < L; row0 += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {   // row i = row0 + u; column c of row i lives in acc[(c + u) % L]
      const int row = row0 + u;
      const bool last = (u == U - 1) && (row == L - 1);
      if (e_out) e_out[row * ST] = (int64_t)q;
      uint64_t hN, hA = 0;
      const double n0 = n[0], n1 = n[1];
      mac_first(acc[u % L], n0, qd, hN);
      {
        const double ph = fma_rz(n1, qd, TWO104);
        const double pl = fma_rz(n1, qd, TWO104P52 - ph);
        const int64_t low = (int64_t)acc[u % L];         // column 0 is complete (a multiple of 2^52, maybe negative)
        acc[(u + 1) % L] += d2u(pl) + hN + (uint64_t)(low >> LW);
        hN = d2u(ph);
        acc[u % L] = 0;                                    // becomes the new top column (column L of row i)
      }
      double bn = 0.0;
      if (GHOST) {
        // no branch around the multiplicand part of the next row: after the last row it runs with b = 0 and only adds
        // exponent-field biases, which the constants below take off again (20 ghost products per pass against a
        // conditional region in the last row of every chunk, where ptxas cannot interleave the quotient-digit chain)
        bn = b[(last ? row : row + 1) * ST];
        if (last) bn = 0.0;
        mac_first(acc[(u + 1) % L], a[0], bn, hA);
        if (!M_UPFRONT && e_in && !last) acc[(u + 1) % L] -= (uint64_t)e_in[(row + 1) * ST];
        q = (acc[(u + 1) % L] * n0inv) & M52;
      } else if (!last) {
        bn = b[(row + 1) * ST];
        mac_first(acc[(u + 1) % L], a[0], bn, hA);
        if (!M_UPFRONT && e_in) acc[(u + 1) % L] -= (uint64_t)e_in[(row + 1) * ST];
        q = (acc[(u + 1) % L] * n0inv) & M52;
      }
      mac_span<L, 2, L, const double*, G>(acc, n, qd, hN, u);
      acc[u % L] += topA + hN + INIT;
      if (GHOST || !last) {
        mac_span<L, 1, L, double[L], G>(acc, a, bn, hA, u + 1);
        topA = hA;
        qd = limb_of(q);
      }
    }
    if (U != L) {   // rotate back: column c returns to acc[c]
      uint64_t t[L];
#pragma unroll
      for (int j = 0; j < L; ++j) t[j] = acc[(j + U) % L];
#pragma unroll
      for (int j = 0; j < L; ++j) acc[j] = t[j];
    }
  }
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] += final_bias(j, GHOST);
  // (This 3 L-deep chain is hidden by the other warp of the scheduler: a one-step parallel carry -- r[j] = low52(acc[j])
  // + (acc[j-1] >> 52), sequential fallback only when some r[j] leaves [0, 2^52), probability ~2^-39 per limb -- is
  // bit-exact and SLOWER, 116.7-117.2 against 115.6 ms for k_dec_pair<20> at 100 000: the kernel is dispatch-bound and
  // the detection costs 10 more instructions per pass.  r02, gpurun_out/r02_carry*.json; not kept.)
  int64_t c = 0;                                            // signed ripple: single columns may be negative, the
#pragma unroll
  for (int j = 0; j < L; ++j) {                             // value is in [0, 2x): no carry out of the top
    const int64_t v = (int64_t)acc[j] + c;
    r_out[j * ST] = limb_of((uint64_t)v & M52);
    c = v >> LW;
  }
}

// ------------------------------------------------------------------------------------------------
// Exact helpers on integer limbs (used once per result, not in the exponentiation loop)
// ------------------------------------------------------------------------------------------------
template <int L> PHE_HD void ints_of(uint64_t (&x)[L], const double (&d)[L]) {
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = int_of(d[j]);
}
template <int L> PHE_HD void limbs_of(double (&d)[L], const uint64_t (&x)[L]) {
#pragma unroll
  for (int j = 0; j < L; ++j) d[j] = limb_of(x[j]);
}

// x (exact) -= y (exact) across the group; returns 1 in every lane if the result went negative
// (then x holds the result mod 2^(52K)).
template <int L, int TPI, class Env> PHE_HD uint32_t sub_exact(uint64_t (&x)[L], const uint64_t (&y)[L]) {
  const int lane = Env::lane();
  uint32_t bw = 0;
#pragma unroll
  for (int j = 0; j < L; ++j) { const uint64_t v = x[j] - y[j] - bw; x[j] = v & M52; bw = (uint32_t)(v >> 63); }
  uint32_t any = bw;   // a lane borrows out at most once over all rounds; the top lane's is the sign
#pragma unroll 1
  for (int s = 1; s < TPI; ++s) {
    uint32_t b2 = Env::from_below(bw);
    if (lane == 0) b2 = 0;
    bw = 0;
    if (Env::any(b2 != 0)) {
#pragma unroll
      for (int j = 0; j < L; ++j) { const uint64_t v = x[j] - b2; x[j] = v & M52; b2 = (uint32_t)(v >> 63); }
      bw = b2;
    }
    any |= bw;
  }
  return Env::bcast(any, TPI - 1);
}

// x (exact) += y (exact) across the group (carry out of the top is dropped).
template <int L, int TPI, class Env> PHE_HD void add_exact(uint64_t (&x)[L], const uint64_t (&y)[L]) {
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] += y[j];
  normalize_exact<L, TPI, Env>(x);
}

// if x >= n: x -= n   (x exact)
template <int L, int TPI, class Env> PHE_HD void cond_sub(uint64_t (&x)[L], const uint64_t (&n)[L]) {
  uint64_t d[L];
#pragma unroll
  for (int j = 0; j < L; ++j) d[j] = x[j];
  const uint32_t neg = sub_exact<L, TPI, Env>(d, n);
  if (!neg) {
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = d[j];
  }
}

// this lane's integer limbs of a padded entry
template <int L, int TPI, class Env> PHE_HD void ints_from_entry(uint64_t (&x)[L], const double* entry) {
  const double* s = entry + Env::lane() * Pad<L>::LP;
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = int_of(s[j]);
}

// montmul result (< 2n, exact limbs) -> canonical integer limbs in [0, n)
template <int L, int TPI, class Env> PHE_HD void canonical_ints(uint64_t (&x)[L], const double (&v)[L], const double* n_entry) {
  uint64_t ni[L];
  ints_of<L>(x, v);
  ints_from_entry<L, TPI, Env>(ni, n_entry);
  cond_sub<L, TPI, Env>(x, ni);
}

// ------------------------------------------------------------------------------------------------
// Layout conversion
// ------------------------------------------------------------------------------------------------

// Integer limbs [lane*L, lane*L+L) of the little-endian u32 word array w[0..nwords) (words beyond are zero).
template <int L, int TPI, class Env>
PHE_HD void ints_from_words(uint64_t (&x)[L], const uint32_t* w, int nwords) {
  const int lane = Env::lane();
#pragma unroll
  for (int j = 0; j < L; ++j) {
    const int bit = (lane * L + j) * LW;
    const int wi = bit >> 5;
    const uint32_t sh = (uint32_t)(bit & 31);
    const uint64_t w0 = (wi < nwords) ? w[wi] : 0u;
    const uint64_t w1 = (wi + 1 < nwords) ? w[wi + 1] : 0u;
    const uint64_t w2 = (wi + 2 < nwords) ? w[wi + 2] : 0u;
    uint64_t v = (w0 | (w1 << 32)) >> sh;
    if (sh) v |= w2 << (64 - sh);
    x[j] = v & M52;
  }
}
template <int L, int TPI, class Env>
PHE_HD void limbs_from_words(double (&x)[L], const uint32_t* w, int nwords) {
  uint64_t t[L];
  ints_from_words<L, TPI, Env>(t, w, nwords);
  limbs_of<L>(x, t);
}

// Lane-block padded limb array ([TPI][LP] doubles, shared or global memory) <- registers.  Caller syncs.
template <int L, int TPI, class Env> PHE_HD void limbs_to_mem(double* dst, const double (&x)[L]) {
  constexpr int LP = Pad<L>::LP;
  double* d = dst + Env::lane() * LP;
#pragma unroll
  for (int j = 0; j < L; ++j) d[j] = x[j];
#pragma unroll
  for (int j = L; j < LP; ++j) d[j] = 0.0;
}
template <int L, int TPI, class Env> PHE_HD void limbs_from_mem(double (&x)[L], const double* src) {
  constexpr int LP = Pad<L>::LP;
  const double* s = src + Env::lane() * LP;
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = s[j];
}
template <int L, int TPI, class Env> PHE_HD void ints_to_mem(uint64_t* dst, const uint64_t (&x)[L]) {
  constexpr int LP = Pad<L>::LP;
  uint64_t* d = dst + Env::lane() * LP;
#pragma unroll
  for (int j = 0; j < L; ++j) d[j] = x[j];
#pragma unroll
  for (int j = L; j < LP; ++j) d[j] = 0ull;
}

// Word v of the number whose exact integer limbs sit in the padded array.
template <int L, int TPI> PHE_HD uint32_t word_from_ints(const uint64_t* limbs, int v) {
  constexpr int LP = Pad<L>::LP;
  constexpr int K = L * TPI;
  const int bit = v * 32;
  const int g = bit / LW;
  const int o = bit - g * LW;
  uint64_t u = 0;
  if (g < K) u = limbs[(g / L) * LP + (g % L)] >> o;
  if (o > LW - 32 && g + 1 < K) u |= limbs[((g + 1) / L) * LP + ((g + 1) % L)] << (LW - o);
  return (uint32_t)u;
}

}  // namespace phe
