// paillier_items.cuh -- per-item (one bignum per lane group) bodies of the Paillier hot path on the radix-2^52
// FP64 Montgomery engine (mont52.cuh).  Each body is a template over Env so that the CUDA kernels
// (phe_kernels.cuh) and the host lockstep emulator (tests/emu) run the same code.
//
// Constant blocks ("entries") are doubles in the padded limb layout [TPI][LP] (KP doubles each), every limb an
// exact integer < 2^52.  Every body funnels its Montgomery products through ONE montmul call site: the unrolled
// product is ~22 KB of code and the L1.5 instruction cache holds 32 KB.
#pragma once
#include "mont52.cuh"

namespace phe {

// Shared-memory regions private to one lane group.
struct GroupSmem {
  double* b0;   // KP doubles: multiplier operand buffer
  double* b1;   // KP doubles: second operand buffer / integer limb staging for stores
};

struct alignas(16) D2 { double x, y; };

// entry (padded limb layout, global or shared) -> this lane's registers
template <int L, int TPI, class Env> PHE_HD void load_entry(double (&x)[L], const double* e) {
  limbs_from_mem<L, TPI, Env>(x, e);
}

// cooperative copy of one padded entry (KP doubles) into a group buffer; each lane copies its own block
template <int L, int TPI, class Env> PHE_HD void copy_entry(double* dst, const double* src) {
  constexpr int LP = Pad<L>::LP;
  const int lane = Env::lane();
  const D2* s = reinterpret_cast<const D2*>(src + lane * LP);
  D2* d = reinterpret_cast<D2*>(dst + lane * LP);
#pragma unroll
  for (int j = 0; j < LP / 2; ++j) d[j] = s[j];
}

// canonical integer limbs (registers) -> little-endian u32 words in global memory
template <int L, int TPI, class Env>
PHE_HD void store_words(uint32_t* out, int nwords, const uint64_t (&x)[L], double* stage) {
  uint64_t* st = reinterpret_cast<uint64_t*>(stage);
  Env::sync();
  ints_to_mem<L, TPI, Env>(st, x);
  Env::sync();
  if (out) {   // null: this group is a padding duplicate of the last item (block-uniform loops), skip the store
    for (int v = Env::lane(); v < nwords; v += TPI) out[v] = word_from_ints<L, TPI>(st, v);
  }
  Env::sync();
}

// bits [pos, pos+width) of a little-endian word array (width <= 16), zero beyond nwords
PHE_HD uint32_t get_bits(const uint32_t* w, int nwords, int pos, int width) {
  const int wi = pos >> 5;
  const uint32_t lo = (wi < nwords) ? w[wi] : 0u;
  const uint32_t hi = (wi + 1 < nwords) ? w[wi + 1] : 0u;
  const uint32_t sh = (uint32_t)(pos & 31);
  const uint64_t two = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(two >> sh) & ((1u << width) - 1u);
}

// x (montmul result, < 2N) += 1, exact limbs again: still a valid montmul operand (no reduction needed here)
template <int L, int TPI, class Env> PHE_HD void plus_one_exact(double (&x)[L]) {
  uint64_t xi[L];
  ints_of<L>(xi, x);
  if (Env::lane() == 0) xi[0] += 1ull;
  normalize_exact<L, TPI, Env>(xi);
  limbs_of<L>(x, xi);
}

// ------------------------------------------------------------------------------------------------
// HE add: out = a * b mod N  (ipcl::CipherText::operator+ -> raw_add; ipcl_bindings_classes.cpp:318-321)
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_modmul(const uint32_t* a_w, const uint32_t* b_w, uint32_t* out_w, int nwords, const double* n_entry,
                        uint64_t n0inv, const double* r2, GroupSmem sm) {
  double x[L];
  limbs_from_words<L, TPI, Env>(x, b_w, nwords);
  Env::sync();
  limbs_to_mem<L, TPI, Env>(sm.b0, x);
  limbs_from_words<L, TPI, Env>(x, a_w, nwords);
  Env::sync();
  const double* bp = sm.b0;
#pragma unroll 1
  for (int step = 0; step < 2; ++step) {   // one montmul call site: a*b*R^-1, then *R^2*R^-1
    montmul<L, TPI, Env>(x, x, bp, n_entry, n0inv);
    bp = r2;
  }
  uint64_t xi[L];
  canonical_ints<L, TPI, Env>(xi, x, n_entry);
  store_words<L, TPI, Env>(out_w, nwords, xi, sm.b1);
}

// ------------------------------------------------------------------------------------------------
// One Montgomery product per element: out = a * b * R^-1 mod N (canonical words).  b is either words (converted here) or
// a constant entry.  Three uses, all of them HE adds that cost ONE product instead of item_modmul's two:
//   * broadcast add (other.size == 1): b = the single ciphertext times R, brought there once (b_entry = R^2 first);
//   * the add tree of sum / dot / matmul (ipcl_python.py:810-827 rotate-and-add): the R^-1 of every level is left in
//     and the root is multiplied by R^width once (a product of w numbers is w - 1 + 1 Montgomery products);
//   * an element that only passes a tree level: b = R mod N (the product by the Montgomery one is the identity).
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_modmul1(const uint32_t* a_w, const uint32_t* b_w, const double* b_entry, uint32_t* out_w, int nwords,
                         const double* n_entry, uint64_t n0inv, GroupSmem sm) {
  double x[L];
  // b_w may differ between the lane groups of a warp (a tree level mixes products and pass-through elements): only the
  // loads are conditional, every group goes through the same synchronisation points
  if (b_w) limbs_from_words<L, TPI, Env>(x, b_w, nwords);
  else load_entry<L, TPI, Env>(x, b_entry);
  Env::sync();
  limbs_to_mem<L, TPI, Env>(sm.b0, x);
  limbs_from_words<L, TPI, Env>(x, a_w, nwords);
  Env::sync();
  montmul<L, TPI, Env>(x, x, sm.b0, n_entry, n0inv);
  uint64_t xi[L];
  canonical_ints<L, TPI, Env>(xi, x, n_entry);
  store_words<L, TPI, Env>(out_w, nwords, xi, sm.b1);
}

// ------------------------------------------------------------------------------------------------
// Batched modular inverse (Montgomery's trick), used for the negative-plaintext rule of HE-mul: the reference inverts
// the ciphertext with gmpy2.invert element by element on the host (ipcl_python.py:272-276, 426-441, 470-479).
// A lane group owns one block of consecutive elements c_first .. c_(first+cnt-1):
//   item_inv_prefix:  P_j = c_first * ... * c_(first+j) (Montgomery form, to global scratch), block total (canonical)
//   item_inv_unwind:  given the inverse of the block total, walks back: c_j^-1 = Inv * P_(j-1), Inv *= c_j
// The block totals are inverted by the same two kernels one level up, until a handful remain for the host's extended
// Euclid.  6 Montgomery products per element instead of one 4096-bit inversion (1.9 ms in Python, ~60 us in GMP).
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_inv_prefix(const uint32_t* c_w, int nwords, int cnt, double* P, uint32_t* total_w,
                            const double* n_entry, uint64_t n0inv, const double* r2, const double* oneM,
                            const double* one_plain, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  double v[L];
  const double* bp = r2;
#pragma unroll 1
  for (int step = 0; step <= 2 * cnt; ++step) {
    const int j = step >> 1;
    if (step == 2 * cnt) {                 // leave the Montgomery domain with the block total
      load_entry<L, TPI, Env>(v, P + (size_t)(cnt - 1) * KP);
      bp = one_plain;
    } else if ((step & 1) == 0) {          // c_j -> Montgomery form
      limbs_from_words<L, TPI, Env>(v, c_w + (size_t)j * nwords, nwords);
      bp = r2;
    } else {                               // P_j = P_(j-1) * c_j
      load_entry<L, TPI, Env>(v, j == 0 ? oneM : P + (size_t)(j - 1) * KP);
      bp = sm.b0;
    }
    montmul<L, TPI, Env>(v, v, bp, n_entry, n0inv);
    if (step == 2 * cnt) break;
    if ((step & 1) == 0) {
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.b0, v);
      Env::sync();
    } else {
      limbs_to_mem<L, TPI, Env>(P + (size_t)j * KP, v);
    }
  }
  uint64_t xi[L];
  canonical_ints<L, TPI, Env>(xi, v, n_entry);
  store_words<L, TPI, Env>(total_w, nwords, xi, sm.b1);
}

template <int L, int TPI, class Env>
PHE_HD void item_inv_unwind(const uint32_t* c_w, int nwords, int cnt, const double* P, const uint32_t* tinv_w,
                            uint32_t* out_w, const double* n_entry, uint64_t n0inv, const double* r2,
                            const double* oneM, const double* one_plain, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  double v[L];
  uint64_t xi[L];
  // steps: -1: Inv = tinv * R;  then for j = cnt-1 .. 0:  0: t = P_(j-1) * Inv;  1: out_j = t * 1;
  //                                                        2: c_j * R^2;         3: Inv = c_jM * Inv
  int j = cnt - 1, phase = -1;
  const double* bp = r2;
  limbs_from_words<L, TPI, Env>(v, tinv_w, nwords);
#pragma unroll 1
  for (;;) {
    montmul<L, TPI, Env>(v, v, bp, n_entry, n0inv);
    if (phase == -1 || phase == 3) {       // v is the running inverse (Montgomery form): park it in b1
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.b1, v);
      Env::sync();
      if (phase == 3) --j;
      if (j < 0) break;
      load_entry<L, TPI, Env>(v, j == 0 ? oneM : P + (size_t)(j - 1) * KP);
      bp = sm.b1;
      phase = 0;
    } else if (phase == 0) {
      bp = one_plain;
      phase = 1;
    } else if (phase == 1) {
      canonical_ints<L, TPI, Env>(xi, v, n_entry);
      store_words<L, TPI, Env>(out_w ? out_w + (size_t)j * nwords : nullptr, nwords, xi, sm.b0);
      if (j == 0) break;                    // the running inverse is not needed any more
      limbs_from_words<L, TPI, Env>(v, c_w + (size_t)j * nwords, nwords);
      bp = r2;
      phase = 2;
    } else {                                // phase 2: v = c_j in Montgomery form
      bp = sm.b1;
      phase = 3;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fixed-window exponentiation, one montmul call site.
//   base: either words (to-Montgomery conversion done here) or a Montgomery-form entry (base_mont).
//   exponent: little-endian words, ebits significant bits (uniform across the launch).
//   table: (1<<WIN) entries of KP doubles in global memory, private to this group.
//   result: canonical words (out_w) after leaving the Montgomery domain.
// Replaces ipcl::modExp element: a^e mod N.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env, int WIN>
PHE_HD void item_powm(const uint32_t* base_w, int base_words, const double* base_mont, const uint32_t* e_w,
                      int e_words, int ebits, uint32_t* out_w, int out_words, const double* n_entry, uint64_t n0inv,
                      const double* r2, const double* oneM, const double* one_plain, double* tbl, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  constexpr int TS = 1 << WIN;
  const int nd = (ebits + WIN - 1) / WIN;   // number of digits (>= 1)
  double x[L];

  enum { P_TOMONT = 0, P_TABLE = 1, P_SQR = 2, P_MUL = 3, P_FROMMONT = 4 };
  int phase, ti = 2, sq = 0, w = nd - 2;

  if (base_mont) {
    load_entry<L, TPI, Env>(x, base_mont);
    phase = -1;
  } else {
    limbs_from_words<L, TPI, Env>(x, base_w, base_words);
    phase = P_TOMONT;
  }

  const double* bp = r2;
#pragma unroll 1
  for (;;) {
    if (phase == -1) {
      // x holds xM: start the table
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.b0, x);                  // b0 = xM for the whole table build
      limbs_to_mem<L, TPI, Env>(tbl + 1 * KP, x);           // T[1]
      copy_entry<L, TPI, Env>(tbl, oneM);                   // T[0] = R mod n
      Env::sync();
      if (TS > 2) { phase = P_TABLE; bp = sm.b0; }
      else phase = -2;
    }
    if (phase == -2) {
      // table complete: load the top digit
      const uint32_t d = get_bits(e_w, e_words, (nd - 1) * WIN, WIN);
      load_entry<L, TPI, Env>(x, tbl + d * KP);
      if (w < 0) { phase = P_FROMMONT; bp = one_plain; }
      else { phase = P_SQR; sq = 0; }
    }
    if (phase == P_SQR) {
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.b1, x);
      Env::sync();
      bp = sm.b1;
    } else if (phase == P_MUL) {
      const uint32_t d = get_bits(e_w, e_words, w * WIN, WIN);
      Env::sync();
      copy_entry<L, TPI, Env>(sm.b1, tbl + d * KP);
      Env::sync();
      bp = sm.b1;
    }

    montmul<L, TPI, Env>(x, x, bp, n_entry, n0inv);

    if (phase == P_TOMONT) {
      phase = -1;
    } else if (phase == P_TABLE) {
      limbs_to_mem<L, TPI, Env>(tbl + ti * KP, x);
      if (++ti == TS) phase = -2;
    } else if (phase == P_SQR) {
      if (++sq == WIN) phase = P_MUL;
    } else if (phase == P_MUL) {
      if (--w < 0) { phase = P_FROMMONT; bp = one_plain; }
      else { phase = P_SQR; sq = 0; }
    } else {  // P_FROMMONT
      break;
    }
  }
  uint64_t xi[L];
  canonical_ints<L, TPI, Env>(xi, x, n_entry);
  store_words<L, TPI, Env>(out_w, out_words, xi, sm.b1);
}

// ------------------------------------------------------------------------------------------------
// Sliding-window exponentiation for an exponent SHARED by every item of the launch (decrypt: p-1 / q-1; classic
// obfuscator: n): the host turns the exponent into a program (phe_api.cu: build_powm_program)
//     prog[0]            table index of the leading window (PROG_ONE: exponent is zero)
//     prog[1..nprog]     (squarings << 8) | table index, PROG_NOMUL as index = trailing squarings only
// over the table of odd powers T[k] = x^(2k+1), k < TS = 2^(WS-1).  For a 1024-bit exponent and WS = 6 that is
// 1 + 32 + 1023 + ~146 products instead of 30 + 1024 + 205 with fixed 5-bit windows.  One montmul call site.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t PROG_NOMUL = 0xffu;
constexpr uint32_t PROG_ONE = 0xffffu;

template <int L, int TPI, class Env, int WS>
PHE_HD void item_powm_prog(const uint32_t* base_w, int base_words, const double* base_mont, const uint32_t* prog,
                           int nprog, uint32_t* out_w, int out_words, const double* n_entry, uint64_t n0inv,
                           const double* r2, const double* oneM, const double* one_plain, double* tbl, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  constexpr int TS = 1 << (WS - 1);
  double x[L];
  enum { P_TOMONT = 0, P_X2 = 1, P_TABLE = 2, P_SQR = 3, P_MUL = 4, P_FROMMONT = 5 };
  int phase, ti = 1, sq = 0, pc = 1;
  uint32_t op = 0;

  if (base_mont) {
    load_entry<L, TPI, Env>(x, base_mont);
    phase = -1;
  } else {
    limbs_from_words<L, TPI, Env>(x, base_w, base_words);
    phase = P_TOMONT;
  }
  const double* bp = r2;
#pragma unroll 1
  for (;;) {
    if (phase == -1) {            // x holds xM: T[0] = xM, then x^2
      Env::sync();
      limbs_to_mem<L, TPI, Env>(tbl, x);
      limbs_to_mem<L, TPI, Env>(sm.b0, x);
      Env::sync();
      if (TS > 1 && prog[0] != PROG_ONE) { phase = P_X2; bp = sm.b0; }
      else phase = -2;
    }
    if (phase == -2) {            // table complete: load the leading window
      const uint32_t i0 = prog[0];
      load_entry<L, TPI, Env>(x, i0 == PROG_ONE ? oneM : tbl + (size_t)i0 * KP);
      pc = 1;
      phase = -3;
    }
    if (phase == -3) {            // fetch the next program entry
      if (pc > nprog) { phase = P_FROMMONT; bp = one_plain; }
      else {
        op = prog[pc++];
        sq = (int)(op >> 8);
        phase = sq > 0 ? P_SQR : P_MUL;
      }
    }
    if (phase == P_SQR) {
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.b1, x);
      Env::sync();
      bp = sm.b1;
    } else if (phase == P_MUL) {
      Env::sync();
      copy_entry<L, TPI, Env>(sm.b1, tbl + (size_t)(op & 0xffu) * KP);
      Env::sync();
      bp = sm.b1;
    }

    montmul<L, TPI, Env>(x, x, bp, n_entry, n0inv);

    if (phase == P_TOMONT) {
      phase = -1;
    } else if (phase == P_X2) {   // x = xM^2: the table multiplier; restart the chain from T[0]
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.b0, x);
      Env::sync();
      load_entry<L, TPI, Env>(x, tbl);
      bp = sm.b0;
      phase = P_TABLE;
    } else if (phase == P_TABLE) {
      limbs_to_mem<L, TPI, Env>(tbl + (size_t)ti * KP, x);
      if (++ti == TS) phase = -2;
    } else if (phase == P_SQR) {
      if (--sq == 0) phase = ((op & 0xffu) == PROG_NOMUL) ? -3 : P_MUL;
    } else if (phase == P_MUL) {
      phase = -3;
    } else {  // P_FROMMONT
      break;
    }
  }
  uint64_t xi[L];
  canonical_ints<L, TPI, Env>(xi, x, n_entry);
  store_words<L, TPI, Env>(out_w, out_words, xi, sm.b1);
}

// ------------------------------------------------------------------------------------------------
// Decrypt pre-reduction: xM = c * R mod x^2 from the double-width ciphertext c = c_lo + 2^(32*hw) c_hi:
//   xM = montmul(c_lo, R^2) + montmul(c_hi, 2^(32 hw) R^2)   (both constants mod x^2).
// Output: Montgomery-form entry (padded exact limbs, value < 4 x^2) in global memory.
// Part of ipcl::PrivateKey::decrypt -> decryptCRT (ipcl_bindings_classes.cpp:127-133): "c mod p^2".
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_dec_prep(const uint32_t* c_w, int hw, double* out_entry, const double* n_entry, uint64_t n0inv,
                          const double* r2, const double* k2, GroupSmem sm) {
  double x[L];
  uint64_t acc[L];
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] = 0;
  const double* bp = r2;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    limbs_from_words<L, TPI, Env>(x, c_w + half * hw, hw);
    montmul<L, TPI, Env>(x, x, bp, n_entry, n0inv);
#pragma unroll
    for (int j = 0; j < L; ++j) acc[j] += int_of(x[j]);
    bp = k2;
  }
  (void)sm;
  normalize_exact<L, TPI, Env>(acc);   // value < 4 x^2 < R: fits
  limbs_of<L>(x, acc);
  limbs_to_mem<L, TPI, Env>(out_entry, x);
}

// ------------------------------------------------------------------------------------------------
// DJN encrypt with a fixed-base comb table (no squarings):
//   obf = prod_j T[j][digit_j(r)],  T[j][d] = hs^(d * 2^(wb j)) * R mod n^2,  digits of wb bits (wb <= 16)
//   ct  = (1 + m n) * obf mod n^2
// ipcl::PublicKey::encrypt + applyObfuscator (DJN) (ipcl_bindings_classes.cpp:53-60, 71-83).
// r_w == nullptr -> make_secure = false (ct = 1 + m n).
// The table lives in HBM (2.7 GB at wb = 16 for a 2048-bit key: 64 windows x 65536 entries x 640 B); the entry of
// window j+1 is fetched with cp.async into the idle operand buffer while window j is being multiplied.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env> PHE_HD void copy_entry_async(double* dst, const double* src) {
  constexpr int LP = Pad<L>::LP;
  const int lane = Env::lane();
#pragma unroll
  for (int j = 0; j < LP / 2; ++j) Env::cp_async16(dst + lane * LP + 2 * j, src + lane * LP + 2 * j);
}

template <int L, int TPI, class Env>
PHE_HD void item_encrypt_comb(const uint32_t* m_w, int m_words, const uint32_t* r_w, int r_words, int nwin, int wb,
                              uint32_t* out_w, int out_words, const double* n_entry, uint64_t n0inv,
                              const double* nR, const double* comb, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  double x[L];
  int j = 1;
  // step kinds: 0 comb multiply, 1 raw (m * nR), 2 final (raw * obf)
  int kind;
  const double* bp;
  double* cur = sm.b0;      // buffer the pending prefetch was issued into
  double* other = sm.b1;
  if (r_w) {
    const uint32_t d0 = get_bits(r_w, r_words, 0, wb);
    load_entry<L, TPI, Env>(x, comb + (size_t)d0 * KP);
    kind = (nwin > 1) ? 0 : 1;
    if (nwin > 1) {
      const uint32_t d1 = get_bits(r_w, r_words, wb, wb);
      Env::sync();
      copy_entry_async<L, TPI, Env>(cur, comb + (((size_t)1 << wb) + d1) * KP);
    }
  } else {
    kind = 1;
  }
#pragma unroll 1
  for (;;) {
    if (kind == 0) {
      Env::cp_async_wait();
      Env::sync();                       // entry j is in `cur`; every lane is done reading `other`
      if (j + 1 < nwin) {
        const uint32_t d = get_bits(r_w, r_words, (j + 1) * wb, wb);
        copy_entry_async<L, TPI, Env>(other, comb + ((((size_t)j + 1) << wb) + d) * KP);
      }
      bp = cur;
      double* t = cur; cur = other; other = t;
    } else if (kind == 1) {
      if (r_w) {   // park obf*R in b1
        Env::sync();
        limbs_to_mem<L, TPI, Env>(sm.b1, x);
        Env::sync();
      }
      limbs_from_words<L, TPI, Env>(x, m_w, m_words);
      bp = nR;
    } else {
      bp = sm.b1;
    }

    montmul<L, TPI, Env>(x, x, bp, n_entry, n0inv);

    if (kind == 0) {
      if (++j == nwin) kind = 1;
    } else if (kind == 1) {
      // x == m*n (mod n^2), < 2 n^2
      if (!r_w) break;
      plus_one_exact<L, TPI, Env>(x);   // == 1 + m n (mod n^2), < 2 n^2 + 1: fine as the next multiplicand
      kind = 2;
    } else {
      break;
    }
  }
  uint64_t xi[L];
  canonical_ints<L, TPI, Env>(xi, x, n_entry);
  if (!r_w) {   // make_secure = false: ct = 1 + m n, and m n mod n^2 <= n^2 - n
    if (Env::lane() == 0) xi[0] += 1ull;
    normalize_exact<L, TPI, Env>(xi);
  }
  store_words<L, TPI, Env>(out_w, out_words, xi, sm.b0);
}

// ------------------------------------------------------------------------------------------------
// ct = (1 + m n) * obf mod n^2 with obf given as canonical words (classic r^n path, or apply_obfuscator
// on an existing ciphertext).
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_encrypt_finish(const uint32_t* m_w, int m_words, const uint32_t* obf_w, uint32_t* out_w,
                                int out_words, const double* n_entry, uint64_t n0inv, const double* nR,
                                const double* r2, GroupSmem sm) {
  double x[L];
  // b1 = obf * R
  limbs_from_words<L, TPI, Env>(x, obf_w, out_words);
  const double* bp = r2;
  int kind = 0;
#pragma unroll 1
  for (;;) {
    montmul<L, TPI, Env>(x, x, bp, n_entry, n0inv);
    if (kind == 0) {
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.b1, x);
      Env::sync();
      limbs_from_words<L, TPI, Env>(x, m_w, m_words);
      bp = nR;
      kind = 1;
    } else if (kind == 1) {
      plus_one_exact<L, TPI, Env>(x);
      bp = sm.b1;
      kind = 2;
    } else {
      break;
    }
  }
  uint64_t xi[L];
  canonical_ints<L, TPI, Env>(xi, x, n_entry);
  store_words<L, TPI, Env>(out_w, out_words, xi, sm.b0);
}

// ------------------------------------------------------------------------------------------------
// Decrypt tail (per ciphertext), everything in the x^2-sized shape:
//   for x in {p, q}:  u_x = c^(x-1) mod x^2 (canonical words, from item_powm)
//       Lx  = (u_x - 1) / x           exact division = -(Montgomery quotient of (u_x - 1) w.r.t. x) mod R
//       m_x = Lx * h_x mod x
//   m = m_p + ((m_q - m_p) * p^-1 mod q) * p
// ipcl::PrivateKey::decryptCRT: computeLfun, *hp/hq, computeCRT.
// Constant entries (cst, KP doubles each): see DecTailConst.  Six Montgomery products through one call site:
//   step 0/2: quotient capture for p / q      step 1/3: * h_p R / * h_q R
//   step 4:   (m_q - m_p) * p^-1 R mod q      step 5:   h * p R mod n
// ------------------------------------------------------------------------------------------------
enum DecTailConst {
  DT_P = 0,       // p padded
  DT_Q,           // q padded
  DT_N,           // n = p q
  DT_HPM,         // hp * R mod p
  DT_HQM,         // hq * R mod q
  DT_PINVM,       // (p^-1 mod q) * R mod q
  DT_PMN,         // p * R mod n
  DT_ONE,         // plain integer 1
  DT_COUNT
};

template <int L, int TPI, class Env>
PHE_HD void item_dec_tail(const uint32_t* up_w, const uint32_t* uq_w, int u_words, uint32_t* m_w, int m_words,
                          const double* cst, const uint64_t* n0invs /* [p, q, n] */, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  constexpr int LP = Pad<L>::LP;
  const int lane = Env::lane();
  double x[L];
  uint64_t xi[L], mp[L], one[L];
  uint64_t* qbuf = reinterpret_cast<uint64_t*>(sm.b1);
  uint32_t u_zero = 0;
  ints_from_entry<L, TPI, Env>(one, cst + DT_ONE * KP);
#pragma unroll
  for (int j = 0; j < L; ++j) mp[j] = 0;

#pragma unroll 1
  for (int step = 0; step < 6; ++step) {
    const int idx = (step >> 1) & 1;                       // 0: p, 1: q   (steps 0..3)
    const double* mod_e = cst + (step < 2 ? DT_P : step < 5 ? DT_Q : DT_N) * KP;
    const uint64_t n0 = n0invs[step < 2 ? 0 : step < 5 ? 1 : 2];
    const double* bp;
    if (step == 0 || step == 2) {
      ints_from_words<L, TPI, Env>(xi, idx ? uq_w : up_w, u_words);
      u_zero = sub_exact<L, TPI, Env>(xi, one);             // u - 1; u = 0 only for a non-unit c: then L = floor(-1/x) = -1
      limbs_of<L>(x, xi);
      bp = cst + DT_ONE * KP;
      Env::sync();
    } else if (step == 1 || step == 3) {
      bp = cst + (idx ? DT_HQM : DT_HPM) * KP;
    } else if (step == 4) {
      bp = cst + DT_PINVM * KP;
    } else {
      bp = cst + DT_PMN * KP;
    }

    if (step == 0 || step == 2) montmul<L, TPI, Env, true>(x, x, bp, mod_e, n0, qbuf);   // q = -(u-1) x^-1 mod R
    else montmul<L, TPI, Env>(x, x, bp, mod_e, n0);

    if (step == 0 || step == 2) {
      // Lx = -q mod R = (~q) + 1 over K limbs
      Env::sync();
#pragma unroll
      for (int j = 0; j < L; ++j) xi[j] = (~qbuf[lane * LP + j]) & M52;
      if (lane == 0) xi[0] += 1ull;
      normalize_exact<L, TPI, Env>(xi);
      {   // u = 0: L = -1 = x - 1 (mod x), as the oracle's floor division.  Branch-free: the shuffles inside sub_exact
          // must stay warp-convergent and u_zero is only uniform within a lane group.
        uint64_t xm1[L];
        ints_from_entry<L, TPI, Env>(xm1, mod_e);
        sub_exact<L, TPI, Env>(xm1, one);
#pragma unroll
        for (int j = 0; j < L; ++j) xi[j] = u_zero ? xm1[j] : xi[j];
      }
      limbs_of<L>(x, xi);
    } else if (step == 1) {
      canonical_ints<L, TPI, Env>(mp, x, mod_e);            // m_p in [0, p)
    } else if (step == 3) {
      canonical_ints<L, TPI, Env>(xi, x, mod_e);            // m_q in [0, q)
      const uint32_t neg = sub_exact<L, TPI, Env>(xi, mp);  // m_q - m_p
      uint64_t addend[L], qi[L];
      ints_from_entry<L, TPI, Env>(qi, mod_e);
#pragma unroll
      for (int j = 0; j < L; ++j) addend[j] = neg ? qi[j] : 0ull;
      add_exact<L, TPI, Env>(xi, addend);                   // + q if negative: wraps back into range
      limbs_of<L>(x, xi);
    } else if (step == 4) {
      canonical_ints<L, TPI, Env>(xi, x, mod_e);            // h in [0, q)
      limbs_of<L>(x, xi);
    } else if (step == 5) {
      canonical_ints<L, TPI, Env>(xi, x, mod_e);            // h * p (< n)
      add_exact<L, TPI, Env>(xi, mp);                       // m
    }
  }
  store_words<L, TPI, Env>(m_w, m_words, xi, sm.b0);
}

// ------------------------------------------------------------------------------------------------
// CRT recombination (ipcl computeCRT): m = m_p + ((m_q - m_p) * p^-1 mod q) * p from the two halves produced by
// item_dec_pair; same constants (DecTailConst) and shape as item_dec_tail, of which this is the last two steps.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_dec_crt(const uint32_t* mp_w, const uint32_t* mq_w, int half_words, uint32_t* m_w, int m_words,
                         const double* cst, const uint64_t* n0invs /* [p, q, n] */, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  double x[L];
  uint64_t xi[L], mp[L];
  ints_from_words<L, TPI, Env>(mp, mp_w, half_words);
  ints_from_words<L, TPI, Env>(xi, mq_w, half_words);
  {
    const uint32_t neg = sub_exact<L, TPI, Env>(xi, mp);   // m_q - m_p
    uint64_t addend[L], qi[L];
    ints_from_entry<L, TPI, Env>(qi, cst + DT_Q * KP);
#pragma unroll
    for (int j = 0; j < L; ++j) addend[j] = neg ? qi[j] : 0ull;
    add_exact<L, TPI, Env>(xi, addend);                    // + q if negative: wraps back into [0, q)
    limbs_of<L>(x, xi);
  }
#pragma unroll 1
  for (int step = 0; step < 2; ++step) {
    const double* mod_e = cst + (step == 0 ? DT_Q : DT_N) * KP;
    montmul<L, TPI, Env>(x, x, cst + (step == 0 ? DT_PINVM : DT_PMN) * KP, mod_e, n0invs[step == 0 ? 1 : 2]);
    canonical_ints<L, TPI, Env>(xi, x, mod_e);             // step 0: h in [0, q); step 1: h * p (< n)
    limbs_of<L>(x, xi);
  }
  add_exact<L, TPI, Env>(xi, mp);                          // m
  store_words<L, TPI, Env>(m_w, m_words, xi, sm.b0);
}

// ------------------------------------------------------------------------------------------------
// CRT half of decrypt on the p-adic pair engine (mont52.cuh: pair_pass), one (ciphertext, x) per lane, x = p or q:
//     m_x = L_x(c^(x-1) mod x^2) * h_x mod x          (ipcl::PrivateKey::decryptCRT)
// driven by a host-built program (phe_api.cu: build_decpair_program) that is the same for every lane:
//     load the four (bits/2)-bit chunks C_k of c, multiply each by the constant pair W_k ~ 2^(k bits/2) R^2, sum
//     -> c in pair form; table of odd powers; sliding-window squarings / multiplications (exponent x - 1);
//     multiply by (1, 0) to leave the Montgomery domain: c^(x-1) = v0 + v1 x with v0 = 1, so L_x = v1 - [v0 = 0];
//     multiply by (h_x R, 0); write the canonical result.
// State per lane: X0 in registers; XS0 (copy of X0), X1, Y0, Y1, E in shared memory (columns, stride PE::STRIDE);
// table slots in global memory (same column layout).  A product is two or three pair_pass calls through ONE call site.
// ------------------------------------------------------------------------------------------------
enum PairOp : uint32_t {
  PO_END = 0, PO_LOADC, PO_YCONST, PO_YT, PO_YX, PO_XT, PO_TX, PO_MUL, PO_SQR, PO_SUM4, PO_FINISH, PO_OUT
};
enum PairConst { PC_W0 = 0, PC_W1, PC_W2, PC_W3, PC_ONE, PC_HR, PC_COUNT };   // constant pairs, [PC_COUNT][2][L] doubles

template <class PE> struct PairSmem {   // all pointers already offset to this lane's column
  double *xs0, *x1, *y0, *y1;
  int64_t* e;                           // L + 1 entries
};

template <int L, class PE> PHE_HD void col_store(double* dst, const double (&x)[L]) {
#pragma unroll
  for (int j = 0; j < L; ++j) dst[j * PE::STRIDE] = x[j];
}
template <int L, class PE> PHE_HD void col_load(double (&x)[L], const double* src) {
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = src[j * PE::STRIDE];
}
template <int L, class PE> PHE_HD void col_copy(double* dst, const double* src) {
#pragma unroll
  for (int j = 0; j < L; ++j) dst[j * PE::STRIDE] = src[j * PE::STRIDE];
}
// global -> shared column copy that does not pass through registers; completed by PE::cp_async_wait()
template <int L, class PE> PHE_HD void col_copy_async(double* dst_shared, const double* src_global) {
#pragma unroll
  for (int j = 0; j < L; ++j) PE::cp_async8(dst_shared + j * PE::STRIDE, src_global + j * PE::STRIDE);
}

// x <- 2 x as exact limbs (x < 2^(52 L - 1))
template <int L> PHE_HD void pair_double(double (&x)[L]) {
  uint64_t t[L];
  ints_of<L>(t, x);
#pragma unroll
  for (int j = 0; j < L; ++j) t[j] <<= 1;
  ripple<L>(t, 0u);
  limbs_of<L>(x, t);
}

template <int L, class PE>
PHE_HD void item_dec_pair(const uint32_t* c_w, int chunk_words, const uint32_t* prog, uint32_t* out_w, int out_words,
                          const double* n, uint64_t n0inv, const double* cst, double* tbl,
                          PairSmem<PE> sm) {
  constexpr int ST = PE::STRIDE;
  double x[L];
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = 0.0;
  int pc = 0, sub = 0, sqleft = 0;
  bool square = false;
  bool y_pending = false;   // the table entry of the next PO_YT is already on its way into y0 / y1

#pragma unroll 1
  for (;;) {
    if (sub == 0 && sqleft == 0) {
      bool done = false;
#pragma unroll 1
      for (;;) {   // data-movement instructions up to the next product
        const uint32_t ins = prog[pc++];
        const uint32_t op = ins & 0xffu, arg = ins >> 8;
        if (op == PO_MUL) { square = false; break; }
        if (op == PO_SQR) {
          square = true; sqleft = (int)arg;
          // The program is the same for every lane and known ahead: if a multiplication by a table entry follows
          // these squarings, start fetching the entry now (asynchronously, straight into shared memory: a square
          // does not touch y0 / y1), so that its DRAM latency -- the per-lane tables do not fit the L2 -- is covered
          // by the squarings instead of stalling the multiplication (ncu r01: long_scoreboard 0.34 per issue).
          const uint32_t nxt = prog[pc];
          if ((nxt & 0xffu) == PO_YT) {
            const double* src = tbl + (size_t)(nxt >> 8) * 2 * L * ST;
            col_copy_async<L, PE>(sm.y0, src);
            col_copy_async<L, PE>(sm.y1, src + L * ST);
            y_pending = true;
          }
          break;
        }
        if (op == PO_END) { done = true; break; }
        if (op == PO_LOADC) {
          limbs_from_words<L, 1, PE>(x, c_w + arg * chunk_words, chunk_words);
          col_store<L, PE>(sm.xs0, x);
#pragma unroll
          for (int j = 0; j < L; ++j) sm.x1[j * ST] = 0.0;
        } else if (op == PO_YCONST) {
          const double* src = cst + (size_t)arg * 2 * L;
#pragma unroll
          for (int j = 0; j < L; ++j) { sm.y0[j * ST] = src[j]; sm.y1[j * ST] = src[L + j]; }
        } else if (op == PO_YT) {
          if (y_pending) {           // prefetched under the preceding squarings
            PE::cp_async_wait();
            y_pending = false;
          } else {
            const double* src = tbl + (size_t)arg * 2 * L * ST;
            col_copy<L, PE>(sm.y0, src);
            col_copy<L, PE>(sm.y1, src + L * ST);
          }
        } else if (op == PO_YX) {
          col_copy<L, PE>(sm.y0, sm.xs0);
          col_copy<L, PE>(sm.y1, sm.x1);
        } else if (op == PO_XT) {
          const double* src = tbl + (size_t)arg * 2 * L * ST;
          col_load<L, PE>(x, src);
          col_store<L, PE>(sm.xs0, x);
          col_copy<L, PE>(sm.x1, src + L * ST);
        } else if (op == PO_TX) {
          double* dst = tbl + (size_t)arg * 2 * L * ST;
          col_store<L, PE>(dst, x);
          col_copy<L, PE>(dst + L * ST, sm.x1);
        } else if (op == PO_SUM4) {   // X = slot 0 + slot 1 + slot 2 + slot 3 (each part < 2x: sums < 8x < R)
#pragma unroll 1
          for (int part = 0; part < 2; ++part) {
            uint64_t t[L];
#pragma unroll
            for (int j = 0; j < L; ++j) {
              uint64_t v = 0;
#pragma unroll
              for (int sidx = 0; sidx < 4; ++sidx) v += int_of(tbl[((size_t)(sidx * 2 + part) * L + j) * ST]);
              t[j] = v;
            }
            ripple<L>(t, 0u);
            if (part == 0) { limbs_of<L>(x, t); col_store<L, PE>(sm.xs0, x); }
            else {
#pragma unroll
              for (int j = 0; j < L; ++j) sm.x1[j * ST] = limb_of(t[j]);
            }
          }
        } else if (op == PO_FINISH) {
          // X = (w, z1) = c^(x-1) itself (out of the Montgomery domain), w < 2x, z1 < 4x.  Canonical digits v0, v1;
          // L = (c^(x-1) - 1) div x = v1 - [v0 == 0]   (v0 == 1 for every unit c; floor semantics otherwise)
          uint64_t v0[L], v1[L], xi[L], one[L];
          ints_of<L>(v0, x);
#pragma unroll
          for (int j = 0; j < L; ++j) { v1[j] = int_of(sm.x1[j * ST]); xi[j] = int_of(n[j]); one[j] = (j == 0) ? 1ull : 0ull; }
          {
            uint64_t d[L];
#pragma unroll
            for (int j = 0; j < L; ++j) d[j] = v0[j];
            if (!sub_exact<L, 1, PE>(d, xi)) {   // w >= x: v0 = w - x, carry one x into the second digit
#pragma unroll
              for (int j = 0; j < L; ++j) v0[j] = d[j];
              add_exact<L, 1, PE>(v1, one);
            }
          }
#pragma unroll 1
          for (int k = 0; k < 4; ++k) cond_sub<L, 1, PE>(v1, xi);   // second digit < 4x + 1
          uint64_t nz = 0;
#pragma unroll
          for (int j = 0; j < L; ++j) nz |= v0[j];
          if (nz == 0) {                         // v1 - 1 mod x
            if (sub_exact<L, 1, PE>(v1, one)) add_exact<L, 1, PE>(v1, xi);
          }
          limbs_of<L>(x, v1);
          col_store<L, PE>(sm.xs0, x);
#pragma unroll
          for (int j = 0; j < L; ++j) sm.x1[j * ST] = 0.0;
        } else if (op == PO_OUT) {
          uint64_t v[L], xi[L];
          ints_of<L>(v, x);
#pragma unroll
          for (int j = 0; j < L; ++j) xi[j] = int_of(n[j]);
          cond_sub<L, 1, PE>(v, xi);
          uint64_t* st = reinterpret_cast<uint64_t*>(sm.e);
#pragma unroll
          for (int j = 0; j < L; ++j) st[j * ST] = v[j];
          if (out_w) {
            for (int w = 0; w < out_words; ++w) {
              const int bit = w * 32, g = bit / LW, o = bit - g * LW;
              uint64_t u = (g < L) ? st[g * ST] >> o : 0ull;
              if (o > LW - 32 && g + 1 < L) u |= st[(g + 1) * ST] << (LW - o);
              out_w[w] = (uint32_t)u;
            }
          }
        }
      }
      if (done) break;
    }

    // a product is two (square) or three (multiplication) passes through the one pair_pass call site:
    //   sub 0: a = X0, b = X0 | Y0        -> Z0 (to XS0), E
    //   sub 1: a = 2 X0 | X0, b = X1 | Y1, E -> Z1 (square: to X1; multiplication: first cross term, parked in E)
    //   sub 2 (multiplication): a = X1, b = Y0 -> second cross term (to X1); Z1 = sum of the two (Y is left intact)
    const double* b;
    const int64_t* ein = nullptr;
    int64_t* eout = nullptr;
    double* rout;
    if (sub == 0) {
      b = square ? sm.xs0 : sm.y0;
      eout = sm.e;
      rout = sm.xs0;
    } else if (sub == 1) {
      ein = sm.e;
      if (square) {                    // a = 2 X0 (exact limbs again), b = X1
        pair_double<L>(x);
        b = sm.x1;
        rout = sm.x1;
      } else {
        b = sm.y1;
        rout = reinterpret_cast<double*>(sm.e);   // E is consumed by this very pass: park the first cross term there
      }
    } else {
      col_load<L, PE>(x, sm.x1);
      b = sm.y0;
      rout = sm.x1;
    }

    pair_pass<L, PE>(rout, x, b, ein, eout, n, n0inv);

    if (sub == 0) {
      sub = 1;
    } else if (sub == 1 && !square) {
      sub = 2;
    } else {
      if (!square) {                   // Z1 = (X0 Y1 - m) R^-1 + X1 Y0 R^-1, < 4x, exact limbs
        uint64_t t[L];
#pragma unroll
        for (int j = 0; j < L; ++j) t[j] = int_of(sm.x1[j * ST]) + int_of(reinterpret_cast<const double*>(sm.e)[j * ST]);
        ripple<L>(t, 0u);
#pragma unroll
        for (int j = 0; j < L; ++j) sm.x1[j * ST] = limb_of(t[j]);
      }
      sub = 0;
      col_load<L, PE>(x, sm.xs0);      // X0 = Z0
      if (square) --sqleft;
    }
  }
}

}  // namespace phe
