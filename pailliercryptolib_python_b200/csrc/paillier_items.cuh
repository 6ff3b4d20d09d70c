// paillier_items.cuh -- per-item (one bignum per lane group) bodies of the Paillier hot path.
// Each body is a template over Env so that the CUDA kernels (phe_kernels.cu) and the host lockstep
// emulator (tests/emu) run the same code.
//
// Constant blocks ("entries") live in the padded limb layout [TPI][LP] (KP = TPI*LP u32 words each).
#pragma once
#include "mont28.cuh"

namespace phe {

// Shared-memory regions private to one lane group.
struct GroupSmem {
  uint32_t* b0;   // KP words: multiplier operand buffer
  uint32_t* b1;   // KP words: second operand buffer / limb staging for stores
};

template <int L, int TPI> struct Shape {
  static constexpr int LP = Pad<L>::LP;
  static constexpr int K = L * TPI;
  static constexpr int KP = LP * TPI;
  static constexpr int BITS = K * LW;
};

// entry (padded limb layout, global or shared) -> this lane's registers
template <int L, int TPI, class Env> PHE_HD void load_entry(uint32_t (&x)[L], const uint32_t* e) {
  limbs_from_smem<L, TPI, Env>(x, e);
}

// cooperative copy of one padded entry (KP words) into a group smem buffer; lane copies its own block
template <int L, int TPI, class Env> PHE_HD void copy_entry(uint32_t* dst, const uint32_t* src) {
  constexpr int LP = Pad<L>::LP;
  const int lane = Env::lane();
  const U4* s = reinterpret_cast<const U4*>(src + lane * LP);
  U4* d = reinterpret_cast<U4*>(dst + lane * LP);
#pragma unroll
  for (int j = 0; j < LP / 4; ++j) d[j] = s[j];
}

// canonical exact limbs (registers) -> little-endian u32 words in global memory
template <int L, int TPI, class Env>
PHE_HD void store_words(uint32_t* out, int nwords, const uint32_t (&x)[L], uint32_t* stage) {
  Env::sync();
  limbs_to_smem<L, TPI, Env>(stage, x);
  Env::sync();
  if (out) {   // null: this group is a padding duplicate of the last item (block-uniform loops), skip the store
    for (int v = Env::lane(); v < nwords; v += TPI) out[v] = word_from_smem_limbs<L, TPI>(stage, v);
  }
  Env::sync();
}

// bits [pos, pos+width) of a little-endian word array (width <= 16), zero beyond nwords
PHE_HD uint32_t get_bits(const uint32_t* w, int nwords, int pos, int width) {
  const int wi = pos >> 5;
  const uint32_t lo = (wi < nwords) ? w[wi] : 0u;
  const uint32_t hi = (wi + 1 < nwords) ? w[wi + 1] : 0u;
  return funnel_r(lo, hi, (uint32_t)(pos & 31)) & ((1u << width) - 1u);
}

// ------------------------------------------------------------------------------------------------
// HE add: out = a * b mod N  (ipcl::CipherText::operator+ -> raw_add; ipcl_bindings_classes.cpp:318-321)
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_modmul(const uint32_t* a_w, const uint32_t* b_w, uint32_t* out_w, int nwords,
                        const uint32_t (&n)[L], uint32_t n0inv, const uint32_t* r2, GroupSmem sm) {
  uint32_t x[L];
  {
    uint32_t y[L];
    limbs_from_words<L, TPI, Env>(y, b_w, nwords);
    Env::sync();
    limbs_to_smem<L, TPI, Env>(sm.b0, y);
  }
  limbs_from_words<L, TPI, Env>(x, a_w, nwords);
  Env::sync();
  const uint32_t* bp = sm.b0;
#pragma unroll 1
  for (int step = 0; step < 2; ++step) {   // one montmul call site: a*b*R^-1, then *R^2*R^-1
    montmul<L, TPI, Env>(x, x, bp, n, n0inv);
    bp = r2;
  }
  canonicalize<L, TPI, Env>(x, n);
  store_words<L, TPI, Env>(out_w, nwords, x, sm.b1);
}

// ------------------------------------------------------------------------------------------------
// Sliding-free fixed-window exponentiation, one montmul call site.
//   base: either words (to-Montgomery conversion done here) or a Montgomery-form entry (base_mont).
//   exponent: little-endian words, ebits significant bits (uniform across the launch).
//   table: (1<<WIN) entries of KP words in global memory, private to this group.
//   result: canonical words (out_w) after leaving the Montgomery domain.
// Replaces ipcl::modExp element: a^e mod N.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env, int WIN>
PHE_HD void item_powm(const uint32_t* base_w, int base_words, const uint32_t* base_mont,
                      const uint32_t* e_w, int e_words, int ebits, uint32_t* out_w, int out_words,
                      const uint32_t (&n)[L], uint32_t n0inv, const uint32_t* r2, const uint32_t* oneM,
                      const uint32_t* one_plain, uint32_t* tbl, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  constexpr int LP = Pad<L>::LP;
  constexpr int TS = 1 << WIN;
  const int lane = Env::lane();
  const int nd = (ebits + WIN - 1) / WIN;   // number of digits (>= 1)
  uint32_t x[L];

  enum { P_TOMONT = 0, P_TABLE = 1, P_SQR = 2, P_MUL = 3, P_FROMMONT = 4 };
  int phase, ti = 2, sq = 0, w = nd - 2;

  if (base_mont) {
    load_entry<L, TPI, Env>(x, base_mont);
    phase = -1;
  } else {
    limbs_from_words<L, TPI, Env>(x, base_w, base_words);
    phase = P_TOMONT;
  }

  const uint32_t* bp = r2;
#pragma unroll 1
  for (;;) {
    if (phase == -1) {
      // x holds xM: start the table
      Env::sync();
      limbs_to_smem<L, TPI, Env>(sm.b0, x);                 // b0 = xM for the whole table build
      limbs_to_smem<L, TPI, Env>(tbl + 1 * KP, x);          // T[1]
      { // T[0] = R mod n
        const U4* s = reinterpret_cast<const U4*>(oneM + lane * LP);
        U4* d = reinterpret_cast<U4*>(tbl + lane * LP);
#pragma unroll
        for (int j = 0; j < LP / 4; ++j) d[j] = s[j];
      }
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
      limbs_to_smem<L, TPI, Env>(sm.b1, x);
      Env::sync();
      bp = sm.b1;
    } else if (phase == P_MUL) {
      const uint32_t d = get_bits(e_w, e_words, w * WIN, WIN);
      Env::sync();
      copy_entry<L, TPI, Env>(sm.b1, tbl + d * KP);
      Env::sync();
      bp = sm.b1;
    }

    montmul<L, TPI, Env>(x, x, bp, n, n0inv);

    if (phase == P_TOMONT) {
      phase = -1;
    } else if (phase == P_TABLE) {
      limbs_to_smem<L, TPI, Env>(tbl + ti * KP, x);
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
  canonicalize<L, TPI, Env>(x, n);
  store_words<L, TPI, Env>(out_w, out_words, x, sm.b1);
}

// ------------------------------------------------------------------------------------------------
// Decrypt pre-reduction: xM = c * R mod x^2 from the double-width ciphertext c = c_lo + 2^(32*hw) c_hi:
//   xM = montmul(c_lo, R^2) + montmul(c_hi, 2^(32 hw) R^2)   (both constants mod x^2).
// Output: Montgomery-form entry (padded limbs, almost normalised, value < 4 x^2) in global memory.
// Part of ipcl::PrivateKey::decrypt -> decryptCRT (ipcl_bindings_classes.cpp:127-133): "c mod p^2".
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_dec_prep(const uint32_t* c_w, int hw, uint32_t* out_entry, const uint32_t (&n)[L], uint32_t n0inv,
                          const uint32_t* r2, const uint32_t* k2, GroupSmem sm) {
  uint32_t x[L], acc[L];
#pragma unroll
  for (int j = 0; j < L; ++j) acc[j] = 0;
  const uint32_t* bp = r2;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    limbs_from_words<L, TPI, Env>(x, c_w + half * hw, hw);
    montmul<L, TPI, Env>(x, x, bp, n, n0inv);
#pragma unroll
    for (int j = 0; j < L; ++j) acc[j] += x[j];
    bp = k2;
  }
  (void)sm;
  normalize_exact<L, TPI, Env>(acc);   // value < 4 x^2 < R: fits; exact limbs keep the column bound of montmul
  limbs_to_smem<L, TPI, Env>(out_entry, acc);
}

// ------------------------------------------------------------------------------------------------
// DJN encrypt with a fixed-base comb table (no squarings):
//   obf = prod_j T[j][digit_j(r)],  T[j][d] = hs^(d * 2^(WB j)) * R mod n^2
//   ct  = (1 + m n) * obf mod n^2
// ipcl::PublicKey::encrypt + applyObfuscator (DJN) (ipcl_bindings_classes.cpp:53-60, 71-83).
// r_w == nullptr -> make_secure = false (ct = 1 + m n).
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env, int WB>
PHE_HD void item_encrypt_comb(const uint32_t* m_w, int m_words, const uint32_t* r_w, int r_words, int nwin,
                              uint32_t* out_w, int out_words, const uint32_t (&n)[L], uint32_t n0inv,
                              const uint32_t* nR, const uint32_t* comb, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  const int lane = Env::lane();
  uint32_t x[L];
  int j = 1;
  // step kinds: 0 comb multiply, 1 raw (m * nR), 2 final (raw * obf)
  int kind;
  const uint32_t* bp;
  if (r_w) {
    const uint32_t d0 = get_bits(r_w, r_words, 0, WB);
    load_entry<L, TPI, Env>(x, comb + (size_t)d0 * KP);
    kind = (nwin > 1) ? 0 : 1;
  } else {
    kind = 1;
  }
#pragma unroll 1
  for (;;) {
    if (kind == 0) {
      const uint32_t d = get_bits(r_w, r_words, j * WB, WB);
      Env::sync();
      copy_entry<L, TPI, Env>(sm.b0, comb + ((size_t)j * (1u << WB) + d) * KP);
      Env::sync();
      bp = sm.b0;
    } else if (kind == 1) {
      if (r_w) {   // park obf*R in b1
        Env::sync();
        limbs_to_smem<L, TPI, Env>(sm.b1, x);
        Env::sync();
      }
      limbs_from_words<L, TPI, Env>(x, m_w, m_words);
      bp = nR;
    } else {
      bp = sm.b1;
    }

    montmul<L, TPI, Env>(x, x, bp, n, n0inv);

    if (kind == 0) {
      if (++j == nwin) kind = 1;
    } else if (kind == 1) {
      // x == m*n (mod n^2), < 2 n^2: make it exact and add 1 (1 + m n < n^2 always)
      canonicalize<L, TPI, Env>(x, n);
      if (lane == 0) x[0] += 1u;
      normalize_exact<L, TPI, Env>(x);
      if (!r_w) break;
      kind = 2;
    } else {
      canonicalize<L, TPI, Env>(x, n);
      break;
    }
  }
  store_words<L, TPI, Env>(out_w, out_words, x, sm.b0);
}

// ------------------------------------------------------------------------------------------------
// ct = (1 + m n) * obf mod n^2 with obf given as canonical words (classic r^n path, or apply_obfuscator
// on an existing ciphertext when m_w == nullptr: ct = ct_in * obf).
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
PHE_HD void item_encrypt_finish(const uint32_t* m_w, int m_words, const uint32_t* obf_w, uint32_t* out_w,
                                int out_words, const uint32_t (&n)[L], uint32_t n0inv, const uint32_t* nR,
                                const uint32_t* r2, GroupSmem sm) {
  const int lane = Env::lane();
  uint32_t x[L];
  // b1 = obf * R
  limbs_from_words<L, TPI, Env>(x, obf_w, out_words);
  const uint32_t* bp = r2;
  int kind = 0;
#pragma unroll 1
  for (;;) {
    montmul<L, TPI, Env>(x, x, bp, n, n0inv);
    if (kind == 0) {
      Env::sync();
      limbs_to_smem<L, TPI, Env>(sm.b1, x);
      Env::sync();
      limbs_from_words<L, TPI, Env>(x, m_w, m_words);
      bp = nR;
      kind = 1;
    } else if (kind == 1) {
      canonicalize<L, TPI, Env>(x, n);
      if (lane == 0) x[0] += 1u;
      normalize_exact<L, TPI, Env>(x);
      bp = sm.b1;
      kind = 2;
    } else {
      canonicalize<L, TPI, Env>(x, n);
      break;
    }
  }
  store_words<L, TPI, Env>(out_w, out_words, x, sm.b0);
}

// ------------------------------------------------------------------------------------------------
// Decrypt tail (per ciphertext), everything in the x^2-sized shape:
//   for x in {p, q}:  u_x = c^(x-1) mod x^2 (canonical words, from item_powm)
//       Lx  = (u_x - 1) / x           exact division = -(Montgomery quotient of (u_x - 1) w.r.t. x) mod R
//       m_x = Lx * h_x mod x
//   m = m_p + ((m_q - m_p) * p^-1 mod q) * p
// ipcl::PrivateKey::decryptCRT: computeLfun, *hp/hq, computeCRT.
// Constant entries (cst, KP words each): see DecTailConst.
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
                          const uint32_t* cst, const uint32_t* n0invs /* [p, q, n] */, GroupSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  const int lane = Env::lane();
  uint32_t mod[L], x[L], mp[L], one[L];
  load_entry<L, TPI, Env>(one, cst + DT_ONE * KP);
#pragma unroll
  for (int j = 0; j < L; ++j) mp[j] = 0;

#pragma unroll 1
  for (int idx = 0; idx < 2; ++idx) {
    load_entry<L, TPI, Env>(mod, cst + (idx ? DT_Q : DT_P) * KP);
    const uint32_t n0 = n0invs[idx];
    limbs_from_words<L, TPI, Env>(x, idx ? uq_w : up_w, u_words);
    sub_exact<L, TPI, Env>(x, one);                                  // u - 1  (u >= 1)
    uint32_t q[L];
#pragma unroll
    for (int j = 0; j < L; ++j) q[j] = 0;
    uint32_t junk[L];
    montmul<L, TPI, Env, true>(junk, x, cst + DT_ONE * KP, mod, n0, q);   // q = -(u-1) x^-1 mod R
    // Lx = -q mod R = (~q) + 1 over K limbs
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = (~q[j]) & LMASK;
    if (lane == 0) x[0] += 1u;
    normalize_exact<L, TPI, Env>(x);
    montmul<L, TPI, Env>(x, x, cst + (idx ? DT_HQM : DT_HPM) * KP, mod, n0);
    canonicalize<L, TPI, Env>(x, mod);                               // m_x in [0, x)
    if (idx == 0) {
#pragma unroll
      for (int j = 0; j < L; ++j) mp[j] = x[j];
    }
  }
  // here: mod = q, x = m_q
  const uint32_t neg = sub_exact<L, TPI, Env>(x, mp);                // m_q - m_p
  {                                                                  // + q if negative (branch-free: shuffles inside)
    uint32_t addend[L];
#pragma unroll
    for (int j = 0; j < L; ++j) addend[j] = neg ? mod[j] : 0u;
    add_exact<L, TPI, Env>(x, addend);                               // wraps back into range
  }
  montmul<L, TPI, Env>(x, x, cst + DT_PINVM * KP, mod, n0invs[1]);
  canonicalize<L, TPI, Env>(x, mod);                                 // h in [0, q)
  load_entry<L, TPI, Env>(mod, cst + DT_N * KP);
  montmul<L, TPI, Env>(x, x, cst + DT_PMN * KP, mod, n0invs[2]);
  canonicalize<L, TPI, Env>(x, mod);                                 // h * p (< n)
  add_exact<L, TPI, Env>(x, mp);                                     // m
  store_words<L, TPI, Env>(m_w, m_words, x, sm.b0);
}

}  // namespace phe
