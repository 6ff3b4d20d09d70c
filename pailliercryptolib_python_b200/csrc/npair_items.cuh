// npair_items.cuh -- n-adic pair engine for arithmetic mod n^2 (HE-mul, DJN encrypt, comb-table construction).
//
// Replaces, for x = n, what ipcl::modExp -> mbx_exp_mb8 does on 2k-limb numbers mod n^2 (SURVEY.md 8a rows a2, a6, a7;
// /root/reference/src/ipcl_python/bindings/ipcl_bindings_classes.cpp:53-60, 71-83 (encrypt + obfuscator), :324-325
// (CipherText * PlainText)).  Same idea as the one-lane p-adic engine of decrypt (mont52.cuh: pair_pass), on the
// multi-lane Montgomery product: a number mod n^2 is a pair (X0, X1), X0, X1 < 4n, meaning (X0 + X1 n) R^-1 mod n^2
// with R = 2^(52 K) the Montgomery radix of the n-sized shape (K = L TPI limbs).  With X0 Y0 = u R - m n (u, m the
// result and the quotient of the Montgomery reduction mod n)
//     (X0 + X1 n)(Y0 + Y1 n) R^-1 = u + (X0 Y1 + X1 Y0 - m) R^-1 n      (mod n^2)
// so a product costs 3 n-sized Montgomery products (6 K^2 limb products) and a square 2 (4 K^2), against 8 K^2 for one
// Montgomery product mod n^2.  "- m" enters the second reduction as the addend E = D - m, D = ceil(R / n) n.
//
// Every pass of every product of a kernel goes through ONE montmul_e call site (npair_run); the per-kernel control
// (which product comes next, data movement) is a small state machine `Ctl` with
//     int next(double (&x)[L], const double*& y0, const double*& y1)     -> NPairKind
// State between products: X0 in registers (x), a copy of X0 in sm.xs0, X1 in sm.x1.
#pragma once
#include "paillier_items.cuh"

namespace phe {

// constant entries of the n context ([KP] doubles each, staged into shared memory)
enum NPairEntry { NE_N = 0, NE_ONE, NE_D, NE_W00, NE_W01, NE_W10, NE_W11, NE_OM0, NE_OM1, NE_COUNT };
//   NE_D: low K limbs of D = ceil(R / n) n (top limb: NPairCtxArgs::d_top)
//   NE_Wc0/1: digits (mod n, div n) of 2^(c * 32 * n_words) R^2 mod n^2: chunk c of a ciphertext -> pair form
//   NE_OM0/1: digits of R mod n^2: the number 1 in pair form

struct NPairCtxArgs {
  const double* entries;   // NE_COUNT entries
  uint64_t n0inv;          // -n^-1 mod 2^52
  uint64_t d_top;          // limb K of D
};

struct NPairSmem {         // per lane group; every buffer KP doubles, e: KP + 2 integers
  double *xs0, *x1, *y0, *y1;
  uint64_t* e;
};

enum NPairKind {
  NK_DONE = -1,
  NK_MUL = 0,    // X <- X * Y                       passes: X0 Y0 | X0 Y1 + E | X1 Y0
  NK_SQR = 1,    // X <- X * X                       passes: X0 X0 | 2 X0 X1 + E
  NK_X1Z = 2,    // X = (x, 0):   X <- X * Y         passes: X0 Y0 | X0 Y1 + E
  NK_Y1Z = 3,    // Y = (y0, 0):  X <- X * Y         passes: X0 Y0 | X1 Y0 + E
  NK_MONT = 4,   // x <- x y0 R^-1 mod n             (one ordinary Montgomery product; X1, xs0 untouched)
  NK_PLAIN = 5   // x <- high half, sm.e <- low half of  x * y0 + E  (E = the K + 1 integers in sm.e)
};

// E = D - m in place: sm.e holds the quotient digits m_i (lane 0 wrote them), becomes the K + 1 exact limbs of E
template <int L, int TPI, class Env>
PHE_HD void npair_make_e(uint64_t* e, const double* d_entry, uint64_t d_top) {
  constexpr int LP = Pad<L>::LP;
  uint64_t d[L], q[L];
  Env::sync();
  ints_from_entry<L, TPI, Env>(d, d_entry);
  const uint64_t* qs = e + Env::lane() * LP;
#pragma unroll
  for (int j = 0; j < L; ++j) q[j] = qs[j];
  const uint32_t neg = sub_exact<L, TPI, Env>(d, q);   // D >= R > m: a borrow out of the low K limbs comes off the top one
  ints_to_mem<L, TPI, Env>(e, d);
  if (Env::lane() == 0) e[TPI * LP] = d_top - neg;
  Env::sync();
}

template <int L, int TPI, class Env, class Ctl>
PHE_HD void npair_run(Ctl& ctl, const double* cst, uint64_t n0inv, uint64_t d_top, NPairSmem sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  constexpr int LP = Pad<L>::LP;
  const double* n_entry = cst + NE_N * KP;
  double x[L], z[L];
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = 0.0;
  int kind = NK_DONE, sub = 0;
  const double *y0 = nullptr, *y1 = nullptr;
#pragma unroll 1
  for (;;) {
    if (sub == 0) {
      kind = ctl.next(x, y0, y1);
      if (kind == NK_DONE) break;
    }
    const double* b;
    const uint64_t* ein = nullptr;
    uint64_t* cap = nullptr;
    bool plain = false;
    if (kind >= NK_MONT) {
      b = y0;
      if (kind == NK_PLAIN) { ein = sm.e; cap = sm.e; plain = true; }
    } else if (sub == 0) {
      b = (kind == NK_SQR) ? sm.xs0 : y0;
      cap = sm.e;
    } else if (sub == 1) {
      ein = sm.e;
      if (kind == NK_SQR) {              // a = 2 X0 (exact limbs again), b = X1
        uint64_t t[L];
        ints_of<L>(t, x);
#pragma unroll
        for (int j = 0; j < L; ++j) t[j] <<= 1;
        normalize_exact<L, TPI, Env>(t);
        limbs_of<L>(x, t);
        b = sm.x1;
      } else {
        b = y1;
      }
    } else {                             // a = X1, b = Y0
      limbs_from_mem<L, TPI, Env>(x, sm.x1);
      b = y0;
      if (kind == NK_Y1Z) ein = sm.e;
    }

    montmul_e<L, TPI, Env>(z, x, b, n_entry, n0inv, ein, cap, plain);

    if (kind >= NK_MONT) {
#pragma unroll
      for (int j = 0; j < L; ++j) x[j] = z[j];
      continue;
    }
    bool finish = false;
    if (sub == 0) {
      Env::sync();                       // every lane is done reading b (xs0 when squaring)
      limbs_to_mem<L, TPI, Env>(sm.xs0, z);   // Z0
      npair_make_e<L, TPI, Env>(sm.e, cst + NE_D * KP, d_top);
      sub = (kind == NK_Y1Z) ? 2 : 1;
    } else if (sub == 1) {
      Env::sync();                       // lane 0 is done reading E, every lane its b
      if (kind == NK_MUL) {              // park the first cross term in E's buffer
        limbs_to_mem<L, TPI, Env>(reinterpret_cast<double*>(sm.e), z);
        Env::sync();
        sub = 2;
      } else {
        limbs_to_mem<L, TPI, Env>(sm.x1, z);
        finish = true;
      }
    } else {
      if (kind == NK_MUL) {              // Z1 = (X0 Y1 - m) R^-1 + X1 Y0 R^-1, < 4n
        uint64_t t[L];
        ints_of<L>(t, z);
        const double* pk = reinterpret_cast<const double*>(sm.e) + Env::lane() * LP;
#pragma unroll
        for (int j = 0; j < L; ++j) t[j] += int_of(pk[j]);
        normalize_exact<L, TPI, Env>(t);
        limbs_of<L>(z, t);
      }
      Env::sync();
      limbs_to_mem<L, TPI, Env>(sm.x1, z);
      finish = true;
    }
    if (finish) {
      Env::sync();
      limbs_from_mem<L, TPI, Env>(x, sm.xs0);   // X0 = Z0
      sub = 0;
    }
  }
}

// ---- helpers shared by the controls ------------------------------------------------------------------------------

// pair entry in memory: [X0 entry (KP)][X1 entry (KP)]
template <int L, int TPI, class Env>
PHE_HD void npair_store(double* dst, const double (&x)[L], const NPairSmem& sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  limbs_to_mem<L, TPI, Env>(dst, x);
  copy_entry<L, TPI, Env>(dst + KP, sm.x1);
}
template <int L, int TPI, class Env>
PHE_HD void npair_load(double (&x)[L], const double* src, const NPairSmem& sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  limbs_from_mem<L, TPI, Env>(x, src);
  Env::sync();
  limbs_to_mem<L, TPI, Env>(sm.xs0, x);
  copy_entry<L, TPI, Env>(sm.x1, src + KP);
  Env::sync();
}
template <int L, int TPI, class Env>
PHE_HD void npair_set_y(const double* src, const NPairSmem& sm) {
  constexpr int KP = Shape<L, TPI>::KP;
  Env::sync();
  copy_entry<L, TPI, Env>(sm.y0, src);
  copy_entry<L, TPI, Env>(sm.y1, src + KP);
  Env::sync();
}

// Ciphertext words -> pair form, three steps of a control: chunk 0 (X1Z by W0), chunk 1 (X1Z by W1), sum.
//   step 0: x <- chunk 0;                      product by (W00, W01)
//   step 1: park the result in y0 / y1; x <- chunk 1;   product by (W10, W11)
//   step 2: X <- parked + X   (digits < 4n)
template <int L, int TPI, class Env>
PHE_HD void npair_conv_step(int step, double (&x)[L], const uint32_t* c_w, int chunk_words, const double* cst,
                            const NPairSmem& sm, const double*& y0, const double*& y1) {
  constexpr int KP = Shape<L, TPI>::KP;
  constexpr int LP = Pad<L>::LP;
  if (step == 0) {
    limbs_from_words<L, TPI, Env>(x, c_w, chunk_words);
    y0 = cst + NE_W00 * KP; y1 = cst + NE_W01 * KP;
  } else if (step == 1) {
    Env::sync();
    copy_entry<L, TPI, Env>(sm.y0, sm.xs0);
    copy_entry<L, TPI, Env>(sm.y1, sm.x1);
    Env::sync();
    limbs_from_words<L, TPI, Env>(x, c_w + chunk_words, chunk_words);
    y0 = cst + NE_W10 * KP; y1 = cst + NE_W11 * KP;
  } else {
    const int lane = Env::lane();
    uint64_t t[L];
    ints_of<L>(t, x);
#pragma unroll
    for (int j = 0; j < L; ++j) t[j] += int_of(sm.y0[lane * LP + j]);
    normalize_exact<L, TPI, Env>(t);
    limbs_of<L>(x, t);
#pragma unroll
    for (int j = 0; j < L; ++j) t[j] = int_of(sm.x1[lane * LP + j]) + int_of(sm.y1[lane * LP + j]);
    normalize_exact<L, TPI, Env>(t);
    Env::sync();
    limbs_to_mem<L, TPI, Env>(sm.xs0, x);
    {
      double w[L];
      limbs_of<L>(w, t);
      limbs_to_mem<L, TPI, Env>(sm.x1, w);
    }
    Env::sync();
  }
}

// After the product by (1, 0): x = Z0, sm.x1 = Z1 with value = Z0 + Z1 n (mod n^2), Z0, Z1 < 2n.  Canonical digits
// v0 = value mod n, v1 = value div n (+ extra, a number < 2n as limbs in `extra`, or null), then the setup of the final
// plain product  v1 * n + v0:  x <- v1, sm.e <- v0.
template <int L, int TPI, class Env>
PHE_HD void npair_canon_setup(double (&x)[L], const double* cst, const NPairSmem& sm, const double* extra) {
  constexpr int KP = Shape<L, TPI>::KP;
  constexpr int LP = Pad<L>::LP;
  const int lane = Env::lane();
  uint64_t v0[L], v1[L], ni[L], d[L];
  ints_of<L>(v0, x);
  ints_from_entry<L, TPI, Env>(ni, cst + NE_N * KP);
#pragma unroll
  for (int j = 0; j < L; ++j) { v1[j] = int_of(sm.x1[lane * LP + j]); d[j] = v0[j]; }
  if (!sub_exact<L, TPI, Env>(d, ni)) {   // Z0 >= n: carry one n into the second digit
#pragma unroll
    for (int j = 0; j < L; ++j) v0[j] = d[j];
    if (lane == 0) v1[0] += 1ull;
  }
  if (extra) {
#pragma unroll
    for (int j = 0; j < L; ++j) v1[j] += int_of(extra[lane * LP + j]);
  }
  normalize_exact<L, TPI, Env>(v1);       // < 4n + 1
#pragma unroll 1
  for (int k = 0; k < 4; ++k) cond_sub<L, TPI, Env>(v1, ni);
  limbs_of<L>(x, v1);
  Env::sync();
  ints_to_mem<L, TPI, Env>(sm.e, v0);
  if (lane == 0) sm.e[TPI * LP] = 0ull;
  Env::sync();
}

// After NK_PLAIN: x = high K limbs, sm.e = low K limbs (integers, padded layout) -> little-endian words.
// peers (or null): n_peers more destinations of the same row (the gather buffers of the other GPUs, mapped peer memory:
// the stores go out over NVLink while the group computes on)
constexpr int NPAIR_MAX_PEERS = 15;
template <int L, int TPI, class Env>
PHE_HD void npair_store_words(uint32_t* out, int nwords, const double (&x)[L], const NPairSmem& sm,
                              uint32_t* const* peers = nullptr, int n_peers = 0, size_t peer_off = 0) {
  constexpr int LP = Pad<L>::LP;
  constexpr int K = L * TPI;
  uint64_t hi[L];
  ints_of<L>(hi, x);
  uint64_t* hs = reinterpret_cast<uint64_t*>(sm.y0);
  Env::sync();
  ints_to_mem<L, TPI, Env>(hs, hi);
  Env::sync();
  if (out) {
    const uint64_t* lo = sm.e;
    for (int v = Env::lane(); v < nwords; v += TPI) {
      const int bit = v * 32, g = bit / LW, o = bit - g * LW;
      auto limb = [&](int gg) -> uint64_t {
        if (gg >= 2 * K) return 0ull;
        const uint64_t* src = gg < K ? lo : hs;
        const int k = gg < K ? gg : gg - K;
        return src[(k / L) * LP + (k % L)];
      };
      uint64_t u = limb(g) >> o;
      if (o > LW - 32) u |= limb(g + 1) << (LW - o);
      out[v] = (uint32_t)u;
      for (int k = 0; k < n_peers; ++k) peers[k][peer_off + v] = (uint32_t)u;
    }
  }
  Env::sync();
}

// ------------------------------------------------------------------------------------------------
// HE mul: out = c^e mod n^2, per-item exponent, fixed window WIN (ipcl::CipherText::operator*(PlainText) -> raw_mul ->
// ipcl::modExp; ipcl_bindings_classes.cpp:324-325).  Table: (1 << WIN) pair entries in global memory per group.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env, int WIN>
struct NPairPowmCtl {
  static constexpr int KP = Shape<L, TPI>::KP;
  static constexpr int TS = 1 << WIN;
  const uint32_t* c_w; int chunk_words;
  const uint32_t* e_w; int e_words; int ebits;
  uint32_t* out_w; int out_words;
  const double* cst; double* tbl; NPairSmem sm;
  int phase = 0, ti = 2, sq = 0, w = 0;

  PHE_HD int next(double (&x)[L], const double*& y0, const double*& y1) {
    int nd = (ebits + WIN - 1) / WIN;
    if (nd < 1) nd = 1;
#pragma unroll 1
    for (;;) {
      switch (phase) {
        case 0: case 1:
          npair_conv_step<L, TPI, Env>(phase, x, c_w, chunk_words, cst, sm, y0, y1);
          ++phase;
          return NK_X1Z;
        case 2:
          npair_conv_step<L, TPI, Env>(2, x, c_w, chunk_words, cst, sm, y0, y1);
          copy_entry<L, TPI, Env>(tbl, cst + NE_OM0 * KP);            // T[0] = 1
          copy_entry<L, TPI, Env>(tbl + KP, cst + NE_OM1 * KP);
          npair_store<L, TPI, Env>(tbl + 2 * KP, x, sm);              // T[1] = c
          npair_set_y<L, TPI, Env>(tbl + 2 * KP, sm);                 // Y = c for the whole table build
          y0 = sm.y0; y1 = sm.y1;
          if (TS > 2) { phase = 3; return NK_MUL; }
          phase = 4;
          break;
        case 3:
          npair_store<L, TPI, Env>(tbl + (size_t)ti * 2 * KP, x, sm);
          if (++ti < TS) return NK_MUL;
          phase = 4;
          break;
        case 4: {
          const uint32_t d = get_bits(e_w, e_words, (nd - 1) * WIN, WIN);
          npair_load<L, TPI, Env>(x, tbl + (size_t)d * 2 * KP, sm);
          w = nd - 2;
          sq = 0;
          phase = (w < 0) ? 7 : 5;
          break;
        }
        case 5:
          if (sq < WIN) { ++sq; return NK_SQR; }
          {
            const uint32_t d = get_bits(e_w, e_words, w * WIN, WIN);
            npair_set_y<L, TPI, Env>(tbl + (size_t)d * 2 * KP, sm);
            y0 = sm.y0; y1 = sm.y1;
          }
          phase = 6;
          return NK_MUL;
        case 6:
          --w; sq = 0;
          phase = (w < 0) ? 7 : 5;
          break;
        case 7:
          y0 = cst + NE_ONE * KP; y1 = nullptr;
          phase = 8;
          return NK_Y1Z;
        case 8:
          npair_canon_setup<L, TPI, Env>(x, cst, sm, nullptr);
          y0 = cst + NE_N * KP;
          phase = 9;
          return NK_PLAIN;
        default:
          npair_store_words<L, TPI, Env>(out_w, out_words, x, sm);
          return NK_DONE;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Exponent alignment: out = c^(2^delta) mod n^2 = delta squarings (the reference multiplies by the plaintext BASE^delta
// with a full modexp: ipcl_python.py:551-560, 602-606, 672-690).  Every lane group of a CTA runs max_delta squarings
// (the shuffles of a warp must stay convergent); a group whose own delta is smaller keeps a copy of its value from
// the moment it was reached (in y0 / y1, which a square does not touch) and takes it back at the end.  The host sorts
// the rows by delta, so the squarings thrown away are few.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
struct NPairScaleCtl {
  static constexpr int KP = Shape<L, TPI>::KP;
  const uint32_t* c_w; int chunk_words;
  int delta, max_delta;
  uint32_t* out_w; int out_words;
  const double* cst; NPairSmem sm;
  int phase = 0, sq = 0;

  PHE_HD int next(double (&x)[L], const double*& y0, const double*& y1) {
#pragma unroll 1
    for (;;) {
      switch (phase) {
        case 0: case 1:
          npair_conv_step<L, TPI, Env>(phase, x, c_w, chunk_words, cst, sm, y0, y1);
          ++phase;
          return NK_X1Z;
        case 2:
          npair_conv_step<L, TPI, Env>(2, x, c_w, chunk_words, cst, sm, y0, y1);
          phase = 3;
          break;
        case 3:
          Env::sync();
          if (sq == delta) {            // this group's result: park it (no shuffles in here: groups may diverge)
            copy_entry<L, TPI, Env>(sm.y0, sm.xs0);
            copy_entry<L, TPI, Env>(sm.y1, sm.x1);
          }
          Env::sync();
          if (sq < max_delta) { ++sq; return NK_SQR; }
          npair_load<L, TPI, Env>(x, sm.y0, sm);     // y0, y1 are adjacent: a pair entry
          y0 = cst + NE_ONE * KP; y1 = nullptr;
          phase = 4;
          return NK_Y1Z;
        case 4:
          npair_canon_setup<L, TPI, Env>(x, cst, sm, nullptr);
          y0 = cst + NE_N * KP;
          phase = 5;
          return NK_PLAIN;
        default:
          npair_store_words<L, TPI, Env>(out_w, out_words, x, sm);
          return NK_DONE;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Sliding-window exponentiation mod n^2 for an exponent SHARED by every item (classic obfuscator r^n; the program
// format is item_powm_prog's, paillier_items.cuh): table of the odd powers T[k] = x^(2k+1), k < 2^(WS-1), as pair
// entries in global memory.  Base: nchunks (1 or 2) chunks of chunk_words words.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env, int WS>
struct NPairProgCtl {
  static constexpr int KP = Shape<L, TPI>::KP;
  static constexpr int TS = 1 << (WS - 1);
  const uint32_t* c_w; int chunk_words, nchunks;
  const uint32_t* prog; int nprog;
  uint32_t* out_w; int out_words;
  const double* cst; double* tbl; NPairSmem sm;
  int phase = 0, ti = 1, sq = 0, pc = 1;
  uint32_t op = 0;

  PHE_HD int next(double (&x)[L], const double*& y0, const double*& y1) {
#pragma unroll 1
    for (;;) {
      switch (phase) {
        case 0:
          npair_conv_step<L, TPI, Env>(0, x, c_w, chunk_words, cst, sm, y0, y1);
          phase = (nchunks > 1) ? 1 : 3;
          return NK_X1Z;
        case 1:
          npair_conv_step<L, TPI, Env>(1, x, c_w, chunk_words, cst, sm, y0, y1);
          phase = 2;
          return NK_X1Z;
        case 2:
          npair_conv_step<L, TPI, Env>(2, x, c_w, chunk_words, cst, sm, y0, y1);
          phase = 3;
          break;
        case 3:   // X = x in pair form: T[0]; then x^2 as the table multiplier
          npair_store<L, TPI, Env>(tbl, x, sm);
          if (TS > 1 && prog[0] != PROG_ONE) { phase = 4; return NK_SQR; }
          phase = 6;
          break;
        case 4:   // X = x^2: park it in slot 1 (rewritten by x^3 later), Y = x^2, restart from T[0]
          npair_store<L, TPI, Env>(tbl + 2 * KP, x, sm);
          npair_set_y<L, TPI, Env>(tbl + 2 * KP, sm);
          npair_load<L, TPI, Env>(x, tbl, sm);
          y0 = sm.y0; y1 = sm.y1;
          phase = 5;
          return NK_MUL;
        case 5:
          npair_store<L, TPI, Env>(tbl + (size_t)ti * 2 * KP, x, sm);
          if (++ti < TS) { y0 = sm.y0; y1 = sm.y1; return NK_MUL; }
          phase = 6;
          break;
        case 6: {  // leading window
          const uint32_t i0 = prog[0];
          if (i0 == PROG_ONE) {
            limbs_from_mem<L, TPI, Env>(x, cst + NE_OM0 * KP);
            Env::sync();
            limbs_to_mem<L, TPI, Env>(sm.xs0, x);
            copy_entry<L, TPI, Env>(sm.x1, cst + NE_OM1 * KP);
            Env::sync();
          } else {
            npair_load<L, TPI, Env>(x, tbl + (size_t)i0 * 2 * KP, sm);
          }
          pc = 1;
          phase = 7;
          break;
        }
        case 7:   // next program entry
          if (pc > nprog) { phase = 10; break; }
          op = prog[pc++];
          sq = (int)(op >> 8);
          phase = 8;
          break;
        case 8:
          if (sq > 0) { --sq; return NK_SQR; }
          phase = 7;
          if ((op & 0xffu) != PROG_NOMUL) {
            npair_set_y<L, TPI, Env>(tbl + (size_t)(op & 0xffu) * 2 * KP, sm);
            y0 = sm.y0; y1 = sm.y1;
            return NK_MUL;
          }
          break;
        case 10:
          y0 = cst + NE_ONE * KP; y1 = nullptr;
          phase = 11;
          return NK_Y1Z;
        case 11:
          npair_canon_setup<L, TPI, Env>(x, cst, sm, nullptr);
          y0 = cst + NE_N * KP;
          phase = 12;
          return NK_PLAIN;
        default:
          npair_store_words<L, TPI, Env>(out_w, out_words, x, sm);
          return NK_DONE;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// DJN encrypt on the pair engine with a fixed-base comb table of pair entries:
//   obf = prod_j T[j][digit_j(r)],  T[j][d] = hs^(d 2^(wb j)) in pair form;  ct = (1 + m n) obf mod n^2
// With obf = V0 + V1 n:  ct = V0 + (V1 + m V0 mod n) n, and m V0 mod n = montmul(m, X0) (X0 = V0 R mod n).
// ipcl::PublicKey::encrypt + applyObfuscator (ipcl_bindings_classes.cpp:53-60, 71-83).  r_w == null: ct = 1 + m n.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
struct NPairEncCtl {
  static constexpr int KP = Shape<L, TPI>::KP;
  const uint32_t* m_w; int m_words;
  const uint32_t* r_w; int r_words; int nwin, wb;
  uint32_t* out_w; int out_words;
  const double* cst; const double* comb; NPairSmem sm;
  uint32_t* const* peers = nullptr;   // n_peers more output matrices (the other GPUs' gather buffers); this row starts
  int n_peers = 0;                    // peer_off words into each
  size_t peer_off = 0;
  int phase = 0, j = 1;

  PHE_HD const double* entry(int win, uint32_t d) const { return comb + ((((size_t)win) << wb) + d) * 2 * KP; }
  // The table is tens of GB of random 2 KP-double entries (640 B at 2048-bit keys): every fetch is a DRAM access.  The
  // digits of r are known up front, so one lane of the group asks the TMA unit to bring the next entry into the L2
  // (cp.async.bulk.prefetch.L2) a whole pair product (~10 us) before npair_set_y copies it.
  PHE_HD void prefetch(int win) const {
    if (Env::lane() == 0) Env::prefetch_l2(entry(win, get_bits(r_w, r_words, win * wb, wb)), (uint32_t)(2 * KP * sizeof(double)));
  }

  PHE_HD int next(double (&x)[L], const double*& y0, const double*& y1) {
#pragma unroll 1
    for (;;) {
      switch (phase) {
        case 0:
          if (!r_w) {   // make_secure = false: 1 + m n as a plain product
            limbs_from_words<L, TPI, Env>(x, m_w, m_words);
            Env::sync();
            {
              uint64_t one[L];
#pragma unroll
              for (int k = 0; k < L; ++k) one[k] = 0ull;
              if (Env::lane() == 0) one[0] = 1ull;
              ints_to_mem<L, TPI, Env>(sm.e, one);
              if (Env::lane() == 0) sm.e[KP] = 0ull;
            }
            Env::sync();
            y0 = cst + NE_N * KP;
            phase = 9;
            return NK_PLAIN;
          }
          if (nwin > 1) prefetch(1);
          npair_load<L, TPI, Env>(x, entry(0, get_bits(r_w, r_words, 0, wb)), sm);
          phase = (nwin > 1) ? 1 : 2;
          break;
        case 1:
          if (j + 1 < nwin) prefetch(j + 1);     // the entry of the next window: on its way out of HBM under this product
          npair_set_y<L, TPI, Env>(entry(j, get_bits(r_w, r_words, j * wb, wb)), sm);
          y0 = sm.y0; y1 = sm.y1;
          if (++j == nwin) phase = 2;
          return NK_MUL;
        case 2: {   // t = m X0 R^-1 = m V0 mod n
          double t[L];
          limbs_from_words<L, TPI, Env>(t, m_w, m_words);
          Env::sync();
          limbs_to_mem<L, TPI, Env>(sm.y0, t);
          Env::sync();
          y0 = sm.y0;
          phase = 3;
          return NK_MONT;
        }
        case 3:
          Env::sync();
          limbs_to_mem<L, TPI, Env>(sm.y1, x);      // park t
          Env::sync();
          limbs_from_mem<L, TPI, Env>(x, sm.xs0);   // X0 again
          y0 = cst + NE_ONE * KP; y1 = nullptr;
          phase = 4;
          return NK_Y1Z;
        case 4:
          npair_canon_setup<L, TPI, Env>(x, cst, sm, sm.y1);
          y0 = cst + NE_N * KP;
          phase = 9;
          return NK_PLAIN;
        default:
          npair_store_words<L, TPI, Env>(out_w, out_words, x, sm, peers, out_w ? n_peers : 0, peer_off);
          return NK_DONE;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Comb table construction in pair form.
//   bases: the chain hs^(2^i), i < wb nwin: element i = wb j + k is T[j][2^k]; T[j][0] = 1 alongside
//   level k: T[j][2^k + e] = T[j][e] * T[j][2^k], 0 < e < 2^k
// ------------------------------------------------------------------------------------------------
template <int L, int TPI, class Env>
struct NPairCombBasesCtl {
  static constexpr int KP = Shape<L, TPI>::KP;
  const uint32_t* hs_w; int chunk_words;
  int nwin, wb; double* comb; bool writer;
  const double* cst; NPairSmem sm;
  int phase = 0, i = 0;

  PHE_HD int next(double (&x)[L], const double*& y0, const double*& y1) {
    if (phase < 2) {
      npair_conv_step<L, TPI, Env>(phase, x, hs_w, chunk_words, cst, sm, y0, y1);
      ++phase;
      return NK_X1Z;
    }
    if (phase == 2) {
      npair_conv_step<L, TPI, Env>(2, x, hs_w, chunk_words, cst, sm, y0, y1);
      phase = 3;
    }
    const int jw = i / wb, k = i - jw * wb;
    double* row = comb + (((size_t)jw) << wb) * 2 * KP;
    if (writer) {
      npair_store<L, TPI, Env>(row + ((size_t)1 << k) * 2 * KP, x, sm);
      if (k == 0) {
        copy_entry<L, TPI, Env>(row, cst + NE_OM0 * KP);
        copy_entry<L, TPI, Env>(row + KP, cst + NE_OM1 * KP);
      }
    }
    if (++i == wb * nwin) return NK_DONE;
    return NK_SQR;
  }
};

template <int L, int TPI, class Env>
struct NPairCombLevelCtl {
  static constexpr int KP = Shape<L, TPI>::KP;
  double* row; int level, e; bool store;
  NPairSmem sm;
  int phase = 0;

  PHE_HD int next(double (&x)[L], const double*& y0, const double*& y1) {
    if (phase == 0) {
      npair_load<L, TPI, Env>(x, row + (size_t)e * 2 * KP, sm);
      npair_set_y<L, TPI, Env>(row + ((size_t)1 << level) * 2 * KP, sm);
      y0 = sm.y0; y1 = sm.y1;
      phase = 1;
      return NK_MUL;
    }
    if (store) npair_store<L, TPI, Env>(row + (((size_t)1 << level) + e) * 2 * KP, x, sm);
    return NK_DONE;
  }
};

}  // namespace phe
