// Host lockstep emulator for the lane-group code in csrc/paillier_items.cuh (TEST INFRASTRUCTURE).
// Each lane of a group is a std::thread; shuffles are barrier-synchronised exchanges.
// Built by tests/emu_util.py: g++ -O2 -std=c++20 -frounding-math -pthread -shared -fPIC.  Every lane thread runs with
// the rounding mode FE_TOWARDZERO so that std::fma reproduces the device's fma.rz.f64 bit for bit.
#include <atomic>
#include <cfenv>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../../pailliercryptolib_python_b200/csrc/paillier_items.cuh"
#include "../../pailliercryptolib_python_b200/csrc/npair_items.cuh"

namespace {

// Sense-reversing spin barrier: the lanes of a group exchange values thousands of times per Montgomery product and a
// futex-based std::barrier costs microseconds per phase.
struct SpinBarrier {
  const int n;
  std::atomic<int> waiting{0};
  std::atomic<int> phase{0};
  explicit SpinBarrier(int n_) : n(n_) {}
  void arrive_and_wait() {
    const int ph = phase.load(std::memory_order_acquire);
    if (waiting.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
      waiting.store(0, std::memory_order_relaxed);
      phase.store(ph + 1, std::memory_order_release);
    } else {
      int spins = 0;
      while (phase.load(std::memory_order_acquire) == ph)
        if (++spins > 2000) { std::this_thread::yield(); spins = 0; }
    }
  }
};

struct Exchange {
  uint32_t slot[32];
  SpinBarrier bar;
  explicit Exchange(int n) : bar(n) {}
};

thread_local int t_lane = 0;
thread_local Exchange* t_ex = nullptr;

template <int TPI> struct EmuEnv {
  static int lane() { return t_lane; }
  static uint32_t xchg(uint32_t v, int src) {
    t_ex->slot[t_lane] = v;
    t_ex->bar.arrive_and_wait();
    const uint32_t r = t_ex->slot[src];
    t_ex->bar.arrive_and_wait();
    return r;
  }
  static uint32_t bcast(uint32_t v, int src) { return TPI == 1 ? v : xchg(v, src); }
  static uint32_t from_above(uint32_t v) { return TPI == 1 ? 0u : xchg(v, (t_lane + 1) % TPI); }
  static uint32_t from_below(uint32_t v) { return TPI == 1 ? 0u : xchg(v, (t_lane + TPI - 1) % TPI); }
  static bool any(bool p) {
    if (TPI == 1) return p;
    bool r = false;
    for (int l = 0; l < TPI; ++l) r = r || xchg(p ? 1u : 0u, l) != 0;
    return r;
  }
  static void sync() { if (TPI > 1) t_ex->bar.arrive_and_wait(); }
  static void cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
  static void cp_async_wait() {}
  static void prefetch_l2(const void*, uint32_t) {}
};

template <int TPI> void run_group(const std::function<void()>& body) {
  Exchange ex(TPI);
  std::vector<std::thread> th;
  for (int l = 0; l < TPI; ++l)
    th.emplace_back([&, l] { t_lane = l; t_ex = &ex; std::fesetround(FE_TOWARDZERO); body(); });
  for (auto& t : th) t.join();
}

template <int L, int TPI> struct Bufs {
  std::vector<double> b0, b1;
  phe::GroupSmem sm;
  Bufs() : b0(phe::Shape<L, TPI>::KP + 2), b1(phe::Shape<L, TPI>::KP + 2) {
    auto al = [](std::vector<double>& v) { uintptr_t p = (uintptr_t)v.data(); return (double*)((p + 15) & ~(uintptr_t)15); };
    sm.b0 = al(b0); sm.b1 = al(b1);
  }
};

// 16-byte aligned copy of caller memory (entries are read with 16-byte vector loads)
struct AlignedCopy {
  std::vector<double> store; double* p;
  AlignedCopy(const double* src, size_t n) : store(n + 2) {
    uintptr_t a = ((uintptr_t)store.data() + 15) & ~(uintptr_t)15; p = (double*)a; std::memcpy(p, src, n * 8);
  }
};

template <int L, int TPI>
int do_modmul(const uint32_t* a, const uint32_t* b, uint32_t* out, int nwords, int count, const double* n_e,
              uint64_t n0inv, const double* r2_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::item_modmul<L, TPI, Env>(a + (size_t)i * nwords, b + (size_t)i * nwords, out + (size_t)i * nwords, nwords,
                                    ne.p, n0inv, r2.p, bufs.sm);
    });
  }
  return 0;
}

// one-product HE add: b as words (b_w) or as a constant entry (b_e)
template <int L, int TPI>
int do_modmul1(const uint32_t* a, const uint32_t* b_w, const double* b_e, uint32_t* out, int nwords, int count,
               const double* n_e, uint64_t n0inv) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP);
  std::vector<double> zero(KP, 0.0);
  AlignedCopy be(b_e ? b_e : zero.data(), KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::item_modmul1<L, TPI, Env>(a + (size_t)i * nwords, b_w ? b_w + (size_t)i * nwords : nullptr, b_e ? be.p : nullptr,
                                     out + (size_t)i * nwords, nwords, ne.p, n0inv, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI, int WIN>
int do_powm(const uint32_t* base, int base_words, const double* base_mont, const uint32_t* e, int e_words,
            int e_stride, int ebits, uint32_t* out, int out_words, int count, const double* n_e, uint64_t n0inv,
            const double* r2_e, const double* oneM_e, const double* one_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP), oneM(oneM_e, KP), one(one_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    std::vector<double> tbl((size_t)(1 << WIN) * KP + 2);
    double* tp = (double*)(((uintptr_t)tbl.data() + 15) & ~(uintptr_t)15);
    const double* bmp = nullptr;
    AlignedCopy* bmc = nullptr;
    if (base_mont) { bmc = new AlignedCopy(base_mont + (size_t)i * KP, KP); bmp = bmc->p; }
    run_group<TPI>([&] {
      phe::item_powm<L, TPI, Env, WIN>(base ? base + (size_t)i * base_words : nullptr, base_words, bmp,
                                       e + (size_t)i * e_stride, e_words, ebits, out + (size_t)i * out_words,
                                       out_words, ne.p, n0inv, r2.p, oneM.p, one.p, tp, bufs.sm);
    });
    delete bmc;
  }
  return 0;
}

template <int L, int TPI>
int do_powm_prog(const uint32_t* base, int base_words, const double* base_mont, const uint32_t* prog, int nprog,
                 uint32_t* out, int out_words, int count, const double* n_e, uint64_t n0inv, const double* r2_e,
                 const double* oneM_e, const double* one_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  constexpr int WS = 6;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP), oneM(oneM_e, KP), one(one_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    std::vector<double> tbl((size_t)(1 << (WS - 1)) * KP + 2);
    double* tp = (double*)(((uintptr_t)tbl.data() + 15) & ~(uintptr_t)15);
    const double* bmp = nullptr;
    AlignedCopy* bmc = nullptr;
    if (base_mont) { bmc = new AlignedCopy(base_mont + (size_t)i * KP, KP); bmp = bmc->p; }
    run_group<TPI>([&] {
      phe::item_powm_prog<L, TPI, Env, WS>(base ? base + (size_t)i * base_words : nullptr, base_words, bmp, prog, nprog,
                                           out + (size_t)i * out_words, out_words, ne.p, n0inv, r2.p, oneM.p, one.p,
                                           tp, bufs.sm);
    });
    delete bmc;
  }
  return 0;
}

template <int L, int TPI>
int do_inv_block(bool unwind, const uint32_t* c, int nwords, int count, int block, double* P, uint32_t* totals,
                 const uint32_t* tinv, uint32_t* out, const double* n_e, uint64_t n0inv, const double* r2_e,
                 const double* oneM_e, const double* one_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP), oneM(oneM_e, KP), one(one_e, KP);
  std::vector<double> Pa((size_t)count * KP + 2);
  double* Pp = (double*)(((uintptr_t)Pa.data() + 15) & ~(uintptr_t)15);
  if (unwind) std::memcpy(Pp, P, (size_t)count * KP * 8);
  for (int b = 0; b < count / block; ++b) {
    Bufs<L, TPI> bufs;
    const size_t first = (size_t)b * block;
    run_group<TPI>([&] {
      if (!unwind)
        phe::item_inv_prefix<L, TPI, Env>(c + first * nwords, nwords, block, Pp + first * KP, totals + (size_t)b * nwords,
                                          ne.p, n0inv, r2.p, oneM.p, one.p, bufs.sm);
      else
        phe::item_inv_unwind<L, TPI, Env>(c + first * nwords, nwords, block, Pp + first * KP, tinv + (size_t)b * nwords,
                                          out + first * nwords, ne.p, n0inv, r2.p, oneM.p, one.p, bufs.sm);
    });
  }
  if (!unwind) std::memcpy(P, Pp, (size_t)count * KP * 8);
  return 0;
}

template <int L, int TPI>
int do_dec_prep(const uint32_t* c, int hw, double* out_entries, int count, const double* n_e, uint64_t n0inv,
                const double* r2_e, const double* k2_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP), k2(k2_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::item_dec_prep<L, TPI, Env>(c + (size_t)i * 2 * hw, hw, out_entries + (size_t)i * KP, ne.p, n0inv, r2.p, k2.p,
                                      bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI>
int do_encrypt_comb(const uint32_t* m, int m_words, const uint32_t* r, int r_words, int nwin, int wb, uint32_t* out,
                    int out_words, int count, const double* n_e, uint64_t n0inv, const double* nR_e,
                    const double* comb, size_t comb_doubles) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), nR(nR_e, KP), cb(comb, comb_doubles);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::item_encrypt_comb<L, TPI, Env>(m + (size_t)i * m_words, m_words, r ? r + (size_t)i * r_words : nullptr,
                                              r_words, nwin, wb, out + (size_t)i * out_words, out_words, ne.p, n0inv, nR.p,
                                              cb.p, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI>
int do_encrypt_finish(const uint32_t* m, int m_words, const uint32_t* obf, uint32_t* out, int out_words, int count,
                      const double* n_e, uint64_t n0inv, const double* nR_e, const double* r2_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), nR(nR_e, KP), r2(r2_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::item_encrypt_finish<L, TPI, Env>(m + (size_t)i * m_words, m_words, obf + (size_t)i * out_words,
                                            out + (size_t)i * out_words, out_words, ne.p, n0inv, nR.p, r2.p, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI>
int do_dec_tail(const uint32_t* up, const uint32_t* uq, int u_words, uint32_t* m, int m_words, int count,
                const double* cst, const uint64_t* n0invs) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy c(cst, (size_t)phe::DT_COUNT * KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::item_dec_tail<L, TPI, Env>(up + (size_t)i * u_words, uq + (size_t)i * u_words, u_words,
                                      m + (size_t)i * m_words, m_words, c.p, n0invs, bufs.sm);
    });
  }
  return 0;
}

// one-bignum-per-lane environment of the p-adic pair engine: a single column (stride 1)
struct EmuPairEnv {
  static constexpr int STRIDE = 1;
  static int lane() { return 0; }
  static int column() { return 0; }
  static uint32_t bcast(uint32_t v, int) { return v; }
  static uint32_t from_above(uint32_t) { return 0u; }
  static uint32_t from_below(uint32_t) { return 0u; }
  static bool any(bool p) { return p; }
  static void sync() {}
  static void cp_async8(void* dst, const void* src) { std::memcpy(dst, src, 8); }
  static void cp_async_wait() {}
};

// nseg > 1: the program is the time-sliced one (k_dec_pair); every segment starts on FRESH per-lane state (as when
// another warp continues the unit), only the unit's table (slots + 1 entries, the last holds the parked pair) persists
template <int L>
int do_dec_pair(const uint32_t* c, int c_words, int chunk_words, const uint32_t* prog, uint32_t* out, int out_words,
                int count, const double* mod, uint64_t n0inv, const double* cst, int slots, int nseg = 1,
                const int* seg_off = nullptr) {
  std::fesetround(FE_TOWARDZERO);
  for (int i = 0; i < count; ++i) {
    std::vector<double> tbl((size_t)(slots + 1) * 2 * L);
    for (int s = 0; s < nseg; ++s) {
      std::vector<double> xs0(L, 7.0), x1(L, 7.0), y0(L, 7.0), y1(L, 7.0);
      std::vector<int64_t> e(L + 1, 7);
      phe::PairSmem<EmuPairEnv> sm{xs0.data(), x1.data(), y0.data(), y1.data(), e.data()};
      phe::item_dec_pair<L, EmuPairEnv>(c + (size_t)i * c_words, chunk_words, prog + (seg_off ? seg_off[s] : 0),
                                        out + (size_t)i * out_words, out_words, mod, n0inv, cst, tbl.data(), sm);
    }
  }
  std::fesetround(FE_TONEAREST);
  return 0;
}

template <int L, int TPI>
int do_dec_crt(const uint32_t* mp, const uint32_t* mq, int half_words, uint32_t* m, int m_words, int count,
               const double* cst, const uint64_t* n0invs) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy c(cst, (size_t)phe::DT_COUNT * KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::item_dec_crt<L, TPI, Env>(mp + (size_t)i * half_words, mq + (size_t)i * half_words, half_words,
                                     m + (size_t)i * m_words, m_words, c.p, n0invs, bufs.sm);
    });
  }
  return 0;
}

// ---- n-adic pair engine (npair_items.cuh) ----
template <int L, int TPI> struct NBufs {
  static constexpr int KP = phe::Shape<L, TPI>::KP;
  std::vector<double> store;
  phe::NPairSmem sm;
  NBufs() : store(5 * KP + 6) {
    double* g = (double*)(((uintptr_t)store.data() + 15) & ~(uintptr_t)15);
    sm.xs0 = g; sm.x1 = g + KP; sm.y0 = g + 2 * KP; sm.y1 = g + 3 * KP; sm.e = reinterpret_cast<uint64_t*>(g + 4 * KP);
  }
};

template <int L, int TPI, int WIN>
int do_mul_npair(const uint32_t* c, int chunk_words, const uint32_t* e, int e_words, int e_stride, int ebits,
                 uint32_t* out, int count, const double* cst_e, uint64_t n0inv, uint64_t d_top) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy cst(cst_e, (size_t)phe::NE_COUNT * KP);
  const int cw = 2 * chunk_words;
  for (int i = 0; i < count; ++i) {
    NBufs<L, TPI> bufs;
    std::vector<double> tbl((size_t)(2 << WIN) * KP + 2);
    double* tp = (double*)(((uintptr_t)tbl.data() + 15) & ~(uintptr_t)15);
    run_group<TPI>([&] {
      phe::NPairPowmCtl<L, TPI, Env, WIN> ctl;
      ctl.c_w = c + (size_t)i * cw; ctl.chunk_words = chunk_words;
      ctl.e_w = e + (size_t)i * e_stride; ctl.e_words = e_words; ctl.ebits = ebits;
      ctl.out_w = out + (size_t)i * cw; ctl.out_words = cw; ctl.cst = cst.p; ctl.tbl = tp; ctl.sm = bufs.sm;
      phe::npair_run<L, TPI, Env>(ctl, cst.p, n0inv, d_top, bufs.sm);
    });
  }
  return 0;
}

// exponent alignment: out = c^(2^delta); every item runs max_delta squarings and keeps the value it had after its own
template <int L, int TPI>
int do_scale_npair(const uint32_t* c, int chunk_words, const int* delta, int max_delta, uint32_t* out, int count,
                   const double* cst_e, uint64_t n0inv, uint64_t d_top) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy cst(cst_e, (size_t)phe::NE_COUNT * KP);
  const int cw = 2 * chunk_words;
  for (int i = 0; i < count; ++i) {
    NBufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::NPairScaleCtl<L, TPI, Env> ctl;
      ctl.c_w = c + (size_t)i * cw; ctl.chunk_words = chunk_words;
      ctl.delta = delta[i]; ctl.max_delta = max_delta;
      ctl.out_w = out + (size_t)i * cw; ctl.out_words = cw; ctl.cst = cst.p; ctl.sm = bufs.sm;
      phe::npair_run<L, TPI, Env>(ctl, cst.p, n0inv, d_top, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI>
int do_powm_prog_npair(const uint32_t* c, int chunk_words, int nchunks, const uint32_t* prog, int nprog, uint32_t* out,
                       int out_words, int count, const double* cst_e, uint64_t n0inv, uint64_t d_top) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  constexpr int WS = 6;
  AlignedCopy cst(cst_e, (size_t)phe::NE_COUNT * KP);
  const int cw = nchunks * chunk_words;
  for (int i = 0; i < count; ++i) {
    NBufs<L, TPI> bufs;
    std::vector<double> tbl((size_t)(2 << (WS - 1)) * KP + 2);
    double* tp = (double*)(((uintptr_t)tbl.data() + 15) & ~(uintptr_t)15);
    run_group<TPI>([&] {
      phe::NPairProgCtl<L, TPI, Env, WS> ctl;
      ctl.c_w = c + (size_t)i * cw; ctl.chunk_words = chunk_words; ctl.nchunks = nchunks; ctl.prog = prog; ctl.nprog = nprog;
      ctl.out_w = out + (size_t)i * out_words; ctl.out_words = out_words; ctl.cst = cst.p; ctl.tbl = tp; ctl.sm = bufs.sm;
      phe::npair_run<L, TPI, Env>(ctl, cst.p, n0inv, d_top, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI>
int do_comb_npair(const uint32_t* hs, int chunk_words, int nwin, int wb, double* comb, const double* cst_e,
                  uint64_t n0inv, uint64_t d_top) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy cst(cst_e, (size_t)phe::NE_COUNT * KP);
  const size_t total = ((size_t)nwin << wb) * 2 * KP;
  std::vector<double> store(total + 2);
  double* cb = (double*)(((uintptr_t)store.data() + 15) & ~(uintptr_t)15);
  {
    NBufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::NPairCombBasesCtl<L, TPI, Env> ctl;
      ctl.hs_w = hs; ctl.chunk_words = chunk_words; ctl.nwin = nwin; ctl.wb = wb; ctl.comb = cb; ctl.writer = true;
      ctl.cst = cst.p; ctl.sm = bufs.sm;
      phe::npair_run<L, TPI, Env>(ctl, cst.p, n0inv, d_top, bufs.sm);
    });
  }
  for (int level = 1; level < wb; ++level)
    for (int j = 0; j < nwin; ++j)
      for (int e = 1; e < (1 << level); ++e) {
        NBufs<L, TPI> bufs;
        run_group<TPI>([&] {
          phe::NPairCombLevelCtl<L, TPI, Env> ctl;
          ctl.row = cb + (((size_t)j) << wb) * 2 * KP; ctl.level = level; ctl.e = e; ctl.store = true; ctl.sm = bufs.sm;
          phe::npair_run<L, TPI, Env>(ctl, cst.p, n0inv, d_top, bufs.sm);
        });
      }
  std::memcpy(comb, cb, total * 8);
  return 0;
}

template <int L, int TPI>
int do_encrypt_npair(const uint32_t* m, int m_words, const uint32_t* r, int r_words, int nwin, int wb, uint32_t* out,
                     int out_words, int count, const double* cst_e, uint64_t n0inv, uint64_t d_top, const double* comb,
                     size_t comb_doubles) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy cst(cst_e, (size_t)phe::NE_COUNT * KP), cb(comb, comb_doubles);
  // two more destinations of every row, as the fused gather of config 4 uses them (peer buffers): must equal `out`
  std::vector<uint32_t> peer_a((size_t)count * out_words, 0xdeadbeefu), peer_b((size_t)count * out_words, 0xdeadbeefu);
  uint32_t* peer_ptrs[2] = {peer_a.data(), peer_b.data()};
  for (int i = 0; i < count; ++i) {
    NBufs<L, TPI> bufs;
    run_group<TPI>([&] {
      phe::NPairEncCtl<L, TPI, Env> ctl;
      ctl.m_w = m + (size_t)i * m_words; ctl.m_words = m_words;
      ctl.r_w = r ? r + (size_t)i * r_words : nullptr; ctl.r_words = r_words; ctl.nwin = nwin; ctl.wb = wb;
      ctl.out_w = out + (size_t)i * out_words; ctl.out_words = out_words; ctl.cst = cst.p; ctl.comb = cb.p; ctl.sm = bufs.sm;
      ctl.peers = peer_ptrs; ctl.n_peers = 2; ctl.peer_off = (size_t)i * out_words;
      phe::npair_run<L, TPI, Env>(ctl, cst.p, n0inv, d_top, bufs.sm);
    });
  }
  if (std::memcmp(peer_a.data(), out, peer_a.size() * 4) || std::memcmp(peer_b.data(), out, peer_b.size() * 4)) return -3;
  return 0;
}

}  // namespace

#define DISPATCH_SHAPE(CALL)                       \
  switch (shape) {                                 \
    case 201: { constexpr int L = 20, TPI = 1; return CALL; } \
    case 202: { constexpr int L = 20, TPI = 2; return CALL; } \
    case 204: { constexpr int L = 20, TPI = 4; return CALL; } \
    case 208: { constexpr int L = 20, TPI = 8; return CALL; } \
    case 154: { constexpr int L = 15, TPI = 4; return CALL; } \
    case 158: { constexpr int L = 15, TPI = 8; return CALL; } \
    case 74:  { constexpr int L = 7, TPI = 4; return CALL; }  \
    default: return -1;                            \
  }

extern "C" {

int emu_modmul(int shape, const uint32_t* a, const uint32_t* b, uint32_t* out, int nwords, int count,
               const double* n_e, uint64_t n0inv, const double* r2_e) {
  DISPATCH_SHAPE((do_modmul<L, TPI>(a, b, out, nwords, count, n_e, n0inv, r2_e)));
}

int emu_modmul1(int shape, const uint32_t* a, const uint32_t* b_w, const double* b_e, uint32_t* out, int nwords, int count,
                const double* n_e, uint64_t n0inv) {
  DISPATCH_SHAPE((do_modmul1<L, TPI>(a, b_w, b_e, out, nwords, count, n_e, n0inv)));
}

int emu_scale_npair(int shape, const uint32_t* c, int chunk_words, const int* delta, int max_delta, uint32_t* out, int count,
                    const double* cst, uint64_t n0inv, uint64_t d_top) {
  DISPATCH_SHAPE((do_scale_npair<L, TPI>(c, chunk_words, delta, max_delta, out, count, cst, n0inv, d_top)));
}

int emu_powm(int shape, int win, const uint32_t* base, int base_words, const double* base_mont, const uint32_t* e,
             int e_words, int e_stride, int ebits, uint32_t* out, int out_words, int count, const double* n_e,
             uint64_t n0inv, const double* r2_e, const double* oneM_e, const double* one_e) {
#define POWM_CALL(W) do_powm<L, TPI, W>(base, base_words, base_mont, e, e_words, e_stride, ebits, out, out_words, count, n_e, n0inv, r2_e, oneM_e, one_e)
  if (win == 5) { DISPATCH_SHAPE((POWM_CALL(5))); }
  if (win == 3) { DISPATCH_SHAPE((POWM_CALL(3))); }
  if (win == 1) { DISPATCH_SHAPE((POWM_CALL(1))); }
  return -2;
}

int emu_powm_prog(int shape, const uint32_t* base, int base_words, const double* base_mont, const uint32_t* prog,
                  int nprog, uint32_t* out, int out_words, int count, const double* n_e, uint64_t n0inv,
                  const double* r2_e, const double* oneM_e, const double* one_e) {
  DISPATCH_SHAPE((do_powm_prog<L, TPI>(base, base_words, base_mont, prog, nprog, out, out_words, count, n_e, n0inv, r2_e, oneM_e, one_e)));
}

int emu_inv_block(int shape, int unwind, const uint32_t* c, int nwords, int count, int block, double* P,
                  uint32_t* totals, const uint32_t* tinv, uint32_t* out, const double* n_e, uint64_t n0inv,
                  const double* r2_e, const double* oneM_e, const double* one_e) {
  DISPATCH_SHAPE((do_inv_block<L, TPI>(unwind != 0, c, nwords, count, block, P, totals, tinv, out, n_e, n0inv, r2_e, oneM_e, one_e)));
}

int emu_dec_prep(int shape, const uint32_t* c, int hw, double* out_entries, int count, const double* n_e,
                 uint64_t n0inv, const double* r2_e, const double* k2_e) {
  DISPATCH_SHAPE((do_dec_prep<L, TPI>(c, hw, out_entries, count, n_e, n0inv, r2_e, k2_e)));
}

int emu_encrypt_comb(int shape, const uint32_t* m, int m_words, const uint32_t* r, int r_words, int nwin, int wb,
                     uint32_t* out, int out_words, int count, const double* n_e, uint64_t n0inv,
                     const double* nR_e, const double* comb, uint64_t comb_doubles) {
  DISPATCH_SHAPE((do_encrypt_comb<L, TPI>(m, m_words, r, r_words, nwin, wb, out, out_words, count, n_e, n0inv, nR_e, comb, comb_doubles)));
}

int emu_encrypt_finish(int shape, const uint32_t* m, int m_words, const uint32_t* obf, uint32_t* out, int out_words,
                       int count, const double* n_e, uint64_t n0inv, const double* nR_e, const double* r2_e) {
  DISPATCH_SHAPE((do_encrypt_finish<L, TPI>(m, m_words, obf, out, out_words, count, n_e, n0inv, nR_e, r2_e)));
}

int emu_dec_pair(int L, const uint32_t* c, int c_words, int chunk_words, const uint32_t* prog, uint32_t* out,
                 int out_words, int count, const double* mod, uint64_t n0inv, const double* cst, int slots) {
  if (L == 10) return do_dec_pair<10>(c, c_words, chunk_words, prog, out, out_words, count, mod, n0inv, cst, slots);
  if (L == 20) return do_dec_pair<20>(c, c_words, chunk_words, prog, out, out_words, count, mod, n0inv, cst, slots);
  if (L == 30) return do_dec_pair<30>(c, c_words, chunk_words, prog, out, out_words, count, mod, n0inv, cst, slots);
  return -1;
}

int emu_dec_pair_segments(int L, const uint32_t* c, int c_words, int chunk_words, const uint32_t* prog, int nseg,
                          const int* seg_off, uint32_t* out, int out_words, int count, const double* mod, uint64_t n0inv,
                          const double* cst, int slots) {
  if (L == 10) return do_dec_pair<10>(c, c_words, chunk_words, prog, out, out_words, count, mod, n0inv, cst, slots, nseg, seg_off);
  if (L == 20) return do_dec_pair<20>(c, c_words, chunk_words, prog, out, out_words, count, mod, n0inv, cst, slots, nseg, seg_off);
  if (L == 30) return do_dec_pair<30>(c, c_words, chunk_words, prog, out, out_words, count, mod, n0inv, cst, slots, nseg, seg_off);
  return -1;
}

int emu_dec_crt(int shape, const uint32_t* mp, const uint32_t* mq, int half_words, uint32_t* m, int m_words, int count,
                const double* cst, const uint64_t* n0invs) {
  DISPATCH_SHAPE((do_dec_crt<L, TPI>(mp, mq, half_words, m, m_words, count, cst, n0invs)));
}

int emu_dec_tail(int shape, const uint32_t* up, const uint32_t* uq, int u_words, uint32_t* m, int m_words, int count,
                 const double* cst, const uint64_t* n0invs) {
  DISPATCH_SHAPE((do_dec_tail<L, TPI>(up, uq, u_words, m, m_words, count, cst, n0invs)));
}
int emu_mul_npair(int shape, int win, const uint32_t* c, int chunk_words, const uint32_t* e, int e_words, int e_stride,
                  int ebits, uint32_t* out, int count, const double* cst, uint64_t n0inv, uint64_t d_top) {
#define NPM_CALL(W) do_mul_npair<L, TPI, W>(c, chunk_words, e, e_words, e_stride, ebits, out, count, cst, n0inv, d_top)
  if (win == 5) { DISPATCH_SHAPE((NPM_CALL(5))); }
  if (win == 3) { DISPATCH_SHAPE((NPM_CALL(3))); }
  if (win == 1) { DISPATCH_SHAPE((NPM_CALL(1))); }
  return -2;
}

int emu_powm_prog_npair(int shape, const uint32_t* c, int chunk_words, int nchunks, const uint32_t* prog, int nprog,
                        uint32_t* out, int out_words, int count, const double* cst, uint64_t n0inv, uint64_t d_top) {
  DISPATCH_SHAPE((do_powm_prog_npair<L, TPI>(c, chunk_words, nchunks, prog, nprog, out, out_words, count, cst, n0inv, d_top)));
}

int emu_comb_npair(int shape, const uint32_t* hs, int chunk_words, int nwin, int wb, double* comb, const double* cst,
                   uint64_t n0inv, uint64_t d_top) {
  DISPATCH_SHAPE((do_comb_npair<L, TPI>(hs, chunk_words, nwin, wb, comb, cst, n0inv, d_top)));
}

int emu_encrypt_npair(int shape, const uint32_t* m, int m_words, const uint32_t* r, int r_words, int nwin, int wb,
                      uint32_t* out, int out_words, int count, const double* cst, uint64_t n0inv, uint64_t d_top,
                      const double* comb, uint64_t comb_doubles) {
  DISPATCH_SHAPE((do_encrypt_npair<L, TPI>(m, m_words, r, r_words, nwin, wb, out, out_words, count, cst, n0inv, d_top, comb, comb_doubles)));
}
}
