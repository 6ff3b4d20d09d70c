// Host lockstep emulator for the lane-group code in csrc/paillier_items.cuh (TEST INFRASTRUCTURE).
// Each lane of a group is a std::thread; shuffles are barrier-synchronised exchanges.
// Built by tests/test_emu.py: g++ -O2 -std=c++20 -pthread -shared -fPIC.
#include <barrier>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../../pailliercryptolib_python_b200/csrc/paillier_items.cuh"

namespace {

struct Exchange {
  uint32_t slot[32];
  std::barrier<> bar;
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
  static void sync() { if (TPI > 1) t_ex->bar.arrive_and_wait(); }
};

template <int TPI> void run_group(const std::function<void()>& body) {
  Exchange ex(TPI);
  std::vector<std::thread> th;
  for (int l = 0; l < TPI; ++l)
    th.emplace_back([&, l] { t_lane = l; t_ex = &ex; body(); });
  for (auto& t : th) t.join();
}

template <int L, int TPI> struct Bufs {
  std::vector<uint32_t> b0, b1;
  phe::GroupSmem sm;
  Bufs() : b0(phe::Shape<L, TPI>::KP + 4), b1(phe::Shape<L, TPI>::KP + 4) {
    // 16-byte align
    auto al = [](std::vector<uint32_t>& v) { uintptr_t p = (uintptr_t)v.data(); return (uint32_t*)((p + 15) & ~(uintptr_t)15); };
    sm.b0 = al(b0); sm.b1 = al(b1);
  }
};

template <int L, int TPI, class F> void with_mod(const uint32_t* n_entry, F f) {
  using Env = EmuEnv<TPI>;
  run_group<TPI>([&] {
    uint32_t n[L];
    phe::load_entry<L, TPI, Env>(n, n_entry);
    f(n);
  });
}

// Aligned copy of caller memory (ctypes numpy buffers are 16B-aligned in practice, but be safe)
struct AlignedCopy {
  std::vector<uint32_t> store; uint32_t* p;
  AlignedCopy(const uint32_t* src, size_t n) : store(n + 4) {
    uintptr_t a = ((uintptr_t)store.data() + 15) & ~(uintptr_t)15; p = (uint32_t*)a; std::memcpy(p, src, n * 4);
  }
};

template <int L, int TPI>
int do_modmul(const uint32_t* a, const uint32_t* b, uint32_t* out, int nwords, int count, const uint32_t* n_e,
              uint32_t n0inv, const uint32_t* r2_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    with_mod<L, TPI>(ne.p, [&](uint32_t (&n)[L]) {
      phe::item_modmul<L, TPI, Env>(a + (size_t)i * nwords, b + (size_t)i * nwords, out + (size_t)i * nwords, nwords,
                                    n, n0inv, r2.p, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI, int WIN>
int do_powm(const uint32_t* base, int base_words, const uint32_t* base_mont, const uint32_t* e, int e_words,
            int e_stride, int ebits, uint32_t* out, int out_words, int count, const uint32_t* n_e, uint32_t n0inv,
            const uint32_t* r2_e, const uint32_t* oneM_e, const uint32_t* one_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP), oneM(oneM_e, KP), one(one_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    std::vector<uint32_t> tbl((size_t)(1 << WIN) * KP + 4);
    uint32_t* tp = (uint32_t*)(((uintptr_t)tbl.data() + 15) & ~(uintptr_t)15);
    std::vector<uint32_t> bm;
    const uint32_t* bmp = nullptr;
    AlignedCopy* bmc = nullptr;
    if (base_mont) { bmc = new AlignedCopy(base_mont + (size_t)i * KP, KP); bmp = bmc->p; }
    with_mod<L, TPI>(ne.p, [&](uint32_t (&n)[L]) {
      phe::item_powm<L, TPI, Env, WIN>(base ? base + (size_t)i * base_words : nullptr, base_words, bmp,
                                       e + (size_t)i * e_stride, e_words, ebits, out + (size_t)i * out_words,
                                       out_words, n, n0inv, r2.p, oneM.p, one.p, tp, bufs.sm);
    });
    delete bmc;
  }
  return 0;
}

template <int L, int TPI>
int do_dec_prep(const uint32_t* c, int hw, uint32_t* out_entries, int count, const uint32_t* n_e, uint32_t n0inv,
                const uint32_t* r2_e, const uint32_t* k2_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), r2(r2_e, KP), k2(k2_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    with_mod<L, TPI>(ne.p, [&](uint32_t (&n)[L]) {
      phe::item_dec_prep<L, TPI, Env>(c + (size_t)i * 2 * hw, hw, out_entries + (size_t)i * KP, n, n0inv, r2.p, k2.p,
                                      bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI, int WB>
int do_encrypt_comb(const uint32_t* m, int m_words, const uint32_t* r, int r_words, int nwin, uint32_t* out,
                    int out_words, int count, const uint32_t* n_e, uint32_t n0inv, const uint32_t* nR_e,
                    const uint32_t* comb, size_t comb_words) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), nR(nR_e, KP), cb(comb, comb_words);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    with_mod<L, TPI>(ne.p, [&](uint32_t (&n)[L]) {
      phe::item_encrypt_comb<L, TPI, Env, WB>(m + (size_t)i * m_words, m_words, r ? r + (size_t)i * r_words : nullptr,
                                              r_words, nwin, out + (size_t)i * out_words, out_words, n, n0inv, nR.p,
                                              cb.p, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI>
int do_encrypt_finish(const uint32_t* m, int m_words, const uint32_t* obf, uint32_t* out, int out_words, int count,
                      const uint32_t* n_e, uint32_t n0inv, const uint32_t* nR_e, const uint32_t* r2_e) {
  using Env = EmuEnv<TPI>;
  constexpr int KP = phe::Shape<L, TPI>::KP;
  AlignedCopy ne(n_e, KP), nR(nR_e, KP), r2(r2_e, KP);
  for (int i = 0; i < count; ++i) {
    Bufs<L, TPI> bufs;
    with_mod<L, TPI>(ne.p, [&](uint32_t (&n)[L]) {
      phe::item_encrypt_finish<L, TPI, Env>(m + (size_t)i * m_words, m_words, obf + (size_t)i * out_words,
                                            out + (size_t)i * out_words, out_words, n, n0inv, nR.p, r2.p, bufs.sm);
    });
  }
  return 0;
}

template <int L, int TPI>
int do_dec_tail(const uint32_t* up, const uint32_t* uq, int u_words, uint32_t* m, int m_words, int count,
                const uint32_t* cst, const uint32_t* n0invs) {
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

}  // namespace

#define DISPATCH_SHAPE(CALL)                       \
  switch (shape) {                                 \
    case 371: { constexpr int L = 37, TPI = 1; return CALL; } \
    case 372: { constexpr int L = 37, TPI = 2; return CALL; } \
    case 374: { constexpr int L = 37, TPI = 4; return CALL; } \
    case 198: { constexpr int L = 19, TPI = 8; return CALL; } \
    case 194: { constexpr int L = 19, TPI = 4; return CALL; } \
    case 282: { constexpr int L = 28, TPI = 2; return CALL; } \
    case 284: { constexpr int L = 28, TPI = 4; return CALL; } \
    case 288: { constexpr int L = 28, TPI = 8; return CALL; } \
    default: return -1;                            \
  }

extern "C" {

int emu_modmul(int shape, const uint32_t* a, const uint32_t* b, uint32_t* out, int nwords, int count,
               const uint32_t* n_e, uint32_t n0inv, const uint32_t* r2_e) {
  DISPATCH_SHAPE((do_modmul<L, TPI>(a, b, out, nwords, count, n_e, n0inv, r2_e)));
}

int emu_powm(int shape, int win, const uint32_t* base, int base_words, const uint32_t* base_mont, const uint32_t* e,
             int e_words, int e_stride, int ebits, uint32_t* out, int out_words, int count, const uint32_t* n_e,
             uint32_t n0inv, const uint32_t* r2_e, const uint32_t* oneM_e, const uint32_t* one_e) {
#define POWM_CALL(W) do_powm<L, TPI, W>(base, base_words, base_mont, e, e_words, e_stride, ebits, out, out_words, count, n_e, n0inv, r2_e, oneM_e, one_e)
  if (win == 5) { DISPATCH_SHAPE((POWM_CALL(5))); }
  if (win == 3) { DISPATCH_SHAPE((POWM_CALL(3))); }
  if (win == 1) { DISPATCH_SHAPE((POWM_CALL(1))); }
  return -2;
}

int emu_dec_prep(int shape, const uint32_t* c, int hw, uint32_t* out_entries, int count, const uint32_t* n_e,
                 uint32_t n0inv, const uint32_t* r2_e, const uint32_t* k2_e) {
  DISPATCH_SHAPE((do_dec_prep<L, TPI>(c, hw, out_entries, count, n_e, n0inv, r2_e, k2_e)));
}

int emu_encrypt_comb(int shape, const uint32_t* m, int m_words, const uint32_t* r, int r_words, int nwin,
                     uint32_t* out, int out_words, int count, const uint32_t* n_e, uint32_t n0inv,
                     const uint32_t* nR_e, const uint32_t* comb, uint64_t comb_words) {
  DISPATCH_SHAPE((do_encrypt_comb<L, TPI, 8>(m, m_words, r, r_words, nwin, out, out_words, count, n_e, n0inv, nR_e, comb, comb_words)));
}

int emu_encrypt_finish(int shape, const uint32_t* m, int m_words, const uint32_t* obf, uint32_t* out, int out_words,
                       int count, const uint32_t* n_e, uint32_t n0inv, const uint32_t* nR_e, const uint32_t* r2_e) {
  DISPATCH_SHAPE((do_encrypt_finish<L, TPI>(m, m_words, obf, out, out_words, count, n_e, n0inv, nR_e, r2_e)));
}

int emu_dec_tail(int shape, const uint32_t* up, const uint32_t* uq, int u_words, uint32_t* m, int m_words, int count,
                 const uint32_t* cst, const uint32_t* n0invs) {
  DISPATCH_SHAPE((do_dec_tail<L, TPI>(up, uq, u_words, m, m_words, count, cst, n0invs)));
}
}
