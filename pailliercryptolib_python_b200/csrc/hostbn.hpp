// hostbn.hpp -- small host-side big integer used ONLY for per-key setup (Montgomery constants, CRT constants,
// key generation).  Not on the data path: everything per-element runs in the CUDA kernels.
// Replaces the parts of the IPP BigNumber class the reference's key constructors rely on
// (ipcl::PublicKey / ipcl::PrivateKey ctors reached from ipcl_bindings_classes.cpp:16-27, 96-101).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace hbn {

struct BN {
  std::vector<uint32_t> w;  // little-endian, no leading zero words (zero = empty)

  BN() {}
  explicit BN(uint64_t v) { while (v) { w.push_back((uint32_t)v); v >>= 32; } }
  static BN from_words(const uint32_t* p, size_t n) { BN r; r.w.assign(p, p + n); r.trim(); return r; }
  void trim() { while (!w.empty() && w.back() == 0) w.pop_back(); }
  bool is_zero() const { return w.empty(); }
  bool is_odd() const { return !w.empty() && (w[0] & 1u); }
  size_t bits() const {
    if (w.empty()) return 0;
    return 32 * (w.size() - 1) + (32 - (size_t)__builtin_clz(w.back()));
  }
  bool bit(size_t i) const { return (i >> 5) < w.size() && ((w[i >> 5] >> (i & 31)) & 1u); }
  void to_words(uint32_t* out, size_t n) const {
    if (w.size() > n) throw std::runtime_error("BN::to_words: value does not fit");
    std::memset(out, 0, n * 4);
    if (!w.empty()) std::memcpy(out, w.data(), w.size() * 4);
  }
  uint32_t low() const { return w.empty() ? 0u : w[0]; }
};

inline int cmp(const BN& a, const BN& b) {
  if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
  for (size_t i = a.w.size(); i-- > 0;)
    if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
  return 0;
}
inline bool operator==(const BN& a, const BN& b) { return cmp(a, b) == 0; }
inline bool operator<(const BN& a, const BN& b) { return cmp(a, b) < 0; }

inline BN add(const BN& a, const BN& b) {
  BN r; const size_t n = std::max(a.w.size(), b.w.size());
  r.w.resize(n + 1);
  uint64_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c += (uint64_t)(i < a.w.size() ? a.w[i] : 0) + (i < b.w.size() ? b.w[i] : 0);
    r.w[i] = (uint32_t)c; c >>= 32;
  }
  r.w[n] = (uint32_t)c; r.trim(); return r;
}

// a - b, requires a >= b
inline BN sub(const BN& a, const BN& b) {
  if (cmp(a, b) < 0) throw std::runtime_error("BN::sub: negative result");
  BN r; r.w.resize(a.w.size());
  int64_t c = 0;
  for (size_t i = 0; i < a.w.size(); ++i) {
    c += (int64_t)a.w[i] - (i < b.w.size() ? b.w[i] : 0);
    r.w[i] = (uint32_t)c; c >>= 32;
  }
  r.trim(); return r;
}

inline BN mul(const BN& a, const BN& b) {
  BN r; if (a.is_zero() || b.is_zero()) return r;
  r.w.assign(a.w.size() + b.w.size(), 0);
  for (size_t i = 0; i < a.w.size(); ++i) {
    uint64_t c = 0; const uint64_t ai = a.w[i];
    for (size_t j = 0; j < b.w.size(); ++j) {
      c += ai * b.w[j] + r.w[i + j];
      r.w[i + j] = (uint32_t)c; c >>= 32;
    }
    r.w[i + b.w.size()] = (uint32_t)c;
  }
  r.trim(); return r;
}

inline BN shl(const BN& a, size_t s) {
  BN r; if (a.is_zero()) return r;
  const size_t ws = s >> 5, bs = s & 31;
  r.w.assign(a.w.size() + ws + 1, 0);
  for (size_t i = 0; i < a.w.size(); ++i) {
    r.w[i + ws] |= a.w[i] << bs;
    if (bs) r.w[i + ws + 1] |= a.w[i] >> (32 - bs);
  }
  r.trim(); return r;
}

inline BN shr(const BN& a, size_t s) {
  BN r; const size_t ws = s >> 5, bs = s & 31;
  if (ws >= a.w.size()) return r;
  r.w.assign(a.w.size() - ws, 0);
  for (size_t i = ws; i < a.w.size(); ++i) {
    r.w[i - ws] = a.w[i] >> bs;
    if (bs && i + 1 < a.w.size()) r.w[i - ws] |= a.w[i + 1] << (32 - bs);
  }
  r.trim(); return r;
}

// Knuth algorithm D.  q = a / b, r = a % b.
inline void divmod(const BN& a, const BN& b, BN* q, BN* r) {
  if (b.is_zero()) throw std::runtime_error("BN::divmod: division by zero");
  if (cmp(a, b) < 0) { if (q) *q = BN(); if (r) *r = a; return; }
  if (b.w.size() == 1) {
    BN qq; qq.w.resize(a.w.size()); uint64_t rem = 0; const uint64_t d = b.w[0];
    for (size_t i = a.w.size(); i-- > 0;) { rem = (rem << 32) | a.w[i]; qq.w[i] = (uint32_t)(rem / d); rem %= d; }
    qq.trim(); if (q) *q = qq; if (r) *r = BN(rem); return;
  }
  const int s = __builtin_clz(b.w.back());
  const BN v = shl(b, s); BN u = shl(a, s);
  const size_t n = v.w.size();
  u.w.resize(a.w.size() + 1, 0);
  const size_t m = u.w.size() - n - 1;
  BN qq; qq.w.assign(m + 1, 0);
  const uint64_t B = 1ull << 32;
  for (size_t j = m + 1; j-- > 0;) {
    const uint64_t num = ((uint64_t)u.w[j + n] << 32) | u.w[j + n - 1];
    uint64_t qhat = num / v.w[n - 1], rhat = num % v.w[n - 1];
    while (qhat >= B || qhat * v.w[n - 2] > ((rhat << 32) | u.w[j + n - 2])) {
      --qhat; rhat += v.w[n - 1];
      if (rhat >= B) break;
    }
    int64_t borrow = 0; uint64_t carry = 0;
    for (size_t i = 0; i < n; ++i) {
      const uint64_t p = qhat * v.w[i] + carry; carry = p >> 32;
      const int64_t t = (int64_t)u.w[i + j] - borrow - (int64_t)(p & 0xffffffffu);
      u.w[i + j] = (uint32_t)t; borrow = (t < 0) ? 1 : 0;
    }
    const int64_t t = (int64_t)u.w[j + n] - borrow - (int64_t)carry;
    u.w[j + n] = (uint32_t)t;
    if (t < 0) {
      --qhat; uint64_t c = 0;
      for (size_t i = 0; i < n; ++i) { c += (uint64_t)u.w[i + j] + v.w[i]; u.w[i + j] = (uint32_t)c; c >>= 32; }
      u.w[j + n] += (uint32_t)c;
    }
    qq.w[j] = (uint32_t)qhat;
  }
  qq.trim(); if (q) *q = qq;
  if (r) { u.trim(); *r = shr(u, s); }
}

inline BN mod(const BN& a, const BN& m) { BN r; divmod(a, m, nullptr, &r); return r; }
inline BN div(const BN& a, const BN& m) { BN q; divmod(a, m, &q, nullptr); return q; }
inline BN mulmod(const BN& a, const BN& b, const BN& m) { return mod(mul(a, b), m); }

// Montgomery (64-bit limbs) modexp for odd moduli; falls back to mul+div for even ones.
inline BN modexp(const BN& base, const BN& e, const BN& m) {
  if (m.is_zero()) throw std::runtime_error("BN::modexp: zero modulus");
  if (m == BN(1)) return BN();
  BN result(1);
  if (!m.is_odd()) {
    BN b = mod(base, m);
    for (size_t i = e.bits(); i-- > 0;) { result = mulmod(result, result, m); if (e.bit(i)) result = mulmod(result, b, m); }
    return result;
  }
  typedef unsigned __int128 u128;
  const size_t k = (m.w.size() + 1) / 2;
  std::vector<uint64_t> n(k, 0);
  for (size_t i = 0; i < m.w.size(); ++i) n[i / 2] |= (uint64_t)m.w[i] << (32 * (i & 1));
  uint64_t n0 = 1; for (int i = 0; i < 6; ++i) n0 *= 2 - n[0] * n0; n0 = ~n0 + 1;  // -n^-1 mod 2^64
  auto to64 = [&](const BN& x) { std::vector<uint64_t> v(k, 0); for (size_t i = 0; i < x.w.size(); ++i) v[i / 2] |= (uint64_t)x.w[i] << (32 * (i & 1)); return v; };
  auto montmul = [&](const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
    std::vector<uint64_t> t(k + 2, 0);
    for (size_t i = 0; i < k; ++i) {
      u128 c = 0;
      for (size_t j = 0; j < k; ++j) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[k]; t[k] = (uint64_t)c; t[k + 1] = (uint64_t)(c >> 64);
      const uint64_t q = t[0] * n0;
      c = (u128)q * n[0] + t[0]; c >>= 64;
      for (size_t j = 1; j < k; ++j) { c += (u128)q * n[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[k]; t[k - 1] = (uint64_t)c; t[k] = t[k + 1] + (uint64_t)(c >> 64);
    }
    bool ge = t[k] != 0;
    if (!ge) { ge = true; for (size_t i = k; i-- > 0;) if (t[i] != n[i]) { ge = t[i] > n[i]; break; } }
    if (ge) { uint64_t bw = 0; for (size_t i = 0; i < k; ++i) { const u128 d = (u128)t[i] - n[i] - bw; t[i] = (uint64_t)d; bw = (uint64_t)(d >> 64) & 1; } }
    t.resize(k); return t;
  };
  const BN R = shl(BN(1), 64 * k);
  const std::vector<uint64_t> r2 = to64(mod(mul(R, R), m));
  std::vector<uint64_t> one(k, 0); one[0] = 1;
  const std::vector<uint64_t> bM = montmul(to64(mod(base, m)), r2);
  std::vector<uint64_t> acc = montmul(one, r2);
  for (size_t i = e.bits(); i-- > 0;) { acc = montmul(acc, acc); if (e.bit(i)) acc = montmul(acc, bM); }
  acc = montmul(acc, one);
  BN out; out.w.resize(2 * k);
  for (size_t i = 0; i < k; ++i) { out.w[2 * i] = (uint32_t)acc[i]; out.w[2 * i + 1] = (uint32_t)(acc[i] >> 32); }
  out.trim(); return out;
}

inline BN gcd(BN a, BN b) { while (!b.is_zero()) { BN r = mod(a, b); a = b; b = r; } return a; }

// inverse of a modulo a PRIME p (Fermat)
inline BN modinv_prime(const BN& a, const BN& p) { return modexp(a, sub(p, BN(2)), p); }

// a^-1 mod m for any m > 1 (extended Euclid); throws if gcd(a, m) != 1
inline BN modinv(const BN& a, const BN& m) {
  BN r0 = m, r1 = mod(a, m);
  BN t0, t1(1);          // coefficients of a, kept as magnitudes with signs
  bool s0 = false, s1 = false;
  while (!r1.is_zero()) {
    BN q, r2;
    divmod(r0, r1, &q, &r2);
    // t2 = t0 - q * t1
    const BN qt = mul(q, t1);
    BN t2; bool s2;
    if (s0 == s1) {       // same sign: t0 - qt
      if (cmp(t0, qt) >= 0) { t2 = sub(t0, qt); s2 = s0; } else { t2 = sub(qt, t0); s2 = !s0; }
    } else {              // opposite signs: magnitudes add, sign of t0
      t2 = add(t0, qt); s2 = s0;
    }
    r0 = r1; r1 = r2; t0 = t1; s0 = s1; t1 = t2; s1 = s2;
  }
  if (!(r0 == BN(1))) throw std::runtime_error("BN::modinv: not invertible");
  BN t = mod(t0, m);
  if (s0 && !t.is_zero()) t = sub(m, t);
  return t;
}

// -n^-1 mod 2^28 for odd n
inline uint32_t neg_inv28(uint32_t n0) {
  uint32_t x = 1; for (int i = 0; i < 5; ++i) x *= 2u - n0 * x;
  return (0u - x) & ((1u << 28) - 1u);
}

}  // namespace hbn
