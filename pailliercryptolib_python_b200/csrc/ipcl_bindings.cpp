// ipcl_bindings.cpp -- pybind11 module `ipcl_bindings`: the reference's binding surface
// (/root/reference/src/ipcl_python/bindings/ipcl_bindings.cpp:21-63, ipcl_bindings_classes.cpp) re-implemented as a thin
// shim over the C ABI of libphe_b200.so (include/phe_b200.h).  No arithmetic on the hot path happens here: containers
// hold packed little-endian uint32 limb arrays and every encrypt / decrypt / + / * is one C-ABI call with the GIL
// released.  The small BigNumber class only serves the scalar conveniences the reference exposes on ipclBigNumber.
//
// Differences from the reference, on purpose:
//   * containers are packed arrays (fixed stride), not vector<BigNumber>; from_packed()/to_packed() move whole
//     batches to and from numpy without per-element Python objects;
//   * numpy-array constructors copy the caller's buffer instead of delete[]-ing it (ipcl_bindings_classes.cpp:182,287 UB);
//   * the bytes constructor never writes into the immutable bytes object (ipcl_bindings.cpp:109-116 UB);
//   * context.initializeContext selects/validates the CUDA device instead of starting Intel QAT; hybridMode is kept
//     as an inert setting.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "phe_b200.h"

namespace py = pybind11;

namespace {

[[noreturn]] void throw_phe(const char* what) {
  throw std::runtime_error(std::string(what) + ": " + phe_last_error());
}

// ------------------------------------------------------------------------------------------------ BigNumber
// Sign-magnitude integer, little-endian 32-bit words, at least one word (IppsBigNumState semantics).
struct BigNumber {
  std::vector<uint32_t> w{0};
  bool neg = false;

  BigNumber() {}
  explicit BigNumber(uint32_t v) : w{v} {}
  BigNumber(const uint32_t* p, size_t n) : w(p, p + (n ? n : 0)) { if (w.empty()) w.push_back(0); trim(); }
  void trim() {
    while (w.size() > 1 && w.back() == 0) w.pop_back();
    if (w.size() == 1 && w[0] == 0) neg = false;
  }
  bool is_zero() const { return w.size() == 1 && w[0] == 0; }
  int bit_size() const {
    if (is_zero()) return 1;  // IPP reports one bit for zero
    return 32 * (int)(w.size() - 1) + (32 - __builtin_clz(w.back()));
  }
  int dword_size() const { return (int)w.size(); }
};

int cmp_mag(const BigNumber& a, const BigNumber& b) {
  if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
  for (size_t i = a.w.size(); i-- > 0;)
    if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
  return 0;
}
int cmp(const BigNumber& a, const BigNumber& b) {
  if (a.neg != b.neg) return a.neg ? -1 : 1;
  const int c = cmp_mag(a, b);
  return a.neg ? -c : c;
}
BigNumber add_mag(const BigNumber& a, const BigNumber& b) {
  BigNumber r;
  const size_t n = std::max(a.w.size(), b.w.size());
  r.w.assign(n + 1, 0);
  uint64_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c += (uint64_t)(i < a.w.size() ? a.w[i] : 0) + (i < b.w.size() ? b.w[i] : 0);
    r.w[i] = (uint32_t)c;
    c >>= 32;
  }
  r.w[n] = (uint32_t)c;
  r.trim();
  return r;
}
BigNumber sub_mag(const BigNumber& a, const BigNumber& b) {  // |a| >= |b|
  BigNumber r;
  r.w.assign(a.w.size(), 0);
  int64_t c = 0;
  for (size_t i = 0; i < a.w.size(); ++i) {
    int64_t v = (int64_t)a.w[i] - (i < b.w.size() ? b.w[i] : 0) + c;
    c = v < 0 ? -1 : 0;
    r.w[i] = (uint32_t)(v & 0xffffffffll);
  }
  r.trim();
  return r;
}
BigNumber add(const BigNumber& a, const BigNumber& b) {
  BigNumber r;
  if (a.neg == b.neg) { r = add_mag(a, b); r.neg = a.neg; }
  else if (cmp_mag(a, b) >= 0) { r = sub_mag(a, b); r.neg = a.neg; }
  else { r = sub_mag(b, a); r.neg = b.neg; }
  r.trim();
  return r;
}
BigNumber negate(BigNumber a) { a.neg = !a.neg; a.trim(); return a; }
BigNumber mul(const BigNumber& a, const BigNumber& b) {
  BigNumber r;
  r.w.assign(a.w.size() + b.w.size(), 0);
  for (size_t i = 0; i < a.w.size(); ++i) {
    uint64_t c = 0;
    for (size_t j = 0; j < b.w.size(); ++j) {
      c += (uint64_t)a.w[i] * b.w[j] + r.w[i + j];
      r.w[i + j] = (uint32_t)c;
      c >>= 32;
    }
    r.w[i + b.w.size()] = (uint32_t)c;
  }
  r.neg = a.neg != b.neg;
  r.trim();
  return r;
}
std::string to_decimal(const BigNumber& v) {
  if (v.is_zero()) return "0";
  std::vector<uint32_t> t(v.w);
  std::string out;
  while (!(t.size() == 1 && t[0] == 0)) {
    uint64_t rem = 0;
    for (size_t i = t.size(); i-- > 0;) {
      const uint64_t cur = (rem << 32) | t[i];
      t[i] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
    }
    while (t.size() > 1 && t.back() == 0) t.pop_back();
    char buf[16];
    snprintf(buf, sizeof buf, "%09u", (unsigned)rem);
    out.insert(0, buf);
  }
  const size_t nz = out.find_first_not_of('0');
  out = out.substr(nz);
  return v.neg ? "-" + out : out;
}
std::string to_hex(const uint32_t* w, size_t n) {  // "0x" + upper-case hex, most significant word first (BigNumber::num2hex)
  static const char* d = "0123456789ABCDEF";
  std::string s = "0x";
  for (size_t i = n; i-- > 0;)
    for (int sh = 28; sh >= 0; sh -= 4) s.push_back(d[(w[i] >> sh) & 15]);
  return s;
}
// pyByte2BN (ipcl_bindings.cpp:100-119): little-endian bytes, any length, zero padded to whole words
BigNumber from_bytes(const std::string& b) {
  std::vector<uint32_t> w((b.size() + 3) / 4, 0);
  if (!b.empty()) std::memcpy(w.data(), b.data(), b.size());
  return BigNumber(w.data(), w.size());
}
// BN2bytes (ipcl_bindings.cpp:121-129): BITSIZE_WORD(bitlen) * 4 bytes
py::bytes to_bytes(const BigNumber& v) {
  return py::bytes(reinterpret_cast<const char*>(v.w.data()), v.w.size() * 4);
}
py::bytes words_to_bytes(const uint32_t* w, size_t n) {
  while (n > 1 && w[n - 1] == 0) --n;
  return py::bytes(reinterpret_cast<const char*>(w), n * 4);
}
std::string addr_tag(const void* p) {
  std::stringstream ss;
  ss << p;
  return std::to_string(std::hash<std::string>{}(ss.str())).substr(0, 10);
}

// ------------------------------------------------------------------------------------------------ keys
struct PublicKey {
  phe_pubkey* h = nullptr;
  int bits = 0, n_words = 0;
  BigNumber n;
  PublicKey(const BigNumber& n_, int bits_, bool djn, const BigNumber* hs = nullptr, int randbits = 0) {
    if (n_.neg || n_.is_zero()) throw std::runtime_error("ipclPublicKey: n must be positive");
    n = n_;
    bits = std::max(bits_, n.bit_size());
    n_words = (bits + 31) / 32;
    std::vector<uint32_t> nw(n_words, 0), hsw;
    std::copy(n.w.begin(), n.w.end(), nw.begin());
    if (hs) {
      if ((int)hs->w.size() > 2 * n_words) throw std::runtime_error("ipclPublicKey: hs larger than n^2");
      hsw.assign(2 * (size_t)n_words, 0);
      std::copy(hs->w.begin(), hs->w.end(), hsw.begin());
    }
    if (phe_pubkey_create(nw.data(), n_words, bits, djn ? 1 : 0, hs ? hsw.data() : nullptr, randbits, &h))
      throw_phe("ipclPublicKey");
  }
  PublicKey(const PublicKey&) = delete;
  ~PublicKey() { phe_pubkey_destroy(h); }
  bool djn() const { return phe_pubkey_is_djn(h) == 1; }
  int randbits() const { return phe_pubkey_randbits(h); }
  BigNumber nsquare() const {
    std::vector<uint32_t> v(2 * (size_t)n_words);
    phe_pubkey_get_nsquare(h, v.data());
    return BigNumber(v.data(), v.size());
  }
  BigNumber hs() const {
    std::vector<uint32_t> v(2 * (size_t)n_words);
    phe_pubkey_get_hs(h, v.data());
    return BigNumber(v.data(), v.size());
  }
};
using PubPtr = std::shared_ptr<PublicKey>;

py::tuple pubkey_state(const PublicKey& pk) {  // getTupleIpclPubKey (ipcl_bindings.cpp:66-81)
  if (pk.djn()) return py::make_tuple(1, to_bytes(pk.n), pk.bits, to_bytes(pk.hs()), pk.randbits());
  return py::make_tuple(0, to_bytes(pk.n), pk.bits, 0, 0);
}
PubPtr pubkey_from_state(const py::tuple& t) {  // setIpclPubKey (ipcl_bindings.cpp:83-98)
  const int scheme = t[0].cast<int>();
  const BigNumber n = from_bytes(t[1].cast<std::string>());
  const int bits = t[2].cast<int>();
  if (scheme == 0) return std::make_shared<PublicKey>(n, bits, false);
  const BigNumber hs = from_bytes(t[3].cast<std::string>());
  return std::make_shared<PublicKey>(n, bits, true, &hs, t[4].cast<int>());
}

// ------------------------------------------------------------------------------------------------ containers
// ipcl::BaseText: `count` numbers, `stride` words each, packed.
// std::vector that does not zero-fill on resize(): output buffers of 25-50 MB are overwritten by the C-ABI call
template <class T> struct DefaultInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new ((void*)p) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
using Words = std::vector<uint32_t, DefaultInitAlloc<uint32_t>>;

struct Packed {
  Words data;
  size_t count = 0, stride = 1;
  const uint32_t* at(size_t i) const { return data.data() + i * stride; }
  uint32_t* at(size_t i) { return data.data() + i * stride; }
  void check(size_t i) const { if (i >= count) throw py::index_error("index out of range"); }
  BigNumber element(size_t i) const { check(i); return BigNumber(at(i), stride); }
  static Packed from_list(const std::vector<BigNumber>& v) {
    Packed p;
    if (v.empty()) throw std::runtime_error("BaseText: empty container");
    p.count = v.size();
    p.stride = 1;
    for (auto& b : v) {
      if (b.neg) throw std::runtime_error("BaseText: negative BigNumber");
      p.stride = std::max(p.stride, b.w.size());
    }
    p.data.assign(p.count * p.stride, 0);
    for (size_t i = 0; i < p.count; ++i) std::copy(v[i].w.begin(), v[i].w.end(), p.at(i));
    return p;
  }
  static Packed from_u32_array(const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& a) {
    // the reference's numpy constructor: every uint32 is one element
    Packed p;
    auto r = a.unchecked<1>();
    if (r.shape(0) == 0) throw std::runtime_error("BaseText: empty container");
    p.count = (size_t)r.shape(0);
    p.stride = 1;
    p.data.resize(p.count);
    for (size_t i = 0; i < p.count; ++i) p.data[i] = r(i);
    return p;
  }
  static Packed from_matrix(const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& a) {
    if (a.ndim() != 2) throw std::runtime_error("from_packed: need a [count, words] uint32 array");
    Packed p;
    p.count = (size_t)a.shape(0);
    p.stride = std::max<size_t>(1, (size_t)a.shape(1));
    p.data.assign(a.data(), a.data() + p.count * (size_t)a.shape(1));
    return p;
  }
  // re-strided copy; throws if a value does not fit
  Words restride(size_t words, const char* what) const {
    Words out(count * words, 0u);
    for (size_t i = 0; i < count; ++i) {
      const uint32_t* s = at(i);
      for (size_t j = words; j < stride; ++j)
        if (s[j]) throw std::runtime_error(std::string(what) + ": value does not fit the key size");
      std::copy(s, s + std::min(words, stride), out.data() + i * words);
    }
    return out;
  }
  py::array_t<uint32_t> to_matrix(size_t words) const {
    py::array_t<uint32_t> a({(py::ssize_t)count, (py::ssize_t)words});
    if (words == stride) {
      if (!data.empty()) std::memcpy(a.mutable_data(), data.data(), data.size() * 4);
    } else {
      const Words v = restride(words, "to_packed");
      if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * 4);
    }
    return a;
  }
  Packed chunk(size_t start, size_t len) const {
    if (start + len > count) throw py::index_error("slice out of range");
    Packed p;
    p.count = len; p.stride = stride;
    p.data.assign(data.begin() + start * stride, data.begin() + (start + len) * stride);
    return p;
  }
  Packed rotated(int shift) const {  // BaseText::rotate: out[i] = in[(i + shift) mod count]
    Packed p;
    p.count = count; p.stride = stride; p.data.resize(data.size());
    if (count == 0) return p;
    const long long c = (long long)count;
    long long s = ((long long)shift % c + c) % c;
    for (size_t i = 0; i < count; ++i) std::copy(at((i + (size_t)s) % count), at((i + (size_t)s) % count) + stride, p.at(i));
    return p;
  }
  py::list texts() const {
    py::list l;
    for (size_t i = 0; i < count; ++i) l.append(std::make_shared<BigNumber>(at(i), stride));
    return l;
  }
  py::list element_vec(size_t i) const {
    check(i);
    size_t n = stride;
    while (n > 1 && at(i)[n - 1] == 0) --n;
    py::list l;
    for (size_t j = 0; j < n; ++j) l.append(at(i)[j]);
    return l;
  }
  std::string element_hex(size_t i) const {
    check(i);
    size_t n = stride;
    while (n > 1 && at(i)[n - 1] == 0) --n;
    return to_hex(at(i), n);
  }
  py::list state_list() const {
    py::list l;
    for (size_t i = 0; i < count; ++i) l.append(words_to_bytes(at(i), stride));
    return l;
  }
  static Packed from_state_list(size_t length, const py::list& l) {
    std::vector<BigNumber> v;
    v.reserve(length);
    for (size_t i = 0; i < length; ++i) v.push_back(from_bytes(l[i].cast<std::string>()));
    return from_list(v);
  }
};

struct PlainText : Packed {
  PlainText() {}
  explicit PlainText(Packed p) : Packed(std::move(p)) {}
};

// Ciphertext batches stay in HBM between operations: encrypt / + / * / modinv leave their result on the device only
// (dev set, host_valid false, data empty) and take operands from wherever they are -- the C ABI accepts host and
// device pointers alike.  Anything that looks at the words (indexing, getTexts, to_packed, pickle ...) goes through
// H(), which brings the batch to the host once.
struct CipherText : Packed {
  PubPtr pk;
  mutable uint32_t* dev = nullptr;
  mutable bool host_valid = true;
  CipherText(PubPtr k, Packed p) : Packed(std::move(p)), pk(std::move(k)) {
    const size_t cw = 2 * (size_t)pk->n_words;
    if (stride != cw) { data = restride(cw, "ipclCipherText"); stride = cw; }
  }
  CipherText(PubPtr k, size_t n, uint32_t* d) : pk(std::move(k)), dev(d), host_valid(false) {   // device-resident result
    count = n; stride = 2 * (size_t)pk->n_words;
  }
  CipherText(const CipherText&) = delete;
  CipherText& operator=(const CipherText&) = delete;
  ~CipherText() { if (dev) phe_dev_free(dev); }
  size_t words() const { return count * stride; }
  void sync_host() const {
    if (host_valid) return;
    Words& d = const_cast<Words&>(data);
    d.resize(words());
    // the GIL stays held: two Python threads looking at the same batch must not both resize `data`
    if (phe_copy(d.data(), dev, words() * 4)) throw_phe("ipclCipherText (device -> host)");
    host_valid = true;
  }
  const uint32_t* operand() const { return host_valid ? data.data() : dev; }   // host or device pointer for the C ABI
  bool on_device() const { return dev != nullptr; }
  // device pointer of the batch, uploading it once if it only lives on the host (the row operations below work on HBM)
  const uint32_t* device() const {
    if (!dev) {
      if (count == 0) throw std::runtime_error("ipclCipherText: empty container");
      uint32_t* d = nullptr;
      if (phe_dev_alloc(pk->h, words(), &d)) throw_phe("ipclCipherText (device allocation)");
      if (phe_copy(d, data.data(), words() * 4)) { phe_dev_free(d); throw_phe("ipclCipherText (host -> device)"); }
      dev = d;
    }
    return dev;
  }
};
inline const CipherText& H(const CipherText& c) { c.sync_host(); return c; }

// Output buffer of a ciphertext-valued operation: device memory when it can be had, else a host vector.
struct CtOut {
  PubPtr pk;
  size_t count;
  uint32_t* dev = nullptr;
  Packed host;
  CtOut(const PubPtr& k, size_t n) : pk(k), count(n) {
    const size_t cw = 2 * (size_t)pk->n_words;
    if (n == 0 || phe_dev_alloc(pk->h, n * cw, &dev) != 0) {
      dev = nullptr;
      host.count = n; host.stride = cw; host.data.resize(n * cw);
    }
  }
  uint32_t* ptr() { return dev ? dev : host.data.data(); }
  void drop() { if (dev) { phe_dev_free(dev); dev = nullptr; } }
  std::shared_ptr<CipherText> finish() {
    if (dev) { uint32_t* d = dev; dev = nullptr; return std::make_shared<CipherText>(pk, count, d); }
    return std::make_shared<CipherText>(pk, std::move(host));
  }
};

// ---- the hot path: one C-ABI call each, GIL released --------------------------------------------------------------
std::shared_ptr<CipherText> encrypt(const PubPtr& pk, const PlainText& pt, bool make_secure) {
  // plaintexts travel at their own stride when it is shorter than the key (53-bit fixed-point mantissas are 2 words,
  // not 64): 32x less to copy and to push over PCIe
  Words wide;
  const uint32_t* m = pt.data.data();
  size_t m_words = pt.stride;
  if (pt.stride > (size_t)pk->n_words) { wide = pt.restride((size_t)pk->n_words, "encrypt"); m = wide.data(); m_words = (size_t)pk->n_words; }
  CtOut out(pk, pt.count);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_encrypt_compact(pk->h, m, (int)m_words, pt.count, nullptr, 0, make_secure ? 1 : 0, out.ptr());
  }
  if (rc) { out.drop(); throw_phe("encrypt"); }
  return out.finish();
}

Packed obfuscate(const PubPtr& pk, Packed ct) {
  const size_t cw = 2 * (size_t)pk->n_words;
  if (ct.stride != cw) { ct.data = ct.restride(cw, "apply_obfuscator"); ct.stride = cw; }
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_obfuscate(pk->h, ct.data.data(), ct.count, nullptr, 0);
  }
  if (rc) throw_phe("apply_obfuscator");
  return ct;
}

struct PrivateKey {
  phe_privkey* h = nullptr;
  PubPtr pk;
  BigNumber p, q;
  PrivateKey(PubPtr pub, const BigNumber& p_, const BigNumber& q_) : pk(std::move(pub)) {
    const int hw = pk->n_words;  // generous: p, q fit n_words words
    std::vector<uint32_t> pw(hw, 0), qw(hw, 0);
    if ((int)p_.w.size() > hw || (int)q_.w.size() > hw) throw std::runtime_error("ipclPrivateKey: p or q larger than n");
    std::copy(p_.w.begin(), p_.w.end(), pw.begin());
    std::copy(q_.w.begin(), q_.w.end(), qw.begin());
    if (phe_privkey_create(pk->h, pw.data(), hw, qw.data(), hw, &h)) throw_phe("ipclPrivateKey");
    // the key stores p < q (ipcl::PrivateKey swaps)
    if (cmp_mag(p_, q_) <= 0) { p = p_; q = q_; } else { p = q_; q = p_; }
  }
  PrivateKey(const PrivateKey&) = delete;
  ~PrivateKey() { phe_privkey_destroy(h); }
  // decrypt + classification of the plaintexts on the device (phe_decrypt_mantissas): (mantissas, classes, words) as
  // numpy arrays; `words` is only meaningful in the rows of class 2
  py::tuple decrypt_mantissas(const CipherText& ct) const {
    if (cmp(ct.pk->n, pk->n) != 0) throw std::runtime_error("decrypt: public key mismatch");
    py::array_t<long long> mant((py::ssize_t)ct.count);
    py::array_t<unsigned char> cls((py::ssize_t)ct.count);
    py::array_t<uint32_t> rows({(py::ssize_t)ct.count, (py::ssize_t)pk->n_words});
    long long* mp = mant.mutable_data();
    unsigned char* cp = cls.mutable_data();
    uint32_t* rp = rows.mutable_data();
    int rc;
    {
      py::gil_scoped_release nogil;
      rc = phe_decrypt_mantissas(h, ct.operand(), ct.count, mp, cp, rp);
    }
    if (rc) throw_phe("decrypt");
    return py::make_tuple(mant, cls, rows);
  }
  PlainText decrypt(const CipherText& ct) const {
    if (cmp(ct.pk->n, pk->n) != 0) throw std::runtime_error("decrypt: public key mismatch");
    Packed out;
    out.count = ct.count; out.stride = (size_t)pk->n_words;
    out.data.resize(out.count * out.stride);
    int rc;
    {
      py::gil_scoped_release nogil;
      rc = phe_decrypt(h, ct.operand(), ct.count, out.data.data());
    }
    if (rc) throw_phe("decrypt");
    return PlainText(std::move(out));
  }
};

std::shared_ptr<CipherText> ct_add(const CipherText& a, const CipherText& b) {
  if (cmp(a.pk->n, b.pk->n) != 0) throw std::runtime_error("CipherText +: two different public keys detected");
  if (b.count != a.count && b.count != 1) throw std::runtime_error("CipherText +: size mismatch");
  CtOut out(a.pk, a.count);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_add(a.pk->h, a.operand(), a.count, b.operand(), b.count, out.ptr());
  }
  if (rc) { out.drop(); throw_phe("CipherText +"); }
  return out.finish();
}

// Row-wise inverse modulo n^2 (not part of the reference module: the reference's Python inverts with gmpy2 one element
// at a time; pailliercryptolib_python_b200/ipcl_python.py calls this instead).
std::shared_ptr<CipherText> ct_modinv(const CipherText& a) {
  CtOut out(a.pk, a.count);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_invert(a.pk->h, a.operand(), a.count, out.ptr());
  }
  if (rc) { out.drop(); throw_phe("CipherText modinv"); }
  return out.finish();
}

std::shared_ptr<CipherText> ct_mul(const CipherText& a, const PlainText& b) {
  if (b.count != a.count && b.count != 1) throw std::runtime_error("CipherText *: size mismatch");
  // exponent words: trim to what is used (the reference's modExp takes any exponent: up to 2 n_words here)
  size_t ew = 1;
  for (size_t i = 0; i < b.count; ++i)
    for (size_t j = b.stride; j-- > ew;)
      if (b.at(i)[j]) { ew = j + 1; break; }
  if (ew > 2 * (size_t)a.pk->n_words) throw std::runtime_error("CipherText *: plaintext larger than n^2");
  const Words e = b.restride(ew, "CipherText *");
  CtOut out(a.pk, a.count);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_mul(a.pk->h, a.operand(), a.count, e.data(), (int)ew, b.count, out.ptr());
  }
  if (rc) { out.drop(); throw_phe("CipherText *"); }
  return out.finish();
}

// ---- row operations: the batch stays in HBM (include/phe_b200.h "row operations") ------------------------------------
// What the reference's Python does with lists of BigNumber objects around + and * (ipcl_python.py:528-741 exponent
// alignment, :272-276 inversion for negative plaintexts, :777-880 matmul operand maps and add trees); called by
// pailliercryptolib_python_b200/ipcl_python.py, not part of the reference module.
using IdxArray = py::array_t<long long, py::array::c_style | py::array::forcecast>;
using DeltaArray = py::array_t<int, py::array::c_style | py::array::forcecast>;

std::shared_ptr<CipherText> device_result(const PubPtr& pk, size_t n) {
  uint32_t* d = nullptr;
  if (phe_dev_alloc(pk->h, n * 2 * (size_t)pk->n_words, &d)) throw_phe("ipclCipherText (device allocation)");
  return std::make_shared<CipherText>(pk, n, d);
}

// out[i] = a[idx[i]]
std::shared_ptr<CipherText> ct_gather(const CipherText& a, const IdxArray& idx) {
  const size_t n = (size_t)idx.size();
  if (n == 0) throw std::runtime_error("ipclCipherText.gather: empty index list");
  const uint32_t* src = a.device();
  auto out = device_result(a.pk, n);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_gather_rows_dev(a.pk->h, src, a.count, idx.data(), n, out->dev, nullptr);
  }
  if (rc) throw_phe("ipclCipherText.gather");
  return out;
}

std::shared_ptr<CipherText> ct_device_copy(const CipherText& a) {
  const uint32_t* src = a.device();
  auto out = device_result(a.pk, a.count);
  if (phe_copy(out->dev, src, a.words() * 4)) throw_phe("ipclCipherText (device copy)");
  return out;
}

// copy of a with rows idx raised to 2^delta (exponent alignment)
std::shared_ptr<CipherText> ct_scaled(const CipherText& a, const IdxArray& idx, const DeltaArray& delta) {
  if (idx.size() != delta.size()) throw std::runtime_error("ipclCipherText.scale_rows: idx / delta size mismatch");
  auto out = ct_device_copy(a);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_scale_rows_dev(a.pk->h, out->dev, out->count, idx.data(), delta.data(), (size_t)idx.size(), nullptr);
  }
  if (rc) throw_phe("ipclCipherText.scale_rows");
  return out;
}

// copy of a with rows idx inverted modulo n^2
std::shared_ptr<CipherText> ct_inverted_rows(const CipherText& a, const IdxArray& idx) {
  auto out = ct_device_copy(a);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_invert_rows_dev(a.pk->h, out->dev, out->count, idx.data(), (size_t)idx.size(), nullptr);
  }
  if (rc) throw_phe("ipclCipherText.invert_rows");
  return out;
}

// [groups * width] -> [groups]: HE-sum of every run of `width` rows
std::shared_ptr<CipherText> ct_segsum(const CipherText& a, size_t groups, size_t width) {
  if (groups == 0 || width == 0 || groups * width != a.count) throw std::runtime_error("ipclCipherText.segsum: groups * width != size");
  const uint32_t* src = a.device();
  auto out = device_result(a.pk, groups);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = phe_segsum_dev(a.pk->h, src, groups, width, out->dev, nullptr);
  }
  if (rc) throw_phe("ipclCipherText.segsum");
  return out;
}

// rows [start, start + len) without leaving the device when the batch lives there
std::shared_ptr<CipherText> ct_chunk(const CipherText& a, size_t start, size_t len) {
  if (start + len > a.count) throw py::index_error("slice out of range");
  if (!a.on_device() || len == 0) return std::make_shared<CipherText>(a.pk, H(a).chunk(start, len));
  auto out = device_result(a.pk, len);
  if (phe_copy(out->dev, a.dev + start * a.stride, len * a.stride * 4)) throw_phe("ipclCipherText (device slice)");
  return out;
}

// ------------------------------------------------------------------------------------------------ context / hybrid
enum class HybridMode { OPTIMAL = 95, QAT = 100, PREF_QAT90 = 90, PREF_QAT80 = 80, PREF_QAT70 = 70, PREF_QAT60 = 60, HALF = 50,
                        PREF_IPP60 = 40, PREF_IPP70 = 30, PREF_IPP80 = 20, PREF_IPP90 = 10, IPP = 0, UNDEFINED = -1 };
HybridMode g_hybrid = HybridMode::UNDEFINED;
struct Context {};
struct HybridControl {};
struct Keypair {};

template <class T> size_t slice_bounds(const T& self, const py::slice& s, size_t* len) {
  size_t start, stop, step;
  if (!s.compute(self.count, &start, &stop, &step, len)) throw py::error_already_set();
  if (step != 1) throw std::runtime_error("Step size not supported");
  return start;
}

}  // namespace

PYBIND11_MODULE(ipcl_bindings, m) {
  m.doc() = "ipcl_bindings surface over libphe_b200.so (B200 / sm_100a Paillier engine)";

  py::class_<Keypair>(m, "ipclKeypair")
      .def_static("generate_keypair", [](int64_t n_length, bool enable_DJN) {
        // ipcl::generateKeypair: 200 <= n_length <= 2048, n_length % 4 == 0; here up to 3072
        if (n_length < 200 || n_length > 3072 || n_length % 4)
          throw std::runtime_error("generateKeyPair: modulus size in bits should belong to either 1Kb, 2Kb, 3Kb or 4Kb range only, key size exceed the range!!! (n_length must be a multiple of 4 in [200, 3072])");
        const int nw = ((int)n_length + 31) / 32, pw = ((int)n_length / 2 + 31) / 32;
        std::vector<uint32_t> n(nw), p(pw), q(pw);
        int rc;
        {
          py::gil_scoped_release nogil;
          rc = phe_keygen((int)n_length, n.data(), p.data(), q.data());
        }
        if (rc) throw_phe("generate_keypair");
        auto pub = std::make_shared<PublicKey>(BigNumber(n.data(), n.size()), (int)n_length, enable_DJN);
        auto pri = std::make_shared<PrivateKey>(pub, BigNumber(p.data(), p.size()), BigNumber(q.data(), q.size()));
        return py::make_tuple(pub, pri);
      });

  py::class_<Context>(m, "context")
      .def_static("initializeContext", [](const std::string& kind) {
        // ipcl::initializeContext("QAT") (ipcl_bindings.hpp:27-35): here = is a CUDA device usable
        (void)kind;
        return phe_device_count() > 0;
      })
      .def_static("terminateContext", []() { return true; })
      .def_static("isQATRunning", []() { return false; })
      .def_static("isQATActive", []() { return false; })
      .def_static("deviceCount", []() { return phe_device_count(); })
      .def_static("setDevice", [](int d) { if (phe_set_device(d)) throw_phe("setDevice"); });

  py::enum_<HybridMode>(m, "hybridMode")
      .value("OPTIMAL", HybridMode::OPTIMAL).value("QAT", HybridMode::QAT)
      .value("PREF_QAT90", HybridMode::PREF_QAT90).value("PREF_QAT80", HybridMode::PREF_QAT80)
      .value("PREF_QAT70", HybridMode::PREF_QAT70).value("PREF_QAT60", HybridMode::PREF_QAT60)
      .value("HALF", HybridMode::HALF).value("PREF_IPP60", HybridMode::PREF_IPP60)
      .value("PREF_IPP70", HybridMode::PREF_IPP70).value("PREF_IPP80", HybridMode::PREF_IPP80)
      .value("PREF_IPP90", HybridMode::PREF_IPP90).value("IPP", HybridMode::IPP)
      .value("UNDEFINED", HybridMode::UNDEFINED)
      .export_values();

  py::class_<HybridControl>(m, "hybridControl")
      .def_static("setHybridMode", [](HybridMode mode) { g_hybrid = mode; })
      .def_static("setHybridOff", []() { g_hybrid = HybridMode::UNDEFINED; })
      .def_static("getHybridMode", []() { return g_hybrid; });

  // ---- ipclBigNumber (ipcl_bindings_classes.cpp:380-491)
  py::class_<BigNumber, std::shared_ptr<BigNumber>>(m, "ipclBigNumber")
      .def(py::init<const BigNumber&>())
      .def(py::init([](uint32_t v) { return std::make_shared<BigNumber>(v); }))
      .def(py::init([](const py::list& data) {
        std::vector<uint32_t> v = data.cast<std::vector<uint32_t>>();
        return std::make_shared<BigNumber>(v.data(), v.size());
      }))
      .def(py::init([](const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& a) {
        return std::make_shared<BigNumber>(a.data(), (size_t)a.size());
      }))
      .def(py::init([](const py::bytes& b) { return std::make_shared<BigNumber>(from_bytes(b)); }))
      .def("__repr__", [](const BigNumber& s) { return "<BigNumber " + addr_tag(&s) + " val: " + to_decimal(s) + ">"; })
      .def("__str__", [](const BigNumber& s) { return to_decimal(s); })
      .def("__getitem__", [](const BigNumber& s, size_t i) {
        if (i >= s.w.size()) throw std::out_of_range("Index is larger than size: " + std::to_string(s.w.size()));
        return s.w[i];
      })
      .def("__eq__", [](const BigNumber& a, const BigNumber& b) { return cmp(a, b) == 0; })
      .def("__ne__", [](const BigNumber& a, const BigNumber& b) { return cmp(a, b) != 0; })
      .def("__lt__", [](const BigNumber& a, const BigNumber& b) { return cmp(a, b) < 0; })
      .def("__le__", [](const BigNumber& a, const BigNumber& b) { return cmp(a, b) <= 0; })
      .def("__gt__", [](const BigNumber& a, const BigNumber& b) { return cmp(a, b) > 0; })
      .def("__ge__", [](const BigNumber& a, const BigNumber& b) { return cmp(a, b) >= 0; })
      .def("__hash__", [](const BigNumber& a) { return std::hash<std::string>{}(to_decimal(a)); })
      .def("__add__", [](const BigNumber& a, const BigNumber& b) { return add(a, b); })
      .def("__sub__", [](const BigNumber& a, const BigNumber& b) { return add(a, negate(b)); })
      .def("__iadd__", [](BigNumber& a, const BigNumber& b) { a = add(a, b); return a; })
      .def("__mul__", [](const BigNumber& a, const BigNumber& b) { return mul(a, b); })
      .def("__mul__", [](const BigNumber& a, uint32_t b) { return mul(a, BigNumber(b)); })
      .def("DwordSize", &BigNumber::dword_size)
      .def("BitSize", &BigNumber::bit_size)
      .def("data", [](const BigNumber& s) {
        py::list l;
        for (uint32_t x : s.w) l.append(x);
        return py::make_tuple((int)s.w.size(), l);
      })
      .def("to_bytes", [](const BigNumber& s) { return to_bytes(s); })
      .def_property_readonly_static("Zero", [](const py::object&) { return std::make_shared<BigNumber>(0u); })
      .def_property_readonly_static("One", [](const py::object&) { return std::make_shared<BigNumber>(1u); })
      .def_property_readonly_static("Two", [](const py::object&) { return std::make_shared<BigNumber>(2u); })
      .def(py::pickle([](const BigNumber& s) { return py::make_tuple(to_bytes(s)); },
                      [](py::tuple t) { return std::make_shared<BigNumber>(from_bytes(t[0].cast<std::string>())); }));

  // ---- ipclPublicKey (ipcl_bindings_classes.cpp:12-91)
  py::class_<PublicKey, PubPtr>(m, "ipclPublicKey")
      .def(py::init([](const BigNumber& n) { return std::make_shared<PublicKey>(n, 1024, false); }))
      .def(py::init([](const BigNumber& n, int bits) { return std::make_shared<PublicKey>(n, bits, false); }))
      .def(py::init([](const BigNumber& n, int bits, bool djn) { return std::make_shared<PublicKey>(n, bits, djn); }))
      .def_static("create", [](const BigNumber& n, int bits, const BigNumber& hs, int randbits) {
        // ipcl::PublicKey::create(n, bits, hs, randbits): a DJN key with a given generator (what unpickling uses)
        return std::make_shared<PublicKey>(n, bits, true, &hs, randbits);
      })
      .def("__repr__", [](const PublicKey& s) { return "<ipclPublicKey " + addr_tag(&s) + ">"; })
      .def("__eq__", [](const PublicKey& a, const PublicKey& b) { return cmp(a.n, b.n) == 0; })
      .def("__hash__", [](const PublicKey& s) { return std::hash<std::string>{}(to_decimal(s.n)); })
      .def_property_readonly("n", [](const PublicKey& s) { return std::make_shared<BigNumber>(s.n); })
      .def_property_readonly("length", [](const PublicKey& s) { return s.bits; })
      .def_property_readonly("nsquare", [](const PublicKey& s) { return std::make_shared<BigNumber>(s.nsquare()); })
      .def_property_readonly("hs", [](const PublicKey& s) { return std::make_shared<BigNumber>(s.hs()); })
      .def_property_readonly("isDJN", &PublicKey::djn)
      .def_property_readonly("randbits", &PublicKey::randbits)
      .def("encrypt", [](const PubPtr& s, const PlainText& pt, bool make_secure) { return encrypt(s, pt, make_secure); })
      .def("encrypt_tolist", [](const PubPtr& s, const PlainText& pt, bool make_secure) { return H(*encrypt(s, pt, make_secure)).texts(); })
      .def("apply_obfuscator", [](const PubPtr& s, const BigNumber& ct) {
        return std::make_shared<BigNumber>(obfuscate(s, Packed::from_list({ct})).element(0));
      })
      .def("apply_obfuscator", [](const PubPtr& s, const CipherText& ct) { return obfuscate(s, H(ct)).texts(); })
      .def("apply_obfuscator_packed", [](const PubPtr& s, const CipherText& ct) {
        return std::make_shared<CipherText>(s, obfuscate(s, H(ct)));
      })
      .def(py::pickle([](const PublicKey& s) { return pubkey_state(s); }, [](py::tuple t) { return pubkey_from_state(t); }));

  // ---- ipclPrivateKey (ipcl_bindings_classes.cpp:93-163)
  py::class_<PrivateKey, std::shared_ptr<PrivateKey>>(m, "ipclPrivateKey")
      .def(py::init([](const PubPtr& pk, const BigNumber& p, const BigNumber& q) { return std::make_shared<PrivateKey>(pk, p, q); }))
      .def("__repr__", [](const PrivateKey& s) { return "<ipclPrivateKey " + addr_tag(&s) + ">"; })
      .def("__eq__", [](const PrivateKey& a, const PrivateKey& b) { return cmp(a.q, b.q) == 0; })
      .def("__hash__", [](const PrivateKey& s) { return std::hash<std::string>{}(addr_tag(&s)); })
      .def_property_readonly("n", [](const PrivateKey& s) { return std::make_shared<BigNumber>(s.pk->n); })
      .def_property_readonly("p", [](const PrivateKey& s) { return std::make_shared<BigNumber>(s.p); })
      .def_property_readonly("q", [](const PrivateKey& s) { return std::make_shared<BigNumber>(s.q); })
      .def_property_readonly("public_key", [](const PrivateKey& s) { return s.pk; })
      .def("decrypt", [](const PrivateKey& s, const CipherText& ct) { return s.decrypt(ct); })
      .def("decrypt_tolist", [](const PrivateKey& s, const CipherText& ct) { return s.decrypt(ct).texts(); })
      .def("decrypt_mantissas", [](const PrivateKey& s, const CipherText& ct) { return s.decrypt_mantissas(ct); },
           "decrypt and classify on the device: (int64 mantissas, uint8 classes 0 positive / 1 negative / 2 see words, uint32 words)")
      .def(py::pickle(
          [](const PrivateKey& s) { return py::make_tuple(to_bytes(s.pk->n), to_bytes(s.p), to_bytes(s.q), pubkey_state(*s.pk)); },
          [](py::tuple t) {
            // reference state is (n, p, q) (ipcl_bindings_classes.cpp:142-162); we append the public-key tuple so a DJN
            // key keeps its hs.  A 3-tuple (reference pickle) rebuilds a classic-scheme key, as upstream does.
            const BigNumber n = from_bytes(t[0].cast<std::string>());
            const BigNumber p = from_bytes(t[1].cast<std::string>()), q = from_bytes(t[2].cast<std::string>());
            PubPtr pk = t.size() > 3 ? pubkey_from_state(t[3].cast<py::tuple>()) : std::make_shared<PublicKey>(n, n.bit_size(), false);
            return std::make_shared<PrivateKey>(pk, p, q);
          }));

  // ---- ipclPlainText (ipcl_bindings_classes.cpp:165-266)
  py::class_<PlainText>(m, "ipclPlainText")
      .def(py::init([](uint32_t v) { return PlainText(Packed::from_list({BigNumber(v)})); }))
      .def(py::init([](const BigNumber& v) { return PlainText(Packed::from_list({v})); }))
      .def(py::init<const PlainText&>())
      .def(py::init([](const py::list& data) { return PlainText(Packed::from_list(data.cast<std::vector<BigNumber>>())); }))
      .def(py::init([](const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& a) { return PlainText(Packed::from_u32_array(a)); }))
      .def_static("from_packed", [](const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& a) { return PlainText(Packed::from_matrix(a)); },
                  "build from a [count, words] little-endian uint32 limb matrix (no per-element objects)")
      .def("to_packed", [](const PlainText& s, py::object words) { return s.to_matrix(words.is_none() ? s.stride : words.cast<size_t>()); },
           py::arg("words") = py::none())
      .def("__repr__", [](const PlainText& s) { return "<ipclPlainText " + addr_tag(&s) + ">"; })
      .def("__str__", [](const PlainText& s) { return "<ipclPlainText " + addr_tag(&s) + ">"; })
      .def("__eq__", [](const PlainText& a, const PlainText& b) {
        if (a.count != b.count) throw std::runtime_error("Size mismatch");
        for (size_t i = 0; i < a.count; ++i)
          if (cmp(a.element(i), b.element(i)) != 0) throw std::runtime_error("PlainText mismatch");
        return true;
      })
      .def("__getitem__", [](const PlainText& s, size_t i) { return std::make_shared<BigNumber>(s.element(i)); })
      .def("__getitem__", [](const PlainText& s, const py::slice& sl) { size_t len; const size_t st = slice_bounds(s, sl, &len); return PlainText(s.chunk(st, len)); })
      .def("__len__", [](const PlainText& s) { return s.count; })
      .def("rotate", [](const PlainText& s, int shift) { return PlainText(s.rotated(shift)); })
      .def("getElementVec", [](const PlainText& s, size_t i) { return s.element_vec(i); })
      .def("getElementHex", [](const PlainText& s, size_t i) { return s.element_hex(i); })
      .def("getTexts", [](const PlainText& s) { return s.texts(); })
      .def("getSize", [](const PlainText& s) { return s.count; })
      .def(py::pickle([](const PlainText& s) { return py::make_tuple(s.count, s.state_list()); },
                      [](const py::tuple& t) { return PlainText(Packed::from_state_list(t[0].cast<size_t>(), t[1].cast<py::list>())); }));

  // ---- ipclCipherText (ipcl_bindings_classes.cpp:268-378)
  py::class_<CipherText, std::shared_ptr<CipherText>>(m, "ipclCipherText")
      .def(py::init([](const PubPtr& pk, uint32_t v) { return std::make_shared<CipherText>(pk, Packed::from_list({BigNumber(v)})); }))
      .def(py::init([](const PubPtr& pk, const BigNumber& v) { return std::make_shared<CipherText>(pk, Packed::from_list({v})); }))
      .def(py::init([](const PubPtr& pk, const py::list& data) { return std::make_shared<CipherText>(pk, Packed::from_list(data.cast<std::vector<BigNumber>>())); }))
      .def(py::init([](const PubPtr& pk, const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& a) { return std::make_shared<CipherText>(pk, Packed::from_u32_array(a)); }))
      // the reference's Python rebuilds a container from what __getitem__(slice) returned (ipcl_python.py:355-356);
      // a slice is a container here, so this is the copy constructor under the same key
      .def(py::init([](const PubPtr& pk, const CipherText& other) {
        if (cmp(pk->n, other.pk->n) != 0) throw std::runtime_error("ipclCipherText: public key mismatch");
        return std::make_shared<CipherText>(pk, H(other).chunk(0, other.count));
      }))
      .def_static("from_packed", [](const PubPtr& pk, const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& a) { return std::make_shared<CipherText>(pk, Packed::from_matrix(a)); })
      .def("to_packed", [](const CipherText& s) { return H(s).to_matrix(s.stride); })
      .def("__repr__", [](const CipherText& s) { return "<ipclCipherText " + addr_tag(&s) + ">"; })
      .def("__str__", [](const CipherText& s) { return "<ipclCipherText " + addr_tag(&s) + ">"; })
      .def("__getitem__", [](const CipherText& s, size_t i) { return std::make_shared<BigNumber>(H(s).element(i)); })
      .def("__getitem__", [](const CipherText& s, const py::slice& sl) { size_t len; const size_t st = slice_bounds(s, sl, &len); return ct_chunk(s, st, len); })
      .def("modinv", [](const CipherText& a) { return ct_modinv(a); })
      .def("gather", &ct_gather, "out[i] = self[idx[i]] (rows may repeat); device resident")
      .def("scale_rows", &ct_scaled, "copy of self with rows idx raised to 2^delta (exponent alignment); device resident")
      .def("invert_rows", &ct_inverted_rows, "copy of self with rows idx inverted modulo n^2; device resident")
      .def("segsum", &ct_segsum, "HE-sum of every run of `width` rows: [groups * width] -> [groups]; device resident")
      .def("wait", [](const CipherText& s) {   // block until the batch has been computed (results are enqueued, not awaited)
        if (s.dev && s.count) { uint32_t w; py::gil_scoped_release nogil; if (phe_copy(&w, s.dev, 4)) throw_phe("ipclCipherText.wait"); }
      })
      .def_property_readonly("on_device", [](const CipherText& s) { return s.on_device(); })
      .def_property_readonly("host_valid", [](const CipherText& s) { return s.host_valid; })
      .def("__add__", [](const CipherText& a, const CipherText& b) { return ct_add(a, b); })
      .def("__add__", [](const CipherText& a, const PlainText& b) { return ct_add(a, *encrypt(a.pk, b, false)); })
      .def("__mul__", [](const CipherText& a, const PlainText& b) { return ct_mul(a, b); })
      .def("__len__", [](const CipherText& s) { return s.count; })
      .def("getCipherText", [](const CipherText& s, size_t i) { s.check(i); return ct_chunk(s, i, 1); })
      .def("rotate", [](const CipherText& s, int shift) { return std::make_shared<CipherText>(s.pk, H(s).rotated(shift)); })
      .def("getElementVec", [](const CipherText& s, size_t i) { return H(s).element_vec(i); })
      .def("getElementHex", [](const CipherText& s, size_t i) { return H(s).element_hex(i); })
      .def_property_readonly("public_key", [](const CipherText& s) { return s.pk; })
      .def("getTexts", [](const CipherText& s) { return H(s).texts(); })
      .def("getSize", [](const CipherText& s) { return s.count; })
      .def(py::pickle([](const CipherText& s) { return py::make_tuple(s.count, H(s).state_list(), pubkey_state(*s.pk)); },
                      [](const py::tuple& t) {
                        PubPtr pk = pubkey_from_state(t[2].cast<py::tuple>());
                        return std::make_shared<CipherText>(pk, Packed::from_state_list(t[0].cast<size_t>(), t[1].cast<py::list>()));
                      }));
}
