// phe_api.cu -- C ABI (include/phe_b200.h) over the sm_100a kernels.  Host side = per-key setup with hostbn,
// buffer management, launches.  There is no CPU fallback for the compute entry points.
#include <sys/random.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/phe_b200.h"
#include "hostbn.hpp"
#include "phe_kernels.cuh"
#include "npair_kernels.cuh"
#include "phe_shapes.hpp"

using hbn::BN;

namespace phe {
cudaError_t chacha20_fill(uint32_t* d_out, size_t words, const uint32_t key[8], const uint32_t nonce[3], uint32_t counter0,
                          int mask_every, uint32_t top_mask, cudaStream_t s);
extern const PairOps g_pair_10, g_pair_20, g_pair_30;
extern const ShapeOps g_ops_20_1, g_ops_20_2, g_ops_20_4, g_ops_20_8, g_ops_15_4, g_ops_15_8;
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_counter() { return g_launches.load(); }
// ---- per-kind launch timing (cudaEvent pairs on the launching stream) -------------------------------------
namespace {
struct TimingState {
  std::mutex mu;
  bool on = false;
  struct Pair { int kind; cudaEvent_t e0, e1; };
  std::vector<Pair> open_pairs;                 // recorded, not yet read
  std::vector<cudaEvent_t> pending_begin[KK_COUNT];
  double ms[KK_COUNT] = {0};
  unsigned long long n[KK_COUNT] = {0};
} g_tm;
std::atomic<bool> g_tm_on{false};
}  // namespace
void timing_begin(int kind, cudaStream_t s) {
  if (!g_tm_on.load(std::memory_order_relaxed)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  std::lock_guard<std::mutex> lk(g_tm.mu);
  g_tm.pending_begin[kind].push_back(e);
}
void timing_end(int kind, cudaStream_t s) {
  if (!g_tm_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_tm.mu);
  if (g_tm.pending_begin[kind].empty()) return;
  cudaEvent_t e0 = g_tm.pending_begin[kind].back();
  g_tm.pending_begin[kind].pop_back();
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return; }
  cudaEventRecord(e1, s);
  g_tm.open_pairs.push_back({kind, e0, e1});
}
static void timing_drain() {   // caller holds no lock; waits for the recorded events
  std::vector<TimingState::Pair> pairs;
  { std::lock_guard<std::mutex> lk(g_tm.mu); pairs.swap(g_tm.open_pairs); }
  for (auto& p : pairs) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      std::lock_guard<std::mutex> lk(g_tm.mu);
      g_tm.ms[p.kind] += ms; g_tm.n[p.kind] += 1;
    }
    cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
  }
}
static void timing_set(bool on) {
  timing_drain();
  std::lock_guard<std::mutex> lk(g_tm.mu);
  for (int k = 0; k < KK_COUNT; ++k) { g_tm.ms[k] = 0; g_tm.n[k] = 0; }
  g_tm_on.store(on);
}
static bool timing_read(int kind, double* ms, unsigned long long* n) {
  if (kind < 0 || kind >= KK_COUNT) return false;
  timing_drain();
  std::lock_guard<std::mutex> lk(g_tm.mu);
  if (ms) *ms = g_tm.ms[kind];
  if (n) *n = g_tm.n[kind];
  return true;
}
const ShapeOps* shape_ops(int L, int TPI) {
  const ShapeOps* all[] = {&g_ops_20_1, &g_ops_20_2, &g_ops_20_4, &g_ops_20_8, &g_ops_15_4, &g_ops_15_8};
  for (auto* o : all) if (o->L == L && o->TPI == TPI) return o;
  return nullptr;
}
const PairOps* pair_ops(int L) {
  const PairOps* all[] = {&g_pair_10, &g_pair_20, &g_pair_30};
  for (auto* o : all) if (o->L == L) return o;
  return nullptr;
}
}  // namespace phe

using namespace phe;

namespace {

thread_local std::string t_err;
int fail(const std::string& m) { t_err = m; return 1; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)
#define PHE_TRY(x) do { int r_ = (x); if (r_) return r_; } while (0)
// joins the key's event chain for the rest of the scope (see StreamChain)
#define KEY_CHAIN(key, stream) ChainScope chain_((key)->chain, (cudaStream_t)(stream)); if (chain_.rc) return chain_.rc

// smallest shape whose capacity covers mod_bits + 8 (R >= 2^8 N keeps every Montgomery product < 2N)
const ShapeOps* shape_for_bits(int mod_bits) {
  static const int order[][2] = {{20, 1}, {20, 2}, {15, 4}, {20, 4}, {15, 8}, {20, 8}};   // 1040 .. 8320 bits
  for (auto& s : order) {
    const ShapeOps* o = shape_ops(s[0], s[1]);
    if (o && o->capacity_bits >= mod_bits + 8) return o;
  }
  return nullptr;
}

// Entries are doubles on the device; host vectors and DevBuf count u32 words: EW(o) words per entry.
inline size_t EW(const ShapeOps* o) { return 2 * (size_t)o->KP; }

// value -> padded limb entry [TPI][LP]: every 52-bit limb as the double holding that integer
void to_entry(const BN& v, const ShapeOps* o, uint32_t* out_words) {
  const int LP = o->KP / o->TPI;
  std::vector<double> ent((size_t)o->KP, 0.0);
  if (v.bits() > (size_t)o->capacity_bits) throw std::runtime_error("to_entry: value exceeds shape capacity");
  for (int g = 0; g < o->L * o->TPI; ++g) {
    const size_t bit = (size_t)g * LW;
    const size_t wi = bit >> 5, sh = bit & 31;
    unsigned __int128 three = 0;
    for (int k = 0; k < 3; ++k)
      if (wi + k < v.w.size()) three |= (unsigned __int128)v.w[wi + k] << (32 * k);
    ent[(g / o->L) * LP + (g % o->L)] = (double)((uint64_t)(three >> sh) & M52);
  }
  std::memcpy(out_words, ent.data(), ent.size() * sizeof(double));
}

// -n^-1 mod 2^52 for odd n
uint64_t neg_inv52(const BN& n) {
  const uint64_t n0 = (uint64_t)n.w[0] | (n.w.size() > 1 ? (uint64_t)n.w[1] << 32 : 0ull);
  uint64_t x = 1;
  for (int i = 0; i < 6; ++i) x *= 2ull - n0 * x;
  return (0ull - x) & M52;
}

// Montgomery block (ME_COUNT entries) for modulus N with context-specific extra value
std::vector<uint32_t> mont_block(const BN& N, const BN& extra, const ShapeOps* o, uint64_t* n0inv) {
  if (!N.is_odd()) throw std::runtime_error("modulus must be odd");
  std::vector<uint32_t> blk((size_t)ME_COUNT * EW(o));
  const BN R = hbn::shl(BN(1), o->capacity_bits);
  const BN Rm = hbn::mod(R, N);
  to_entry(N, o, &blk[(size_t)ME_N * EW(o)]);
  to_entry(hbn::mulmod(Rm, Rm, N), o, &blk[(size_t)ME_R2 * EW(o)]);
  to_entry(Rm, o, &blk[(size_t)ME_ONEM * EW(o)]);
  to_entry(BN(1), o, &blk[(size_t)ME_ONE * EW(o)]);
  to_entry(extra, o, &blk[(size_t)ME_X0 * EW(o)]);
  *n0inv = neg_inv52(N);
  return blk;
}

// Constant block of the n-adic pair engine (npair_items.cuh: NPairEntry) for modulus n on shape o.
std::vector<uint32_t> npair_block(const BN& n, int n_words, const ShapeOps* o, uint64_t* n0inv, uint64_t* d_top) {
  std::vector<uint32_t> blk((size_t)NE_COUNT * EW(o));
  const BN R = hbn::shl(BN(1), o->capacity_bits);
  const BN n2 = hbn::mul(n, n);
  const BN D = hbn::mul(hbn::div(hbn::add(R, hbn::sub(n, BN(1))), n), n);   // ceil(R / n) n
  BN dq, dr;
  hbn::divmod(D, R, &dq, &dr);
  if (dq.bits() > 52) throw std::runtime_error("npair_block: top limb of D out of range");
  *d_top = dq.is_zero() ? 0ull : ((uint64_t)dq.w[0] | (dq.w.size() > 1 ? (uint64_t)dq.w[1] << 32 : 0ull));
  auto put = [&](int idx, const BN& v) { to_entry(v, o, &blk[(size_t)idx * EW(o)]); };
  put(NE_N, n); put(NE_ONE, BN(1)); put(NE_D, dr);
  const BN R2 = hbn::mod(hbn::mul(hbn::mod(R, n2), hbn::mod(R, n2)), n2);
  for (int c = 0; c < 2; ++c) {
    const BN w = hbn::mulmod(hbn::mod(hbn::shl(BN(1), (size_t)c * 32 * n_words), n2), R2, n2);
    BN q, r;
    hbn::divmod(w, n, &q, &r);
    put(NE_W00 + 2 * c, r); put(NE_W01 + 2 * c, q);
  }
  {
    BN q, r;
    hbn::divmod(hbn::mod(R, n2), n, &q, &r);
    put(NE_OM0, r); put(NE_OM1, q);
  }
  *n0inv = neg_inv52(n);
  return blk;
}

struct DevBuf {
  uint32_t* p = nullptr;
  size_t words = 0;
  int ensure(size_t w) {
    if (w <= words) return 0;
    if (p) cudaFree(p);
    p = nullptr; words = 0;
    // grow geometrically to avoid re-allocating on slowly increasing batch sizes
    size_t want = w + w / 8;
    cudaError_t e = cudaMalloc(&p, want * 4);
    if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, w * 4); want = w; }   // retry without the slack
    if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    words = want;
    return 0;
  }
  int ensure_exact(size_t w) {   // large long-lived tables: no geometric slack
    if (w <= words) return 0;
    release();
    cudaError_t e = cudaMalloc(&p, w * 4);
    if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    words = w;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; words = 0; }
};

// The workspaces of a key (window tables, intermediates, scheduler counters) are shared by every call on that key.
// The *_dev entry points only enqueue on the caller's stream, so two calls on different streams would overlap on the
// GPU and overwrite each other's scratch.  Every call therefore joins a per-key event chain: before it touches the
// scratch its stream waits for the event recorded after the previous call on the key, and it records a new one when
// its own work is enqueued.  Calls on one key thus execute in the order they were made, whatever their streams (no
// host synchronisation); calls on different keys stay independent.
struct StreamChain {
  cudaEvent_t ev = nullptr;
  cudaStream_t last = nullptr;
  bool used = false;
  int enter(cudaStream_t s) {
    if (used && s != last) CUDA_TRY(cudaStreamWaitEvent(s, ev, 0));
    return 0;
  }
  int leave(cudaStream_t s) {
    if (!ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ev, s));
    last = s; used = true;
    return 0;
  }
  void release() { if (ev) cudaEventDestroy(ev); ev = nullptr; used = false; }
};
// RAII over one call: enter() at construction (status in rc), leave() at scope exit
struct ChainScope {
  StreamChain& c; cudaStream_t s; int rc;
  ChainScope(StreamChain& chain, cudaStream_t st) : c(chain), s(st), rc(chain.enter(st)) {}
  ~ChainScope() { if (!rc) c.leave(s); }
};

int upload(DevBuf& b, const std::vector<uint32_t>& v) {
  PHE_TRY(b.ensure(v.size()));
  CUDA_TRY(cudaMemcpy(b.p, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

int random_words(uint32_t* out, size_t words) {
  uint8_t* p = reinterpret_cast<uint8_t*>(out);
  size_t left = words * 4;
  while (left) {
    const ssize_t got = getrandom(p, left > (1u << 24) ? (1u << 24) : left, 0);
    if (got <= 0) return fail("getrandom failed");
    p += got; left -= (size_t)got;
  }
  return 0;
}

BN random_bits(size_t bits) {
  std::vector<uint32_t> w((bits + 31) / 32);
  if (random_words(w.data(), w.size())) throw std::runtime_error("getrandom failed");
  if (bits & 31) w.back() &= (1u << (bits & 31)) - 1u;
  return BN::from_words(w.data(), w.size());
}

int max_bits_host(const uint32_t* e, int e_words, size_t n) {
  int best = 1;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t* row = e + i * (size_t)e_words;
    for (int j = e_words - 1; j >= 0; --j)
      if (row[j]) { const int b = 32 * j + 32 - __builtin_clz(row[j]); if (b > best) best = b; break; }
  }
  return best;
}

// Sliding-window program of a shared exponent (paillier_items.cuh: item_powm_prog), window width PROG_WS.
std::vector<uint32_t> build_powm_program(const BN& e) {
  std::vector<uint32_t> prog;
  if (e.is_zero()) { prog.push_back(PROG_ONE); return prog; }
  long i = (long)e.bits() - 1;
  bool first = true;
  uint32_t pending_sq = 0;
  while (i >= 0) {
    if (!e.bit((size_t)i)) { ++pending_sq; --i; continue; }
    long j = std::max<long>(i - PROG_WS + 1, 0);
    while (!e.bit((size_t)j)) ++j;
    uint32_t digit = 0;
    for (long k = i; k >= j; --k) digit = (digit << 1) | (e.bit((size_t)k) ? 1u : 0u);
    const uint32_t idx = (digit - 1) >> 1;
    if (first) { prog.push_back(idx); first = false; }
    else {
      uint32_t sq = pending_sq + (uint32_t)(i - j + 1);
      while (sq > 0xffffffu) { prog.push_back((0xffffffu << 8) | PROG_NOMUL); sq -= 0xffffffu; }
      prog.push_back((sq << 8) | idx);
    }
    pending_sq = 0;
    i = j - 1;
  }
  if (pending_sq) prog.push_back((pending_sq << 8) | PROG_NOMUL);
  return prog;
}

// ---- p-adic pair engine (decrypt halves): constants and program -------------------------------------------------
// L limbs of 52 bits as doubles
void to_limbs52(const BN& v, int L, double* out) {
  for (int g = 0; g < L; ++g) {
    const size_t bit = (size_t)g * 52, wi = bit >> 5, sh = bit & 31;
    unsigned __int128 three = 0;
    for (int k = 0; k < 3; ++k)
      if (wi + k < v.w.size()) three |= (unsigned __int128)v.w[wi + k] << (32 * k);
    out[g] = (double)((uint64_t)(three >> sh) & M52);
  }
  if (v.bits() > (size_t)L * 52) throw std::runtime_error("to_limbs52: value exceeds L limbs");
}

struct PairBlock {
  int L = 0;
  uint64_t n0inv = 0;
  std::vector<double> mod;       // [L] limbs of x
  std::vector<double> cst;       // [PC_COUNT][2][L]
  std::vector<uint32_t> prog;    // the whole program in one piece
  std::vector<uint32_t> segprog; // the same program cut into time slices (k_dec_pair), segments back to back
  std::vector<int> segoff;       // start of every segment inside segprog
};

// Segments a unit of k_dec_pair is time-sliced into when a launch has more units than resident warps (phe_kernels.cuh).
// 16: the ragged end of a launch is at most 1/16 of an exponentiation; a boundary costs one store + load of the running
// pair (measured 4 / 8 / 16 over batch sizes: profiles/r02_dec_sweep_ready_queue.json).
constexpr int PAIR_SEGMENTS = 16;
static_assert(PAIR_SEGMENTS <= PAIR_MAX_SEG, "DecPairArgs::seg_off is too short");

// Cuts the pair-engine program P into at most `nseg` segments of (nearly) equal cost -- a multiplication is 3 passes, a
// run of n squarings 2 n (runs are split) -- at points from instruction `zone` on where the only live state is the
// running pair (X0, X1): before a run of squarings or before the load of a multiplier.  A segment ends with
// [PO_TX park, PO_END] (the pair goes to table slot `park`), the next one starts with [PO_XT park].
void split_pair_program(const std::vector<uint32_t>& P, size_t zone, int nseg, uint32_t park, std::vector<uint32_t>* out,
                        std::vector<int>* off) {
  auto ins = [](uint32_t op, uint32_t arg) { return op | (arg << 8); };
  auto cost_of = [](uint32_t w) -> uint64_t {
    const uint32_t op = w & 0xffu, arg = w >> 8;
    return op == PO_MUL ? 3u : (op == PO_SQR ? 2ull * arg : 0u);
  };
  uint64_t total = 0;
  for (uint32_t w : P) total += cost_of(w);
  out->clear(); off->clear();
  off->push_back(0);
  uint64_t acc = 0;
  int k = 1;                                   // next boundary: acc >= k * total / nseg
  size_t seg_start_cost_marker = 0;            // cost at the start of the current segment (no empty segments)
  auto target = [&](int kk) { return (total * (uint64_t)kk + nseg - 1) / nseg; };
  auto cut = [&]() {
    out->push_back(ins(PO_TX, park)); out->push_back(ins(PO_END, 0));
    off->push_back((int)out->size());
    out->push_back(ins(PO_XT, park));
    seg_start_cost_marker = acc;
    while (k < nseg && target(k) <= acc) ++k;
  };
  for (size_t i = 0; i < P.size(); ++i) {
    const uint32_t op = P[i] & 0xffu, arg = P[i] >> 8;
    const bool cuttable = i >= zone && (op == PO_SQR || op == PO_YT || op == PO_YCONST);
    if (cuttable && k < nseg && acc >= target(k) && acc > seg_start_cost_marker) cut();
    if (op == PO_SQR && i >= zone) {
      uint32_t left = arg;
      while (k < nseg && left > 0 && acc + 2ull * left > target(k)) {   // the boundary falls inside this run
        const uint64_t need = target(k) > acc ? target(k) - acc : 0;
        uint32_t a = (uint32_t)((need + 1) / 2);
        if (a >= left) break;
        if (a > 0) { out->push_back(ins(PO_SQR, a)); acc += 2ull * a; left -= a; }
        if (acc > seg_start_cost_marker) cut(); else break;
      }
      if (left > 0) { out->push_back(ins(PO_SQR, left)); acc += 2ull * left; }
    } else {
      out->push_back(P[i]);
      acc += cost_of(P[i]);
    }
  }
}

// x: p or q (bits == chunk_bits), hx = h_x, chunk_bits = key bits / 2.  Returns false if no pair shape fits.
bool build_pair_block(const BN& x, const BN& hx, int chunk_bits, PairBlock* out) {
  int L = 0;
  for (int cand : {10, 20, 30}) if ((size_t)cand * 52 >= x.bits() + 8 && pair_ops(cand)) { L = cand; break; }
  if (!L || (int)x.bits() != chunk_bits) return false;
  out->L = L;
  out->n0inv = neg_inv52(x);
  const BN R = hbn::shl(BN(1), 52 * (size_t)L);
  const BN x2 = hbn::mul(x, x);
  out->mod.assign((size_t)L, 0.0);
  to_limbs52(x, L, out->mod.data());
  out->cst.assign((size_t)PC_COUNT * 2 * L, 0.0);
  auto put_pair = [&](int idx, const BN& v0, const BN& v1) {
    to_limbs52(v0, L, &out->cst[((size_t)idx * 2 + 0) * L]);
    to_limbs52(v1, L, &out->cst[((size_t)idx * 2 + 1) * L]);
  };
  const BN R2 = hbn::mod(hbn::mul(R, R), x2);
  for (int c = 0; c < 4; ++c) {   // W_c ~ 2^(c chunk_bits) R^2 mod x^2 as p-adic digits
    const BN w = hbn::mulmod(hbn::mod(hbn::shl(BN(1), (size_t)c * chunk_bits), x2), R2, x2);
    BN q, r;
    hbn::divmod(w, x, &q, &r);
    put_pair(PC_W0 + c, r, q);
  }
  put_pair(PC_ONE, BN(1), BN());
  put_pair(PC_HR, hbn::mulmod(hx, hbn::mod(R, x), x), BN());
  // program
  auto ins = [](uint32_t op, uint32_t arg) { return op | (arg << 8); };
  std::vector<uint32_t>& P = out->prog;
  P.clear();
  for (int c = 3; c >= 0; --c) { P.push_back(ins(PO_LOADC, c)); P.push_back(ins(PO_YCONST, PC_W0 + c)); P.push_back(ins(PO_MUL, 0)); P.push_back(ins(PO_TX, c)); }
  P.push_back(ins(PO_SUM4, 0));
  const std::vector<uint32_t> sw = build_powm_program(hbn::sub(x, BN(1)));
  constexpr int TS = 1 << (PROG_WS - 1);
  P.push_back(ins(PO_TX, 0)); P.push_back(ins(PO_YX, 0)); P.push_back(ins(PO_SQR, 1)); P.push_back(ins(PO_YX, 0)); P.push_back(ins(PO_XT, 0));
  for (int t = 1; t < TS; ++t) { P.push_back(ins(PO_MUL, 0)); P.push_back(ins(PO_TX, t)); }
  P.push_back(ins(PO_XT, sw[0]));
  const size_t zone = P.size();   // from here on only the running pair is live between products
  for (size_t i = 1; i < sw.size(); ++i) {
    const uint32_t sq = sw[i] >> 8, idx = sw[i] & 0xffu;
    if (sq) P.push_back(ins(PO_SQR, sq));
    if (idx != PROG_NOMUL) { P.push_back(ins(PO_YT, idx)); P.push_back(ins(PO_MUL, 0)); }
  }
  P.push_back(ins(PO_YCONST, PC_ONE)); P.push_back(ins(PO_MUL, 0)); P.push_back(ins(PO_FINISH, 0));
  P.push_back(ins(PO_YCONST, PC_HR)); P.push_back(ins(PO_MUL, 0)); P.push_back(ins(PO_OUT, 0)); P.push_back(ins(PO_END, 0));
  int nseg = PAIR_SEGMENTS;   // PHE_DEC_SEGMENTS=<2..16> at key creation: another slicing (A/B runs); 1 is handled at launch
  if (const char* e = getenv("PHE_DEC_SEGMENTS")) { const int v = atoi(e); if (v >= 2 && v <= PAIR_MAX_SEG) nseg = v; }
  split_pair_program(P, zone, nseg, (uint32_t)TS, &out->segprog, &out->segoff);
  return true;
}

int window_for_bits(int ebits) { return ebits <= 8 ? 1 : (ebits <= 160 ? 3 : 5); }

}  // namespace

// Fixed-base comb table of a DJN key in device memory.  Tables are shared: every phe_pubkey object of the same key
// (same n, hs, randbits, engine, device) and digit width points at ONE table through a process-wide registry, so
// unpickling a key twice -- or holding it once through the C ABI and once through the Python classes -- does not build
// or hold the 35 GB twice.  The table lives as long as some key object uses it.
struct CombTable {
  DevBuf buf;
  int wb = 0, nwin = 0, device = 0;
  double build_ms = 0.0;
  size_t bytes = 0;
  ~CombTable() { buf.release(); }
};
namespace {
std::mutex g_comb_mu;
std::map<std::string, std::weak_ptr<CombTable>> g_comb_reg;
// bytes of comb tables alive on `device` (caller holds g_comb_mu)
size_t comb_live_bytes(int device) {
  size_t tot = 0;
  for (auto it = g_comb_reg.begin(); it != g_comb_reg.end();) {
    if (auto t = it->second.lock()) { if (t->device == device) tot += t->bytes; ++it; }
    else it = g_comb_reg.erase(it);
  }
  return tot;
}
}  // namespace

struct phe_pubkey {
  int bits = 0, n_words = 0, djn = 0, randbits = 0, device = 0;
  BN n, nsq, hs;
  const ShapeOps* ops = nullptr;  // shape of the n^2 context
  const ShapeOps* nops = nullptr; // shape of the n-sized numbers of the n-adic pair engine (npair_items.cuh)
  bool use_npair = false;         // HE mul, DJN encrypt and the comb table run on (X0, X1) pairs mod n
  mutable NPairCtxArgs nctx{};
  mutable DevBuf d_nctx;
  std::vector<uint32_t> h_nctx;
  mutable MontCtxArgs ctx{};
  mutable DevBuf d_ctx;
  std::vector<uint32_t> h_ctx;    // Montgomery block, uploaded lazily
  mutable std::shared_ptr<CombTable> comb;   // the (shared) fixed-base table once built
  int comb_bits_wanted = 0;       // 0: choose from the free device memory
  mutable bool comb_wide = false;
  mutable size_t comb_seen = 0;   // elements encrypted so far under the automatic width (promotion counter)
  mutable DevBuf ws_a, ws_b, ws_c, ws_d, ws_r, ws_tbl, ws_inv;  // op workspaces
  mutable DevBuf ws_idx, ws_rows, ws_rows2, ws_tree[2], ws_ent; // row operations (alignment, inverse of rows, add trees)
  mutable DevBuf d_prog_n;                         // classic scheme: sliding-window program of the exponent n
  std::vector<uint32_t> h_prog_n;
  mutable std::mutex mu;
  mutable StreamChain chain;       // orders the calls that share this key's scratch across streams
  mutable bool dev_ready = false;  // device state (Montgomery block, comb table) is built on first compute call
};

struct phe_privkey {
  const phe_pubkey* pk = nullptr;
  BN p, q;
  const ShapeOps* ops = nullptr;  // shape of the x^2 contexts (also used for p, q, n in the tail)
  mutable DevBuf d_ctx[2], d_exp[2], d_prog[2], d_tail;
  // p-adic pair engine (balanced keys): per x = p, q
  bool use_pair = false;
  PairBlock pairb[2];
  mutable DevBuf d_pair_cst[2], d_pair_prog[2], d_pair_segprog[2];
  mutable MontCtxArgs ctx[2]{};
  int ebits[2] = {0, 0};
  int hw = 0;  // words of x^2 (= n_words)
  uint64_t n0invs[3] = {0, 0, 0};
  mutable DevBuf ws_in, ws_out, ws_mont[2], ws_u[2], ws_tbl, ws_sched, ws_cls, d_n;
  mutable std::mutex mu;
  mutable StreamChain chain;
  mutable bool dev_ready = false;
  std::vector<uint32_t> h_ctx[2], h_exp[2], h_prog[2], h_tail;   // host copies uploaded on first compute call
};

namespace {

// Device state of a key is built on the first compute call (so key objects can be created, inspected and pickled on
// a host without a GPU); every compute entry point goes through these and fails loudly when no device exists.
int pk_ensure_device(const phe_pubkey* pk) {
  if (pk->dev_ready) return 0;
  if (phe_device_count() <= 0) return fail("no CUDA device (the compute path has no CPU fallback)");
  CUDA_TRY(cudaGetDevice(const_cast<int*>(&pk->device)));
  PHE_TRY(upload(pk->d_ctx, pk->h_ctx));
  pk->ctx.entries = reinterpret_cast<const double*>(pk->d_ctx.p);
  if (pk->use_npair) {
    PHE_TRY(upload(pk->d_nctx, pk->h_nctx));
    pk->nctx.entries = reinterpret_cast<const double*>(pk->d_nctx.p);
  }
  pk->dev_ready = true;
  return 0;
}

// Comb table of the DJN obfuscator.  The number of Montgomery products per encrypt is randbits / wb + 2, so the digit
// width wb is as wide as the device memory comfortably allows: nwin * 2^wb entries, i.e. 35 GB at wb = 20 for a
// 2048-bit key (52 + 2 products, ~0.2 s to build) on an idle 180 GB B200, 2.7 GB at wb = 16 (64 + 2 products, 15 ms).
// A key starts on a small table (wb = 12: 225 MB, < 1 ms) and is promoted to the wide one once it has encrypted
// COMB_PROMOTE elements, so that a caller with a handful of values never waits for (or holds) the large table.
// Budget across keys: the automatic width is the widest whose table fits a quarter of the free memory, 40 GB, AND keeps
// the tables of ALL keys on the device within the comb budget (60 % of the device memory, PHE_COMB_BUDGET_GB to
// change): a process with many keys gets narrower tables for the later ones by rule, not by allocation failure.
// PHE_COMB_BITS or phe_pubkey_set_comb_bits pin the width (no budget check then).
constexpr size_t COMB_PROMOTE = 32768;
std::string comb_key(const phe_pubkey* pk, int wb) {
  std::string k;
  auto put = [&](const void* p, size_t n) { k.append(static_cast<const char*>(p), n); };
  const int hdr[5] = {pk->device, pk->use_npair ? 1 : 0, pk->randbits, wb, pk->n_words};
  put(hdr, sizeof(hdr));
  std::vector<uint32_t> w(2 * (size_t)pk->n_words);
  pk->n.to_words(w.data(), pk->n_words); put(w.data(), (size_t)pk->n_words * 4);
  pk->hs.to_words(w.data(), w.size()); put(w.data(), w.size() * 4);
  return k;
}
int pk_ensure_comb(const phe_pubkey* pk, size_t count) {
  if (!pk->djn) return fail("comb table requested for a non-DJN key");
  int wb = pk->comb_bits_wanted;
  if (wb <= 0) { const char* e = getenv("PHE_COMB_BITS"); if (e) wb = atoi(e); }
  const size_t entry_bytes = pk->use_npair ? 2 * EW(pk->nops) * 4 : EW(pk->ops) * 4;
  auto table_bytes = [&](int w) { return (size_t)((pk->randbits + w - 1) / w) * ((size_t)1 << w) * entry_bytes; };
  std::lock_guard<std::mutex> reg_lock(g_comb_mu);
  const bool pinned = wb > 0;
  if (pinned) {
    if (pk->comb) return 0;
  } else {
    pk->comb_seen += count;
    const bool wide = pk->comb_seen >= COMB_PROMOTE;
    if (pk->comb && (pk->comb_wide || !wide)) return 0;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = total_b / 10 * 6;
    if (const char* e = getenv("PHE_COMB_BUDGET_GB")) budget = (size_t)(atof(e) * 1073741824.0);
    size_t live = comb_live_bytes(pk->device);
    if (pk->comb && pk->comb.use_count() == 1) { free_b += pk->comb->bytes; live -= pk->comb->bytes; }   // released first
    wb = wide ? 20 : 12;
    auto shared_already = [&](int w) { auto it = g_comb_reg.find(comb_key(pk, w)); return it != g_comb_reg.end() && !it->second.expired(); };
    while (wb > 8 && !shared_already(wb) &&
           (table_bytes(wb) > free_b / 4 || table_bytes(wb) > ((size_t)40 << 30) || live + table_bytes(wb) > budget)) wb -= 2;
    if (wide) pk->comb_wide = true;
    if (pk->comb && wb <= pk->comb->wb) return 0;
    if (pk->comb) { CUDA_TRY(cudaDeviceSynchronize()); pk->comb.reset(); }
  }
  if (wb < 1) wb = 1;
  if (wb > 22) wb = 22;
  for (;;) {
    {   // another key object of the same key already holds this table: share it
      auto it = g_comb_reg.find(comb_key(pk, wb));
      if (it != g_comb_reg.end()) if (auto t = it->second.lock()) { pk->comb = t; return 0; }
    }
    auto t = std::make_shared<CombTable>();
    // an allocation that fails (fragmented or shared device) falls back to narrower tables instead of failing the encrypt
    if (t->buf.ensure_exact(table_bytes(wb) / 4) != 0) {
      cudaGetLastError();
      if (wb <= 8 || pinned) return fail("comb table: out of device memory");
      wb -= 2;
      continue;
    }
    const auto t_build0 = std::chrono::steady_clock::now();
    t->wb = wb; t->nwin = (pk->randbits + wb - 1) / wb; t->device = pk->device; t->bytes = table_bytes(wb);
    std::vector<uint32_t> hsw(2 * (size_t)pk->n_words);
    pk->hs.to_words(hsw.data(), hsw.size());
    DevBuf dhs;
    PHE_TRY(upload(dhs, hsw));
    cudaError_t e;
    if (pk->use_npair) {
      CombNPairArgs ca{};
      ca.hs_w = dhs.p; ca.chunk_words = pk->n_words; ca.nwin = t->nwin; ca.wb = wb;
      ca.comb = reinterpret_cast<double*>(t->buf.p); ca.ctx = pk->nctx;
      e = pk->nops->comb_build_npair(ca, 0);
    } else {
      CombArgs ca{};
      ca.hs_w = dhs.p; ca.hs_words = 2 * pk->n_words; ca.nwin = t->nwin; ca.wb = wb;
      ca.comb = reinterpret_cast<double*>(t->buf.p); ca.ctx = pk->ctx;
      e = pk->ops->comb_build(ca, 0);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    dhs.release();
    if (e != cudaSuccess) return fail(std::string("comb table build: ") + cudaGetErrorString(e));
    t->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build0).count();
    g_comb_reg[comb_key(pk, wb)] = t;
    pk->comb = t;
    return 0;
  }
}

int sk_ensure_device(const phe_privkey* sk) {
  if (sk->dev_ready) return 0;
  {
    std::lock_guard<std::mutex> lk(sk->pk->mu);
    PHE_TRY(pk_ensure_device(sk->pk));
  }
  for (int y = 0; y < 2; ++y) {
    PHE_TRY(upload(sk->d_ctx[y], sk->h_ctx[y]));
    sk->ctx[y].entries = reinterpret_cast<const double*>(sk->d_ctx[y].p);
    PHE_TRY(upload(sk->d_exp[y], sk->h_exp[y]));
    PHE_TRY(upload(sk->d_prog[y], sk->h_prog[y]));
  }
  PHE_TRY(upload(sk->d_tail, sk->h_tail));
  if (sk->use_pair) {
    for (int y = 0; y < 2; ++y) {
      const PairBlock& b = sk->pairb[y];
      std::vector<uint32_t> tmp(b.cst.size() * 2, 0);
      std::memcpy(tmp.data(), b.cst.data(), b.cst.size() * 8);
      PHE_TRY(upload(sk->d_pair_cst[y], tmp));
      PHE_TRY(upload(sk->d_pair_prog[y], b.prog));
      PHE_TRY(upload(sk->d_pair_segprog[y], b.segprog));
    }
  }
  sk->dev_ready = true;
  return 0;
}

// shared exponent / generic powm launch on the n^2 context of a public key
int launch_powm(const ShapeOps* o, const MontCtxArgs& ctx, const uint32_t* d_base, int base_words, const uint32_t* d_e,
                int e_words, size_t e_stride, int ebits, uint32_t* d_out, int out_words, int count, DevBuf& tbl,
                cudaStream_t s) {
  const int win = window_for_bits(ebits);
  PHE_TRY(tbl.ensure(o->powm_tbl_words(win, 1, count)));
  PowmArgs p{};
  p.base_w = d_base; p.base_words = base_words;
  p.base_mont[0] = p.base_mont[1] = nullptr;
  p.e_w[0] = d_e; p.e_w[1] = nullptr;
  p.e_words = e_words; p.e_stride = e_stride;
  p.ebits[0] = ebits; p.ebits[1] = 0;
  p.out_w[0] = d_out; p.out_w[1] = nullptr;
  p.out_words = out_words; p.count = count;
  p.ctx[0] = ctx; p.ctx[1] = ctx;
  p.tbl = reinterpret_cast<double*>(tbl.p);
  CUDA_TRY(o->powm(win, p, 1, s));
  return 0;
}

constexpr size_t CHUNK = 1u << 20;  // items per launch (bounds scratch and int indexing)

// one DJN comb launch (n^2 Montgomery engine or n-adic pair engine)
int launch_encrypt_comb(const phe_pubkey* pk, const uint32_t* m_w, int m_words, const uint32_t* r_w, int r_words,
                        uint32_t* out_w, int count, cudaStream_t s, uint32_t* const* peers = nullptr, int n_peers = 0,
                        size_t peer_row0 = 0) {
  const int cw = 2 * pk->n_words;
  if (n_peers && !pk->use_npair) return fail("phe_encrypt_dev_multi needs the n-adic pair engine");
  if (pk->use_npair) {
    EncNPairArgs a{};
    a.n_peers = n_peers;
    for (int k = 0; k < n_peers; ++k) a.peer_out[k] = peers[k] + peer_row0 * cw;
    a.m_w = m_w; a.m_words = m_words; a.r_w = r_w; a.r_words = r_words;
    a.nwin = pk->comb ? pk->comb->nwin : 0; a.wb = pk->comb ? pk->comb->wb : 0;
    a.out_w = out_w; a.out_words = cw; a.count = count; a.ctx = pk->nctx;
    a.comb = pk->comb ? reinterpret_cast<const double*>(pk->comb->buf.p) : nullptr;
    CUDA_TRY(pk->nops->encrypt_npair(a, s));
    return 0;
  }
  EncCombArgs a{};
  a.m_w = m_w; a.m_words = m_words; a.r_w = r_w; a.r_words = r_words;
  a.nwin = pk->comb ? pk->comb->nwin : 0; a.wb = pk->comb ? pk->comb->wb : 0;
  a.out_w = out_w; a.out_words = cw; a.count = count; a.ctx = pk->ctx;
  a.comb = pk->comb ? reinterpret_cast<const double*>(pk->comb->buf.p) : nullptr;
  CUDA_TRY(pk->ops->encrypt_comb(a, s));
  return 0;
}

// obf[i] for r[i] into d_obf (canonical ciphertext words).  d_r: device, r_words per item.
int obfuscators_dev(const phe_pubkey* pk, const uint32_t* d_r, int r_words, size_t count, uint32_t* d_obf,
                    cudaStream_t s) {
  const int cw = 2 * pk->n_words;
  if (pk->djn) {
    // comb kernel with m = 0 gives (1 + 0) * obf
    PHE_TRY(pk_ensure_comb(pk, count));
    PHE_TRY(pk->ws_d.ensure((size_t)pk->n_words));
    CUDA_TRY(cudaMemsetAsync(pk->ws_d.p, 0, (size_t)pk->n_words * 4, s));
    for (size_t off = 0; off < count; off += CHUNK) {
      const int c = (int)std::min(CHUNK, count - off);
      // zero words: m = 0 for every item
      PHE_TRY(launch_encrypt_comb(pk, pk->ws_d.p, 0, d_r + off * r_words, r_words, d_obf + off * cw, c, s));
    }
    return 0;
  }
  // classic: r^n mod n^2, shared exponent n -> sliding-window program
  if (!pk->d_prog_n.p) PHE_TRY(upload(pk->d_prog_n, pk->h_prog_n));
  if (pk->use_npair && r_words <= pk->n_words) {   // r < n: one chunk
    for (size_t off = 0; off < count; off += CHUNK) {
      const int c = (int)std::min(CHUNK, count - off);
      PHE_TRY(pk->ws_tbl.ensure(pk->nops->powm_prog_npair_tbl_words(c)));
      ProgNPairArgs a{};
      a.c_w = d_r + off * r_words; a.chunk_words = r_words; a.nchunks = 1;
      a.prog = pk->d_prog_n.p; a.nprog = (int)pk->h_prog_n.size() - 1;
      a.out_w = d_obf + off * cw; a.out_words = cw; a.count = c; a.ctx = pk->nctx;
      a.tbl = reinterpret_cast<double*>(pk->ws_tbl.p);
      CUDA_TRY(pk->nops->powm_prog_npair(a, s));
    }
    return 0;
  }
  for (size_t off = 0; off < count; off += CHUNK) {
    const int c = (int)std::min(CHUNK, count - off);
    PHE_TRY(pk->ws_tbl.ensure(pk->ops->powm_prog_tbl_words(1, c)));
    PowmArgs p{};
    p.base_w = d_r + off * r_words; p.base_words = r_words;
    p.out_w[0] = d_obf + off * cw; p.out_words = cw; p.count = c;
    p.ctx[0] = pk->ctx; p.ctx[1] = pk->ctx;
    p.tbl = reinterpret_cast<double*>(pk->ws_tbl.p);
    p.prog[0] = pk->d_prog_n.p; p.nprog[0] = (int)pk->h_prog_n.size() - 1;
    CUDA_TRY(pk->ops->powm_prog(p, 1, s));
  }
  return 0;
}

int make_r_host(const phe_pubkey* pk, size_t count, std::vector<uint32_t>& r, int* r_words) {
  if (pk->djn) {
    *r_words = (pk->randbits + 31) / 32;
    r.resize(count * (size_t)*r_words);
    PHE_TRY(random_words(r.data(), r.size()));
    if (pk->randbits & 31) {
      const uint32_t mask = (1u << (pk->randbits & 31)) - 1u;
      for (size_t i = 0; i < count; ++i) r[i * *r_words + *r_words - 1] &= mask;
    }
    return 0;
  }
  // classic: r = random(bits) mod (n - 1) + 1   (ipcl getNormalObfuscator)
  *r_words = pk->n_words;
  r.assign(count * (size_t)pk->n_words, 0);
  const BN nm1 = hbn::sub(pk->n, BN(1));
  for (size_t i = 0; i < count; ++i) {
    const BN v = hbn::add(hbn::mod(random_bits(pk->bits), nm1), BN(1));
    v.to_words(&r[i * pk->n_words], pk->n_words);
  }
  return 0;
}

// DJN obfuscator exponents drawn on the device: ChaCha20 keystream under a key and nonce fresh from getrandom(2) for
// every call (csrc/chacha20.cu), top word of every r masked to randbits.  Result in pk->ws_r.
int random_r_dev(const phe_pubkey* pk, size_t count, int* r_words, cudaStream_t s) {
  *r_words = (pk->randbits + 31) / 32;
  const size_t words = count * (size_t)*r_words;
  PHE_TRY(pk->ws_r.ensure(words));
  uint32_t seed[11];
  PHE_TRY(random_words(seed, 11));
  const uint32_t mask = (pk->randbits & 31) ? ((1u << (pk->randbits & 31)) - 1u) : 0xffffffffu;
  CUDA_TRY(chacha20_fill(pk->ws_r.p, words, seed, seed + 8, 0u, *r_words, mask, s));
  return 0;
}

int encrypt_dev_impl(const phe_pubkey* pk, const uint32_t* d_m, size_t count, const uint32_t* d_r, int r_words,
                     uint32_t* d_ct, cudaStream_t s, int m_words = 0, uint32_t* const* peers = nullptr, int n_peers = 0) {
  const int cw = 2 * pk->n_words;
  if (m_words <= 0) m_words = pk->n_words;   // stride of the plaintext rows
  if (count == 0) return 0;
  if (!d_r || pk->djn) {   // d_r == nullptr: make_secure = 0, ct = 1 + m n
    if (d_r) {
      PHE_TRY(pk_ensure_comb(pk, count));
      if ((size_t)r_words * 32 < (size_t)pk->randbits) return fail("phe_encrypt: r_words too small for randbits");
    }
    for (size_t off = 0; off < count; off += CHUNK) {
      const int c = (int)std::min(CHUNK, count - off);
      PHE_TRY(launch_encrypt_comb(pk, d_m + off * (size_t)m_words, m_words, d_r ? d_r + off * r_words : nullptr, r_words,
                                  d_ct + off * cw, c, s, peers, n_peers, off));
    }
    return 0;
  }
  // classic scheme: obf = r^n, then ct = (1 + m n) * obf
  PHE_TRY(pk->ws_c.ensure(std::min(count, CHUNK) * cw));
  for (size_t off = 0; off < count; off += CHUNK) {
    const int c = (int)std::min(CHUNK, count - off);
    PHE_TRY(obfuscators_dev(pk, d_r + off * r_words, r_words, c, pk->ws_c.p, s));
    EncFinishArgs f{};
    f.m_w = d_m + off * (size_t)m_words; f.m_words = m_words; f.obf_w = pk->ws_c.p;
    f.out_w = d_ct + off * cw; f.out_words = cw; f.count = c; f.ctx = pk->ctx;
    CUDA_TRY(pk->ops->encrypt_finish(f, s));
  }
  return 0;
}

int add_dev_impl(const phe_pubkey* pk, const uint32_t* d_a, size_t na, const uint32_t* d_b, size_t nb, uint32_t* d_out,
                 cudaStream_t s) {
  if (nb != na && nb != 1) return fail("phe_add: size mismatch (b must have na or 1 elements)");
  const int cw = 2 * pk->n_words;
  if (nb == 1 && na > 1) {
    // broadcast: bring the single operand into the Montgomery domain once (b R mod n^2), then every element is ONE
    // Montgomery product a[i] * (b R) * R^-1 instead of two
    PHE_TRY(pk->ws_ent.ensure((size_t)cw));
    Modmul1Args c{};
    c.a = d_b; c.b_w = nullptr; c.b_stride = 0; c.b_entry = pk->ctx.entries + (size_t)ME_R2 * pk->ops->KP;
    c.out = pk->ws_ent.p; c.nwords = cw; c.count = 1; c.ctx = pk->ctx;
    CUDA_TRY(pk->ops->modmul1(c, s));
    for (size_t off = 0; off < na; off += CHUNK) {
      Modmul1Args a{};
      a.a = d_a + off * cw; a.b_w = pk->ws_ent.p; a.b_stride = 0; a.b_entry = nullptr;
      a.out = d_out + off * cw; a.nwords = cw; a.count = (int)std::min(CHUNK, na - off); a.ctx = pk->ctx;
      CUDA_TRY(pk->ops->modmul1(a, s));
    }
    return 0;
  }
  for (size_t off = 0; off < na; off += CHUNK) {
    const int c = (int)std::min(CHUNK, na - off);
    const bool bc = (nb == 1 && na != 1);
    CUDA_TRY(pk->ops->modmul(d_a + off * cw, bc ? d_b : d_b + off * cw, bc ? 0 : (size_t)cw, d_out + off * cw, cw, c,
                             pk->ctx, s));
  }
  return 0;
}

// out[i] = a[i]^-1 mod n^2 for count device rows (Montgomery's trick, recursive over block totals; the last <= 16
// values are inverted on the host).  Fails with "not invertible" if any a[i] shares a factor with n.
// Batched modular inverse (Montgomery's trick), level by level: every level turns `count` elements into count / block
// block totals (k_inv_block, prefix), the totals are inverted one level down, and the way back (unwind) turns the inverse
// of a total into the inverses of its elements.  Block sizes: ~count / resident groups (>= 4) while there is parallel
// work, then ~sqrt(count), then one block, so that exactly ONE number is left for the host's extended Euclid (3 ms at
// 4096 bits).  All levels live in one persistent workspace of the key (no cudaMalloc / cudaFree per call).
int invert_dev_impl(const phe_pubkey* pk, const uint32_t* d_a, size_t count, uint32_t* d_out, cudaStream_t s) {
  const ShapeOps* o = pk->ops;
  const size_t cw = 2 * (size_t)pk->n_words;
  if (count == 0) return 0;
  struct Lvl { size_t count, block, nblocks, padded, in_pad, out_pad, P, totals, tinv; };
  std::vector<Lvl> lv;
  const size_t groups = (size_t)o->resident_groups();
  size_t words = 0;
  auto carve = [&](size_t n) { const size_t at = words; words += (n + 3) & ~(size_t)3; return at; };
  for (size_t c = count; c > 1;) {
    Lvl l{};
    l.count = c;
    if (c <= 32) l.block = c;
    else if (c <= 1024) { l.block = 1; while (l.block * l.block < c) ++l.block; }
    else { l.block = (c + groups - 1) / groups; if (l.block < 4) l.block = 4; if (l.block > 256) l.block = 256; }
    l.nblocks = (c + l.block - 1) / l.block;
    l.padded = l.nblocks * l.block;
    if (l.padded != c) { l.in_pad = carve(l.padded * cw); l.out_pad = carve(l.padded * cw); }
    l.P = carve(l.padded * EW(o));
    l.totals = carve(l.nblocks * cw);
    l.tinv = carve(l.nblocks * cw);
    lv.push_back(l);
    c = l.nblocks;
  }
  const size_t base = carve(cw);   // the single value the host inverts when count == 1
  PHE_TRY(pk->ws_inv.ensure(words));
  uint32_t* W = pk->ws_inv.p;
  // forward: prefix products and block totals
  const uint32_t* src = d_a;
  std::vector<const uint32_t*> srcs(lv.size());
  std::vector<uint32_t*> dsts(lv.size());
  for (size_t k = 0; k < lv.size(); ++k) {
    const Lvl& l = lv[k];
    uint32_t* dst = (k == 0) ? d_out : W + lv[k - 1].tinv;
    if (l.padded != l.count) {   // pad with ones so that every block has the same length
      std::vector<uint32_t> ones((l.padded - l.count) * cw, 0);
      for (size_t i = 0; i < l.padded - l.count; ++i) ones[i * cw] = 1;
      CUDA_TRY(cudaMemcpyAsync(W + l.in_pad, src, l.count * cw * 4, cudaMemcpyDeviceToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(W + l.in_pad + l.count * cw, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));   // `ones` is a stack-lifetime staging buffer
      src = W + l.in_pad;
    }
    srcs[k] = src; dsts[k] = dst;
    InvArgs a{};
    a.c_w = src; a.nwords = (int)cw; a.count = (int)l.padded; a.block = (int)l.block;
    a.P = reinterpret_cast<double*>(W + l.P); a.total_w = W + l.totals; a.ctx = pk->ctx;
    if (o->inv_block(a, false, s) != cudaSuccess) return fail("phe_invert: prefix kernel launch failed");
    src = W + l.totals;
  }
  // the one inverse left: on the host
  {
    const uint32_t* one_src = lv.empty() ? d_a : W + lv.back().totals;
    uint32_t* one_dst = lv.empty() ? d_out : W + lv.back().tinv;
    std::vector<uint32_t> h(cw);
    CUDA_TRY(cudaMemcpyAsync(h.data(), one_src, cw * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    try {
      hbn::modinv(BN::from_words(h.data(), cw), pk->nsq).to_words(h.data(), cw);
    } catch (const std::exception&) { return fail("phe_invert: an element is not invertible modulo n^2"); }
    CUDA_TRY(cudaMemcpyAsync(W + base, h.data(), cw * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(one_dst, W + base, cw * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));     // h is a stack-lifetime staging buffer
  }
  // backward: unwind every level
  for (size_t k = lv.size(); k-- > 0;) {
    const Lvl& l = lv[k];
    InvArgs a{};
    a.c_w = srcs[k]; a.nwords = (int)cw; a.count = (int)l.padded; a.block = (int)l.block;
    a.P = reinterpret_cast<double*>(W + l.P); a.tinv_w = W + l.tinv;
    a.out_w = (l.padded != l.count) ? W + l.out_pad : dsts[k]; a.ctx = pk->ctx;
    if (o->inv_block(a, true, s) != cudaSuccess) return fail("phe_invert: unwind kernel launch failed");
    if (l.padded != l.count)
      CUDA_TRY(cudaMemcpyAsync(dsts[k], W + l.out_pad, l.count * cw * 4, cudaMemcpyDeviceToDevice, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

// ---- row operations on device-resident ciphertext matrices ----------------------------------------------------------
// host index list -> device (int64), in the key's index workspace at word offset `at`
int upload_idx(const phe_pubkey* pk, const long long* idx, size_t n, size_t rows_limit, size_t at, cudaStream_t s) {
  for (size_t i = 0; i < n; ++i)
    if (idx[i] < 0 || (rows_limit && (size_t)idx[i] >= rows_limit)) return fail("row index out of range");
  PHE_TRY(pk->ws_idx.ensure(at + 2 * n + 2));
  CUDA_TRY(cudaMemcpyAsync(pk->ws_idx.p + at, idx, n * 8, cudaMemcpyHostToDevice, s));
  return 0;
}

int move_rows_impl(const phe_pubkey* pk, const uint32_t* d_src, const long long* idx, size_t n, size_t rows_limit,
                   uint32_t* d_dst, int scatter, cudaStream_t s) {
  if (n == 0) return 0;
  PHE_TRY(upload_idx(pk, idx, n, rows_limit, 0, s));
  CUDA_TRY(rows_move(d_src, d_dst, reinterpret_cast<const long long*>(pk->ws_idx.p), (long long)n, 2 * pk->n_words, scatter, s));
  CUDA_TRY(cudaStreamSynchronize(s));   // idx may be a temporary of the caller (pageable copy)
  return 0;
}

int mul_dev_impl(const phe_pubkey* pk, const uint32_t* d_ct, size_t n, const uint32_t* d_e, int e_words, size_t ne,
                 int exp_bits, uint32_t* d_out, cudaStream_t s);

// ct[idx[i]] <- ct[idx[i]]^(2^delta[i]) in place (distinct rows)
int scale_rows_impl(const phe_pubkey* pk, uint32_t* d_ct, size_t rows, const long long* idx, const int* delta, size_t n,
                    cudaStream_t s) {
  if (n == 0) return 0;
  const int cw = 2 * pk->n_words;
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) { order[i] = i; if (delta[i] < 0) return fail("phe_scale_rows: negative delta"); }
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return delta[a] > delta[b]; });
  std::vector<long long> sidx(n);
  std::vector<int> sdelta(n);
  for (size_t i = 0; i < n; ++i) { sidx[i] = idx[order[i]]; sdelta[i] = delta[order[i]]; }
  {   // distinct rows: two lane groups must not work on one row
    std::vector<long long> chk(sidx);
    std::sort(chk.begin(), chk.end());
    if (std::adjacent_find(chk.begin(), chk.end()) != chk.end()) return fail("phe_scale_rows: duplicate row index");
  }
  const size_t dat = 2 * n + 2;
  PHE_TRY(pk->ws_idx.ensure(dat + n));        // both lists: a later ensure() would move the indices already uploaded
  PHE_TRY(upload_idx(pk, sidx.data(), n, rows, 0, s));
  CUDA_TRY(cudaMemcpyAsync(pk->ws_idx.p + dat, sdelta.data(), n * 4, cudaMemcpyHostToDevice, s));
  const long long* d_idx = reinterpret_cast<const long long*>(pk->ws_idx.p);
  const int* d_delta = reinterpret_cast<const int*>(pk->ws_idx.p + dat);
  if (pk->use_npair) {
    for (size_t off = 0; off < n; off += CHUNK) {
      ScaleNPairArgs a{};
      a.ct = d_ct; a.chunk_words = pk->n_words; a.idx = d_idx + off; a.delta = d_delta + off;
      a.count = (int)std::min(CHUNK, n - off); a.ctx = pk->nctx;
      CUDA_TRY(pk->nops->scale_npair(a, s));
    }
  } else {   // n^2 Montgomery engine: gather, raise to the plaintext 2^delta, scatter
    const int ew = sdelta[0] / 32 + 1;
    std::vector<uint32_t> e(n * (size_t)ew, 0u);
    for (size_t i = 0; i < n; ++i) e[i * ew + sdelta[i] / 32] = 1u << (sdelta[i] % 32);
    PHE_TRY(pk->ws_rows.ensure(n * cw));
    PHE_TRY(pk->ws_rows2.ensure(n * cw));
    PHE_TRY(pk->ws_d.ensure(n * (size_t)ew));
    CUDA_TRY(cudaMemcpyAsync(pk->ws_d.p, e.data(), e.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(rows_move(d_ct, pk->ws_rows.p, d_idx, (long long)n, cw, 0, s));
    PHE_TRY(mul_dev_impl(pk, pk->ws_rows.p, n, pk->ws_d.p, ew, n, sdelta[0] + 1, pk->ws_rows2.p, s));
    CUDA_TRY(rows_move(pk->ws_rows2.p, d_ct, d_idx, (long long)n, cw, 1, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));   // sidx / sdelta are stack-lifetime staging buffers
  return 0;
}

int invert_dev_impl(const phe_pubkey* pk, const uint32_t* d_a, size_t count, uint32_t* d_out, cudaStream_t s);

// ct[idx[i]] <- ct[idx[i]]^-1 mod n^2 in place
int invert_rows_impl(const phe_pubkey* pk, uint32_t* d_ct, size_t rows, const long long* idx, size_t n, cudaStream_t s) {
  if (n == 0) return 0;
  const int cw = 2 * pk->n_words;
  PHE_TRY(upload_idx(pk, idx, n, rows, 0, s));
  const long long* d_idx = reinterpret_cast<const long long*>(pk->ws_idx.p);
  for (size_t off = 0; off < n; off += CHUNK) {
    const size_t c = std::min(CHUNK, n - off);
    PHE_TRY(pk->ws_rows.ensure(c * cw));
    PHE_TRY(pk->ws_rows2.ensure(c * cw));
    CUDA_TRY(rows_move(d_ct, pk->ws_rows.p, d_idx + off, (long long)c, cw, 0, s));
    PHE_TRY(invert_dev_impl(pk, pk->ws_rows.p, c, pk->ws_rows2.p, s));
    CUDA_TRY(rows_move(pk->ws_rows2.p, d_ct, d_idx + off, (long long)c, cw, 1, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

// out[g] = prod_j ct[g * width + j] mod n^2: a tree of one-product levels, the R^-1 factors repaid at the root
int segsum_impl(const phe_pubkey* pk, const uint32_t* d_ct, size_t groups, size_t width, uint32_t* d_out, cudaStream_t s) {
  if (groups == 0) return 0;
  if (width == 0) return fail("phe_segsum: width must be >= 1");
  const int cw = 2 * pk->n_words;
  if (groups > (size_t)1 << 30 || width > (size_t)1 << 30 || groups * ((width + 1) / 2) > (size_t)1 << 31)
    return fail("phe_segsum: too many rows for one call");
  if (width == 1) {
    CUDA_TRY(cudaMemcpyAsync(d_out, d_ct, groups * cw * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const ShapeOps* o = pk->ops;
  const size_t half_rows = groups * ((width + 1) / 2);
  PHE_TRY(pk->ws_tree[0].ensure(half_rows * cw));
  if (width > 2) PHE_TRY(pk->ws_tree[1].ensure(groups * (((width + 1) / 2 + 1) / 2) * cw));
  const uint32_t* cur = d_ct;
  size_t w = width;
  int flip = 0;
  while (w > 1) {
    TreeLevelArgs a{};
    a.src = cur; a.dst = pk->ws_tree[flip].p; a.nwords = cw; a.groups = (int)groups; a.w = (int)w; a.ctx = pk->ctx;
    CUDA_TRY(o->tree_level(a, s));
    cur = pk->ws_tree[flip].p;
    flip ^= 1;
    w = w / 2 + (w & 1);
  }
  // root: width - 1 products left R^-(width - 1) behind; one more product by R^width mod n^2 repays them
  const BN Rm = hbn::mod(hbn::shl(BN(1), o->capacity_bits), pk->nsq);
  std::vector<uint32_t> ww(2, 0u);
  ww[0] = (uint32_t)(width & 0xffffffffu); ww[1] = (uint32_t)((uint64_t)width >> 32);
  const BN Rw = hbn::modexp(Rm, BN::from_words(ww.data(), 2), pk->nsq);
  std::vector<uint32_t> ent(EW(o));
  to_entry(Rw, o, ent.data());
  PHE_TRY(pk->ws_ent.ensure(EW(o) + 2 * (size_t)cw));
  uint32_t* d_ent = pk->ws_ent.p + 2 * (size_t)cw;   // (the first rows of ws_ent serve the broadcast add)
  d_ent += ((4 - (reinterpret_cast<uintptr_t>(d_ent) / 4) % 4) % 4);   // 16-byte aligned
  CUDA_TRY(cudaMemcpyAsync(d_ent, ent.data(), ent.size() * 4, cudaMemcpyHostToDevice, s));
  Modmul1Args f{};
  f.a = cur; f.b_w = nullptr; f.b_stride = 0; f.b_entry = reinterpret_cast<const double*>(d_ent);
  f.out = d_out; f.nwords = cw; f.count = (int)groups; f.ctx = pk->ctx;
  CUDA_TRY(o->modmul1(f, s));
  CUDA_TRY(cudaStreamSynchronize(s));   // ent is a stack-lifetime staging buffer
  return 0;
}

int mul_dev_impl(const phe_pubkey* pk, const uint32_t* d_ct, size_t n, const uint32_t* d_e, int e_words, size_t ne,
                 int exp_bits, uint32_t* d_out, cudaStream_t s) {
  if (ne != n && ne != 1) return fail("phe_mul: size mismatch (exponents must have n or 1 elements)");
  if (e_words < 1) return fail("phe_mul: e_words must be >= 1");
  const int cw = 2 * pk->n_words;
  int ebits = exp_bits > 0 ? exp_bits : e_words * 32;
  if (ebits > e_words * 32) ebits = e_words * 32;
  const bool bc = (ne == 1 && n != 1);
  if (pk->use_npair) {
    const int win = window_for_bits(ebits);
    for (size_t off = 0; off < n; off += CHUNK) {
      const int c = (int)std::min(CHUNK, n - off);
      PHE_TRY(pk->ws_tbl.ensure(pk->nops->mul_npair_tbl_words(win, c)));
      MulNPairArgs a{};
      a.c_w = d_ct + off * cw; a.chunk_words = pk->n_words;
      a.e_w = bc ? d_e : d_e + off * e_words; a.e_words = e_words; a.e_stride = bc ? 0 : (size_t)e_words; a.ebits = ebits;
      a.out_w = d_out + off * cw; a.count = c; a.ctx = pk->nctx; a.tbl = reinterpret_cast<double*>(pk->ws_tbl.p);
      CUDA_TRY(pk->nops->mul_npair(win, a, s));
    }
    return 0;
  }
  for (size_t off = 0; off < n; off += CHUNK) {
    const int c = (int)std::min(CHUNK, n - off);
    PHE_TRY(launch_powm(pk->ops, pk->ctx, d_ct + off * cw, cw, bc ? d_e : d_e + off * e_words, e_words,
                        bc ? 0 : (size_t)e_words, ebits, d_out + off * cw, cw, c, pk->ws_tbl, s));
  }
  return 0;
}

// decrypt = two CRT halves m_p, m_q on the pair engine (one lane each) + recombination.
// Launches of more units than resident warps run time-sliced (PAIR_SEGMENTS segments per unit, phe_kernels.cuh); the
// window tables are then per unit (10 KB per ciphertext and modulus at 2048-bit keys), so a launch covers at most
// PAIR_CHUNK ciphertexts (2.8 GB of tables).  PHE_DEC_SEGMENTS=1 runs whole units, any other value forces the slicing
// for every batch size (tests; the number of segments itself is fixed when the key is created).
constexpr size_t PAIR_CHUNK = 1u << 17;
int decrypt_pair_impl(const phe_privkey* sk, const uint32_t* d_ct, size_t count, uint32_t* d_m, cudaStream_t s) {
  const ShapeOps* o = sk->ops;
  const PairOps* po = pair_ops(sk->pairb[0].L);
  const int hw = sk->hw, cw = 2 * hw, half = hw / 2;
  size_t chunk = std::min(count, PAIR_CHUNK);
  constexpr int slots = 1 << (PROG_WS - 1);
  // the per-unit tables are the one large scratch of the path: on a device that cannot spare them (other keys' comb
  // tables, the caller's own buffers) run shorter launches rather than fail
  while (sk->ws_tbl.ensure(po->tbl_words((int)chunk, slots)) != 0) {
    if (chunk <= 4096) return 1;    // (the error text of the last cudaMalloc stands)
    chunk = (chunk / 2 + 31) & ~(size_t)31;
  }
  for (int y = 0; y < 2; ++y) PHE_TRY(sk->ws_u[y].ensure(chunk * half));
  PHE_TRY(sk->ws_sched.ensure(po->sched_ints((int)chunk)));
  int want_seg = 0;
  if (const char* e = getenv("PHE_DEC_SEGMENTS")) want_seg = atoi(e);
  for (size_t off = 0; off < count; off += chunk) {
    const int c = (int)std::min(chunk, count - off);
    const int units = 2 * ((c + 31) / 32);
    const bool sliced = want_seg > 0 ? want_seg > 1 : units > po->warps(c);
    DecPairArgs a{};
    a.c_w = d_ct + off * cw; a.c_words = cw; a.chunk_words = half; a.out_words = half; a.count = c; a.slots = slots;
    a.nseg = sliced ? (int)sk->pairb[0].segoff.size() : 1;   // (equal for p and q: phe_privkey_create pads)
    for (int y = 0; y < 2; ++y) {
      a.prog[y] = sliced ? sk->d_pair_segprog[y].p : sk->d_pair_prog[y].p; a.out_w[y] = sk->ws_u[y].p;
      for (int k = 0; k < a.nseg; ++k) a.seg_off[y][k] = sliced ? sk->pairb[y].segoff[k] : 0;
      a.cst[y] = reinterpret_cast<const double*>(sk->d_pair_cst[y].p);
      a.n0inv[y] = sk->pairb[y].n0inv;
    }
    a.tbl = reinterpret_cast<double*>(sk->ws_tbl.p);
    a.sched = reinterpret_cast<int*>(sk->ws_sched.p);
    CUDA_TRY(po->dec_pair(a, sk->pairb[0].mod.data(), sk->pairb[1].mod.data(), s));
    DecCrtArgs t{};
    t.mp_w = sk->ws_u[0].p; t.mq_w = sk->ws_u[1].p; t.half_words = half; t.m_w = d_m + off * hw; t.m_words = hw;
    t.count = c; t.cst = reinterpret_cast<const double*>(sk->d_tail.p);
    for (int i = 0; i < 3; ++i) t.n0invs[i] = sk->n0invs[i];
    CUDA_TRY(o->dec_crt(t, s));
  }
  return 0;
}

int decrypt_dev_impl(const phe_privkey* sk, const uint32_t* d_ct, size_t count, uint32_t* d_m, cudaStream_t s) {
  if (sk->use_pair) return decrypt_pair_impl(sk, d_ct, count, d_m, s);
  const ShapeOps* o = sk->ops;
  const int hw = sk->hw, cw = 2 * hw;
  const size_t chunk = std::min(count, CHUNK);
  for (int y = 0; y < 2; ++y) {
    PHE_TRY(sk->ws_mont[y].ensure(chunk * EW(o)));
    PHE_TRY(sk->ws_u[y].ensure(chunk * hw));
  }
  PHE_TRY(sk->ws_tbl.ensure(o->powm_prog_tbl_words(2, (int)chunk)));
  for (size_t off = 0; off < count; off += CHUNK) {
    const int c = (int)std::min(CHUNK, count - off);
    DecPrepArgs dp{};
    dp.c_w = d_ct + off * cw; dp.hw = hw; dp.count = c;
    for (int y = 0; y < 2; ++y) { dp.out[y] = reinterpret_cast<double*>(sk->ws_mont[y].p); dp.ctx[y] = sk->ctx[y]; }
    CUDA_TRY(o->dec_prep(dp, s));
    PowmArgs p{};
    p.base_w = nullptr; p.base_words = 0; p.e_words = hw; p.e_stride = 0; p.out_words = hw; p.count = c;
    for (int y = 0; y < 2; ++y) {
      p.base_mont[y] = reinterpret_cast<const double*>(sk->ws_mont[y].p); p.e_w[y] = sk->d_exp[y].p; p.ebits[y] = sk->ebits[y];
      p.out_w[y] = sk->ws_u[y].p; p.ctx[y] = sk->ctx[y];
      p.prog[y] = sk->d_prog[y].p; p.nprog[y] = (int)sk->h_prog[y].size() - 1;
    }
    p.tbl = reinterpret_cast<double*>(sk->ws_tbl.p);
    CUDA_TRY(o->powm_prog(p, 2, s));
    DecTailArgs t{};
    t.up_w = sk->ws_u[0].p; t.uq_w = sk->ws_u[1].p; t.u_words = hw; t.m_w = d_m + off * hw; t.m_words = hw;
    t.count = c; t.cst = reinterpret_cast<const double*>(sk->d_tail.p);
    for (int i = 0; i < 3; ++i) t.n0invs[i] = sk->n0invs[i];
    CUDA_TRY(o->dec_tail(t, s));
  }
  return 0;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

const char* phe_last_error(void) { return t_err.c_str(); }
const char* phe_version(void) { return "phe_b200 0.2 (sm_100a, radix-2^52 DFMA Montgomery)"; }

int phe_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int phe_set_device(int device) { CUDA_TRY(cudaSetDevice(device)); return 0; }
int phe_get_device(void) { int d = -1; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; } return d; }
unsigned long long phe_kernel_launches(void) { return launch_counter(); }

int phe_timing_enable(int on) { phe::timing_set(on != 0); return 0; }
int phe_timing_read(int kind, double* ms_total, unsigned long long* launches) {
  if (!phe::timing_read(kind, ms_total, launches)) return fail("phe_timing_read: unknown kernel kind");
  return 0;
}
const char* phe_timing_kind_name(int kind) {
  static const char* names[KK_COUNT] = {"k_modmul", "k_powm", "k_dec_prep", "k_dec_tail", "k_encrypt_comb",
                                        "k_encrypt_finish", "k_comb_build", "k_dec_pair", "k_dec_crt", "k_encrypt_npair",
                                        "k_mul_npair", "k_rows_move"};
  return (kind >= 0 && kind < KK_COUNT) ? names[kind] : nullptr;
}

int phe_pubkey_create(const uint32_t* n, int n_words, int bits, int djn, const uint32_t* hs, int randbits,
                      phe_pubkey** out) {
  try {
    if (!n || !out || n_words <= 0) return fail("phe_pubkey_create: bad arguments");
    BN N = BN::from_words(n, n_words);
    if (!N.is_odd() || N.bits() < 16) return fail("phe_pubkey_create: n must be odd and non-trivial");
    if ((int)N.bits() > bits) bits = (int)N.bits();
    if (n_words * 32 < bits) return fail("phe_pubkey_create: n_words too small");
    std::unique_ptr<phe_pubkey> pk(new phe_pubkey);
    pk->bits = bits; pk->n_words = n_words; pk->djn = djn ? 1 : 0; pk->n = N; pk->nsq = hbn::mul(N, N);
    pk->ops = shape_for_bits(2 * n_words * 32);
    if (!pk->ops) return fail("phe_pubkey_create: key too large (n^2 up to 8192 bits supported)");
    const BN R = hbn::shl(BN(1), pk->ops->capacity_bits);
    pk->h_ctx = mont_block(pk->nsq, hbn::mulmod(N, hbn::mod(R, pk->nsq), pk->nsq), pk->ops, &pk->ctx.n0inv);
    {   // n-adic pair engine: needs n to fill its words (the chunks of a ciphertext are n_words words each)
      const char* off = getenv("PHE_NO_NPAIR_ENGINE");
      pk->nops = shape_for_bits(n_words * 32);
      if (pk->nops && !(off && off[0] == '1')) {
        pk->h_nctx = npair_block(N, n_words, pk->nops, &pk->nctx.n0inv, &pk->nctx.d_top);
        pk->use_npair = true;
      }
    }
    if (djn) {
      pk->randbits = randbits > 0 ? randbits : bits / 2;
      if (hs) {
        pk->hs = BN::from_words(hs, 2 * (size_t)n_words);
        if (!(pk->hs < pk->nsq)) return fail("phe_pubkey_create: hs >= n^2");
      } else {
        // ipcl PublicKey::enableDJN: x random with gcd(x, n) = 1, hs = (-x^2 mod n)^n mod n^2 (once per key, host)
        BN x, g;
        do { x = random_bits((size_t)bits + 128); g = hbn::gcd(x, N); } while (!(g == BN(1)));
        const BN xm = hbn::mod(x, N);
        const BN h = hbn::sub(N, hbn::mulmod(xm, xm, N));
        pk->hs = hbn::modexp(h, N, pk->nsq);
      }
    } else {
      pk->h_prog_n = build_powm_program(N);
    }
    *out = pk.release();
    return 0;
  } catch (const std::exception& e) { return fail(std::string("phe_pubkey_create: ") + e.what()); }
}

void phe_pubkey_destroy(phe_pubkey* pk) {
  if (!pk) return;
  for (DevBuf* b : {&pk->ws_idx, &pk->ws_rows, &pk->ws_rows2, &pk->ws_tree[0], &pk->ws_tree[1], &pk->ws_ent, &pk->ws_inv, &pk->d_ctx, &pk->d_nctx, &pk->d_prog_n, &pk->ws_r, &pk->ws_a, &pk->ws_b, &pk->ws_c, &pk->ws_d, &pk->ws_tbl}) b->release();
  pk->chain.release();
  { std::lock_guard<std::mutex> rl(g_comb_mu); pk->comb.reset(); }
  delete pk;
}
int phe_pubkey_set_comb_bits(phe_pubkey* pk, int bits) {
  if (!pk) return fail("null");
  if (bits < 0 || bits > 22) return fail("phe_pubkey_set_comb_bits: bits must be in [0, 22] (0 = automatic)");
  std::lock_guard<std::mutex> lk(pk->mu);
  if (pk->comb && bits != pk->comb->wb) { cudaDeviceSynchronize(); std::lock_guard<std::mutex> rl(g_comb_mu); pk->comb.reset(); }
  pk->comb_bits_wanted = bits;
  return 0;
}
int phe_pubkey_comb_bits(const phe_pubkey* pk) { return pk ? (pk->comb ? pk->comb->wb : 0) : -1; }
int phe_pubkey_comb_info(const phe_pubkey* pk, unsigned long long* table_bytes, double* build_ms) {
  if (!pk) return fail("null");
  if (table_bytes) *table_bytes = pk->comb ? (unsigned long long)pk->comb->bytes : 0ull;
  if (build_ms) *build_ms = pk->comb ? pk->comb->build_ms : 0.0;
  return 0;
}
int phe_pubkey_bits(const phe_pubkey* pk) { return pk ? pk->bits : -1; }
int phe_pubkey_n_words(const phe_pubkey* pk) { return pk ? pk->n_words : -1; }
int phe_pubkey_is_djn(const phe_pubkey* pk) { return pk ? pk->djn : -1; }
int phe_pubkey_randbits(const phe_pubkey* pk) { return pk ? pk->randbits : -1; }
int phe_pubkey_get_n(const phe_pubkey* pk, uint32_t* o) { if (!pk || !o) return fail("null"); pk->n.to_words(o, pk->n_words); return 0; }
int phe_pubkey_get_nsquare(const phe_pubkey* pk, uint32_t* o) { if (!pk || !o) return fail("null"); pk->nsq.to_words(o, 2 * (size_t)pk->n_words); return 0; }
int phe_pubkey_get_hs(const phe_pubkey* pk, uint32_t* o) { if (!pk || !o) return fail("null"); pk->hs.to_words(o, 2 * (size_t)pk->n_words); return 0; }

int phe_privkey_create(const phe_pubkey* pk, const uint32_t* p, int p_words, const uint32_t* q, int q_words,
                       phe_privkey** out) {
  try {
    if (!pk || !p || !q || !out) return fail("phe_privkey_create: bad arguments");
    BN P = BN::from_words(p, p_words), Q = BN::from_words(q, q_words);
    if (!(hbn::mul(P, Q) == pk->n)) return fail("phe_privkey_create: p * q != n");
    if (Q < P) std::swap(P, Q);
    std::unique_ptr<phe_privkey> sk(new phe_privkey);
    sk->pk = pk; sk->p = P; sk->q = Q; sk->hw = pk->n_words;
    sk->ops = shape_for_bits(pk->n_words * 32);
    if (!sk->ops) return fail("phe_privkey_create: unsupported key size");
    if (2 * Q.bits() > (size_t)pk->n_words * 32)
      return fail("phe_privkey_create: primes too unbalanced (q^2 must fit the n_words words of n)");
    const ShapeOps* o = sk->ops;
    const BN R = hbn::shl(BN(1), o->capacity_bits);
    const BN g = hbn::add(pk->n, BN(1));
    BN hx[2];
    const BN X[2] = {P, Q};
    for (int y = 0; y < 2; ++y) {
      const BN X2 = hbn::mul(X[y], X[y]);
      const BN Rm = hbn::mod(R, X2);
      // extra = 2^(32 hw) * R^2 mod x^2  (high half of the ciphertext in the pre-reduction)
      const BN k2 = hbn::mulmod(hbn::mod(hbn::shl(BN(1), 32 * (size_t)sk->hw), X2), hbn::mulmod(Rm, Rm, X2), X2);
      sk->h_ctx[y] = mont_block(X2, k2, o, &sk->ctx[y].n0inv);
      const BN e = hbn::sub(X[y], BN(1));
      sk->h_exp[y].assign(sk->hw, 0);
      e.to_words(sk->h_exp[y].data(), sk->h_exp[y].size());
      sk->ebits[y] = (int)e.bits();
      sk->h_prog[y] = build_powm_program(e);
      // hx = (L_x(g^(x-1) mod x^2))^-1 mod x
      const BN u = hbn::modexp(hbn::mod(g, X2), e, X2);
      const BN Lx = hbn::div(hbn::sub(u, BN(1)), X[y]);
      hx[y] = hbn::modinv_prime(hbn::mod(Lx, X[y]), X[y]);
    }
    const BN pinv = hbn::modinv_prime(hbn::mod(P, Q), Q);
    std::vector<uint32_t> tail((size_t)DT_COUNT * EW(o));
    auto put = [&](int idx, const BN& v) { to_entry(v, o, &tail[(size_t)idx * EW(o)]); };
    put(DT_P, P); put(DT_Q, Q); put(DT_N, pk->n);
    put(DT_HPM, hbn::mulmod(hx[0], hbn::mod(R, P), P));
    put(DT_HQM, hbn::mulmod(hx[1], hbn::mod(R, Q), Q));
    put(DT_PINVM, hbn::mulmod(pinv, hbn::mod(R, Q), Q));
    put(DT_PMN, hbn::mulmod(P, hbn::mod(R, pk->n), pk->n));
    put(DT_ONE, BN(1));
    sk->h_tail = tail;
    sk->n0invs[0] = neg_inv52(P);
    sk->n0invs[1] = neg_inv52(Q);
    sk->n0invs[2] = neg_inv52(pk->n);
    {   // balanced keys run the CRT halves on the p-adic pair engine
      const int chunk_bits = 16 * sk->hw;
      const char* off = getenv("PHE_NO_PAIR_ENGINE");
      sk->use_pair = !(off && off[0] == '1') && build_pair_block(P, hx[0], chunk_bits, &sk->pairb[0]) &&
                     build_pair_block(Q, hx[1], chunk_bits, &sk->pairb[1]) && sk->pairb[0].L == sk->pairb[1].L;
      if (sk->use_pair) {   // the programs of p and q may cut into different numbers of time slices: k_dec_pair takes ONE count,
        for (int y = 0; y < 2; ++y) {   // so the shorter one gets empty segments ([PO_END]) at its end
          PairBlock& b = sk->pairb[y];
          while (b.segoff.size() < sk->pairb[1 - y].segoff.size()) { b.segoff.push_back((int)b.segprog.size()); b.segprog.push_back(PO_END); }
        }
      }
    }
    *out = sk.release();
    return 0;
  } catch (const std::exception& e) { return fail(std::string("phe_privkey_create: ") + e.what()); }
}

void phe_privkey_destroy(phe_privkey* sk) {
  if (!sk) return;
  for (DevBuf* b : {&sk->d_ctx[0], &sk->d_ctx[1], &sk->d_exp[0], &sk->d_exp[1], &sk->d_prog[0], &sk->d_prog[1], &sk->d_tail, &sk->d_pair_cst[0], &sk->d_pair_cst[1], &sk->d_pair_prog[0], &sk->d_pair_prog[1], &sk->d_pair_segprog[0], &sk->d_pair_segprog[1], &sk->ws_in, &sk->ws_out,
                    &sk->ws_mont[0], &sk->ws_mont[1], &sk->ws_u[0], &sk->ws_u[1], &sk->ws_tbl, &sk->ws_sched, &sk->ws_cls, &sk->d_n}) b->release();
  sk->chain.release();
  delete sk;
}
int phe_privkey_get_p(const phe_privkey* sk, uint32_t* o) { if (!sk || !o) return fail("null"); sk->p.to_words(o, sk->hw); return 0; }
int phe_privkey_get_q(const phe_privkey* sk, uint32_t* o) { if (!sk || !o) return fail("null"); sk->q.to_words(o, sk->hw); return 0; }

// ---- key generation (host) ---------------------------------------------------------------------------
static bool probably_prime(const BN& c) {
  static const std::vector<uint32_t> small = [] {   // initialised once, thread-safe (keygen runs without the GIL)
    std::vector<uint32_t> v;
    std::vector<bool> sieve(8192, true);
    for (uint32_t i = 2; i < 8192; ++i) if (sieve[i]) { v.push_back(i); for (uint32_t j = i * i; j < 8192; j += i) sieve[j] = false; }
    return v;
  }();
  for (uint32_t sp : small) {
    uint64_t rem = 0;
    for (size_t i = c.w.size(); i-- > 0;) rem = ((rem << 32) | c.w[i]) % sp;
    if (rem == 0) return c == BN(sp);
  }
  const BN cm1 = hbn::sub(c, BN(1));
  size_t s = 0;
  while (!cm1.bit(s)) ++s;
  const BN d = hbn::shr(cm1, s);
  for (int round = 0; round < 12; ++round) {
    BN a = hbn::add(hbn::mod(random_bits(c.bits() + 64), hbn::sub(c, BN(3))), BN(2));  // [2, c-2]
    BN x = hbn::modexp(a, d, c);
    if (x == BN(1) || x == cm1) continue;
    bool witness = true;
    for (size_t i = 1; i < s; ++i) {
      x = hbn::mulmod(x, x, c);
      if (x == cm1) { witness = false; break; }
    }
    if (witness) return false;
  }
  return true;
}

static BN random_prime_3mod4(int bits) {
  for (;;) {
    BN c = random_bits(bits);
    std::vector<uint32_t> w((bits + 31) / 32, 0);
    c.to_words(w.data(), w.size());
    w[0] |= 3u;
    const int top = bits - 1, top2 = bits - 2;
    w[top >> 5] |= 1u << (top & 31);
    w[top2 >> 5] |= 1u << (top2 & 31);
    c = BN::from_words(w.data(), w.size());
    if (probably_prime(c)) return c;
  }
}

int phe_keygen(int bits, uint32_t* n_out, uint32_t* p_out, uint32_t* q_out) {
  try {
    // ipcl::generateKeypair accepts 200 <= bits <= 2048 with bits % 4 == 0; 3072 is config 5
    if (bits < 200 || bits > 3072 || (bits % 4) != 0) return fail("phe_keygen: modulus bit length must be a multiple of 4 in [200, 3072]");
    if (!n_out || !p_out || !q_out) return fail("phe_keygen: null output");
    for (;;) {
      const BN p = random_prime_3mod4(bits / 2), q = random_prime_3mod4(bits / 2);
      if (p == q) continue;
      if (!(hbn::gcd(hbn::sub(p, BN(1)), hbn::sub(q, BN(1))) == BN(2))) continue;
      const BN n = hbn::mul(p, q);
      if ((int)n.bits() != bits) continue;
      n.to_words(n_out, (bits + 31) / 32);
      p.to_words(p_out, (bits / 2 + 31) / 32);
      q.to_words(q_out, (bits / 2 + 31) / 32);
      return 0;
    }
  } catch (const std::exception& e) { return fail(std::string("phe_keygen: ") + e.what()); }
}

// ---- device-buffer entry points --------------------------------------------------------------------------
// Obfuscator exponents for a device-path encrypt whose caller passed none: DJN keys draw them on the stream (ChaCha20,
// as the host path does); classic keys draw r in [1, n-1] on the host and upload.  Result in pk->ws_r.
static int draw_r_for_dev(const phe_pubkey* pk, size_t count, int* r_words, cudaStream_t s) {
  if (pk->djn) return random_r_dev(pk, count, r_words, s);
  std::vector<uint32_t> rgen;
  PHE_TRY(make_r_host(pk, count, rgen, r_words));
  PHE_TRY(pk->ws_r.ensure(rgen.size()));
  CUDA_TRY(cudaMemcpyAsync(pk->ws_r.p, rgen.data(), rgen.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));   // rgen is a stack-lifetime staging buffer
  return 0;
}

int phe_encrypt_dev(const phe_pubkey* pk, const uint32_t* d_m, size_t count, const uint32_t* d_r, int r_words,
                    int make_secure, uint32_t* d_ct_out, void* stream) {
  if (!pk || !d_m || !d_ct_out) return fail("phe_encrypt_dev: null argument");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(pk->mu);
  PHE_TRY(pk_ensure_device(pk));
  KEY_CHAIN(pk, stream);
  if (!make_secure) d_r = nullptr;
  else if (!d_r) { PHE_TRY(draw_r_for_dev(pk, count, &r_words, (cudaStream_t)stream)); d_r = pk->ws_r.p; }
  return encrypt_dev_impl(pk, d_m, count, d_r, r_words, d_ct_out, (cudaStream_t)stream);
}
// DJN encrypt fused with the gather of BASELINE config 4: every ciphertext row is stored to d_ct_out and to the same row
// of n_peers more buffers (the other ranks' gather buffers, opened through CUDA IPC and reachable over NVLink).
int phe_encrypt_dev_multi(const phe_pubkey* pk, const uint32_t* d_m, size_t count, const uint32_t* d_r, int r_words,
                          int make_secure, uint32_t* d_ct_out, uint32_t* const* d_peer_out, int n_peers, void* stream) {
  if (!pk || !d_m || !d_ct_out || (n_peers && !d_peer_out)) return fail("phe_encrypt_dev_multi: null argument");
  if (n_peers < 0 || n_peers > NPAIR_MAX_PEERS) return fail("phe_encrypt_dev_multi: at most 15 peer buffers");
  if (!pk->djn) return fail("phe_encrypt_dev_multi: DJN keys only");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(pk->mu);
  PHE_TRY(pk_ensure_device(pk));
  KEY_CHAIN(pk, stream);
  if (!make_secure) d_r = nullptr;
  else if (!d_r) { PHE_TRY(draw_r_for_dev(pk, count, &r_words, (cudaStream_t)stream)); d_r = pk->ws_r.p; }
  return encrypt_dev_impl(pk, d_m, count, d_r, r_words, d_ct_out, (cudaStream_t)stream, 0, d_peer_out, n_peers);
}
// CUDA IPC for the gather buffers of the fused encrypt: export a buffer from phe_dev_alloc (a whole cudaMalloc block),
// open it in another process of the node on that process's current device.
int phe_ipc_export(const uint32_t* d_ptr, unsigned char handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!d_ptr || !handle_out) return fail("phe_ipc_export: null argument");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<uint32_t*>(d_ptr)));
  std::memcpy(handle_out, &h, 64);
  return 0;
}
int phe_ipc_open(const unsigned char handle[64], uint32_t** out) {
  if (!handle || !out) return fail("phe_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = static_cast<uint32_t*>(p);
  return 0;
}
int phe_ipc_close(uint32_t* p) {
  if (p) CUDA_TRY(cudaIpcCloseMemHandle(p));
  return 0;
}
int phe_enable_peer_access(int peer_device) {
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("phe_enable_peer_access: ") + cudaGetErrorString(e)); }
  return 0;
}
int phe_decrypt_dev(const phe_privkey* sk, const uint32_t* d_ct, size_t count, uint32_t* d_m_out, void* stream) {
  if (!sk || !d_ct || !d_m_out) return fail("phe_decrypt_dev: null argument");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(sk->mu);
    PHE_TRY(sk_ensure_device(sk));
  KEY_CHAIN(sk, stream);
  return decrypt_dev_impl(sk, d_ct, count, d_m_out, (cudaStream_t)stream);
}
int phe_add_dev(const phe_pubkey* pk, const uint32_t* d_a, size_t na, const uint32_t* d_b, size_t nb, uint32_t* d_out,
                void* stream) {
  if (!pk || !d_a || !d_b || !d_out) return fail("phe_add_dev: null argument");
  if (na == 0) return 0;
  std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
  KEY_CHAIN(pk, stream);
  return add_dev_impl(pk, d_a, na, d_b, nb, d_out, (cudaStream_t)stream);
}
int phe_mul_dev(const phe_pubkey* pk, const uint32_t* d_ct, size_t n, const uint32_t* d_e, int e_words, size_t ne,
                int exp_bits, uint32_t* d_out, void* stream) {
  if (!pk || !d_ct || !d_e || !d_out) return fail("phe_mul_dev: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
  KEY_CHAIN(pk, stream);
  return mul_dev_impl(pk, d_ct, n, d_e, e_words, ne, exp_bits, d_out, (cudaStream_t)stream);
}

// ---- row operations (device-resident ciphertext matrices; index lists are host arrays) ------------------------------
int phe_gather_rows_dev(const phe_pubkey* pk, const uint32_t* d_src, size_t src_rows, const long long* idx, size_t n,
                        uint32_t* d_dst, void* stream) {
  if (!pk || !d_src || !d_dst || (n && !idx)) return fail("phe_gather_rows_dev: null argument");
  std::lock_guard<std::mutex> lk(pk->mu);
  PHE_TRY(pk_ensure_device(pk));
  KEY_CHAIN(pk, stream);
  return move_rows_impl(pk, d_src, idx, n, src_rows, d_dst, 0, (cudaStream_t)stream);
}
int phe_scatter_rows_dev(const phe_pubkey* pk, const uint32_t* d_src, const long long* idx, size_t n, uint32_t* d_dst,
                         size_t dst_rows, void* stream) {
  if (!pk || !d_src || !d_dst || (n && !idx)) return fail("phe_scatter_rows_dev: null argument");
  std::lock_guard<std::mutex> lk(pk->mu);
  PHE_TRY(pk_ensure_device(pk));
  KEY_CHAIN(pk, stream);
  return move_rows_impl(pk, d_src, idx, n, dst_rows, d_dst, 1, (cudaStream_t)stream);
}
int phe_scale_rows_dev(const phe_pubkey* pk, uint32_t* d_ct, size_t rows, const long long* idx, const int* delta, size_t n,
                       void* stream) {
  if (!pk || !d_ct || (n && (!idx || !delta))) return fail("phe_scale_rows_dev: null argument");
  try {
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    KEY_CHAIN(pk, stream);
    return scale_rows_impl(pk, d_ct, rows, idx, delta, n, (cudaStream_t)stream);
  } catch (const std::exception& e) { return fail(std::string("phe_scale_rows_dev: ") + e.what()); }
}
int phe_invert_rows_dev(const phe_pubkey* pk, uint32_t* d_ct, size_t rows, const long long* idx, size_t n, void* stream) {
  if (!pk || !d_ct || (n && !idx)) return fail("phe_invert_rows_dev: null argument");
  try {
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    KEY_CHAIN(pk, stream);
    return invert_rows_impl(pk, d_ct, rows, idx, n, (cudaStream_t)stream);
  } catch (const std::exception& e) { return fail(std::string("phe_invert_rows_dev: ") + e.what()); }
}
int phe_segsum_dev(const phe_pubkey* pk, const uint32_t* d_ct, size_t groups, size_t width, uint32_t* d_out, void* stream) {
  if (!pk || !d_ct || !d_out) return fail("phe_segsum_dev: null argument");
  try {
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    KEY_CHAIN(pk, stream);
    return segsum_impl(pk, d_ct, groups, width, d_out, (cudaStream_t)stream);
  } catch (const std::exception& e) { return fail(std::string("phe_segsum_dev: ") + e.what()); }
}

// ---- device memory for callers that keep ciphertexts resident between calls (the pybind11 shim) -----------------
// Result buffers of the Python layer come and go with every operator (25-130 MB each): cudaMalloc / cudaFree per
// operator costs milliseconds and a device-wide synchronisation each (the API's encrypt varied 17-35 ms around a 16 ms
// kernel).  Freed blocks are kept, by device and rounded size, with an event recorded on the default stream at the
// moment of the free; a block is handed out again once that event has passed.  At most 16 GB stay cached per process.
namespace {
struct CachedBlock { void* p; cudaEvent_t ev; };
struct BlockCache {
  std::mutex mu;
  std::map<std::pair<int, size_t>, std::vector<CachedBlock>> free_blocks;   // (device, bytes) -> blocks
  std::map<void*, std::pair<int, size_t>> live;                               // blocks handed out
  size_t cached_bytes = 0;
  static size_t round_up(size_t bytes) {
    if (bytes <= (1u << 20)) { size_t b = 4096; while (b < bytes) b <<= 1; return b; }
    return (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
  }
  void flush(int device) {   // give everything cached on `device` back to the driver
    for (auto it = free_blocks.begin(); it != free_blocks.end();) {
      if (it->first.first == device) {
        for (auto& b : it->second) { cudaEventDestroy(b.ev); cudaFree(b.p); cached_bytes -= it->first.second; }
        it = free_blocks.erase(it);
      } else ++it;
    }
  }
} g_blocks;
constexpr size_t BLOCK_CACHE_CAP = (size_t)16 << 30;
}  // namespace

int phe_dev_alloc(const phe_pubkey* pk, size_t words, uint32_t** out) {
  if (!pk || !out) return fail("phe_dev_alloc: null argument");
  {
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
  }
  CUDA_TRY(cudaSetDevice(pk->device));
  const size_t bytes = BlockCache::round_up((words ? words : 1) * 4);
  std::lock_guard<std::mutex> lk(g_blocks.mu);
  auto it = g_blocks.free_blocks.find({pk->device, bytes});
  if (it != g_blocks.free_blocks.end() && !it->second.empty()) {
    CachedBlock b = it->second.back();
    it->second.pop_back();
    g_blocks.cached_bytes -= bytes;
    cudaEventSynchronize(b.ev);     // everything enqueued before the free has finished with the block
    cudaEventDestroy(b.ev);
    g_blocks.live[b.p] = {pk->device, bytes};
    *out = static_cast<uint32_t*>(b.p);
    return 0;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {           // out of memory: return the cached blocks to the driver and try once more
    cudaGetLastError();
    g_blocks.flush(pk->device);
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("phe_dev_alloc: ") + cudaGetErrorString(e)); }
  g_blocks.live[p] = {pk->device, bytes};
  *out = static_cast<uint32_t*>(p);
  return 0;
}
int phe_dev_free(uint32_t* p) {
  if (!p) return 0;
  std::lock_guard<std::mutex> lk(g_blocks.mu);
  auto it = g_blocks.live.find(p);
  if (it == g_blocks.live.end()) {   // not one of ours (or freed twice)
    if (cudaFree(p) != cudaSuccess) { cudaGetLastError(); return fail("phe_dev_free: cudaFree failed"); }
    return 0;
  }
  const std::pair<int, size_t> key = it->second;
  g_blocks.live.erase(it);
  int cur = -1;
  cudaGetDevice(&cur);
  cudaEvent_t ev = nullptr;
  if (g_blocks.cached_bytes + key.second > BLOCK_CACHE_CAP || cur != key.first ||
      cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ev, 0) != cudaSuccess) {
    if (ev) cudaEventDestroy(ev);
    cudaGetLastError();
    if (cudaFree(p) != cudaSuccess) { cudaGetLastError(); return fail("phe_dev_free: cudaFree failed"); }
    return 0;
  }
  g_blocks.free_blocks[key].push_back({p, ev});
  g_blocks.cached_bytes += key.second;
  return 0;
}
int phe_copy(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return 0;
  if (!dst || !src) return fail("phe_copy: null argument");
  CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
  return 0;
}

// ---- host-buffer entry points (every input / output pointer may also be a device pointer: unified addressing) ------------------------------------------------------------------------------
int phe_encrypt(const phe_pubkey* pk, const uint32_t* m, size_t count, const uint32_t* r, int r_words,
                int make_secure, uint32_t* ct_out) {
  return phe_encrypt_compact(pk, m, pk ? pk->n_words : 0, count, r, r_words, make_secure, ct_out);
}

int phe_encrypt_compact(const phe_pubkey* pk, const uint32_t* m, int m_words, size_t count, const uint32_t* r,
                        int r_words, int make_secure, uint32_t* ct_out) {
  try {
    if (!pk || !m || !ct_out) return fail("phe_encrypt: null argument");
    if (m_words < 1 || m_words > pk->n_words) return fail("phe_encrypt: m_words must be in [1, n_words]");
    if (count == 0) return 0;
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    CUDA_TRY(cudaSetDevice(pk->device));
    KEY_CHAIN(pk, 0);
    const int cw = 2 * pk->n_words;
    std::vector<uint32_t> rgen;
    const bool device_r = make_secure && !r && pk->djn;   // DJN: r uniform in [0, 2^randbits), drawn on the device
    if (make_secure && !r && !device_r) { PHE_TRY(make_r_host(pk, count, rgen, &r_words)); r = rgen.data(); }
    if (!make_secure) r = nullptr;
    PHE_TRY(pk->ws_a.ensure(count * (size_t)m_words));
    PHE_TRY(pk->ws_b.ensure(count * cw));
    CUDA_TRY(cudaMemcpyAsync(pk->ws_a.p, m, count * (size_t)m_words * 4, cudaMemcpyDefault, 0));
    const uint32_t* d_r = nullptr;
    if (device_r) {
      PHE_TRY(random_r_dev(pk, count, &r_words, 0));
      d_r = pk->ws_r.p;
    } else if (r) {   // workspace kept with the key: a cudaMalloc/cudaFree pair per call costs milliseconds and a device sync
      PHE_TRY(pk->ws_r.ensure(count * (size_t)r_words));
      CUDA_TRY(cudaMemcpyAsync(pk->ws_r.p, r, count * (size_t)r_words * 4, cudaMemcpyDefault, 0));
      d_r = pk->ws_r.p;
    }
    int rc = encrypt_dev_impl(pk, pk->ws_a.p, count, d_r, r_words, pk->ws_b.p, 0, m_words);
    if (!rc) {
      cudaError_t e = cudaMemcpy(ct_out, pk->ws_b.p, count * cw * 4, cudaMemcpyDefault);
      if (e != cudaSuccess) rc = fail(std::string("phe_encrypt D2H: ") + cudaGetErrorString(e));
    }
    return rc;
  } catch (const std::exception& e) { return fail(std::string("phe_encrypt: ") + e.what()); }
}

int phe_obfuscate(const phe_pubkey* pk, uint32_t* ct, size_t count, const uint32_t* r, int r_words) {
  try {
    if (!pk || !ct) return fail("phe_obfuscate: null argument");
    if (count == 0) return 0;
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    CUDA_TRY(cudaSetDevice(pk->device));
    KEY_CHAIN(pk, 0);
    const int cw = 2 * pk->n_words;
    std::vector<uint32_t> rgen;
    if (!r && pk->djn) {
      PHE_TRY(random_r_dev(pk, count, &r_words, 0));
    } else {
      if (!r) { PHE_TRY(make_r_host(pk, count, rgen, &r_words)); r = rgen.data(); }
      PHE_TRY(pk->ws_r.ensure(count * (size_t)r_words));
      CUDA_TRY(cudaMemcpyAsync(pk->ws_r.p, r, count * (size_t)r_words * 4, cudaMemcpyDefault, 0));
    }
    PHE_TRY(pk->ws_a.ensure(count * cw));   // obfuscators
    PHE_TRY(pk->ws_b.ensure(count * cw));
    CUDA_TRY(cudaMemcpyAsync(pk->ws_b.p, ct, count * cw * 4, cudaMemcpyDefault, 0));
    PHE_TRY(obfuscators_dev(pk, pk->ws_r.p, r_words, count, pk->ws_a.p, 0));
    PHE_TRY(add_dev_impl(pk, pk->ws_b.p, count, pk->ws_a.p, count, pk->ws_b.p, 0));
    CUDA_TRY(cudaMemcpy(ct, pk->ws_b.p, count * cw * 4, cudaMemcpyDefault));
    return 0;
  } catch (const std::exception& e) { return fail(std::string("phe_obfuscate: ") + e.what()); }
}

int phe_decrypt(const phe_privkey* sk, const uint32_t* ct, size_t count, uint32_t* m_out) {
  try {
    if (!sk || !ct || !m_out) return fail("phe_decrypt: null argument");
    if (count == 0) return 0;
    std::lock_guard<std::mutex> lk(sk->mu);
    PHE_TRY(sk_ensure_device(sk));
    CUDA_TRY(cudaSetDevice(sk->pk->device));
    KEY_CHAIN(sk, 0);
    const int hw = sk->hw, cw = 2 * hw;
    PHE_TRY(sk->ws_in.ensure(count * cw));
    PHE_TRY(sk->ws_out.ensure(count * hw));
    CUDA_TRY(cudaMemcpyAsync(sk->ws_in.p, ct, count * cw * 4, cudaMemcpyDefault, 0));
    PHE_TRY(decrypt_dev_impl(sk, sk->ws_in.p, count, sk->ws_out.p, 0));
    CUDA_TRY(cudaMemcpy(m_out, sk->ws_out.p, count * hw * 4, cudaMemcpyDefault));
    return 0;
  } catch (const std::exception& e) { return fail(std::string("phe_decrypt: ") + e.what()); }
}

// decrypt + the classification half of the fixed-point decode (see include/phe_b200.h)
int phe_decrypt_mantissas(const phe_privkey* sk, const uint32_t* ct, size_t count, long long* mant_out,
                          unsigned char* cls_out, uint32_t* m_rows_out) {
  try {
    if (!sk || !ct || !mant_out || !cls_out) return fail("phe_decrypt_mantissas: null argument");
    if (count == 0) return 0;
    std::lock_guard<std::mutex> lk(sk->mu);
    PHE_TRY(sk_ensure_device(sk));
    CUDA_TRY(cudaSetDevice(sk->pk->device));
    KEY_CHAIN(sk, 0);
    const int hw = sk->hw, cw = 2 * hw;
    PHE_TRY(sk->ws_in.ensure(count * cw));
    PHE_TRY(sk->ws_out.ensure(count * hw));
    PHE_TRY(sk->ws_cls.ensure(count * 2 + (count + 3) / 4 + 4));     // int64 mantissas, then the class bytes
    if (!sk->d_n.p) {
      std::vector<uint32_t> nw(hw);
      sk->pk->n.to_words(nw.data(), hw);
      PHE_TRY(upload(sk->d_n, nw));
    }
    CUDA_TRY(cudaMemcpyAsync(sk->ws_in.p, ct, count * cw * 4, cudaMemcpyDefault, 0));
    PHE_TRY(decrypt_dev_impl(sk, sk->ws_in.p, count, sk->ws_out.p, 0));
    long long* d_mant = reinterpret_cast<long long*>(sk->ws_cls.p);
    unsigned char* d_cls = reinterpret_cast<unsigned char*>(sk->ws_cls.p + count * 2);
    CUDA_TRY(classify_plain(sk->ws_out.p, sk->d_n.p, hw, (long long)count, d_mant, d_cls, 0));
    CUDA_TRY(cudaMemcpyAsync(mant_out, d_mant, count * 8, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpy(cls_out, d_cls, count, cudaMemcpyDeviceToHost));
    if (m_rows_out) {   // the rows the host has to decode from their words: usually none
      for (size_t i = 0; i < count; ++i)
        if (cls_out[i] == 2) {
          size_t j = i;
          while (j + 1 < count && cls_out[j + 1] == 2) ++j;         // one copy per run of such rows
          CUDA_TRY(cudaMemcpy(m_rows_out + i * hw, sk->ws_out.p + i * hw, (j - i + 1) * (size_t)hw * 4, cudaMemcpyDeviceToHost));
          i = j;
        }
    }
    return 0;
  } catch (const std::exception& e) { return fail(std::string("phe_decrypt_mantissas: ") + e.what()); }
}

int phe_add(const phe_pubkey* pk, const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out) {
  try {
    if (!pk || !a || !b || !out) return fail("phe_add: null argument");
    if (nb != na && nb != 1) return fail("phe_add: size mismatch (b must have na or 1 elements)");
    if (na == 0) return 0;
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    CUDA_TRY(cudaSetDevice(pk->device));
    KEY_CHAIN(pk, 0);
    const int cw = 2 * pk->n_words;
    PHE_TRY(pk->ws_a.ensure(na * cw));
    PHE_TRY(pk->ws_b.ensure(nb * cw));
    PHE_TRY(pk->ws_c.ensure(na * cw));
    CUDA_TRY(cudaMemcpyAsync(pk->ws_a.p, a, na * cw * 4, cudaMemcpyDefault, 0));
    CUDA_TRY(cudaMemcpyAsync(pk->ws_b.p, b, nb * cw * 4, cudaMemcpyDefault, 0));
    PHE_TRY(add_dev_impl(pk, pk->ws_a.p, na, pk->ws_b.p, nb, pk->ws_c.p, 0));
    CUDA_TRY(cudaMemcpy(out, pk->ws_c.p, na * cw * 4, cudaMemcpyDefault));
    return 0;
  } catch (const std::exception& e) { return fail(std::string("phe_add: ") + e.what()); }
}

int phe_mul(const phe_pubkey* pk, const uint32_t* ct, size_t n, const uint32_t* e, int e_words, size_t ne,
            uint32_t* out) {
  try {
    if (!pk || !ct || !e || !out) return fail("phe_mul: null argument");
    if (ne != n && ne != 1) return fail("phe_mul: size mismatch (exponents must have n or 1 elements)");
    // ipcl::modExp takes exponents of any size (the reference's exponent alignment multiplies by 2^delta, which can
    // exceed n under small keys: ipcl_python.py:551-560); two key lengths of words are accepted here
    if (e_words < 1 || e_words > 2 * pk->n_words) return fail("phe_mul: e_words out of range (1 .. 2 n_words)");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    CUDA_TRY(cudaSetDevice(pk->device));
    KEY_CHAIN(pk, 0);
    const int cw = 2 * pk->n_words;
    const int ebits = max_bits_host(e, e_words, ne);
    PHE_TRY(pk->ws_a.ensure(n * cw));
    PHE_TRY(pk->ws_b.ensure(ne * (size_t)e_words));
    PHE_TRY(pk->ws_c.ensure(n * cw));
    CUDA_TRY(cudaMemcpyAsync(pk->ws_a.p, ct, n * cw * 4, cudaMemcpyDefault, 0));
    CUDA_TRY(cudaMemcpyAsync(pk->ws_b.p, e, ne * (size_t)e_words * 4, cudaMemcpyDefault, 0));
    PHE_TRY(mul_dev_impl(pk, pk->ws_a.p, n, pk->ws_b.p, e_words, ne, ebits, pk->ws_c.p, 0));
    CUDA_TRY(cudaMemcpy(out, pk->ws_c.p, n * cw * 4, cudaMemcpyDefault));
    return 0;
  } catch (const std::exception& e) { return fail(std::string("phe_mul: ") + e.what()); }
}

int phe_invert(const phe_pubkey* pk, const uint32_t* ct, size_t count, uint32_t* out) {
  try {
    if (!pk || !ct || !out) return fail("phe_invert: null argument");
    if (count == 0) return 0;
    std::lock_guard<std::mutex> lk(pk->mu);
    PHE_TRY(pk_ensure_device(pk));
    CUDA_TRY(cudaSetDevice(pk->device));
    KEY_CHAIN(pk, 0);
    const int cw = 2 * pk->n_words;
    int rc = 0;
    for (size_t off = 0; off < count && !rc; off += CHUNK) {
      const size_t c = std::min(CHUNK, count - off);
      PHE_TRY(pk->ws_a.ensure(c * cw));
      PHE_TRY(pk->ws_b.ensure(c * cw));
      CUDA_TRY(cudaMemcpyAsync(pk->ws_a.p, ct + off * cw, c * cw * 4, cudaMemcpyDefault, 0));
      rc = invert_dev_impl(pk, pk->ws_a.p, c, pk->ws_b.p, 0);
      if (!rc) CUDA_TRY(cudaMemcpy(out + off * cw, pk->ws_b.p, c * cw * 4, cudaMemcpyDefault));
    }
    return rc;
  } catch (const std::exception& e) { return fail(std::string("phe_invert: ") + e.what()); }
}

int phe_modexp(const uint32_t* base, const uint32_t* exp, const uint32_t* modulus, int words, size_t count,
               uint32_t* out) {
  try {
    if (!base || !exp || !modulus || !out || words <= 0) return fail("phe_modexp: bad arguments");
    if (count == 0) return 0;
    if (phe_device_count() <= 0) return fail("phe_modexp: no CUDA device (there is no CPU fallback)");
    const BN N = BN::from_words(modulus, words);
    if (!N.is_odd()) return fail("phe_modexp: modulus must be odd");
    const ShapeOps* o = shape_for_bits((int)N.bits());
    if (!o) return fail("phe_modexp: modulus too large (up to 8312 bits)");
    MontCtxArgs ctx{};
    std::vector<uint32_t> blk = mont_block(N, BN(), o, &ctx.n0inv);
    // bases must be < modulus for the Montgomery bound; reduce on the host only if needed
    std::vector<uint32_t> bred;
    for (size_t i = 0; i < count; ++i) {
      BN b = BN::from_words(base + i * words, words);
      if (!(b < N)) {
        if (bred.empty()) bred.assign(base, base + count * (size_t)words);
        hbn::mod(b, N).to_words(&bred[i * words], words);
      }
    }
    const uint32_t* bsrc = bred.empty() ? base : bred.data();
    DevBuf dctx, db, de, dout, tbl;
    int rc = upload(dctx, blk);
    ctx.entries = reinterpret_cast<const double*>(dctx.p);
    if (!rc) rc = db.ensure(count * words);
    if (!rc) rc = de.ensure(count * words);
    if (!rc) rc = dout.ensure(count * words);
    if (!rc && (cudaMemcpy(db.p, bsrc, count * words * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(de.p, exp, count * words * 4, cudaMemcpyHostToDevice) != cudaSuccess)) rc = fail("phe_modexp: H2D failed");
    const int ebits = max_bits_host(exp, words, count);
    for (size_t off = 0; !rc && off < count; off += CHUNK) {
      const int c = (int)std::min(CHUNK, count - off);
      rc = launch_powm(o, ctx, db.p + off * words, words, de.p + off * words, words, words, ebits, dout.p + off * words,
                       words, c, tbl, 0);
    }
    if (!rc && cudaMemcpy(out, dout.p, count * words * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail("phe_modexp: D2H failed");
    for (DevBuf* b : {&dctx, &db, &de, &dout, &tbl}) b->release();
    return rc;
  } catch (const std::exception& e) { return fail(std::string("phe_modexp: ") + e.what()); }
}

// ---- host-only helpers ---------------------------------------------------------------------------------------
int phe_host_shape_for_bits(int mod_bits, int* L_out, int* TPI_out) {
  const ShapeOps* o = shape_for_bits(mod_bits);
  if (!o) return fail("no shape for this size");
  if (L_out) *L_out = o->L;
  if (TPI_out) *TPI_out = o->TPI;
  return 0;
}

int phe_host_mont_block(const uint32_t* mod, int mod_words, int L, int TPI, double* out, uint64_t* n0inv_out) {
  try {
    const ShapeOps* o = shape_ops(L, TPI);
    if (!o) { fail("unknown shape"); return -1; }
    if (!out) return o->KP;
    uint64_t n0 = 0;
    std::vector<uint32_t> blk = mont_block(BN::from_words(mod, mod_words), BN(), o, &n0);
    std::memcpy(out, blk.data(), blk.size() * 4);
    if (n0inv_out) *n0inv_out = n0;
    return o->KP;
  } catch (const std::exception& e) { fail(e.what()); return -1; }
}

int phe_pubkey_npair_block(const phe_pubkey* pk, int* L_out, int* TPI_out, double* out, uint64_t* n0inv_out,
                           uint64_t* d_top_out) {
  if (!pk) { fail("phe_pubkey_npair_block: null key"); return -1; }
  if (!pk->use_npair) return 0;
  if (L_out) *L_out = pk->nops->L;
  if (TPI_out) *TPI_out = pk->nops->TPI;
  if (out) std::memcpy(out, pk->h_nctx.data(), pk->h_nctx.size() * 4);
  if (n0inv_out) *n0inv_out = pk->nctx.n0inv;
  if (d_top_out) *d_top_out = pk->nctx.d_top;
  return pk->nops->KP;
}

int phe_privkey_pair_block(const phe_privkey* sk, int y, int* L_out, uint64_t* n0inv_out, double* mod_out,
                           double* cst_out, uint32_t* prog_out, int prog_cap) {
  if (!sk || y < 0 || y > 1) { fail("phe_privkey_pair_block: bad arguments"); return -1; }
  if (!sk->use_pair) return 0;
  const PairBlock& b = sk->pairb[y];
  if (L_out) *L_out = b.L;
  if (n0inv_out) *n0inv_out = b.n0inv;
  if (mod_out) std::memcpy(mod_out, b.mod.data(), b.mod.size() * 8);
  if (cst_out) std::memcpy(cst_out, b.cst.data(), b.cst.size() * 8);
  if (prog_out) {
    if ((int)b.prog.size() > prog_cap) { fail("phe_privkey_pair_block: program buffer too small"); return -1; }
    std::memcpy(prog_out, b.prog.data(), b.prog.size() * 4);
  }
  return (int)b.prog.size();
}

int phe_privkey_pair_segments(const phe_privkey* sk, int y, uint32_t* prog_out, int prog_cap, int* off_out, int off_cap,
                              int* prog_len_out) {
  if (!sk || y < 0 || y > 1) { fail("phe_privkey_pair_segments: bad arguments"); return -1; }
  if (!sk->use_pair) return 0;
  const PairBlock& b = sk->pairb[y];
  if (prog_len_out) *prog_len_out = (int)b.segprog.size();
  if (prog_out) {
    if ((int)b.segprog.size() > prog_cap) { fail("phe_privkey_pair_segments: program buffer too small"); return -1; }
    std::memcpy(prog_out, b.segprog.data(), b.segprog.size() * 4);
  }
  if (off_out) {
    if ((int)b.segoff.size() > off_cap) { fail("phe_privkey_pair_segments: offset buffer too small"); return -1; }
    std::memcpy(off_out, b.segoff.data(), b.segoff.size() * sizeof(int));
  }
  return (int)b.segoff.size();
}

int phe_chacha20_keystream(const uint32_t key[8], const uint32_t nonce[3], uint32_t counter0, uint32_t* out, size_t words) {
  if (!key || !nonce || !out) return fail("phe_chacha20_keystream: null argument");
  if (phe_device_count() <= 0) return fail("phe_chacha20_keystream: no CUDA device");
  DevBuf b;
  PHE_TRY(b.ensure(words ? words : 1));
  int rc = 0;
  if (chacha20_fill(b.p, words, key, nonce, counter0, 0, 0xffffffffu, 0) != cudaSuccess ||
      cudaMemcpy(out, b.p, words * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail("phe_chacha20_keystream: CUDA error");
  b.release();
  return rc;
}

int phe_host_powm_program(const uint32_t* e, int e_words, uint32_t* out, int out_cap) {
  try {
    const std::vector<uint32_t> prog = build_powm_program(BN::from_words(e, e_words));
    if (out) {
      if ((int)prog.size() > out_cap) { fail("phe_host_powm_program: buffer too small"); return -1; }
      std::memcpy(out, prog.data(), prog.size() * 4);
    }
    return (int)prog.size();
  } catch (const std::exception& ex) { fail(ex.what()); return -1; }
}

int phe_host_modexp(const uint32_t* base, const uint32_t* exp, const uint32_t* modulus, int words, uint32_t* out) {
  try {
    hbn::modexp(BN::from_words(base, words), BN::from_words(exp, words), BN::from_words(modulus, words)).to_words(out, words);
    return 0;
  } catch (const std::exception& e) { return fail(e.what()); }
}

}  // extern "C"
