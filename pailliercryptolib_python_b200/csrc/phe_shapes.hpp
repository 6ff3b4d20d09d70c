// phe_shapes.hpp -- type-erased launch table, one instance per (L, TPI) lane-group shape.
// Each shape is compiled in its own translation unit (shape_*.cu) so the build parallelises.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace phe {

struct MontCtxArgs;
struct PowmArgs;
struct DecPrepArgs;
struct DecTailArgs;
struct EncCombArgs;
struct EncFinishArgs;
struct CombArgs;
struct DecPairArgs;
struct DecCrtArgs;
struct InvArgs;
struct MulNPairArgs;
struct EncNPairArgs;
struct CombNPairArgs;
struct ProgNPairArgs;
struct Modmul1Args;
struct TreeLevelArgs;
struct ScaleNPairArgs;

struct ShapeOps {
  int L, TPI, KP, GPB;   // KP: doubles per padded entry
  int capacity_bits;     // 52 * L * TPI
  // every launcher returns the launch error (cudaGetLastError) and counts one kernel launch
  cudaError_t (*modmul)(const uint32_t* a, const uint32_t* b, size_t b_stride, uint32_t* out, int nwords, int count,
                        const MontCtxArgs& ctx, cudaStream_t s);
  cudaError_t (*powm)(int win, const PowmArgs& p, int ny, cudaStream_t s);
  size_t (*powm_tbl_words)(int win, int ny, int count);   // scratch size (u32 words) for a powm launch
  cudaError_t (*powm_prog)(const PowmArgs& p, int ny, cudaStream_t s);   // shared-exponent sliding-window program
  size_t (*powm_prog_tbl_words)(int ny, int count);
  cudaError_t (*dec_prep)(const DecPrepArgs& p, cudaStream_t s);
  cudaError_t (*dec_tail)(const DecTailArgs& p, cudaStream_t s);
  cudaError_t (*dec_crt)(const DecCrtArgs& p, cudaStream_t s);
  cudaError_t (*inv_block)(const InvArgs& p, bool unwind, cudaStream_t s);   // batched modular inverse, one level
  int (*resident_groups)();
  cudaError_t (*encrypt_comb)(const EncCombArgs& p, cudaStream_t s);
  cudaError_t (*encrypt_finish)(const EncFinishArgs& p, cudaStream_t s);
  cudaError_t (*comb_build)(const CombArgs& p, cudaStream_t s);
  // n-adic pair engine on this shape as the n-sized one (npair_kernels.cuh)
  cudaError_t (*mul_npair)(int win, const MulNPairArgs& p, cudaStream_t s);
  size_t (*mul_npair_tbl_words)(int win, int count);
  cudaError_t (*encrypt_npair)(const EncNPairArgs& p, cudaStream_t s);
  cudaError_t (*comb_build_npair)(const CombNPairArgs& p, cudaStream_t s);
  cudaError_t (*powm_prog_npair)(const ProgNPairArgs& p, cudaStream_t s);
  size_t (*powm_prog_npair_tbl_words)(int count);
  // one-product HE adds (broadcast / constant second operand), add-tree levels, exponent alignment by squarings
  cudaError_t (*modmul1)(const Modmul1Args& p, cudaStream_t s);
  cudaError_t (*tree_level)(const TreeLevelArgs& p, cudaStream_t s);
  cudaError_t (*scale_npair)(const ScaleNPairArgs& p, cudaStream_t s);
};
// dst[i] = src[idx[i]] (scatter == 0) or dst[idx[i]] = src[i]; rows of `words` u32 words (a multiple of 4)
cudaError_t rows_move(const uint32_t* src, uint32_t* dst, const long long* idx, long long n, int words, int scatter,
                      cudaStream_t s);

const ShapeOps* shape_ops(int L, int TPI);   // nullptr if not built

// p-adic pair engine (decrypt halves, one bignum of L limbs per lane)
struct PairOps {
  int L;
  cudaError_t (*dec_pair)(const DecPairArgs& p, const double* mod_p, const double* mod_q, cudaStream_t s);
  size_t (*tbl_words)(int count, int slots);   // table scratch (u32 words) for a launch
  int (*warps)(int count);                     // resident warps of that launch
  size_t (*sched_ints)(int count);             // ints of the scheduler block (piece counter + per-unit progress)
};
const PairOps* pair_ops(int L);   // nullptr if not built
// per plaintext row: signed 63-bit mantissa and class (0 positive, 1 negative = -(n - m), 2 neither); see pair_shapes.cu
cudaError_t classify_plain(const uint32_t* m, const uint32_t* n, int nw, long long count, long long* mant, unsigned char* cls,
                           cudaStream_t s);
unsigned long long launch_counter();          // kernels launched by this library so far
void count_launch();

// Per-kernel-kind device timing (phe_timing_* in the C ABI): when enabled every launch is bracketed by a
// cudaEvent pair on its own stream.  Off by default (no events recorded).
enum KernelKind { KK_MODMUL = 0, KK_POWM, KK_DEC_PREP, KK_DEC_TAIL, KK_ENC_COMB, KK_ENC_FINISH, KK_COMB_BUILD, KK_DEC_PAIR, KK_DEC_CRT, KK_ENC_NPAIR, KK_MUL_NPAIR, KK_ROWS, KK_COUNT };
void timing_begin(int kind, cudaStream_t s);
void timing_end(int kind, cudaStream_t s);
struct TimedLaunch {   // RAII: brackets one kernel launch, counts it
  int kind; cudaStream_t s;
  TimedLaunch(int k, cudaStream_t st) : kind(k), s(st) { timing_begin(k, st); }
  ~TimedLaunch() { timing_end(kind, s); count_launch(); }
};

}  // namespace phe
