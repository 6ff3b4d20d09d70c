/* phe_b200.h -- C ABI of libphe_b200.so: the B200-native replacement for the arithmetic the reference's
 * pybind11 glue calls in ipcl:: (intel/pailliercryptolib, not vendored in /root/reference).
 *
 * Every entry point cites the reference interface it replaces (file:line under /root/reference).
 * Conventions:
 *   - all big numbers are little-endian arrays of uint32_t words at a fixed stride (the layout
 *     BN2bytes/pyByte2BN produce, src/ipcl_python/bindings/ipcl_bindings.cpp:100-138, zero padded);
 *   - for a key of n_words = bits/32 words: plaintext stride = n_words, ciphertext stride = 2*n_words;
 *   - results are canonical residues in [0, modulus);
 *   - every function returns 0 on success, non-zero on error; phe_last_error() gives the message of the
 *     last error on the calling thread (the pybind11 shim maps it to RuntimeError, as ipcl's ERROR_CHECK does);
 *   - *_dev variants take CUDA device pointers and a cudaStream_t (as void*), enqueue only, and never
 *     synchronise; the plain variants take host pointers and include H2D/D2H.
 *   - stream rule: the scratch of a key (window tables, intermediates) is shared by all calls on that key.  Calls on
 *     one key may come from any streams and threads: they are ordered on the GPU in the order they were made (each
 *     call's stream waits on an event recorded after the previous call on that key; no host synchronisation), so they
 *     never overlap each other.  Calls on different keys are independent and do overlap.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef PHE_B200_H_
#define PHE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct phe_pubkey phe_pubkey;   /* replaces ipcl::PublicKey  (ipcl_bindings_classes.cpp:14-91)  */
typedef struct phe_privkey phe_privkey; /* replaces ipcl::PrivateKey (ipcl_bindings_classes.cpp:93-163) */

const char* phe_last_error(void);
const char* phe_version(void);

/* ipcl::initializeContext / terminateContext / isQATRunning (ipcl_bindings.hpp:27-35): here the "context"
 * is the CUDA device.  phe_device_count() == 0 means no usable GPU. */
int phe_device_count(void);
int phe_set_device(int device);
int phe_get_device(void);

/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches evidence). */
unsigned long long phe_kernel_launches(void);

/* Measurement hooks (bench.py): with timing enabled every kernel launch is bracketed by a cudaEvent pair on
 * the stream it is launched on; phe_timing_read waits for the recorded events and returns the summed device
 * time and launch count of one kernel kind since the last phe_timing_enable.  Kinds: 0 k_modmul, 1 k_powm,
 * 2 k_dec_prep, 3 k_dec_tail, 4 k_encrypt_comb, 5 k_encrypt_finish, 6 k_comb_build, 7 k_dec_pair, 8 k_dec_crt,
 * 9 k_encrypt_npair, 10 k_mul_npair (and k_scale_npair), 11 k_rows_move. */
int phe_timing_enable(int on);
int phe_timing_read(int kind, double* ms_total, unsigned long long* launches);
const char* phe_timing_kind_name(int kind);
/* Integer-pipe roofline denominator measured on the current device: sustained IMAD.WIDE.U32 issue rate in
 * multiply-accumulates per second (all SMs, independent chains), best of `reps` runs. */
int phe_int_pipe_peak(int reps, double* mac_per_s);
/* FP64-pipe denominator: sustained DFMA.RZ issue rate in lane operations per second (the Montgomery kernels run on
 * this pipe: csrc/mont52.cuh). */
int phe_fp64_pipe_peak(int reps, double* dfma_per_s);
/* Rate of the bare instruction mix of one 52x52-bit limb product (2 DFMA + 1 DADD + IADD3 + IADD3.X, 20 independent
 * products in flight, nothing else) in limb products per second: the practical ceiling of the Montgomery kernels. */
int phe_product_mix_peak(int reps, double* products_per_s);

/* ---- keys ---------------------------------------------------------------------------------------------- */

/* ipcl::PublicKey(n, bits, enableDJN) and PublicKey::create(n, bits, hs, randbits)
 * (ipcl_bindings_classes.cpp:16-27; ipcl_bindings.cpp:76-98 setIpclPubKey).
 *   n: n_words words.  bits: key length (n_words*32 >= bits).
 *   djn != 0: DJN scheme.  hs == NULL: draw x and compute hs = (-x^2 mod n)^n mod n^2 (enableDJN);
 *             hs != NULL: 2*n_words words, randbits as given (create()).  randbits <= 0 means bits/2.
 * Builds n^2, the Montgomery constants and (DJN) the fixed-base comb table on the current device. */
int phe_pubkey_create(const uint32_t* n, int n_words, int bits, int djn, const uint32_t* hs, int randbits,
                      phe_pubkey** out);
void phe_pubkey_destroy(phe_pubkey* pk);
/* Digit width (1..22 bits, 0 = automatic) of the DJN fixed-base comb table T[j][d] = hs^(d 2^(bits j)): an obfuscated
 * encrypt costs randbits / bits table products and the table nwin * 2^bits entries of HBM (2.7 GB at 16 bits, 35 GB at
 * 20 bits for a 2048-bit key).  Automatic: a 12-bit table at first, promoted to the widest one that fits a quarter of
 * the free device memory (at most 20 bits / 40 GB) once the key has encrypted 32768 elements.  A pinned width takes
 * effect at the next table build (the first obfuscated encrypt, or immediately if the width changes).  The environment
 * variable PHE_COMB_BITS sets the default.  phe_pubkey_comb_bits returns the width in use (0 before a table exists).
 * Tables are shared and budgeted across keys: all phe_pubkey objects of one key (same n, hs, randbits, device) use ONE
 * table per width (a second object neither builds nor holds another copy; the table is freed with its last user), and
 * the automatic width keeps the tables of all keys on a device within 60 % of its memory (PHE_COMB_BUDGET_GB changes
 * that), so a process with many keys gets narrower tables for the later ones by rule rather than by running out. */
int phe_pubkey_set_comb_bits(phe_pubkey* pk, int bits);
int phe_pubkey_comb_bits(const phe_pubkey* pk);
/* What the table in use costs: its bytes of device memory and the wall time of its build in ms (0, 0 before one exists). */
int phe_pubkey_comb_info(const phe_pubkey* pk, unsigned long long* table_bytes, double* build_ms);
int phe_pubkey_bits(const phe_pubkey* pk);
int phe_pubkey_n_words(const phe_pubkey* pk);
int phe_pubkey_is_djn(const phe_pubkey* pk);
int phe_pubkey_randbits(const phe_pubkey* pk);
int phe_pubkey_get_n(const phe_pubkey* pk, uint32_t* n_out);            /* n_words   (PublicKey::getN)   */
int phe_pubkey_get_nsquare(const phe_pubkey* pk, uint32_t* nsq_out);    /* 2*n_words (PublicKey::getNSQ) */
int phe_pubkey_get_hs(const phe_pubkey* pk, uint32_t* hs_out);          /* 2*n_words (PublicKey::getHS)  */

/* ipcl::PrivateKey(pk, p, q) (ipcl_bindings_classes.cpp:96-101).  Checks p*q == n, orders p < q, derives
 * p^2, q^2, hp, hq, p^-1 mod q and their Montgomery forms. */
int phe_privkey_create(const phe_pubkey* pk, const uint32_t* p, int p_words, const uint32_t* q, int q_words,
                       phe_privkey** out);
void phe_privkey_destroy(phe_privkey* sk);
int phe_privkey_get_p(const phe_privkey* sk, uint32_t* p_out);  /* n_words words, zero padded (PrivateKey::getP; p <= q) */
int phe_privkey_get_q(const phe_privkey* sk, uint32_t* q_out);  /* n_words words, zero padded (PrivateKey::getQ) */

/* ipcl::generateKeypair(n_length, enable_DJN) (ipcl_bindings.cpp:12-15): host prime search.
 * p, q = 3 (mod 4), top two bits set, gcd(p-1, q-1) = 2.  bits: a multiple of 4 in [200, 3072] (the reference's rule,
 * 200 <= bits <= 2048 and bits % 4 == 0, with the upper limit lifted for BASELINE config 5).
 * Outputs: n (ceil(bits / 32) words), p, q (ceil(bits / 64) words). */
int phe_keygen(int bits, uint32_t* n_out, uint32_t* p_out, uint32_t* q_out);

/* ---- the hot path, host buffers -------------------------------------------------------------------------- */
/* Every plaintext / ciphertext pointer of this section (m, ct, a, b, out ...) may be a host pointer or a pointer into the
 * key's device (unified virtual addressing: the copies inside are cudaMemcpyDefault), so a caller can leave results
 * in HBM between calls; with a device output the call returns once the work is enqueued on the default stream.  The
 * exponents of phe_mul and explicit obfuscator exponents r are read on the host and must be host pointers. */
/* Device memory on the key's device.  Freed blocks are cached (by device and rounded size, at most 16 GB per process)
 * and handed out again once everything enqueued on the default stream -- and on the blocking streams it orders with --
 * before the free has finished; a block must not be freed while work on a NON-BLOCKING stream still uses it (stricter
 * than cudaFree, which waits for the whole device).  Every block is a whole cudaMalloc allocation (phe_ipc_export). */
int phe_dev_alloc(const phe_pubkey* pk, size_t words, uint32_t** out);
int phe_dev_free(uint32_t* p);
int phe_copy(void* dst, const void* src, size_t bytes);                  /* host or device on either side; synchronous */

/* ipcl::PublicKey::encrypt(PlainText, make_secure) (ipcl_bindings_classes.cpp:53-60).
 *   m: count x n_words.  ct_out: count x 2*n_words.
 *   make_secure == 0: ct = 1 + m n.
 *   make_secure != 0: ct = (1 + m n) * obf; r == NULL draws r from the OS CSPRNG, otherwise r is
 *   count x r_words (DJN: r < 2^randbits, r_words*32 >= randbits; classic: r in [1, n-1], r_words = n_words)
 *   -- the deterministic hook used by the parity tests. */
int phe_encrypt(const phe_pubkey* pk, const uint32_t* m, size_t count, const uint32_t* r, int r_words,
                int make_secure, uint32_t* ct_out);
/* Same with plaintext rows of m_words <= n_words words each (upper words are zero): a batch of 53-bit fixed-point
 * mantissas is 2 words per element instead of n_words = 64, i.e. 32 times less to copy to the device. */
int phe_encrypt_compact(const phe_pubkey* pk, const uint32_t* m, int m_words, size_t count, const uint32_t* r,
                        int r_words, int make_secure, uint32_t* ct_out);

/* ipcl::PublicKey::applyObfuscator(vector<BigNumber>&) (ipcl_bindings_classes.cpp:71-83): ct *= obf(r). */
int phe_obfuscate(const phe_pubkey* pk, uint32_t* ct_inout, size_t count, const uint32_t* r, int r_words);

/* ipcl::PrivateKey::decrypt(CipherText) -> decryptCRT (ipcl_bindings_classes.cpp:127-133).
 *   ct: count x 2*n_words.  m_out: count x n_words.
 * Device scratch held by the private key after the first call: the window tables of the exponentiations, 10.25 KB per
 * ciphertext and CRT half at 2048-bit keys (15.5 KB at 3072, 5.1 KB at 1024), for at most 2^17 ciphertexts per launch
 * (2.8 GB; larger batches run as several launches, and launches shrink further if the device cannot spare that). */
int phe_decrypt(const phe_privkey* sk, const uint32_t* ct, size_t count, uint32_t* m_out);

/* phe_decrypt followed, on the device, by the classification half of the reference's FixedPointNumber.decode
 * (bindings/fixedpoint.py:97-115), for callers that decode fixed-point numbers: per element the signed mantissa when it
 * fits 63 bits and a class byte -- 0: the plaintext m itself is < 2^63 (mantissa = m); 1: n - m < 2^63, a negative value
 * (mantissa = -(n - m)); 2: neither (mantissa 0; the plaintext words of exactly those rows are written to
 * m_rows_out[i * n_words ..], a host buffer of count x n_words words, if it is not NULL).  9 bytes per element cross
 * PCIe instead of 4 n_words, and the host never scans the words.  mant_out: count int64, cls_out: count bytes (host). */
int phe_decrypt_mantissas(const phe_privkey* sk, const uint32_t* ct, size_t count, long long* mant_out,
                          unsigned char* cls_out, uint32_t* m_rows_out);

/* ipcl::CipherText::operator+(CipherText) -> raw_add (ipcl_bindings_classes.cpp:318-321):
 *   out[i] = a[i] * b[i] mod n^2; nb is na or 1 (broadcast). */
int phe_add(const phe_pubkey* pk, const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out);

/* ipcl::CipherText::operator*(PlainText) -> raw_mul -> ipcl::modExp (ipcl_bindings_classes.cpp:324-325):
 *   out[i] = ct[i] ^ e[i] mod n^2; e: ne x e_words (e_words <= 2 n_words: ipcl::modExp takes exponents beyond n, which
 *   the reference's exponent alignment produces under small keys), ne is n or 1 (broadcast). */
int phe_mul(const phe_pubkey* pk, const uint32_t* ct, size_t n, const uint32_t* e, int e_words, size_t ne,
            uint32_t* out);

/* out[i] = ct[i]^-1 mod n^2 (count rows of 2 n_words words).  The reference inverts ciphertexts one by one on the host
 * for HE-mul by a negative plaintext (gmpy2.invert, ipcl_python.py:272-276, 426-441, 470-479); here the whole batch
 * costs 6 Montgomery products per element on the device (Montgomery's trick) plus at most 16 host inversions.  Fails
 * (and writes nothing useful) if some ct[i] is not invertible modulo n^2. */
int phe_invert(const phe_pubkey* pk, const uint32_t* ct, size_t count, uint32_t* out);

/* ipcl::modExp(base, exp, mod) element-wise with one shared odd modulus:
 *   out[i] = base[i] ^ exp[i] mod modulus; all operands `words` words; supports moduli up to 8312 bits. */
int phe_modexp(const uint32_t* base, const uint32_t* exp, const uint32_t* modulus, int words, size_t count,
               uint32_t* out);

/* ---- the hot path, device buffers (zero-copy chaining, multi-GPU gather) ----------------------------------- */

/* make_secure == 0: ct = 1 + m n (d_r ignored).  make_secure != 0: ct = (1 + m n) * obf(r) with r = d_r (count x r_words,
 * device memory) or, when d_r == NULL, drawn by the library exactly as phe_encrypt does (DJN: ChaCha20 keystream on the
 * stream, keyed from getrandom(2); classic: host CSPRNG, uploaded).  Caller-supplied r must come from a CSPRNG; the
 * explicit form exists for the parity tests. */
int phe_encrypt_dev(const phe_pubkey* pk, const uint32_t* d_m, size_t count, const uint32_t* d_r, int r_words,
                    int make_secure, uint32_t* d_ct_out, void* stream);
/* The same fused with the gather of ciphertext shards (BASELINE config 4): every row is also stored to the same row of
 * n_peers (<= 15) more buffers -- the other ranks' gather buffers, opened with CUDA IPC -- by the encrypt kernel itself,
 * so the transfer over NVLink overlaps the arithmetic.  d_ct_out and every d_peer_out[k] already point at this rank's
 * first row.  DJN keys.  phe_enable_peer_access(dev) = cudaDeviceEnablePeerAccess on the current device. */
int phe_encrypt_dev_multi(const phe_pubkey* pk, const uint32_t* d_m, size_t count, const uint32_t* d_r, int r_words,
                          int make_secure, uint32_t* d_ct_out, uint32_t* const* d_peer_out, int n_peers, void* stream);
int phe_enable_peer_access(int peer_device);
/* CUDA IPC for those buffers: export a phe_dev_alloc block as a 64-byte handle; open it in another process of the node
 * (on its current device, peer access enabled lazily); close the mapping. */
int phe_ipc_export(const uint32_t* d_ptr, unsigned char handle_out[64]);
int phe_ipc_open(const unsigned char handle[64], uint32_t** out);
int phe_ipc_close(uint32_t* p);
int phe_decrypt_dev(const phe_privkey* sk, const uint32_t* d_ct, size_t count, uint32_t* d_m_out, void* stream);
int phe_add_dev(const phe_pubkey* pk, const uint32_t* d_a, size_t na, const uint32_t* d_b, size_t nb,
                uint32_t* d_out, void* stream);
/* exp_bits: upper bound on the bit length of every exponent (uniform loop count); 0 means e_words*32. */
int phe_mul_dev(const phe_pubkey* pk, const uint32_t* d_ct, size_t n, const uint32_t* d_e, int e_words, size_t ne,
                int exp_bits, uint32_t* d_out, void* stream);

/* ---- row operations on device-resident ciphertext matrices ------------------------------------------------------------
 * What the reference's Python does around the hot path with per-element lists -- exponent alignment
 * (ipcl_python.py:528-741), inversion of the ciphertexts that meet a negative plaintext (:272-276, 426-441, 470-479,
 * 851-857), the operand maps of matmul (:777-808) and the add trees of sum / dot / matmul (:746-775, 810-880) -- as
 * whole-batch operations on [rows, 2 n_words] matrices that stay in HBM.  Index and delta lists are HOST arrays
 * (int64 row numbers); d_* are device pointers.  All of them enqueue on `stream` and return after the index lists
 * have been consumed. */
/* d_dst[i] = d_src[idx[i]] (i < n; idx[i] < src_rows; rows may repeat: a broadcast is a gather of zeros) */
int phe_gather_rows_dev(const phe_pubkey* pk, const uint32_t* d_src, size_t src_rows, const long long* idx, size_t n,
                        uint32_t* d_dst, void* stream);
/* d_dst[idx[i]] = d_src[i] (i < n; idx[i] < dst_rows, distinct) */
int phe_scatter_rows_dev(const phe_pubkey* pk, const uint32_t* d_src, const long long* idx, size_t n, uint32_t* d_dst,
                         size_t dst_rows, void* stream);
/* d_ct[idx[i]] <- d_ct[idx[i]] ^ (2^delta[i]) mod n^2 in place (distinct rows, delta >= 0 of any size): delta squarings
 * each on the n-adic pair engine -- the reference's HE-mul by the plaintext BASE^delta of the exponent alignment. */
int phe_scale_rows_dev(const phe_pubkey* pk, uint32_t* d_ct, size_t rows, const long long* idx, const int* delta, size_t n,
                       void* stream);
/* d_ct[idx[i]] <- d_ct[idx[i]]^-1 mod n^2 in place (Montgomery's trick on the gathered rows; fails if one of them is
 * not invertible, leaving d_ct untouched) -- gmpy2.invert of ipcl_python.py:272-276. */
int phe_invert_rows_dev(const phe_pubkey* pk, uint32_t* d_ct, size_t rows, const long long* idx, size_t n, void* stream);
/* d_out[g] = HE-sum of the rows d_ct[g * width .. (g + 1) * width) = their product mod n^2 (g < groups): a log-depth tree
 * with ONE Montgomery product per addition (the factors R^-1 the levels leave behind are repaid by one product with
 * R^width at the root), one launch per level.  d_ct is not modified.  ipcl_python.py:810-827 (__padded_ct). */
int phe_segsum_dev(const phe_pubkey* pk, const uint32_t* d_ct, size_t groups, size_t width, uint32_t* d_out, void* stream);

/* The obfuscator exponents r of a DJN key, when the caller passes r = NULL, are drawn on the device from a ChaCha20
 * keystream (RFC 8439 block function) keyed with 256 + 96 fresh bits of getrandom(2) per call.  This exposes the
 * keystream generator for known-answer tests: `words` 32-bit words of keystream starting at block counter0. */
int phe_chacha20_keystream(const uint32_t key[8], const uint32_t nonce[3], uint32_t counter0, uint32_t* out, size_t words);

/* ---- host-only helpers exposed for the CPU test-suite (no GPU needed) --------------------------------------- */

/* Montgomery context block exactly as uploaded to the device for modulus `mod` (mod_words words) in the
 * lane-group shape (L, TPI): 5 entries x KP doubles (each the exact integer value of one 52-bit limb, padded
 * [TPI][LP] layout): N, R^2 mod N, R mod N, 1, extra (0 here), R = 2^(52 L TPI).  Returns KP (>0) or <0 on
 * error.  out may be NULL to query KP.  n0inv_out = -N^-1 mod 2^52. */
int phe_host_mont_block(const uint32_t* mod, int mod_words, int L, int TPI, double* out, uint64_t* n0inv_out);
/* HE mul, DJN encrypt and the comb table run on the n-adic pair engine (csrc/npair_items.cuh): numbers mod n^2 as
 * pairs (X0, X1) of n-sized Montgomery operands.  This dumps the constant block the engine is given for the key:
 * 9 entries x KP doubles (n, 1, D mod R, W00, W01, W10, W11, OM0, OM1 -- csrc/npair_items.cuh: NPairEntry) in the
 * shape (L, TPI) of n, plus n0inv = -n^-1 mod 2^52 and the top limb of D = ceil(R / n) n.  Returns KP (> 0), 0 if the
 * key does not use the engine, < 0 on error.  out may be NULL to query. */
int phe_pubkey_npair_block(const phe_pubkey* pk, int* L_out, int* TPI_out, double* out, uint64_t* n0inv_out,
                           uint64_t* d_top_out);
/* Decrypt runs its two CRT halves on the p-adic pair engine when p and q both have exactly bits/2 bits (every key
 * made by a keygen).  This dumps what the engine is given for x = p (y = 0) or q (y = 1): L (limbs per number),
 * n0inv = -x^-1 mod 2^52, mod_out = [L] limbs of x (doubles), cst_out =
 * [6][2][L] constant pairs (W_0..W_3, (1, 0), (h_x R mod x, 0)) and the program (csrc/paillier_items.cuh: PairOp).
 * Returns the program length, 0 if the key does not use the engine, < 0 on error.  Buffers may be NULL. */
int phe_privkey_pair_block(const phe_privkey* sk, int y, int* L_out, uint64_t* n0inv_out, double* mod_out,
                           double* cst_out, uint32_t* prog_out, int prog_cap);
/* The same program cut into time slices, as k_dec_pair runs it when a launch has more work units than resident warps
 * (a unit -- 32 ciphertexts, one modulus -- is then continued segment by segment by whichever warp is free, so the end
 * of a launch is ragged by one segment instead of one whole exponentiation): segments back to back in prog_out, every
 * one but the last ending with [PO_TX park, PO_END] and every one but the first starting with [PO_XT park], park = the
 * table slot after the window table; off_out[k] = start of segment k.  Returns the number of segments (0 if the key
 * does not use the engine, < 0 on error); *prog_len_out = total length.  Buffers may be NULL to query. */
int phe_privkey_pair_segments(const phe_privkey* sk, int y, uint32_t* prog_out, int prog_cap, int* off_out, int off_cap,
                              int* prog_len_out);
/* Sliding-window program of a shared exponent as executed by k_powm_prog (decrypt: p-1, q-1; classic scheme: n):
 * out[0] = table index of the leading window (0xffff: exponent is zero), out[k>=1] = (squarings << 8) | index into
 * the table of odd powers x^(2 index + 1), index 0xff = no multiplication.  Returns the number of entries (or <0);
 * out may be NULL to query. */
int phe_host_powm_program(const uint32_t* e, int e_words, uint32_t* out, int out_cap);
/* base^exp mod modulus on the host bignum (key-setup arithmetic), words words each. */
int phe_host_modexp(const uint32_t* base, const uint32_t* exp, const uint32_t* modulus, int words, uint32_t* out);
/* Shape selection: writes L, TPI for a modulus of `mod_bits` bits; returns 0 or error. */
int phe_host_shape_for_bits(int mod_bits, int* L_out, int* TPI_out);

#ifdef __cplusplus
}
#endif
#endif /* PHE_B200_H_ */
