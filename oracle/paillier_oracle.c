/* paillier_oracle.c -- CPU restatement of the reference's Paillier hot path in plain C over OpenSSL BIGNUM.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (pailliercryptolib_python_b200) never does.
 *
 * PARITY STATUS: parity unpinned at the ciphertext-bit level (see oracle/paillier_oracle.py header): the
 * reference's arithmetic lives in intel/pailliercryptolib (branch `development`, unpinned; version 2.0.0,
 * /root/reference/lib/ipcl.cmake:6-7) and intel/ipp-crypto, neither vendored nor buildable offline.  Every
 * function here has a unique canonical answer in [0, modulus), and this file is cross-checked against the
 * Python-int oracle and the committed golden vectors (tests/test_oracle.py).
 *
 * Algorithm followed (call sites in /root/reference/src/ipcl_python/bindings):
 *   encrypt  ipcl_bindings_classes.cpp:53-60   ct = (1 + m n) * obf mod n^2, obf = hs^r (DJN) or r^n (classic):
 *            one Montgomery modexp per element (ipcl::modExp -> mbx_exp_mb8 in the reference; here
 *            BN_mod_exp_mont, the same Montgomery windowed exponentiation, scalar instead of 8-lane IFMA)
 *   decrypt  ipcl_bindings_classes.cpp:127-133 decryptCRT: two modexps mod p^2, q^2 with exponents p-1, q-1,
 *            L function, * hp / hq, CRT recombination
 *   add      ipcl_bindings_classes.cpp:318-321 a * b mod n^2
 *   mul      ipcl_bindings_classes.cpp:324-325 a ^ e mod n^2
 * Data layout = the C ABI's: little-endian uint32 limbs, fixed stride (BN2bytes layout, ipcl_bindings.cpp:121-129).
 * Threading: `threads` pthreads over contiguous slices (the reference's optional `omp parallel for`).
 *
 * Build: gcc -O2 -shared -fPIC -pthread oracle/paillier_oracle.c -lcrypto -o oracle/libpaillier_oracle.so
 */
#include <openssl/bn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int op; /* 0 encrypt, 1 decrypt, 2 add, 3 mul, 4 modexp */
  /* key material */
  const uint32_t *n, *hs, *p, *q, *modulus;
  int n_words, djn;
  /* operands */
  const uint32_t *a, *b;
  int a_words, b_words;
  size_t b_count; /* broadcast when 1 */
  uint32_t* out;
  int out_words;
  size_t begin, end;
  int rc;
} job_t;

static BIGNUM* load(const uint32_t* w, int words) { return BN_lebin2bn((const unsigned char*)w, words * 4, NULL); }
static int store(const BIGNUM* v, uint32_t* w, int words) {
  return BN_bn2lebinpad(v, (unsigned char*)w, words * 4) == words * 4 ? 0 : 1;
}

static void* run(void* arg) {
  job_t* j = (job_t*)arg;
  BN_CTX* ctx = BN_CTX_new();
  BIGNUM *x = BN_new(), *y = BN_new(), *t = BN_new(), *u = BN_new();
  j->rc = 0;
  if (j->op == 0 || j->op == 2 || j->op == 3) {
    BIGNUM* n = load(j->n, j->n_words);
    BIGNUM* nsq = BN_new();
    BN_sqr(nsq, n, ctx);
    BN_MONT_CTX* mont = BN_MONT_CTX_new();
    BN_MONT_CTX_set(mont, nsq, ctx);
    BIGNUM* hs = (j->op == 0 && j->djn && j->hs) ? load(j->hs, 2 * j->n_words) : NULL;
    for (size_t i = j->begin; i < j->end && !j->rc; ++i) {
      if (j->op == 0) {
        /* raw_encrypt: 1 + m n (< n^2 since m < n) */
        BIGNUM* m = load(j->a + i * (size_t)j->a_words, j->a_words);
        BN_mul(x, m, n, ctx);
        BN_add_word(x, 1);
        BN_nnmod(x, x, nsq, ctx);
        BN_free(m);
        if (j->b) { /* applyObfuscator */
          BIGNUM* r = load(j->b + i * (size_t)j->b_words, j->b_words);
          if (j->djn) BN_mod_exp_mont(y, hs, r, nsq, ctx, mont);
          else BN_mod_exp_mont(y, r, n, nsq, ctx, mont);
          BN_mod_mul(x, x, y, nsq, ctx);
          BN_free(r);
        }
      } else if (j->op == 2) {
        BIGNUM* a = load(j->a + i * (size_t)j->a_words, j->a_words);
        BIGNUM* b = load(j->b + (j->b_count == 1 ? 0 : i) * (size_t)j->b_words, j->b_words);
        BN_mod_mul(x, a, b, nsq, ctx);
        BN_free(a); BN_free(b);
      } else {
        BIGNUM* a = load(j->a + i * (size_t)j->a_words, j->a_words);
        BIGNUM* e = load(j->b + (j->b_count == 1 ? 0 : i) * (size_t)j->b_words, j->b_words);
        BN_nnmod(a, a, nsq, ctx);
        BN_mod_exp_mont(x, a, e, nsq, ctx, mont);
        BN_free(a); BN_free(e);
      }
      j->rc |= store(x, j->out + i * (size_t)j->out_words, j->out_words);
    }
    if (hs) BN_free(hs);
    BN_MONT_CTX_free(mont);
    BN_free(nsq); BN_free(n);
  } else if (j->op == 1) {
    BIGNUM* n = load(j->n, j->n_words);
    BIGNUM* P = load(j->p, j->n_words / 2);
    BIGNUM* Q = load(j->q, j->n_words / 2);
    if (BN_cmp(P, Q) > 0) { BIGNUM* s = P; P = Q; Q = s; } /* key stores p < q */
    BIGNUM* X[2] = {P, Q};
    BIGNUM *xsq[2], *xm1[2], *hx[2];
    BN_MONT_CTX* mont[2];
    BIGNUM* g = BN_dup(n);
    BN_add_word(g, 1);
    for (int k = 0; k < 2; ++k) {
      xsq[k] = BN_new(); xm1[k] = BN_dup(X[k]); hx[k] = BN_new();
      BN_sqr(xsq[k], X[k], ctx);
      BN_sub_word(xm1[k], 1);
      mont[k] = BN_MONT_CTX_new();
      BN_MONT_CTX_set(mont[k], xsq[k], ctx);
      /* hx = (L_x(g^(x-1) mod x^2))^-1 mod x  (computeHfun) */
      BN_nnmod(t, g, xsq[k], ctx);
      BN_mod_exp_mont(t, t, xm1[k], xsq[k], ctx, mont[k]);
      BN_sub_word(t, 1);
      BN_div(t, NULL, t, X[k], ctx);
      BN_nnmod(t, t, X[k], ctx);
      BN_mod_inverse(hx[k], t, X[k], ctx);
    }
    BIGNUM* pinv = BN_new();
    BN_mod_inverse(pinv, P, Q, ctx);
    BIGNUM* mx[2] = {BN_new(), BN_new()};
    for (size_t i = j->begin; i < j->end && !j->rc; ++i) {
      BIGNUM* c = load(j->a + i * (size_t)j->a_words, j->a_words);
      for (int k = 0; k < 2; ++k) {
        BN_nnmod(t, c, xsq[k], ctx);
        BN_mod_exp_mont(t, t, xm1[k], xsq[k], ctx, mont[k]);
        BN_sub_word(t, 1); /* computeLfun: (u - 1) / x, exact; u = 0 only for non-units, then -1/x floors like Python */
        if (!BN_is_negative(t)) BN_div(t, NULL, t, X[k], ctx); /* else t = -1 = floor(-1 / x) already */
        BN_mod_mul(mx[k], t, hx[k], X[k], ctx);
      }
      BN_free(c);
      /* computeCRT: m = mp + ((mq - mp) * pinv mod q) * p */
      BN_mod_sub(t, mx[1], mx[0], Q, ctx);
      BN_mod_mul(t, t, pinv, Q, ctx);
      BN_mul(t, t, P, ctx);
      BN_add(x, t, mx[0]);
      j->rc |= store(x, j->out + i * (size_t)j->out_words, j->out_words);
    }
    for (int k = 0; k < 2; ++k) { BN_free(xsq[k]); BN_free(xm1[k]); BN_free(hx[k]); BN_MONT_CTX_free(mont[k]); BN_free(mx[k]); }
    BN_free(pinv); BN_free(g); BN_free(P); BN_free(Q); BN_free(n);
  } else { /* generic modexp, shared odd modulus */
    BIGNUM* mod = load(j->modulus, j->a_words);
    BN_MONT_CTX* mont = BN_MONT_CTX_new();
    BN_MONT_CTX_set(mont, mod, ctx);
    for (size_t i = j->begin; i < j->end && !j->rc; ++i) {
      BIGNUM* a = load(j->a + i * (size_t)j->a_words, j->a_words);
      BIGNUM* e = load(j->b + i * (size_t)j->b_words, j->b_words);
      BN_nnmod(a, a, mod, ctx);
      BN_mod_exp_mont(x, a, e, mod, ctx, mont);
      j->rc |= store(x, j->out + i * (size_t)j->out_words, j->out_words);
      BN_free(a); BN_free(e);
    }
    BN_MONT_CTX_free(mont); BN_free(mod);
  }
  BN_free(x); BN_free(y); BN_free(t); BN_free(u);
  BN_CTX_free(ctx);
  return NULL;
}

static int dispatch(job_t proto, size_t count, int threads) {
  if (count == 0) return 0;
  if (threads < 1) threads = 1;
  if ((size_t)threads > count) threads = (int)count;
  job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  int rc = 0;
  for (int t = 0; t < threads; ++t) {
    jobs[t] = proto;
    jobs[t].begin = count * (size_t)t / (size_t)threads;
    jobs[t].end = count * (size_t)(t + 1) / (size_t)threads;
    if (threads == 1) run(&jobs[t]);
    else pthread_create(&th[t], NULL, run, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) {
    if (threads > 1) pthread_join(th[t], NULL);
    rc |= jobs[t].rc;
  }
  free(jobs); free(th);
  return rc;
}

/* ct[i] = (1 + m[i] n) * obf(r[i]) mod n^2; r == NULL: make_secure = false.  hs != NULL selects DJN. */
int oracle_encrypt(const uint32_t* n, int n_words, const uint32_t* hs, const uint32_t* m, size_t count,
                   const uint32_t* r, int r_words, uint32_t* ct_out, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.op = 0; j.n = n; j.n_words = n_words; j.hs = hs; j.djn = hs != NULL;
  j.a = m; j.a_words = n_words; j.b = r; j.b_words = r_words; j.out = ct_out; j.out_words = 2 * n_words;
  return dispatch(j, count, threads);
}

int oracle_decrypt(const uint32_t* n, int n_words, const uint32_t* p, const uint32_t* q, const uint32_t* ct,
                   size_t count, uint32_t* m_out, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.op = 1; j.n = n; j.n_words = n_words; j.p = p; j.q = q;
  j.a = ct; j.a_words = 2 * n_words; j.out = m_out; j.out_words = n_words;
  return dispatch(j, count, threads);
}

int oracle_add(const uint32_t* n, int n_words, const uint32_t* a, size_t na, const uint32_t* b, size_t nb,
               uint32_t* out, int threads) {
  if (nb != na && nb != 1) return 2;
  job_t j; memset(&j, 0, sizeof j);
  j.op = 2; j.n = n; j.n_words = n_words; j.a = a; j.a_words = 2 * n_words; j.b = b; j.b_words = 2 * n_words;
  j.b_count = nb == 1 && na != 1 ? 1 : na; j.out = out; j.out_words = 2 * n_words;
  return dispatch(j, na, threads);
}

int oracle_mul(const uint32_t* n, int n_words, const uint32_t* ct, size_t count, const uint32_t* e, int e_words,
               size_t ne, uint32_t* out, int threads) {
  if (ne != count && ne != 1) return 2;
  job_t j; memset(&j, 0, sizeof j);
  j.op = 3; j.n = n; j.n_words = n_words; j.a = ct; j.a_words = 2 * n_words; j.b = e; j.b_words = e_words;
  j.b_count = ne == 1 && count != 1 ? 1 : count; j.out = out; j.out_words = 2 * n_words;
  return dispatch(j, count, threads);
}

int oracle_modexp(const uint32_t* base, const uint32_t* exp, const uint32_t* modulus, int words, size_t count,
                  uint32_t* out, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.op = 4; j.modulus = modulus; j.a = base; j.a_words = words; j.b = exp; j.b_words = words;
  j.out = out; j.out_words = words;
  return dispatch(j, count, threads);
}
