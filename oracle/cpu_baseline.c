/* CPU baseline / C oracle for the ModpGroup hot path.  TEST + BENCH INFRASTRUCTURE ONLY
 * (see oracle/groups.py header): never linked into libmpvss_b200.so.
 *
 * The reference's arithmetic lives in num-bigint 0.2 (Cargo.toml:15), which is not
 * vendored under /root/reference and cannot be built here (no Rust).  This file restates the
 * reference's *loop structure* for verify_distribution_shares on OpenSSL libcrypto big
 * numbers ("OpenSSL proxy for num-bigint"):
 *   schedule 0 (reference): X_i = prod_j C_j^(i^j mod (q-1)), t full-width exponentiations and
 *       t multiplications per participant (src/participant.rs:423-434, mpvss.rs:114-123), then
 *       a1 = g^r * X^c, a2 = y^r * Y^c as four separate exponentiations (src/dleq.rs:66-84).
 *   schedule 1 (same algorithm as the GPU): Horner in the exponent.
 * One participant per thread.  All values are 256-byte big-endian.
 *
 * build:  gcc -O2 -shared -fPIC -pthread -o oracle/_build/libcpu_baseline.so oracle/cpu_baseline.c -lcrypto
 */
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/obj_mac.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EB 256

typedef struct {
  const uint8_t *q, *commitments, *pk, *y, *r, *c;
  const int64_t* positions;
  size_t t, s, first, stride;
  int schedule;
  uint8_t *x_out, *a1_out, *a2_out;
} job_t;

static void put(const BIGNUM* v, uint8_t* out) { BN_bn2binpad(v, out, EB); }
static void put_pt(const EC_GROUP* g, const EC_POINT* p, uint8_t* out, BN_CTX* ctx);

static void* worker(void* arg) {
  job_t* J = (job_t*)arg;
  BN_CTX* ctx = BN_CTX_new();
  BIGNUM *q = BN_bin2bn(J->q, EB, NULL), *qm1 = BN_dup(q), *g = BN_new(), *c = BN_bin2bn(J->c, EB, NULL);
  BN_sub_word(qm1, 1);
  BN_set_word(g, 4); /* subgroup generator, src/groups/modp.rs:65-66 */
  BN_MONT_CTX* mont = BN_MONT_CTX_new();
  BN_MONT_CTX_set(mont, q, ctx);
  BIGNUM** C = malloc(J->t * sizeof(BIGNUM*));
  for (size_t j = 0; j < J->t; ++j) C[j] = BN_bin2bn(J->commitments + j * EB, EB, NULL);
  BIGNUM *x = BN_new(), *e = BN_new(), *tmp = BN_new(), *pos = BN_new(), *a = BN_new(), *b = BN_new();
  for (size_t i = J->first; i < J->s; i += J->stride) {
    BN_set_word(pos, (BN_ULONG)J->positions[i]);
    if (J->schedule == 0) {
      BN_one(x);
      BN_one(e);
      for (size_t j = 0; j < J->t; ++j) { /* participant.rs:423-434 */
        BN_mod_exp_mont(tmp, C[j], e, q, ctx, mont);
        BN_mod_mul(x, x, tmp, q, ctx);
        BN_mod_mul(e, e, pos, qm1, ctx);
      }
    } else {
      BN_copy(x, C[J->t - 1]);
      for (size_t j = J->t - 1; j-- > 0;) {
        BN_mod_exp_mont(x, x, pos, q, ctx, mont);
        BN_mod_mul(x, x, C[j], q, ctx);
      }
    }
    put(x, J->x_out + i * EB);
    BIGNUM *r = BN_bin2bn(J->r + i * EB, EB, NULL), *pk = BN_bin2bn(J->pk + i * EB, EB, NULL),
           *y = BN_bin2bn(J->y + i * EB, EB, NULL);
    /* dleq.rs:66-84 */
    BN_mod_exp_mont(a, g, r, q, ctx, mont);
    BN_mod_exp_mont(b, x, c, q, ctx, mont);
    BN_mod_mul(a, a, b, q, ctx);
    put(a, J->a1_out + i * EB);
    BN_mod_exp_mont(a, pk, r, q, ctx, mont);
    BN_mod_exp_mont(b, y, c, q, ctx, mont);
    BN_mod_mul(a, a, b, q, ctx);
    put(a, J->a2_out + i * EB);
    BN_free(r); BN_free(pk); BN_free(y);
  }
  for (size_t j = 0; j < J->t; ++j) BN_free(C[j]);
  free(C);
  BN_free(x); BN_free(e); BN_free(tmp); BN_free(pos); BN_free(a); BN_free(b);
  BN_free(q); BN_free(qm1); BN_free(g); BN_free(c);
  BN_MONT_CTX_free(mont);
  BN_CTX_free(ctx);
  return NULL;
}

/* X_i, a1_i, a2_i for `s` participants (arrays of s entries; positions 1-based), `nthreads`
 * host threads, participant i handled by thread i % nthreads. */
int cpu_modp_verify(const uint8_t* q, const uint8_t* commitments, size_t t, const int64_t* positions,
                    const uint8_t* pk, const uint8_t* y, const uint8_t* r, const uint8_t* c, size_t s, int nthreads,
                    int schedule, uint8_t* x_out, uint8_t* a1_out, uint8_t* a2_out) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > s) nthreads = (int)s;
  pthread_t* th = malloc(nthreads * sizeof(pthread_t));
  job_t* jobs = malloc(nthreads * sizeof(job_t));
  for (int k = 0; k < nthreads; ++k) {
    job_t j = {q, commitments, pk, y, r, c, positions, t, s, (size_t)k, (size_t)nthreads, schedule, x_out, a1_out, a2_out};
    jobs[k] = j;
    pthread_create(&th[k], NULL, worker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
  free(th);
  free(jobs);
  return 0;
}


/* out[i] = base[i]^exp[i] mod q (base_stride = 0: one shared base), i % nthreads per thread.  Used to
 * synthesise a box for the CPU reference arm without touching the GPU library. */
typedef struct {
  const uint8_t *q, *base, *exp;
  size_t base_stride, n, first, stride;
  uint8_t* out;
} ejob_t;
static void* eworker(void* arg) {
  ejob_t* J = (ejob_t*)arg;
  BN_CTX* ctx = BN_CTX_new();
  BIGNUM *q = BN_bin2bn(J->q, EB, NULL), *b = BN_new(), *e = BN_new(), *r = BN_new();
  BN_MONT_CTX* mont = BN_MONT_CTX_new();
  BN_MONT_CTX_set(mont, q, ctx);
  for (size_t i = J->first; i < J->n; i += J->stride) {
    BN_bin2bn(J->base + i * J->base_stride, EB, b);
    BN_bin2bn(J->exp + i * EB, EB, e);
    BN_mod_exp_mont(r, b, e, q, ctx, mont);
    put(r, J->out + i * EB);
  }
  BN_free(q); BN_free(b); BN_free(e); BN_free(r);
  BN_MONT_CTX_free(mont);
  BN_CTX_free(ctx);
  return NULL;
}
int cpu_modp_exp(const uint8_t* q, const uint8_t* base, size_t base_stride, const uint8_t* exp, size_t n, int nthreads,
                 uint8_t* out) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = (int)n;
  pthread_t* th = malloc(nthreads * sizeof(pthread_t));
  ejob_t* jobs = malloc(nthreads * sizeof(ejob_t));
  for (int k = 0; k < nthreads; ++k) {
    ejob_t j = {q, base, exp, base_stride, n, (size_t)k, (size_t)nthreads, out};
    jobs[k] = j;
    pthread_create(&th[k], NULL, eworker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
  free(th);
  free(jobs);
  return 0;
}

/* out[i] = scalar[i] * P[i] on secp256k1 (points == NULL: the generator), single thread: box synthesis */
int cpu_secp_mul(const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out) {
  BN_CTX* ctx = BN_CTX_new();
  EC_GROUP* g = EC_GROUP_new_by_curve_name(NID_secp256k1);
  EC_POINT *p = EC_POINT_new(g), *r = EC_POINT_new(g);
  BIGNUM* k = BN_new();
  for (size_t i = 0; i < n; ++i) {
    BN_bin2bn(scalars + i * 32, 32, k);
    if (points) {
      EC_POINT_oct2point(g, p, points + i * 33, 33, ctx);
      EC_POINT_mul(g, r, NULL, p, k, ctx);
    } else {
      EC_POINT_mul(g, r, k, NULL, NULL, ctx);
    }
    put_pt(g, r, out + i * 33, ctx);
  }
  BN_free(k);
  EC_POINT_free(p); EC_POINT_free(r);
  EC_GROUP_free(g);
  BN_CTX_free(ctx);
  return 0;
}

/* ---- secp256k1 (OpenSSL proxy for k256 0.13, Cargo.toml:24) -------------------------------
 * schedule 0 (reference): X_i = sum_j (i^j mod n) * C_j, t variable-base scalar multiplications
 * each converted back to affine like k256's `.into()` (src/groups/secp256k1.rs:99,106), then
 * a1 = r*G + c*X, a2 = r*y + c*Y as four multiplications and two additions (src/dleq.rs:66-84).
 * schedule 1: Horner in the group.  Points: 33-byte SEC1 compressed; scalars: 32-byte big-endian. */
typedef struct {
  const uint8_t *commitments, *pk, *y, *r, *c;
  const int64_t* positions;
  size_t t, s, first, stride;
  int schedule;
  uint8_t *x_out, *a1_out, *a2_out;
} sjob_t;

static void put_pt(const EC_GROUP* g, const EC_POINT* p, uint8_t* out, BN_CTX* ctx) {
  if (EC_POINT_is_at_infinity(g, p)) { memset(out, 0, 33); return; }
  EC_POINT_point2oct(g, p, POINT_CONVERSION_COMPRESSED, out, 33, ctx);
}

static void* sworker(void* arg) {
  sjob_t* J = (sjob_t*)arg;
  BN_CTX* ctx = BN_CTX_new();
  EC_GROUP* g = EC_GROUP_new_by_curve_name(NID_secp256k1);
  BIGNUM* n = BN_new();
  EC_GROUP_get_order(g, n, ctx);
  const EC_POINT* G = EC_GROUP_get0_generator(g);
  EC_POINT** C = malloc(J->t * sizeof(EC_POINT*));
  for (size_t j = 0; j < J->t; ++j) {
    C[j] = EC_POINT_new(g);
    EC_POINT_oct2point(g, C[j], J->commitments + j * 33, 33, ctx);
  }
  BIGNUM *e = BN_new(), *pos = BN_new(), *c = BN_bin2bn(J->c, 32, NULL);
  EC_POINT *x = EC_POINT_new(g), *tmp = EC_POINT_new(g), *a = EC_POINT_new(g), *b = EC_POINT_new(g),
           *pk = EC_POINT_new(g), *y = EC_POINT_new(g);
  for (size_t i = J->first; i < J->s; i += J->stride) {
    BN_set_word(pos, (BN_ULONG)J->positions[i]);
    if (J->schedule == 0) {
      EC_POINT_set_to_infinity(g, x);
      BN_one(e);
      for (size_t j = 0; j < J->t; ++j) { /* participant.rs:1411-1421 */
        EC_POINT_mul(g, tmp, NULL, C[j], e, ctx);
        EC_POINT_make_affine(g, tmp, ctx);
        EC_POINT_add(g, x, x, tmp, ctx);
        EC_POINT_make_affine(g, x, ctx);
        BN_mod_mul(e, e, pos, n, ctx);
      }
    } else {
      EC_POINT_copy(x, C[J->t - 1]);
      for (size_t j = J->t - 1; j-- > 0;) {
        EC_POINT_mul(g, x, NULL, x, pos, ctx);
        EC_POINT_add(g, x, x, C[j], ctx);
      }
    }
    put_pt(g, x, J->x_out + i * 33, ctx);
    BIGNUM* r = BN_bin2bn(J->r + i * 32, 32, NULL);
    EC_POINT_oct2point(g, pk, J->pk + i * 33, 33, ctx);
    EC_POINT_oct2point(g, y, J->y + i * 33, 33, ctx);
    EC_POINT_mul(g, a, NULL, G, r, ctx);
    EC_POINT_mul(g, b, NULL, x, c, ctx);
    EC_POINT_add(g, a, a, b, ctx);
    put_pt(g, a, J->a1_out + i * 33, ctx);
    EC_POINT_mul(g, a, NULL, pk, r, ctx);
    EC_POINT_mul(g, b, NULL, y, c, ctx);
    EC_POINT_add(g, a, a, b, ctx);
    put_pt(g, a, J->a2_out + i * 33, ctx);
    BN_free(r);
  }
  for (size_t j = 0; j < J->t; ++j) EC_POINT_free(C[j]);
  free(C);
  EC_POINT_free(x); EC_POINT_free(tmp); EC_POINT_free(a); EC_POINT_free(b); EC_POINT_free(pk); EC_POINT_free(y);
  BN_free(e); BN_free(pos); BN_free(c); BN_free(n);
  EC_GROUP_free(g);
  BN_CTX_free(ctx);
  return NULL;
}

int cpu_secp_verify(const uint8_t* commitments, size_t t, const int64_t* positions, const uint8_t* pk,
                    const uint8_t* y, const uint8_t* r, const uint8_t* c, size_t s, int nthreads, int schedule,
                    uint8_t* x_out, uint8_t* a1_out, uint8_t* a2_out) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > s) nthreads = (int)s;
  pthread_t* th = malloc(nthreads * sizeof(pthread_t));
  sjob_t* jobs = malloc(nthreads * sizeof(sjob_t));
  for (int k = 0; k < nthreads; ++k) {
    sjob_t j = {commitments, pk, y, r, c, positions, t, s, (size_t)k, (size_t)nthreads, schedule, x_out, a1_out, a2_out};
    jobs[k] = j;
    pthread_create(&th[k], NULL, sworker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
  free(th);
  free(jobs);
  return 0;
}
