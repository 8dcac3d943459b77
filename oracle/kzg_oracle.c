/*
 * kzg_oracle.c -- CPU restatement of go-kzg's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker the CUDA path is compared against and the "port" CPU
 * baseline timed by bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; nothing under go_kzg_b200/ does.
 *
 * The reference (protolambda/go-kzg @ c91cee5e) delegates all Fr / Fp / G1
 * arithmetic to github.com/kilic/bls12-381 v0.1.1-0.20220929213557-ca162e8a70f4
 * (go.mod:8), which is NOT vendored and cannot be built here (no Go toolchain).
 * We therefore restate that library's published behaviour (Montgomery fields mod
 * r and mod p, Jacobian short-Weierstrass y^2 = x^3 + 4, ZCash 48-byte
 * compression, bucket-method MultiExp) and restate go-kzg's own algorithms line by
 * line, each function citing the reference file:line it follows (paths relative to
 * /root/reference).  Pinned by tests/test_oracle_pins.py against every golden
 * vector the reference holds for this path (tests/golden/reference_goldens.json,
 * tests/golden/trusted_setup_g1.bin) and against oracle/pyref.py.
 *
 * External encoding (same as include/b200_kzg.h so buffers can be compared
 * byte for byte): Fr = 4 x u64 little-endian limbs, canonical (non-Montgomery);
 * G1 = Jacobian X,Y,Z each 6 x u64 little-endian canonical limbs, infinity <=> Z==0.
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ---- tiny pthread parallel-for (the image has no libgomp) ---- */
typedef void (*pf_body)(long long i, void *ctx);
typedef struct { pf_body fn; void *ctx; long long n; long long *next; pthread_mutex_t *mu; } pf_job;
static void *pf_worker(void *arg)
{
    pf_job *j = arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        long long i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->n) break;
        j->fn(i, j->ctx);
    }
    return NULL;
}
static int g_threads = 0; /* 0 = all online cores */
int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
void orc_set_threads(int n) { g_threads = n; }
static __thread int t_in_parallel = 0; /* no nested thread teams: inner loops run serially */
static void *pf_worker_outer(void *arg)
{
    t_in_parallel = 1;
    return pf_worker(arg);
}
static void parallel_for(long long n, pf_body fn, void *ctx)
{
    int nt = g_threads > 0 ? g_threads : orc_max_threads();
    if (nt > n) nt = (int)n;
    if (nt <= 1 || t_in_parallel) { for (long long i = 0; i < n; i++) fn(i, ctx); return; }
    long long next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pf_job job = {fn, ctx, n, &next, &mu};
    pthread_t *th = malloc((size_t)nt * sizeof(pthread_t));
    for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, pf_worker_outer, &job);
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------- */
/* generic Montgomery arithmetic on N 64-bit limbs                            */
/* ------------------------------------------------------------------------- */
#define FRN 4
#define FPN 6

/* r: bls/globals.go:9 ModulusStr */
static const u64 FR_P[FRN] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
/* base prime of BLS12-381 */
static const u64 FP_P[FPN] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                              0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static u64 FR_INV, FP_INV;           /* -p^-1 mod 2^64 */
static u64 FR_R2[FRN], FP_R2[FPN];   /* R^2 mod p */
static u64 FR_ONE[FRN], FP_ONE[FPN]; /* R mod p */

static inline int ge_n(const u64 *a, const u64 *b, int n)
{
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline u64 sub_n(u64 *r, const u64 *a, const u64 *b, int n)
{
    u64 borrow = 0;
    for (int i = 0; i < n; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
static inline u64 add_n(u64 *r, const u64 *a, const u64 *b, int n)
{
    u64 carry = 0;
    for (int i = 0; i < n; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        r[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    return carry;
}
static inline int is_zero_n(const u64 *a, int n)
{
    u64 acc = 0;
    for (int i = 0; i < n; i++) acc |= a[i];
    return acc == 0;
}
static inline void mod_add(u64 *r, const u64 *a, const u64 *b, const u64 *p, int n)
{
    u64 t[FPN];
    add_n(t, a, b, n); /* both moduli leave a spare top bit: no carry out */
    if (ge_n(t, p, n)) sub_n(t, t, p, n);
    memcpy(r, t, n * 8);
}
static inline void mod_sub(u64 *r, const u64 *a, const u64 *b, const u64 *p, int n)
{
    u64 t[FPN];
    if (sub_n(t, a, b, n)) add_n(t, t, p, n);
    memcpy(r, t, n * 8);
}
static inline void mont_mul(u64 *r, const u64 *a, const u64 *b, const u64 *p, u64 inv, int n)
{
    u64 t[FPN + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < n; i++) {
        u64 carry = 0;
        for (int j = 0; j < n; j++) {
            u128 s = (u128)a[j] * b[i] + t[j] + carry;
            t[j] = (u64)s;
            carry = (u64)(s >> 64);
        }
        u128 s = (u128)t[n] + carry;
        t[n] = (u64)s;
        t[n + 1] = (u64)(s >> 64);
        u64 m = t[0] * inv;
        s = (u128)m * p[0] + t[0];
        carry = (u64)(s >> 64);
        for (int j = 1; j < n; j++) {
            s = (u128)m * p[j] + t[j] + carry;
            t[j - 1] = (u64)s;
            carry = (u64)(s >> 64);
        }
        s = (u128)t[n] + carry;
        t[n - 1] = (u64)s;
        t[n] = t[n + 1] + (u64)(s >> 64);
    }
    if (t[n] || ge_n(t, p, n)) sub_n(t, t, p, n);
    memcpy(r, t, n * 8);
}

/* ---- Fr ---- (semantics: bls/bignum_kilic.go:95-115, values kept in Montgomery form) */
typedef struct { u64 l[FRN]; } fr_t;
static inline void fr_add(fr_t *r, const fr_t *a, const fr_t *b) { mod_add(r->l, a->l, b->l, FR_P, FRN); }
static inline void fr_sub(fr_t *r, const fr_t *a, const fr_t *b) { mod_sub(r->l, a->l, b->l, FR_P, FRN); }
static inline void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) { mont_mul(r->l, a->l, b->l, FR_P, FR_INV, FRN); }
static inline int fr_is_zero(const fr_t *a) { return is_zero_n(a->l, FRN); }
static inline int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a, b, sizeof(fr_t)) == 0; }
static void fr_to_mont(fr_t *r, const u64 *canon) { mont_mul(r->l, canon, FR_R2, FR_P, FR_INV, FRN); }
static void fr_from_mont(u64 *canon, const fr_t *a)
{
    u64 one[FRN] = {1, 0, 0, 0};
    mont_mul(canon, a->l, one, FR_P, FR_INV, FRN);
}
static void fr_from_u64(fr_t *r, u64 v) /* bls/bignum_kilic.go:61 AsFr */
{
    u64 c[FRN] = {v, 0, 0, 0};
    fr_to_mont(r, c);
}
static void fr_pow(fr_t *r, const fr_t *a, const u64 *e, int n)
{
    fr_t acc;
    memcpy(acc.l, FR_ONE, sizeof acc.l);
    for (int i = n * 64 - 1; i >= 0; i--) {
        fr_mul(&acc, &acc, &acc);
        if (e[i / 64] >> (i % 64) & 1) fr_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void fr_inv(fr_t *r, const fr_t *a) /* bls/bignum_kilic.go:113 InvModFr (0 -> 0) */
{
    u64 e[FRN];
    u64 two[FRN] = {2, 0, 0, 0};
    sub_n(e, FR_P, two, FRN);
    fr_pow(r, a, e, FRN);
}

/* ---- Fp ---- */
typedef struct { u64 l[FPN]; } fp_t;
static inline void fp_add(fp_t *r, const fp_t *a, const fp_t *b) { mod_add(r->l, a->l, b->l, FP_P, FPN); }
static inline void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) { mod_sub(r->l, a->l, b->l, FP_P, FPN); }
static inline void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) { mont_mul(r->l, a->l, b->l, FP_P, FP_INV, FPN); }
static inline void fp_sqr(fp_t *r, const fp_t *a) { mont_mul(r->l, a->l, a->l, FP_P, FP_INV, FPN); }
static inline int fp_is_zero(const fp_t *a) { return is_zero_n(a->l, FPN); }
static inline int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a, b, sizeof(fp_t)) == 0; }
static inline void fp_dbl(fp_t *r, const fp_t *a) { fp_add(r, a, a); }
static void fp_to_mont(fp_t *r, const u64 *canon) { mont_mul(r->l, canon, FP_R2, FP_P, FP_INV, FPN); }
static void fp_from_mont(u64 *canon, const fp_t *a)
{
    u64 one[FPN] = {1, 0, 0, 0, 0, 0};
    mont_mul(canon, a->l, one, FP_P, FP_INV, FPN);
}
static void fp_pow(fp_t *r, const fp_t *a, const u64 *e, int n)
{
    fp_t acc;
    memcpy(acc.l, FP_ONE, sizeof acc.l);
    for (int i = n * 64 - 1; i >= 0; i--) {
        fp_sqr(&acc, &acc);
        if (e[i / 64] >> (i % 64) & 1) fp_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void fp_inv(fp_t *r, const fp_t *a)
{
    u64 e[FPN];
    u64 two[FPN] = {2, 0, 0, 0, 0, 0};
    sub_n(e, FP_P, two, FPN);
    fp_pow(r, a, e, FPN);
}

static int g_init_done = 0;
static void compute_consts(const u64 *p, int n, u64 *inv, u64 *one, u64 *r2)
{
    /* -p^-1 mod 2^64 by Newton iteration */
    u64 x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p[0] * x;
    *inv = (u64)0 - x;
    /* R mod p and R^2 mod p by repeated doubling */
    u64 t[FPN];
    memset(t, 0, sizeof t);
    t[0] = 1;
    for (int i = 0; i < 2 * 64 * n; i++) {
        mod_add(t, t, t, p, n);
        if (i == 64 * n - 1) memcpy(one, t, n * 8);
    }
    memcpy(r2, t, n * 8);
}
void orc_init(void)
{
    if (g_init_done) return;
    compute_consts(FR_P, FRN, &FR_INV, FR_ONE, FR_R2);
    compute_consts(FP_P, FPN, &FP_INV, FP_ONE, FP_R2);
    g_init_done = 1;
}

/* ------------------------------------------------------------------------- */
/* G1 (Jacobian; restates kilic PointG1 behaviour used via bls/bls_kilic.go)  */
/* ------------------------------------------------------------------------- */
typedef struct { fp_t x, y, z; } g1_t;

static inline int g1_is_inf(const g1_t *p) { return fp_is_zero(&p->z); }
static inline void g1_set_inf(g1_t *p) { memset(p, 0, sizeof *p); } /* bls/bls_kilic.go:33 ClearG1 */

static void g1_dbl(g1_t *r, const g1_t *p)
{
    if (g1_is_inf(p)) { g1_set_inf(r); return; }
    fp_t a, b, c, d, e, f, t;
    fp_sqr(&a, &p->x);
    fp_sqr(&b, &p->y);
    fp_sqr(&c, &b);
    fp_add(&d, &p->x, &b);
    fp_sqr(&d, &d);
    fp_sub(&d, &d, &a);
    fp_sub(&d, &d, &c);
    fp_dbl(&d, &d);
    fp_dbl(&e, &a);
    fp_add(&e, &e, &a);
    fp_sqr(&f, &e);
    fp_mul(&t, &p->y, &p->z); /* before x/y are overwritten (r may alias p) */
    fp_dbl(&r->z, &t);
    fp_dbl(&t, &d);
    fp_sub(&r->x, &f, &t);
    fp_sub(&t, &d, &r->x);
    fp_mul(&t, &e, &t);
    fp_dbl(&c, &c);
    fp_dbl(&c, &c);
    fp_dbl(&c, &c);
    fp_sub(&r->y, &t, &c);
}

static void g1_add(g1_t *r, const g1_t *p, const g1_t *q) /* bls/bls_kilic.go:47 AddG1 */
{
    if (g1_is_inf(p)) { *r = *q; return; }
    if (g1_is_inf(q)) { *r = *p; return; }
    fp_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    fp_sqr(&z1z1, &p->z);
    fp_sqr(&z2z2, &q->z);
    fp_mul(&u1, &p->x, &z2z2);
    fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&s1, &p->y, &q->z);
    fp_mul(&s1, &s1, &z2z2);
    fp_mul(&s2, &q->y, &p->z);
    fp_mul(&s2, &s2, &z1z1);
    if (fp_eq(&u1, &u2)) {
        if (fp_eq(&s1, &s2)) { g1_dbl(r, p); return; }
        g1_set_inf(r);
        return;
    }
    fp_sub(&h, &u2, &u1);
    fp_dbl(&i, &h);
    fp_sqr(&i, &i);
    fp_mul(&j, &h, &i);
    fp_sub(&rr, &s2, &s1);
    fp_dbl(&rr, &rr);
    fp_mul(&v, &u1, &i);
    fp_add(&t, &p->z, &q->z);
    fp_sqr(&t, &t);
    fp_sub(&t, &t, &z1z1);
    fp_sub(&t, &t, &z2z2);
    fp_t x3, y3;
    fp_sqr(&x3, &rr);
    fp_sub(&x3, &x3, &j);
    fp_sub(&x3, &x3, &v);
    fp_sub(&x3, &x3, &v);
    fp_sub(&y3, &v, &x3);
    fp_mul(&y3, &rr, &y3);
    fp_mul(&s1, &s1, &j);
    fp_dbl(&s1, &s1);
    fp_sub(&y3, &y3, &s1);
    fp_mul(&r->z, &t, &h);
    r->x = x3;
    r->y = y3;
}
static void g1_neg(g1_t *r, const g1_t *p) /* bls/bls_kilic.go:63 NegG1 */
{
    *r = *p;
    if (!fp_is_zero(&r->y)) sub_n(r->y.l, FP_P, r->y.l, FPN);
}
static void g1_sub(g1_t *r, const g1_t *p, const g1_t *q) /* bls/bls_kilic.go:51 SubG1 */
{
    g1_t nq;
    g1_neg(&nq, q);
    g1_add(r, p, &nq);
}
/* bls/bls_kilic.go:41-45 MulG1: scalar is de-Montgomerised then MulScalar.  The kilic
 * routine itself is not in the tree; restated as a fixed 4-bit window ladder. */
static void g1_mul_canon(g1_t *r, const g1_t *p, const u64 *k /* canonical 4 limbs */)
{
    g1_t tab[16];
    g1_set_inf(&tab[0]);
    tab[1] = *p;
    for (int i = 2; i < 16; i++) {
        if (i & 1) g1_add(&tab[i], &tab[i - 1], p);
        else g1_dbl(&tab[i], &tab[i / 2]);
    }
    g1_t acc;
    g1_set_inf(&acc);
    for (int w = 63; w >= 0; w--) {
        for (int d = 0; d < 4; d++) g1_dbl(&acc, &acc);
        unsigned dig = (unsigned)(k[w / 16] >> ((w % 16) * 4)) & 15;
        if (dig) g1_add(&acc, &acc, &tab[dig]);
    }
    *r = acc;
}
static void g1_mul(g1_t *r, const g1_t *p, const fr_t *k)
{
    u64 c[FRN];
    fr_from_mont(c, k);
    g1_mul_canon(r, p, c);
}
static int g1_equal(const g1_t *a, const g1_t *b) /* bls/bls_kilic.go:106 EqualG1 (projective) */
{
    if (g1_is_inf(a) || g1_is_inf(b)) return g1_is_inf(a) && g1_is_inf(b);
    fp_t za, zb, l, r;
    fp_sqr(&za, &a->z);
    fp_sqr(&zb, &b->z);
    fp_mul(&l, &a->x, &zb);
    fp_mul(&r, &b->x, &za);
    if (!fp_eq(&l, &r)) return 0;
    fp_mul(&za, &za, &a->z);
    fp_mul(&zb, &zb, &b->z);
    fp_mul(&l, &a->y, &zb);
    fp_mul(&r, &b->y, &za);
    return fp_eq(&l, &r);
}
static void g1_to_affine(fp_t *x, fp_t *y, const g1_t *p)
{
    fp_t zi, zi2;
    fp_inv(&zi, &p->z);
    fp_sqr(&zi2, &zi);
    fp_mul(x, &p->x, &zi2);
    fp_mul(&zi2, &zi2, &zi);
    fp_mul(y, &p->y, &zi2);
}
static void g1_generator(g1_t *g) /* decimal coordinates at bls/bls_hbls.go:23-24 */
{
    static const u64 gx[FPN] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
    static const u64 gy[FPN] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
    fp_to_mont(&g->x, gx);
    fp_to_mont(&g->y, gy);
    memcpy(g->z.l, FP_ONE, sizeof g->z.l);
}

/* external (canonical) <-> internal (Montgomery) */
static void g1_load(g1_t *r, const u64 *ext)
{
    fp_to_mont(&r->x, ext);
    fp_to_mont(&r->y, ext + FPN);
    fp_to_mont(&r->z, ext + 2 * FPN);
}
static void g1_store(u64 *ext, const g1_t *p)
{
    fp_from_mont(ext, &p->x);
    fp_from_mont(ext + FPN, &p->y);
    fp_from_mont(ext + 2 * FPN, &p->z);
}
static void fr_load_vec(fr_t *dst, const u64 *ext, size_t n)
{
    for (size_t i = 0; i < n; i++) fr_to_mont(&dst[i], ext + i * FRN);
}
static void fr_store_vec(u64 *ext, const fr_t *src, size_t n)
{
    for (size_t i = 0; i < n; i++) fr_from_mont(ext + i * FRN, &src[i]);
}

/* 48-byte compression (bls/bls_kilic.go:114 ToCompressedG1; format pinned by
 * bls/bls_test.go:18 and eth/trusted_setup.json) */
static void g1_compress(uint8_t out[48], const g1_t *p)
{
    memset(out, 0, 48);
    if (g1_is_inf(p)) { out[0] = 0xC0; return; }
    fp_t x, y;
    g1_to_affine(&x, &y, p);
    u64 xc[FPN], yc[FPN], ny[FPN];
    fp_from_mont(xc, &x);
    fp_from_mont(yc, &y);
    for (int i = 0; i < 48; i++) out[i] = (uint8_t)(xc[(47 - i) / 8] >> (((47 - i) % 8) * 8));
    out[0] |= 0x80;
    sub_n(ny, FP_P, yc, FPN);                 /* y > -y  <=>  y > (p-1)/2 */
    if (!ge_n(ny, yc, FPN)) out[0] |= 0x20;
}
static int g1_decompress(g1_t *p, const uint8_t in[48]) /* bls/bls_kilic.go:118 FromCompressedG1 */
{
    if (!(in[0] & 0x80)) return -1;
    if (in[0] & 0x40) { g1_set_inf(p); return 0; }
    u64 xc[FPN] = {0};
    for (int i = 0; i < 48; i++) {
        uint8_t b = in[i];
        if (i == 0) b &= 0x1F;
        xc[(47 - i) / 8] |= (u64)b << (((47 - i) % 8) * 8);
    }
    if (ge_n(xc, FP_P, FPN)) return -1;
    fp_t x, y, t, four;
    fp_to_mont(&x, xc);
    fp_sqr(&t, &x);
    fp_mul(&t, &t, &x);
    u64 f[FPN] = {4, 0, 0, 0, 0, 0};
    fp_to_mont(&four, f);
    fp_add(&t, &t, &four);
    /* sqrt = t^((p+1)/4) */
    u64 e[FPN], one[FPN] = {1, 0, 0, 0, 0, 0};
    add_n(e, FP_P, one, FPN);
    for (int i = 0; i < FPN; i++) e[i] = (e[i] >> 2) | (i + 1 < FPN ? e[i + 1] << 62 : 0);
    fp_pow(&y, &t, e, FPN);
    fp_t chk;
    fp_sqr(&chk, &y);
    if (!fp_eq(&chk, &t)) return -2;
    u64 yc[FPN], ny[FPN];
    fp_from_mont(yc, &y);
    sub_n(ny, FP_P, yc, FPN);
    int y_big = !ge_n(ny, yc, FPN);
    if (y_big != !!(in[0] & 0x20)) fp_to_mont(&y, ny);
    p->x = x;
    p->y = y;
    memcpy(p->z.l, FP_ONE, sizeof p->z.l);
    /* kilic's FromCompressed also rejects points outside the prime-order subgroup: r * P must be infinity */
    g1_t rp;
    g1_mul_canon(&rp, p, FR_P);
    if (!g1_is_inf(&rp)) return -3;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* FFT settings (fft.go:21-61)                                                */
/* ------------------------------------------------------------------------- */
typedef struct {
    u64 max_width;
    fr_t *expanded; /* max_width+1 entries, first == last == 1 */
    fr_t *reverse;
} orc_fs;

/* bls/globals.go:27-60: Scale2RootOfUnity[k] = 7^((r-1)/2^k) */
static void scale2_root(fr_t *out, unsigned k)
{
    u64 e[FRN], one[FRN] = {1, 0, 0, 0};
    sub_n(e, FR_P, one, FRN);
    for (unsigned s = 0; s < k; s++)
        for (int i = 0; i < FRN; i++) e[i] = (e[i] >> 1) | (i + 1 < FRN ? e[i + 1] << 63 : 0);
    fr_t seven;
    fr_from_u64(&seven, 7);
    fr_pow(out, &seven, e, FRN);
}

orc_fs *orc_fs_new(unsigned max_scale) /* fft.go:44-61 */
{
    orc_init();
    orc_fs *fs = calloc(1, sizeof *fs);
    fs->max_width = (u64)1 << max_scale;
    fs->expanded = malloc((fs->max_width + 1) * sizeof(fr_t));
    fs->reverse = malloc((fs->max_width + 1) * sizeof(fr_t));
    fr_t root;
    scale2_root(&root, max_scale);
    memcpy(fs->expanded[0].l, FR_ONE, sizeof(fr_t)); /* fft.go:21-32 */
    for (u64 i = 1; i <= fs->max_width; i++) fr_mul(&fs->expanded[i], &fs->expanded[i - 1], &root);
    for (u64 i = 0; i <= fs->max_width; i++) fs->reverse[i] = fs->expanded[fs->max_width - i];
    return fs;
}
void orc_fs_free(orc_fs *fs)
{
    if (!fs) return;
    free(fs->expanded);
    free(fs->reverse);
    free(fs);
}
u64 orc_fs_max_width(const orc_fs *fs) { return fs->max_width; }
void orc_fs_expanded_root(const orc_fs *fs, u64 i, u64 *out4) { fr_from_mont(out4, &fs->expanded[i]); }

/* fft_fr.go:8-28 simpleFT */
static void simple_ft(const fr_t *vals, u64 off, u64 stride, const fr_t *roots, u64 rstride, fr_t *out, u64 l)
{
    fr_t v, last;
    for (u64 i = 0; i < l; i++) {
        fr_mul(&v, &vals[off], &roots[0]);
        last = v;
        for (u64 j = 1; j < l; j++) {
            fr_mul(&v, &vals[off + j * stride], &roots[((i * j) % l) * rstride]);
            fr_add(&last, &last, &v);
        }
        out[i] = last;
    }
}
/* fft_fr.go:30-53 _fft */
static void fft_rec(const fr_t *vals, u64 off, u64 stride, const fr_t *roots, u64 rstride, fr_t *out, u64 l)
{
    if (l <= 4) { simple_ft(vals, off, stride, roots, rstride, out, l); return; }
    u64 half = l >> 1;
    fft_rec(vals, off, stride << 1, roots, rstride << 1, out, half);
    fft_rec(vals, off + stride, stride << 1, roots, rstride << 1, out + half, half);
    fr_t t, x, y;
    for (u64 i = 0; i < half; i++) {
        x = out[i];
        y = out[i + half];
        fr_mul(&t, &y, &roots[i * rstride]);
        fr_add(&out[i], &x, &t);
        fr_sub(&out[i + half], &x, &t);
    }
}
/* fft_fr.go:76-105 InplaceFFT (n must be a power of two <= MaxWidth; vals untouched) */
static int inplace_fft(const orc_fs *fs, const fr_t *vals, fr_t *out, u64 n, int inv)
{
    if (n > fs->max_width) return 1;
    if (n & (n - 1)) return 2;
    u64 stride = fs->max_width / n;
    if (inv) {
        fr_t inv_len;
        fr_from_u64(&inv_len, n);
        fr_inv(&inv_len, &inv_len);
        fft_rec(vals, 0, 1, fs->reverse, stride, out, n);
        for (u64 i = 0; i < n; i++) fr_mul(&out[i], &out[i], &inv_len);
    } else {
        fft_rec(vals, 0, 1, fs->expanded, stride, out, n);
    }
    return 0;
}
static u64 next_pow2(u64 v) /* fft.go:11-16 */
{
    if (v == 0) return 1;
    u64 p = 1;
    while (p < v) p <<= 1;
    return p;
}
/* fft_fr.go:55-74 FFT: zero-pad to next power of two; out has next_pow2(n) entries.
 * returns 0 ok, 1 too large */
int orc_fft_fr(const orc_fs *fs, const u64 *in, u64 n, int inv, u64 *out)
{
    if (n > fs->max_width) return 1;
    u64 np = next_pow2(n);
    fr_t *v = calloc(np, sizeof(fr_t)), *o = malloc(np * sizeof(fr_t));
    fr_load_vec(v, in, n);
    int rc = inplace_fft(fs, v, o, np, inv);
    if (!rc) fr_store_vec(out, o, np);
    free(v);
    free(o);
    return rc;
}

/* fft_g1.go:11-31 simpleFTG1 (note: also multiplies by the root 1, as the Go does) */
static void simple_ft_g1(const g1_t *vals, u64 off, u64 stride, const fr_t *roots, u64 rstride, g1_t *out, u64 l)
{
    g1_t v, last;
    for (u64 i = 0; i < l; i++) {
        g1_mul(&v, &vals[off], &roots[0]);
        last = v;
        for (u64 j = 1; j < l; j++) {
            g1_mul(&v, &vals[off + j * stride], &roots[((i * j) % l) * rstride]);
            g1_add(&last, &last, &v);
        }
        out[i] = last;
    }
}
/* fft_g1.go:33-56 _fftG1 */
static void fft_g1_rec(const g1_t *vals, u64 off, u64 stride, const fr_t *roots, u64 rstride, g1_t *out, u64 l)
{
    if (l <= 4) { simple_ft_g1(vals, off, stride, roots, rstride, out, l); return; }
    u64 half = l >> 1;
    fft_g1_rec(vals, off, stride << 1, roots, rstride << 1, out, half);
    fft_g1_rec(vals, off + stride, stride << 1, roots, rstride << 1, out + half, half);
    g1_t t, x, y;
    for (u64 i = 0; i < half; i++) {
        x = out[i];
        y = out[i + half];
        g1_mul(&t, &y, &roots[i * rstride]);
        g1_add(&out[i], &x, &t);
        g1_sub(&out[i + half], &x, &t);
    }
}
/* The same recursion, unrolled at the top so that independent pieces can run on several host
 * threads (the reference is single-threaded; the arithmetic performed is identical): the 2^d
 * sub-transforms of depth d are independent, then each of the d merge levels is a loop of
 * independent butterflies. */
typedef struct { const g1_t *vals; const fr_t *roots; u64 rstride; g1_t *out; u64 n; unsigned d; unsigned lev; } fftp_ctx;
static void fftp_leaf(long long j, void *c_)
{
    fftp_ctx *c = c_;
    u64 sub = c->n >> c->d, pos = 0;
    for (unsigned b = 0; b < c->d; b++) if ((u64)j & ((u64)1 << b)) pos |= (u64)1 << (c->d - 1 - b);
    fft_g1_rec(c->vals, (u64)j, (u64)1 << c->d, c->roots, c->rstride << c->d, c->out + pos * sub, sub);
}
static void fftp_merge(long long q, void *c_)
{
    fftp_ctx *c = c_;
    u64 len = c->n >> c->lev, half = len >> 1;
    u64 blk = (u64)q / half, i = (u64)q % half;
    g1_t *o = c->out + blk * len;
    g1_t t, x = o[i], y = o[i + half];
    g1_mul(&t, &y, &c->roots[i * (c->rstride << c->lev)]);
    g1_add(&o[i], &x, &t);
    g1_sub(&o[i + half], &x, &t);
}
static void fft_g1_top(const g1_t *vals, const fr_t *roots, u64 rstride, g1_t *out, u64 n)
{
    unsigned logn = 0, d = 0;
    while (((u64)1 << logn) < n) logn++;
    if (logn > 5) d = logn - 3 < 6 ? logn - 3 : 6;   /* sub-transforms stay >= 8 points (leaf rule l <= 4 untouched) */
    if (d == 0) { fft_g1_rec(vals, 0, 1, roots, rstride, out, n); return; }
    fftp_ctx c = {vals, roots, rstride, out, n, d, 0};
    parallel_for((long long)1 << d, fftp_leaf, &c);
    for (unsigned lev = d; lev-- > 0;) {
        c.lev = lev;
        parallel_for((long long)(n / 2), fftp_merge, &c);
    }
}
typedef struct { g1_t *out; const g1_t *in; const fr_t *k; int per_elem; } mulv_ctx;
static void mulv_body(long long i, void *c_)
{
    mulv_ctx *c = c_;
    g1_mul(&c->out[i], &c->in[i], c->per_elem ? &c->k[i] : c->k);
}
/* fft_g1.go:58-94 FFTG1 on internal points. returns 0 ok, 1 too large, 2 not pow2 */
static int fft_g1_int(const orc_fs *fs, const g1_t *vals, u64 n, int inv, g1_t *out)
{
    if (n > fs->max_width) return 1;
    if (n & (n - 1)) return 2;
    if (n == 0) return 0;
    u64 stride = fs->max_width / n;
    if (inv) {
        fr_t inv_len;
        fr_from_u64(&inv_len, n);
        fr_inv(&inv_len, &inv_len);
        fft_g1_top(vals, fs->reverse, stride, out, n);
        mulv_ctx mc = {out, out, &inv_len, 0};   /* fft_g1.go:81-84 */
        parallel_for((long long)n, mulv_body, &mc);
    } else {
        fft_g1_top(vals, fs->expanded, stride, out, n);
    }
    return 0;
}
int orc_fft_g1(const orc_fs *fs, const u64 *in, u64 n, int inv, u64 *out)
{
    if (n > fs->max_width) return 1;
    if (n & (n - 1)) return 2;
    g1_t *v = malloc((n + 1) * sizeof(g1_t)), *o = malloc((n + 1) * sizeof(g1_t));
    for (u64 i = 0; i < n; i++) g1_load(&v[i], in + i * 18);
    int rc = fft_g1_int(fs, v, n, inv, o);
    if (!rc) for (u64 i = 0; i < n; i++) g1_store(out + i * 18, &o[i]);
    free(v);
    free(o);
    return rc;
}

/* reverse_bit_order.go:81-101 */
static uint32_t reverse_bits_limited(uint32_t length, uint32_t value)
{
    unsigned bits = 0;
    while (((uint32_t)1 << bits) < length) bits++;
    uint32_t out = 0;
    for (unsigned i = 0; i < bits; i++)
        if (value >> i & 1) out |= (uint32_t)1 << (bits - 1 - i);
    return out;
}
static void reverse_bit_order_g1(g1_t *v, uint32_t n) /* fft_g1.go:97-107 */
{
    for (uint32_t i = 0; i < n; i++) {
        uint32_t r = reverse_bits_limited(n, i);
        if (r > i) { g1_t t = v[i]; v[i] = v[r]; v[r] = t; }
    }
}

/* ------------------------------------------------------------------------- */
/* level-1 helpers exposed for tests                                          */
/* ------------------------------------------------------------------------- */
void orc_g1_generator(u64 *out18) { orc_init(); g1_t g; g1_generator(&g); g1_store(out18, &g); }
void orc_g1_mul(u64 *out18, const u64 *p18, const u64 *k4)
{
    orc_init();
    g1_t p, r;
    g1_load(&p, p18);
    g1_mul_canon(&r, &p, k4);
    g1_store(out18, &r);
}
void orc_g1_add(u64 *out18, const u64 *a18, const u64 *b18)
{
    orc_init();
    g1_t a, b, r;
    g1_load(&a, a18);
    g1_load(&b, b18);
    g1_add(&r, &a, &b);
    g1_store(out18, &r);
}
void orc_g1_sub(u64 *out18, const u64 *a18, const u64 *b18)
{
    orc_init();
    g1_t a, b, r;
    g1_load(&a, a18);
    g1_load(&b, b18);
    g1_sub(&r, &a, &b);
    g1_store(out18, &r);
}
int orc_g1_equal(const u64 *a18, const u64 *b18)
{
    orc_init();
    g1_t a, b;
    g1_load(&a, a18);
    g1_load(&b, b18);
    return g1_equal(&a, &b);
}
void orc_g1_compress(uint8_t *out48, const u64 *p18)
{
    orc_init();
    g1_t p;
    g1_load(&p, p18);
    g1_compress(out48, &p);
}
typedef struct { uint8_t *bytes; u64 *pts; const u64 *scal; const uint8_t *cbytes; const u64 *cpts; int bad; } many_ctx;
static void compress_body(long long i, void *c_)
{
    many_ctx *c = c_;
    g1_t p;
    g1_load(&p, c->cpts + i * 18);
    g1_compress(c->bytes + i * 48, &p);
}
void orc_g1_compress_many(uint8_t *out, const u64 *pts, u64 n)
{
    orc_init();
    many_ctx c = {.bytes = out, .cpts = pts};
    parallel_for((long long)n, compress_body, &c);
}
int orc_g1_decompress(u64 *out18, const uint8_t *in48)
{
    orc_init();
    g1_t p;
    int rc = g1_decompress(&p, in48);
    if (!rc) g1_store(out18, &p);
    return rc;
}
static void decompress_body(long long i, void *c_)
{
    many_ctx *c = c_;
    g1_t p;
    if (g1_decompress(&p, c->cbytes + i * 48)) { c->bad = 1; return; }
    g1_store(c->pts + i * 18, &p);
}
int orc_g1_decompress_many(u64 *out, const uint8_t *in, u64 n)
{
    orc_init();
    many_ctx c = {.pts = out, .cbytes = in};
    parallel_for((long long)n, decompress_body, &c);
    return c.bad;
}
/* k[i]*G for many canonical scalars (exponent-domain oracle: setup secret known) */
static void mul_gen_body(long long i, void *c_)
{
    many_ctx *c = c_;
    g1_t g, r;
    g1_generator(&g);
    g1_mul_canon(&r, &g, c->scal + i * 4);
    g1_store(c->pts + i * 18, &r);
}
void orc_g1_mul_gen_many(u64 *out, const u64 *scalars, u64 n)
{
    orc_init();
    many_ctx c = {.pts = out, .scal = scalars};
    parallel_for((long long)n, mul_gen_body, &c);
}

/* setup.go:9-26 GenerateTestingSetup (G1 half) */
void orc_generate_setup_g1(const u64 *secret4, u64 n, u64 *out)
{
    orc_init();
    fr_t s, spow;
    fr_to_mont(&s, secret4);
    memcpy(spow.l, FR_ONE, sizeof spow.l);
    g1_t g;
    g1_generator(&g);
    u64 *pows = malloc((n + 1) * 4 * sizeof(u64));
    for (u64 i = 0; i < n; i++) { fr_from_mont(pows + i * 4, &spow); fr_mul(&spow, &spow, &s); }
    (void)g;
    many_ctx c = {.pts = out, .scal = pows};
    parallel_for((long long)n, mul_gen_body, &c);
    free(pows);
}

/* ------------------------------------------------------------------------- */
/* LinCombG1 (bls/bls_kilic.go:132-150 -> kilic G1.MultiExp, bucket method)   */
/* ------------------------------------------------------------------------- */
static void lincomb_int(g1_t *out, const g1_t *pts, const u64 *scal /* canonical n x 4 */, u64 n)
{
    g1_set_inf(out);
    if (n == 0) return; /* bls/bls_test.go:69-77: empty => infinity */
    /* kilic MultiExp (recalled): c = 3 for n < 32 else ceil(ln n) */
    int c = 3;
    if (n >= 32) {
        double ln = 0;
        /* ceil(log(n)) without libm: e^k table */
        static const double ek[] = {1, 2.718281828459045, 7.38905609893065, 20.085536923187668, 54.598150033144236,
                                    148.4131591025766, 403.4287934927351, 1096.6331584284585, 2980.9579870417283,
                                    8103.083927575384, 22026.465794806718, 59874.14171519782, 162754.79141900392,
                                    442413.3920089205, 1202604.2841647768, 3269017.3724721107, 8886110.520507872,
                                    24154952.7535753, 65659969.13733051, 178482300.96318725, 485165195.4097903};
        c = 0;
        while (c < 20 && ek[c] < (double)n) c++;
        (void)ln;
    }
    int bucket_size = (1 << c) - 1;
    int nwin = 255 / c + 1;
    g1_t *bucket = malloc((size_t)bucket_size * sizeof(g1_t));
    g1_t *windows = malloc((size_t)nwin * sizeof(g1_t));
    for (int j = 0; j < nwin; j++) {
        for (int i = 0; i < bucket_size; i++) g1_set_inf(&bucket[i]);
        for (u64 i = 0; i < n; i++) {
            int bit = j * c;
            const u64 *k = scal + i * 4;
            u64 idx = 0;
            if (bit < 256) {
                idx = k[bit / 64] >> (bit % 64);
                if (bit % 64 + c > 64 && bit / 64 + 1 < 4) idx |= k[bit / 64 + 1] << (64 - bit % 64);
                idx &= (u64)bucket_size;
            }
            if (idx) g1_add(&bucket[idx - 1], &bucket[idx - 1], &pts[i]);
        }
        g1_t acc, sum;
        g1_set_inf(&acc);
        g1_set_inf(&sum);
        for (int i = bucket_size - 1; i >= 0; i--) {
            g1_add(&sum, &sum, &bucket[i]);
            g1_add(&acc, &acc, &sum);
        }
        windows[j] = acc;
    }
    g1_t acc;
    g1_set_inf(&acc);
    for (int i = nwin - 1; i >= 0; i--) {
        for (int j = 0; j < c; j++) g1_dbl(&acc, &acc);
        g1_add(&acc, &acc, &windows[i]);
    }
    *out = acc;
    free(bucket);
    free(windows);
}
/* kzg_single_proofs.go:17-19 CommitToPoly = LinCombG1(SecretG1[:len], coeffs) */
void orc_lincomb_g1(const u64 *pts, const u64 *scalars, u64 n, u64 *out18)
{
    orc_init();
    g1_t *p = malloc((n + 1) * sizeof(g1_t));
    for (u64 i = 0; i < n; i++) g1_load(&p[i], pts + i * 18);
    g1_t r;
    lincomb_int(&r, p, scalars, n);
    g1_store(out18, &r);
    free(p);
}

/* ------------------------------------------------------------------------- */
/* FK20 (kzg.go:38-116, fk20_single.go, fk20_multi.go)                        */
/* ------------------------------------------------------------------------- */
/* fk20_single.go:40-56 toeplitzPart1 */
static g1_t *toeplitz_part1(const orc_fs *fs, const g1_t *x, u64 n)
{
    u64 n2 = n * 2;
    g1_t *ext = calloc(n2, sizeof(g1_t)); /* zero == infinity */
    memcpy(ext, x, n * sizeof(g1_t));
    g1_t *out = malloc(n2 * sizeof(g1_t));
    if (fft_g1_int(fs, ext, n2, 0, out)) { free(out); out = NULL; }
    free(ext);
    return out;
}
/* fk20_single.go:59-77 ToeplitzPart2 */
static g1_t *toeplitz_part2(const orc_fs *fs, const fr_t *coeffs, const g1_t *x_ext_fft, u64 n2)
{
    fr_t *cf = malloc(n2 * sizeof(fr_t));
    inplace_fft(fs, coeffs, cf, n2, 0);
    g1_t *h = malloc(n2 * sizeof(g1_t));
    mulv_ctx mc = {h, x_ext_fft, cf, 1};   /* fk20_single.go:72-74 */
    parallel_for((long long)n2, mulv_body, &mc);
    free(cf);
    return h;
}
/* fk20_single.go:80-87 ToeplitzPart3 (returns full 2n array; caller keeps first half) */
static g1_t *toeplitz_part3(const orc_fs *fs, const g1_t *h_ext_fft, u64 n2)
{
    g1_t *out = malloc(n2 * sizeof(g1_t));
    fft_g1_int(fs, h_ext_fft, n2, 1, out);
    return out;
}
/* fk20_single.go:106-119 */
static void toeplitz_coeffs_step(fr_t *out /*2n*/, const fr_t *poly, u64 n)
{
    memset(out, 0, 2 * n * sizeof(fr_t));
    out[0] = poly[n - 1];
    for (u64 i = n + 2, j = 1; i < 2 * n; i++, j++) out[i] = poly[j];
}
/* fk20_single.go:89-103 */
static void toeplitz_coeffs_step_strided(fr_t *out /*2k*/, const fr_t *poly, u64 n, u64 offset, u64 stride)
{
    u64 k = n / stride, k2 = 2 * k;
    memset(out, 0, k2 * sizeof(fr_t));
    out[0] = poly[n - 1 - offset];
    for (u64 i = k + 2, j = 2 * stride - offset - 1; i < k2; i++, j += stride) out[i] = poly[j];
}

typedef struct {
    orc_fs *fs;
    u64 n2;           /* extended size */
    u64 chunk_len;    /* 1 for FK20 single */
    g1_t **x_ext_fft; /* chunk_len files of 2k points */
    g1_t *secret_g1;
    u64 secret_len;
} orc_fk;

/* kzg.go:21-36 NewKZGSettings checks + kzg.go:43-64 NewFK20SingleSettings /
 * kzg.go:73-116 NewFK20MultiSettings.  secret_g1: external points.
 * returns NULL on the conditions the Go panics on. */
orc_fk *orc_fk20_new(unsigned max_scale, const u64 *secret_g1, u64 secret_len, u64 n2, u64 chunk_len)
{
    orc_init();
    orc_fs *fs = orc_fs_new(max_scale);
    if (secret_len < fs->max_width || n2 > fs->max_width || (n2 & (n2 - 1)) || n2 < 2 || chunk_len > n2 / 2 ||
        (chunk_len & (chunk_len - 1)) || chunk_len < 1) {
        orc_fs_free(fs);
        return NULL;
    }
    orc_fk *fk = calloc(1, sizeof *fk);
    fk->fs = fs;
    fk->n2 = n2;
    fk->chunk_len = chunk_len;
    fk->secret_len = secret_len;
    fk->secret_g1 = malloc(secret_len * sizeof(g1_t));
    for (u64 i = 0; i < secret_len; i++) g1_load(&fk->secret_g1[i], secret_g1 + i * 18);
    fk->x_ext_fft = calloc(chunk_len, sizeof(g1_t *));
    u64 n = n2 / 2, k = n / chunk_len;
    for (u64 off = 0; off < chunk_len; off++) {
        g1_t *x = calloc(k, sizeof(g1_t));
        /* kzg.go:103-109 (chunk_len == 1 reduces to kzg.go:57-61) */
        u64 start = n - chunk_len - 1 - off;
        for (u64 i = 0, j = start; i + 1 < k; i++, j -= chunk_len) x[i] = fk->secret_g1[j];
        g1_set_inf(&x[k - 1]);
        fk->x_ext_fft[off] = toeplitz_part1(fs, x, k);
        free(x);
    }
    return fk;
}
void orc_fk20_free(orc_fk *fk)
{
    if (!fk) return;
    for (u64 i = 0; i < fk->chunk_len; i++) free(fk->x_ext_fft[i]);
    free(fk->x_ext_fft);
    free(fk->secret_g1);
    orc_fs_free(fk->fs);
    free(fk);
}
/* copy of xExtFFT file `file` (2k external points) */
void orc_fk20_x_ext_fft(const orc_fk *fk, u64 file, u64 *out)
{
    u64 k2 = fk->n2 / fk->chunk_len;
    for (u64 i = 0; i < k2; i++) g1_store(out + i * 18, &fk->x_ext_fft[file][i]);
}

/* kzg_single_proofs.go:17-19 */
void orc_commit_to_poly(const orc_fk *fk, const u64 *coeffs, u64 n, u64 *out18)
{
    g1_t r;
    lincomb_int(&r, fk->secret_g1, coeffs, n);
    g1_store(out18, &r);
}

/* fk20_single.go:122-134 FK20Single (da=0: n proofs natural order)
 * fk20_single.go:139-196 DAUsingFK20 (da=1: 2n proofs, bit-reversed) */
int orc_fk20_single(const orc_fk *fk, const u64 *poly, u64 n, int da, u64 *out)
{
    if (fk->chunk_len != 1 || 2 * n != fk->n2 || (n & (n - 1))) return 1;
    const orc_fs *fs = fk->fs;
    u64 n2 = 2 * n;
    fr_t *p = malloc(n * sizeof(fr_t)), *c = malloc(n2 * sizeof(fr_t));
    fr_load_vec(p, poly, n);
    toeplitz_coeffs_step(c, p, n);
    g1_t *h_ext_fft = toeplitz_part2(fs, c, fk->x_ext_fft[0], n2);
    g1_t *h = toeplitz_part3(fs, h_ext_fft, n2);
    g1_t *res;
    u64 nout;
    if (da) {
        for (u64 i = n; i < n2; i++) g1_set_inf(&h[i]); /* fk20_single.go:163-166 */
        nout = n2;
    } else {
        nout = n;
    }
    res = malloc(nout * sizeof(g1_t));
    fft_g1_int(fs, h, nout, 0, res);
    if (da) reverse_bit_order_g1(res, (uint32_t)nout); /* fk20_single.go:194 */
    for (u64 i = 0; i < nout; i++) g1_store(out + i * 18, &res[i]);
    free(p); free(c); free(h_ext_fft); free(h); free(res);
    return 0;
}

/* fk20_multi.go:58-133 DAUsingFK20Multi: poly has n coefficients, returns 2k proofs bit-reversed */
int orc_fk20_multi_da(const orc_fk *fk, const u64 *poly, u64 n, u64 *out)
{
    if (2 * n != fk->n2 || (n & (n - 1))) return 1;
    const orc_fs *fs = fk->fs;
    u64 l = fk->chunk_len, k = n / l, k2 = 2 * k;
    fr_t *p = malloc(n * sizeof(fr_t)), *c = malloc(k2 * sizeof(fr_t));
    fr_load_vec(p, poly, n);
    g1_t *h_ext_fft = calloc(k2, sizeof(g1_t));
    for (u64 i = 0; i < l; i++) {
        toeplitz_coeffs_step_strided(c, p, n, i, l);
        g1_t *file = toeplitz_part2(fs, c, fk->x_ext_fft[i], k2);
        for (u64 j = 0; j < k2; j++) g1_add(&h_ext_fft[j], &h_ext_fft[j], &file[j]); /* fk20_multi.go:86-89 */
        free(file);
    }
    g1_t *h = toeplitz_part3(fs, h_ext_fft, k2);
    for (u64 i = k; i < k2; i++) g1_set_inf(&h[i]); /* fk20_multi.go:100-103 */
    g1_t *res = malloc(k2 * sizeof(g1_t));
    fft_g1_int(fs, h, k2, 0, res);
    reverse_bit_order_g1(res, (uint32_t)k2); /* fk20_multi.go:131 */
    for (u64 i = 0; i < k2; i++) g1_store(out + i * 18, &res[i]);
    free(p); free(c); free(h_ext_fft); free(h); free(res);
    return 0;
}

/* The headline unit of work: CommitToPoly + FK20Single for `nblobs` polynomials of n
 * coefficients.  With several blobs the blobs are spread over host threads (one blob per
 * thread); with fewer blobs than threads the independent butterflies inside each transform are
 * spread instead.  The reference itself is single-threaded. */
typedef struct { const orc_fk *fk; const u64 *polys; u64 n; u64 *commits, *proofs; int rc; } batch_ctx;
static void batch_body(long long b, void *c_)
{
    batch_ctx *c = c_;
    orc_commit_to_poly(c->fk, c->polys + b * c->n * 4, c->n, c->commits + b * 18);
    if (orc_fk20_single(c->fk, c->polys + b * c->n * 4, c->n, 0, c->proofs + b * c->n * 18)) c->rc = 1;
}
int orc_commit_fk20_batch(const orc_fk *fk, const u64 *polys, u64 n, u64 nblobs, u64 *commits, u64 *proofs, int nthreads)
{
    batch_ctx c = {fk, polys, n, commits, proofs, 0};
    int saved = g_threads;
    if (nthreads > 0) g_threads = nthreads;
    int nt = g_threads > 0 ? g_threads : orc_max_threads();
    if ((long long)nblobs >= nt) parallel_for((long long)nblobs, batch_body, &c);
    else for (u64 b = 0; b < nblobs; b++) batch_body((long long)b, &c);   /* threads go inside the transforms */
    g_threads = saved;
    return c.rc;
}

/* ------------------------------------------------------------------------- */
/* DAS extension (das_extension.go:7-84)                                      */
/* ------------------------------------------------------------------------- */
static void das_ext_rec(const orc_fs *fs, fr_t *ab, u64 len, u64 stride)
{
    if (len == 2) { /* das_extension.go:8-20 */
        fr_t x, y, t;
        fr_add(&x, &ab[0], &ab[1]);
        fr_sub(&y, &ab[0], &ab[1]);
        fr_mul(&t, &y, &fs->expanded[stride]);
        fr_add(&ab[0], &x, &t);
        fr_sub(&ab[1], &x, &t);
        return;
    }
    u64 hh = len >> 1;
    fr_t *a0 = ab, *a1 = ab + hh, t1, t2;
    for (u64 i = 0; i < hh; i++) { /* das_extension.go:34-41 */
        fr_add(&t1, &a0[i], &a1[i]);
        fr_sub(&t2, &a0[i], &a1[i]);
        fr_mul(&a1[i], &t2, &fs->reverse[i * 2 * stride]);
        a0[i] = t1;
    }
    das_ext_rec(fs, a0, hh, stride << 1);
    das_ext_rec(fs, a1, hh, stride << 1);
    fr_t x, y, t;
    for (u64 i = 0; i < hh; i++) { /* das_extension.go:55-65 */
        x = a0[i];
        y = a1[i];
        fr_mul(&t, &y, &fs->expanded[(1 + 2 * i) * stride]);
        fr_add(&a0[i], &x, &t);
        fr_sub(&a1[i], &x, &t);
    }
}
/* das_extension.go:71-84; in place. returns 1 if the Go would panic (domain too small),
 * 2 for n < 2 ("bad usage") */
int orc_das_fft_extension(const orc_fs *fs, u64 *vals, u64 n)
{
    if (n * 2 > fs->max_width) return 1;
    if (n < 2 || (n & (n - 1))) return 2;
    fr_t *v = malloc(n * sizeof(fr_t));
    fr_load_vec(v, vals, n);
    das_ext_rec(fs, v, n, 1);
    fr_t inv_len;
    fr_from_u64(&inv_len, n);
    fr_inv(&inv_len, &inv_len);
    for (u64 i = 0; i < n; i++) fr_mul(&v[i], &v[i], &inv_len);
    fr_store_vec(vals, v, n);
    free(v);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Zero polynomial (zero_poly.go)                                             */
/* ------------------------------------------------------------------------- */
/* zero_poly.go:17-39 */
static void make_zero_poly_mul_leaf(const orc_fs *fs, fr_t *dst, u64 dst_len, const u64 *indices, u64 cnt, u64 stride)
{
    for (u64 i = cnt + 1; i < dst_len; i++) memset(&dst[i], 0, sizeof(fr_t));
    memcpy(dst[cnt].l, FR_ONE, sizeof(fr_t));
    fr_t neg, zero;
    memset(&zero, 0, sizeof zero);
    for (u64 i = 0; i < cnt; i++) {
        fr_sub(&neg, &zero, &fs->expanded[indices[i] * stride]);
        dst[i] = neg;
        if (i > 0) {
            fr_add(&dst[i], &dst[i], &dst[i - 1]);
            for (u64 j = i - 1; j > 0; j--) {
                fr_mul(&dst[j], &dst[j], &neg);
                fr_add(&dst[j], &dst[j], &dst[j - 1]);
            }
            fr_mul(&dst[0], &dst[0], &neg);
        }
    }
}
typedef struct { u64 off, len; } span_t;
/* zero_poly.go:58-107 reduceLeaves; dst/ps are spans of `out`; returns resulting length */
static u64 reduce_leaves(const orc_fs *fs, fr_t *scratch, fr_t *out, span_t dst, const span_t *ps, u64 nps)
{
    u64 n = dst.len;
    u64 out_degree = 0;
    for (u64 i = 0; i < nps; i++) out_degree += ps[i].len - 1;
    fr_t *p_padded = scratch, *mul_eval = scratch + n, *p_eval = scratch + 2 * n;
    u64 last = nps - 1;
    /* zero_poly.go:42-49 padPoly */
    for (u64 i = 0; i < ps[last].len; i++) p_padded[i] = out[ps[last].off + i];
    for (u64 i = ps[last].len; i < n; i++) memset(&p_padded[i], 0, sizeof(fr_t));
    inplace_fft(fs, p_padded, mul_eval, n, 0);
    for (u64 i = 0; i < last; i++) {
        for (u64 j = 0; j < ps[i].len; j++) p_padded[j] = out[ps[i].off + j]; /* partial overwrite, as the Go */
        inplace_fft(fs, p_padded, p_eval, n, 0);
        for (u64 j = 0; j < n; j++) fr_mul(&mul_eval[j], &mul_eval[j], &p_eval[j]);
    }
    inplace_fft(fs, mul_eval, out + dst.off, n, 1);
    return out_degree + 1;
}
/* zero_poly.go:116-217 ZeroPolyViaMultiplication -> zero_eval[length], zero_poly[length]
 * returns 0 ok; 1 domain too small; 2 length not pow2; 3 internal size panic */
static int zero_poly_int(const orc_fs *fs, const u64 *missing, u64 nmiss, u64 length, fr_t *zero_eval, fr_t *zero_poly)
{
    if (nmiss == 0) {
        memset(zero_eval, 0, length * sizeof(fr_t));
        memset(zero_poly, 0, length * sizeof(fr_t));
        return 0;
    }
    if (length > fs->max_width) return 1;
    if (length & (length - 1)) return 2;
    u64 stride = fs->max_width / length;
    const u64 per_leaf_poly = 64, per_leaf = 63;
    if (nmiss <= per_leaf) {
        memset(zero_poly, 0, length * sizeof(fr_t));
        if (nmiss + 1 > length) return 3;
        make_zero_poly_mul_leaf(fs, zero_poly, nmiss + 1, missing, nmiss, stride);
        return inplace_fft(fs, zero_poly, zero_eval, length, 0);
    }
    u64 leaf_count = (nmiss + per_leaf - 1) / per_leaf;
    u64 n = next_pow2(leaf_count * per_leaf_poly);
    fr_t *out = calloc(n, sizeof(fr_t));
    span_t *leaves = malloc(leaf_count * sizeof(span_t));
    u64 off = 0;
    for (u64 i = 0; i < leaf_count; i++) {
        u64 end = off + per_leaf;
        if (end > nmiss) end = nmiss;
        leaves[i].off = i * per_leaf_poly;
        leaves[i].len = per_leaf_poly;
        make_zero_poly_mul_leaf(fs, out + leaves[i].off, per_leaf_poly, missing + off, end - off, stride);
        off += per_leaf;
    }
    const u64 rf = 4;
    fr_t *scratch = malloc(3 * n * sizeof(fr_t));
    u64 nleaves = leaf_count;
    int rc = 0;
    while (nleaves > 1 && !rc) {
        u64 reduced_count = (nleaves + rf - 1) / rf;
        u64 leaf_size = next_pow2(leaves[0].len);
        for (u64 i = 0; i < reduced_count; i++) {
            u64 start = i * rf, end = start + rf;
            u64 out_end = end * leaf_size;
            if (out_end > n) out_end = n;
            span_t reduced = {start * leaf_size, out_end - start * leaf_size};
            if (end > nleaves) end = nleaves;
            if (end > start + 1) {
                u64 deg = 0;
                for (u64 q = start; q < end; q++) deg += leaves[q].len - 1;
                if ((reduced.len & (reduced.len - 1)) || deg + 1 > reduced.len) { rc = 3; break; }
                reduced.len = reduce_leaves(fs, scratch, out, reduced, leaves + start, end - start);
            }
            leaves[i] = reduced;
        }
        nleaves = reduced_count;
    }
    if (!rc) {
        if (leaves[0].len > length) rc = 3;
        else {
            memset(zero_poly, 0, length * sizeof(fr_t));
            memcpy(zero_poly, out + leaves[0].off, leaves[0].len * sizeof(fr_t));
            rc = inplace_fft(fs, zero_poly, zero_eval, length, 0);
        }
    }
    free(out); free(leaves); free(scratch);
    return rc;
}
int orc_zero_poly(const orc_fs *fs, const u64 *missing, u64 nmiss, u64 length, u64 *zero_eval, u64 *zero_poly)
{
    fr_t *ze = malloc(length * sizeof(fr_t)), *zp = malloc(length * sizeof(fr_t));
    int rc = zero_poly_int(fs, missing, nmiss, length, ze, zp);
    if (!rc) { fr_store_vec(zero_eval, ze, length); fr_store_vec(zero_poly, zp, length); }
    free(ze); free(zp);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Recovery (recover_from_samples.go)                                         */
/* ------------------------------------------------------------------------- */
static void shift_poly(fr_t *poly, u64 n, int unshift) /* recover_from_samples.go:9-40 */
{
    fr_t factor, power;
    fr_from_u64(&factor, 5);
    if (!unshift) fr_inv(&factor, &factor);
    memcpy(power.l, FR_ONE, sizeof power.l);
    for (u64 i = 0; i < n; i++) {
        fr_mul(&poly[i], &poly[i], &power);
        fr_mul(&power, &power, &factor);
    }
}
/* recover_from_samples.go:42-109. present[i]==0 <=> samples[i]==nil.
 * returns 0 ok; 1 FFT too large; 4 "bad zero eval" panic; 5 reconstruct mismatch error;
 * other codes from zero poly. n must be a power of two (FFT pads otherwise; unsupported here). */
int orc_recover_poly_from_samples(const orc_fs *fs, const u64 *samples, const uint8_t *present, u64 n, u64 *out)
{
    if (n > fs->max_width) return 1;
    if (n & (n - 1)) return 2;
    u64 *missing = malloc((n + 1) * sizeof(u64)), nmiss = 0;
    for (u64 i = 0; i < n; i++) if (!present[i]) missing[nmiss++] = i;
    fr_t *s = malloc(n * sizeof(fr_t)), *ze = malloc(n * sizeof(fr_t)), *zp = malloc(n * sizeof(fr_t));
    fr_t *a = malloc(n * sizeof(fr_t)), *b = malloc(n * sizeof(fr_t)), *c = malloc(n * sizeof(fr_t));
    fr_load_vec(s, samples, n);
    int rc = zero_poly_int(fs, missing, nmiss, n, ze, zp);
    if (!rc) {
        for (u64 i = 0; i < n; i++)
            if ((!present[i]) != fr_is_zero(&ze[i])) { rc = 4; break; }
    }
    if (!rc) {
        for (u64 i = 0; i < n; i++) {
            if (!present[i]) memset(&a[i], 0, sizeof(fr_t));
            else fr_mul(&a[i], &s[i], &ze[i]);
        }
        inplace_fft(fs, a, b, n, 1);   /* polyWithZero */
        shift_poly(b, n, 0);
        shift_poly(zp, n, 0);
        inplace_fft(fs, b, a, n, 0);   /* evalShiftedPolyWithZero */
        inplace_fft(fs, zp, c, n, 0);  /* evalShiftedZeroPoly */
        for (u64 i = 0; i < n; i++) {  /* recover_from_samples.go:89-91 DivModFr */
            fr_t inv;
            fr_inv(&inv, &c[i]);
            fr_mul(&a[i], &inv, &a[i]);
        }
        inplace_fft(fs, a, b, n, 1);
        shift_poly(b, n, 1);
        inplace_fft(fs, b, a, n, 0);
        for (u64 i = 0; i < n; i++)
            if (present[i] && !fr_eq(&a[i], &s[i])) { rc = 5; break; }
        if (!rc) fr_store_vec(out, a, n);
    }
    free(missing); free(s); free(ze); free(zp); free(a); free(b); free(c);
    return rc;
}
