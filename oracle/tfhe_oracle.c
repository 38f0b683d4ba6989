/*
 * oracle/tfhe_oracle.c — CPU ORACLE (test infrastructure, NOT the product).
 * See tfhe_oracle.h for scope, provenance and parity status ("parity unpinned"
 * at the raw-coefficient level; decrypted identities of main.cpp are pinned).
 *
 * Every function cites the step of SURVEY.md §8c it restates and, where one
 * exists, the reference call site that reaches it.  [UPSTREAM-RECALL] marks
 * behaviour of the un-vendored TFHE v1.1 library restated from its published
 * algorithm.
 */
#include "tfhe_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct oracle_keys {
    oracle_params p;
    int32_t* lwe_key;  /* n */
    int32_t* tlwe_key; /* N */
    int32_t* bk;       /* [n][(k+1)l][k+1][N] */
    int32_t* ksk;      /* [N][t][base][n+1] */
    /* lazily built FP64 transform of bk: [n][(k+1)l][k+1][N/2] re, im (bit-reversed bin order) */
    double* bkfft_re;
    double* bkfft_im;
    pthread_mutex_t fft_lock;
};

/* ------------------------------------------------------------------ params */

/* [UPSTREAM-RECALL] new_default_gate_bootstrapping_parameters; called at main.cpp:21 with 128. */
int oracle_default_params(int32_t minimum_lambda, oracle_params* out) {
    if (minimum_lambda > 128) return -1;
    out->N = 1024;
    out->k = 1;
    out->ks_t = 8;
    out->ks_basebit = 2;
    if (minimum_lambda > 80) {
        out->n = 630;
        out->l = 3;
        out->Bgbit = 7;
        out->ks_stdev = ldexp(1.0, -15);
        out->bk_stdev = ldexp(1.0, -25);
    } else {
        out->n = 500;
        out->l = 2;
        out->Bgbit = 10;
        out->ks_stdev = 2.44e-5;
        out->bk_stdev = 7.18e-9;
    }
    return 0;
}

size_t oracle_bk_words(const oracle_params* p) {
    return (size_t)p->n * (size_t)((p->k + 1) * p->l) * (size_t)(p->k + 1) * (size_t)p->N;
}
size_t oracle_ksk_words(const oracle_params* p) {
    return (size_t)p->N * (size_t)p->ks_t * (size_t)(1 << p->ks_basebit) * (size_t)(p->n + 1);
}

/* --------------------------------------------------------------------- rng */

uint64_t oracle_rng_next(uint64_t* s) { /* splitmix64 */
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double rng_uniform01(uint64_t* s) { return ((double)(oracle_rng_next(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double rng_gaussian(uint64_t* s, double sigma) { /* Box-Muller */
    double u1 = rng_uniform01(s), u2 = rng_uniform01(s);
    return sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}
/* [UPSTREAM-RECALL] dtot32: fractional part of d scaled to 2^32. */
static Torus32o dtot32(double d) { return (int32_t)(int64_t)((d - (double)(int64_t)d) * 4294967296.0); }
static Torus32o gaussian32(uint64_t* s, Torus32o message, double sigma) {
    return (Torus32o)((uint32_t)message + (uint32_t)dtot32(rng_gaussian(s, sigma)));
}

/* --------------------------------------------------------- torus utilities */

/* §8c step 1: interval = floor(2^63/M)*2, result = high 32 bits of mu*interval. */
Torus32o oracle_mod_switch_to_torus32(int32_t mu, int32_t Msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)Msize) * 2;
    uint64_t phase64 = (uint64_t)(int64_t)mu * interv;
    return (Torus32o)(phase64 >> 32);
}
/* §8c step 2: ((uint64(x)<<32) + interval/2) / interval. */
int32_t oracle_mod_switch_from_torus32(Torus32o phase, int32_t Msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)Msize) * 2;
    uint64_t half = interv / 2;
    uint64_t phase64 = ((uint64_t)(uint32_t)phase << 32) + half;
    return (int32_t)(phase64 / interv);
}

/* ------------------------------------------------------------------ keygen */

const oracle_params* oracle_keys_params(const oracle_keys* k) { return &k->p; }
const int32_t* oracle_lwe_key(const oracle_keys* k) { return k->lwe_key; }
const int32_t* oracle_tlwe_key(const oracle_keys* k) { return k->tlwe_key; }
const int32_t* oracle_bk(const oracle_keys* k) { return k->bk; }
const int32_t* oracle_ksk(const oracle_keys* k) { return k->ksk; }

static oracle_keys* keys_alloc(const oracle_params* p) {
    oracle_keys* K = (oracle_keys*)calloc(1, sizeof(oracle_keys));
    K->p = *p;
    K->lwe_key = (int32_t*)calloc((size_t)p->n, sizeof(int32_t));
    K->tlwe_key = (int32_t*)calloc((size_t)p->N, sizeof(int32_t));
    K->bk = (int32_t*)calloc(oracle_bk_words(p), sizeof(int32_t));
    K->ksk = (int32_t*)calloc(oracle_ksk_words(p), sizeof(int32_t));
    pthread_mutex_init(&K->fft_lock, NULL);
    return K;
}

/* negacyclic product of a torus polynomial by a binary key polynomial, added into out (mod 2^32). */
static void poly_addmul_binary(uint32_t* out, const int32_t* key_bits, const uint32_t* a, int N) {
    for (int m = 0; m < N; ++m) {
        if (!key_bits[m]) continue;
        for (int j = 0; j < N; ++j) {
            int idx = j - m;
            if (idx >= 0) out[j] += a[idx];
            else out[j] -= a[idx + N];
        }
    }
}

/* SURVEY.md §3.3 [UPSTREAM-RECALL]: LWE key, ring key, keyswitch key (noise recentred), then
   n TGSW encryptions of the LWE key bits.  Reached from main.cpp:22. */
oracle_keys* oracle_keygen(const oracle_params* p, uint64_t seed) {
    oracle_keys* K = keys_alloc(p);
    uint64_t s = seed;
    const int n = p->n, N = p->N, l = p->l, t = p->ks_t, base = 1 << p->ks_basebit;
    for (int i = 0; i < n; ++i) K->lwe_key[i] = (int32_t)(oracle_rng_next(&s) & 1);
    for (int i = 0; i < N; ++i) K->tlwe_key[i] = (int32_t)(oracle_rng_next(&s) & 1);

    /* keyswitch key: KSK[i][j][h] encrypts h * s'_i * 2^(32-(j+1)basebit) under the LWE key, sigma = ks_stdev;
       h = 0 is the noiseless zero sample; the noises are recentred to zero mean as upstream does. */
    {
        const size_t sizeks = (size_t)N * t * (base - 1);
        double* noise = (double*)malloc(sizeks * sizeof(double));
        double err = 0;
        for (size_t q = 0; q < sizeks; ++q) { noise[q] = rng_gaussian(&s, p->ks_stdev); err += noise[q]; }
        err /= (double)sizeks;
        for (size_t q = 0; q < sizeks; ++q) noise[q] -= err;
        size_t index = 0;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < t; ++j)
                for (int h = 1; h < base; ++h) {
                    int32_t* ct = K->ksk + (((size_t)i * t + j) * base + h) * (size_t)(n + 1);
                    uint32_t mess = (uint32_t)(K->tlwe_key[i] * h) * (UINT32_C(1) << (32 - (j + 1) * p->ks_basebit));
                    uint32_t b = mess + (uint32_t)dtot32(noise[index++]);
                    for (int q = 0; q < n; ++q) {
                        uint32_t a = (uint32_t)oracle_rng_next(&s);
                        ct[q] = (int32_t)a;
                        if (K->lwe_key[q]) b += a;
                    }
                    ct[n] = (int32_t)b;
                }
        free(noise);
    }
    /* bootstrapping key: BK_i = TGSW(s_i): (k+1)l rows, each a TLWE encryption of 0 (b = a*s + e), plus
       s_i * 2^(32-(p+1)Bgbit) on coefficient 0 of polynomial `bloc` in row bloc*l+p. */
    {
        const int rows = (p->k + 1) * l;
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < rows; ++r) {
                uint32_t* a = (uint32_t*)K->bk + (((size_t)i * rows + r) * 2 + 0) * (size_t)N;
                uint32_t* b = (uint32_t*)K->bk + (((size_t)i * rows + r) * 2 + 1) * (size_t)N;
                for (int j = 0; j < N; ++j) b[j] = (uint32_t)gaussian32(&s, 0, p->bk_stdev);
                for (int j = 0; j < N; ++j) a[j] = (uint32_t)oracle_rng_next(&s);
                poly_addmul_binary(b, K->tlwe_key, a, N);
                const int bloc = r / l, pp = r % l;
                uint32_t h = UINT32_C(1) << (32 - (pp + 1) * p->Bgbit);
                uint32_t* tgt = bloc == 0 ? a : b;
                tgt[0] += (uint32_t)K->lwe_key[i] * h;
            }
    }
    return K;
}

oracle_keys* oracle_keys_from_arrays(const oracle_params* p, const int32_t* lwe_key, const int32_t* tlwe_key,
                                     const int32_t* bk, const int32_t* ksk) {
    oracle_keys* K = keys_alloc(p);
    if (lwe_key) memcpy(K->lwe_key, lwe_key, (size_t)p->n * sizeof(int32_t));
    if (tlwe_key) memcpy(K->tlwe_key, tlwe_key, (size_t)p->N * sizeof(int32_t));
    memcpy(K->bk, bk, oracle_bk_words(p) * sizeof(int32_t));
    memcpy(K->ksk, ksk, oracle_ksk_words(p) * sizeof(int32_t));
    return K;
}

void oracle_keys_free(oracle_keys* K) {
    if (!K) return;
    free(K->lwe_key); free(K->tlwe_key); free(K->bk); free(K->ksk);
    free(K->bkfft_re); free(K->bkfft_im);
    pthread_mutex_destroy(&K->fft_lock);
    free(K);
}

/* --------------------------------------------------------- encrypt/decrypt */

/* bootsSymEncrypt (main.cpp:64) [UPSTREAM-RECALL]: mu = +-1/8, sigma = ks_stdev, b = noise + mu + <a,s>. */
void oracle_encrypt_bit(const oracle_keys* K, int32_t bit, uint64_t* rng, int32_t* ct) {
    const int n = K->p.n;
    Torus32o mu = oracle_mod_switch_to_torus32(1, 8);
    uint32_t b = (uint32_t)gaussian32(rng, bit ? mu : -mu, K->p.ks_stdev);
    for (int i = 0; i < n; ++i) {
        uint32_t a = (uint32_t)oracle_rng_next(rng);
        ct[i] = (int32_t)a;
        if (K->lwe_key[i]) b += a;
    }
    ct[n] = (int32_t)b;
}
Torus32o oracle_phase(const oracle_keys* K, const int32_t* ct) {
    const int n = K->p.n;
    uint32_t axs = 0;
    for (int i = 0; i < n; ++i) if (K->lwe_key[i]) axs += (uint32_t)ct[i];
    return (Torus32o)((uint32_t)ct[n] - axs);
}
/* bootsSymDecrypt (main.cpp:78): 1 iff phase > 0. */
int32_t oracle_decrypt_bit(const oracle_keys* K, const int32_t* ct) { return oracle_phase(K, ct) > 0 ? 1 : 0; }

/* bootsCONSTANT (Math.cpp:58): (a=0, b=+-1/8). */
void oracle_trivial(const oracle_params* p, int32_t bit, int32_t* ct) {
    Torus32o mu = oracle_mod_switch_to_torus32(1, 8);
    memset(ct, 0, (size_t)p->n * sizeof(int32_t));
    ct[p->n] = bit ? mu : -mu;
}
/* bootsNOT (Math.cpp:153): negate all n+1 words. */
void oracle_not(const oracle_params* p, const int32_t* in, int32_t* out) {
    for (int i = 0; i <= p->n; ++i) out[i] = (int32_t)(0u - (uint32_t)in[i]);
}

/* ------------------------------------------------------------ gate prologue */

/* §8a/§8c step 1: (c0 in eighths, sa, sb). */
static int gate_consts(int op, int* c0_8, int* sa, int* sb) {
    switch (op) {
        case ORACLE_OP_NAND:  *c0_8 = 1;  *sa = -1; *sb = -1; return 0;
        case ORACLE_OP_AND:   *c0_8 = -1; *sa = 1;  *sb = 1;  return 0;
        case ORACLE_OP_OR:    *c0_8 = 1;  *sa = 1;  *sb = 1;  return 0;
        case ORACLE_OP_XOR:   *c0_8 = 2;  *sa = 2;  *sb = 2;  return 0;
        case ORACLE_OP_XNOR:  *c0_8 = -2; *sa = -2; *sb = -2; return 0;
        case ORACLE_OP_NOR:   *c0_8 = -1; *sa = -1; *sb = -1; return 0;
        case ORACLE_OP_ANDNY: *c0_8 = -1; *sa = -1; *sb = 1;  return 0;
        case ORACLE_OP_ANDYN: *c0_8 = -1; *sa = 1;  *sb = -1; return 0;
        case ORACLE_OP_ORNY:  *c0_8 = 1;  *sa = -1; *sb = 1;  return 0;
        case ORACLE_OP_ORYN:  *c0_8 = 1;  *sa = 1;  *sb = -1; return 0;
        default: return -1;
    }
}
static void lincomb(const oracle_params* p, int c0_8, int sa, const int32_t* ca, int sb, const int32_t* cb, int32_t* t) {
    const int n = p->n;
    for (int i = 0; i < n; ++i)
        t[i] = (int32_t)((uint32_t)sa * (uint32_t)ca[i] + (uint32_t)sb * (uint32_t)cb[i]);
    t[n] = (int32_t)((uint32_t)oracle_mod_switch_to_torus32(c0_8, 8) + (uint32_t)sa * (uint32_t)ca[n] +
                     (uint32_t)sb * (uint32_t)cb[n]);
}
int oracle_gate_prologue(const oracle_params* p, int op, const int32_t* ca, const int32_t* cb, int32_t* t) {
    int c0, sa, sb;
    if (gate_consts(op, &c0, &sa, &sb)) return -1;
    lincomb(p, c0, sa, ca, sb, cb, t);
    return 0;
}

/* ------------------------------------------------ folded negacyclic FP64 FFT */

/* A real polynomial p of degree < N is evaluated at zeta^(4k+1), zeta = exp(i*pi/N), k < N/2, as the
   N/2-point DFT of z_j = (p_j + i p_{j+N/2}) * zeta^j.  Forward: decimation in frequency (natural in,
   bit-reversed out); inverse: decimation in time (bit-reversed in, natural out); pointwise products are
   taken in bit-reversed order so no reordering pass exists. */
typedef struct {
    int N, M;            /* M = N/2 */
    double *tw_re, *tw_im; /* exp(+2 pi i j / M), j < M/2 */
    double *zt_re, *zt_im; /* zeta^j, j < M */
} fft_tables;

static fft_tables* g_tab = NULL;
static pthread_mutex_t g_tab_lock = PTHREAD_MUTEX_INITIALIZER;

static const fft_tables* get_tables(int N) {
    pthread_mutex_lock(&g_tab_lock);
    if (!g_tab || g_tab->N != N) {
        fft_tables* T = (fft_tables*)calloc(1, sizeof(fft_tables));
        T->N = N; T->M = N / 2;
        T->tw_re = (double*)malloc(sizeof(double) * (size_t)T->M / 2);
        T->tw_im = (double*)malloc(sizeof(double) * (size_t)T->M / 2);
        T->zt_re = (double*)malloc(sizeof(double) * (size_t)T->M);
        T->zt_im = (double*)malloc(sizeof(double) * (size_t)T->M);
        const long double PI = 3.14159265358979323846264338327950288L;
        for (int j = 0; j < T->M / 2; ++j) {
            T->tw_re[j] = (double)cosl(2 * PI * j / T->M);
            T->tw_im[j] = (double)sinl(2 * PI * j / T->M);
        }
        for (int j = 0; j < T->M; ++j) {
            T->zt_re[j] = (double)cosl(PI * j / N);
            T->zt_im[j] = (double)sinl(PI * j / N);
        }
        g_tab = T; /* an older table (different N) is leaked deliberately: other threads may hold it */
    }
    const fft_tables* r = g_tab;
    pthread_mutex_unlock(&g_tab_lock);
    return r;
}

static void fft_forward(const fft_tables* T, double* re, double* im) {
    const int M = T->M;
    for (int len = M / 2, step = 1; len >= 1; len >>= 1, step <<= 1)
        for (int blk = 0; blk < M; blk += 2 * len)
            for (int j = 0; j < len; ++j) {
                const double wr = T->tw_re[j * step], wi = T->tw_im[j * step];
                double* ar = re + blk + j; double* ai = im + blk + j;
                const double ur = ar[0], ui = ai[0], vr = ar[len], vi = ai[len];
                ar[0] = ur + vr; ai[0] = ui + vi;
                const double dr = ur - vr, di = ui - vi;
                ar[len] = dr * wr - di * wi; ai[len] = dr * wi + di * wr;
            }
}
static void fft_inverse(const fft_tables* T, double* re, double* im) { /* unscaled */
    const int M = T->M;
    for (int len = 1, step = M / 2; len <= M / 2; len <<= 1, step >>= 1)
        for (int blk = 0; blk < M; blk += 2 * len)
            for (int j = 0; j < len; ++j) {
                const double wr = T->tw_re[j * step], wi = -T->tw_im[j * step];
                double* ar = re + blk + j; double* ai = im + blk + j;
                const double ur = ar[0], ui = ai[0];
                const double vr = ar[len] * wr - ai[len] * wi, vi = ar[len] * wi + ai[len] * wr;
                ar[0] = ur + vr; ai[0] = ui + vi;
                ar[len] = ur - vr; ai[len] = ui - vi;
            }
}
static void poly_to_fft(const fft_tables* T, const int32_t* p, double* re, double* im) {
    const int M = T->M;
    for (int j = 0; j < M; ++j) {
        const double a = (double)p[j], b = (double)p[j + M];
        re[j] = a * T->zt_re[j] - b * T->zt_im[j];
        im[j] = a * T->zt_im[j] + b * T->zt_re[j];
    }
    fft_forward(T, re, im);
}
/* inverse transform, round to nearest, add into acc (mod 2^32). */
static void fft_to_poly_add(const fft_tables* T, double* re, double* im, uint32_t* acc) {
    const int M = T->M;
    fft_inverse(T, re, im);
    const double inv = 1.0 / (double)M;
    for (int j = 0; j < M; ++j) {
        const double zr = re[j] * inv, zi = im[j] * inv;
        const double a = zr * T->zt_re[j] + zi * T->zt_im[j];
        const double b = zi * T->zt_re[j] - zr * T->zt_im[j];
        acc[j] += (uint32_t)(uint64_t)llrint(a);
        acc[j + M] += (uint32_t)(uint64_t)llrint(b);
    }
}

static void ensure_bkfft(oracle_keys* K) {
    pthread_mutex_lock(&K->fft_lock);
    if (!K->bkfft_re) {
        const oracle_params* p = &K->p;
        const fft_tables* T = get_tables(p->N);
        const size_t polys = (size_t)p->n * (size_t)((p->k + 1) * p->l) * 2;
        const size_t M = (size_t)p->N / 2;
        double* re = (double*)malloc(polys * M * sizeof(double));
        double* im = (double*)malloc(polys * M * sizeof(double));
        for (size_t q = 0; q < polys; ++q) poly_to_fft(T, K->bk + q * (size_t)p->N, re + q * M, im + q * M);
        K->bkfft_im = im;
        K->bkfft_re = re;
    }
    pthread_mutex_unlock(&K->fft_lock);
}

/* --------------------------------------------------------- external product */

typedef struct {
    int32_t* T;      /* [2][N] rotated difference */
    int32_t* dig;    /* [(k+1)l][N] digits */
    uint32_t* ext;   /* [2N] sign-extended key row */
    double *fre, *fim; /* [(k+1)l][N/2] */
    double *are, *aim; /* [2][N/2] */
} scratch;

static scratch* scratch_new(const oracle_params* p) {
    scratch* s = (scratch*)calloc(1, sizeof(scratch));
    const size_t N = (size_t)p->N, rows = (size_t)(p->k + 1) * p->l;
    s->T = (int32_t*)malloc(2 * N * sizeof(int32_t));
    s->dig = (int32_t*)malloc(rows * N * sizeof(int32_t));
    s->ext = (uint32_t*)malloc(2 * N * sizeof(uint32_t));
    s->fre = (double*)malloc(rows * N / 2 * sizeof(double));
    s->fim = (double*)malloc(rows * N / 2 * sizeof(double));
    s->are = (double*)malloc(N * sizeof(double));
    s->aim = (double*)malloc(N * sizeof(double));
    return s;
}
static void scratch_free(scratch* s) {
    free(s->T); free(s->dig); free(s->ext); free(s->fre); free(s->fim); free(s->are); free(s->aim); free(s);
}

/* §8c step 3/4: out = X^a * src (negacyclic), 0 <= a < 2N.  [UPSTREAM-RECALL] torusPolynomialMulByXai. */
static void poly_mul_by_xai(int32_t* out, int a, const int32_t* src, int N) {
    if (a < N) {
        for (int i = 0; i < a; ++i) out[i] = (int32_t)(0u - (uint32_t)src[i - a + N]);
        for (int i = a; i < N; ++i) out[i] = src[i - a];
    } else {
        const int aa = a - N;
        for (int i = 0; i < aa; ++i) out[i] = src[i - aa + N];
        for (int i = aa; i < N; ++i) out[i] = (int32_t)(0u - (uint32_t)src[i - aa]);
    }
}

/* §8c step 4: acc <- acc + BK_i (.) ((X^abar - 1) acc). */
static void cmux_step(oracle_keys* K, int mode, scratch* s, int i, int abar, int32_t* acc) {
    const oracle_params* p = &K->p;
    const int N = p->N, l = p->l, Bgbit = p->Bgbit, rows = (p->k + 1) * l;
    const uint32_t halfBg = UINT32_C(1) << (Bgbit - 1), maskMod = (UINT32_C(1) << Bgbit) - 1;
    uint32_t offset = 0;
    for (int q = 1; q <= l; ++q) offset += halfBg << (32 - q * Bgbit);

    for (int c = 0; c < 2; ++c) {
        poly_mul_by_xai(s->T + c * N, abar, acc + c * N, N);
        for (int j = 0; j < N; ++j) s->T[c * N + j] = (int32_t)((uint32_t)s->T[c * N + j] - (uint32_t)acc[c * N + j]);
        for (int q = 0; q < l; ++q) {
            const int decal = 32 - (q + 1) * Bgbit;
            int32_t* d = s->dig + (size_t)(c * l + q) * N;
            for (int j = 0; j < N; ++j)
                d[j] = (int32_t)((((uint32_t)s->T[c * N + j] + offset) >> decal) & maskMod) - (int32_t)halfBg;
        }
    }
    const int32_t* bki = K->bk + (size_t)i * rows * 2 * N;
    if (mode == ORACLE_EXACT) {
        uint32_t* uacc = (uint32_t*)acc;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < 2; ++c) {
                const uint32_t* key = (const uint32_t*)bki + ((size_t)r * 2 + c) * N;
                for (int x = 0; x < N; ++x) { s->ext[x + N] = key[x]; s->ext[x] = 0u - key[x]; }
                const int32_t* d = s->dig + (size_t)r * N;
                uint32_t* out = uacc + c * N;
                for (int m = 0; m < N; ++m) {
                    const uint32_t dm = (uint32_t)d[m];
                    if (!dm) continue;
                    const uint32_t* e = s->ext + (N - m);
                    for (int j = 0; j < N; ++j) out[j] += dm * e[j];
                }
            }
    } else {
        ensure_bkfft(K);
        const fft_tables* T = get_tables(N);
        const int M = N / 2;
        for (int r = 0; r < rows; ++r) poly_to_fft(T, s->dig + (size_t)r * N, s->fre + (size_t)r * M, s->fim + (size_t)r * M);
        for (int c = 0; c < 2; ++c) {
            double* are = s->are + c * M; double* aim = s->aim + c * M;
            for (int q = 0; q < M; ++q) { are[q] = 0; aim[q] = 0; }
            for (int r = 0; r < rows; ++r) {
                const double* kr = K->bkfft_re + (((size_t)i * rows + r) * 2 + c) * M;
                const double* ki = K->bkfft_im + (((size_t)i * rows + r) * 2 + c) * M;
                const double* dr = s->fre + (size_t)r * M; const double* di = s->fim + (size_t)r * M;
                for (int q = 0; q < M; ++q) {
                    are[q] += dr[q] * kr[q] - di[q] * ki[q];
                    aim[q] += dr[q] * ki[q] + di[q] * kr[q];
                }
            }
            fft_to_poly_add(T, are, aim, (uint32_t*)acc + c * N);
        }
    }
}

void oracle_cmux_step(const oracle_keys* K, int mode, int32_t i, int32_t abar, int32_t* acc) {
    scratch* s = scratch_new(&K->p);
    cmux_step((oracle_keys*)K, mode, s, i, abar, acc);
    scratch_free(s);
}

/* ------------------------------------------------------- bootstrap + keyswitch */

static void bootstrap_woks(oracle_keys* K, int mode, scratch* s, const int32_t* in, int32_t* out) {
    const oracle_params* p = &K->p;
    const int n = p->n, N = p->N, Nx2 = 2 * N;
    const Torus32o mu = oracle_mod_switch_to_torus32(1, 8);
    int32_t* acc = (int32_t*)malloc(2 * (size_t)N * sizeof(int32_t));
    int32_t* tv = (int32_t*)malloc((size_t)N * sizeof(int32_t));
    /* step 2: mod switch */
    const int barb = oracle_mod_switch_from_torus32(in[n], Nx2);
    /* step 3: ACC = (0, X^(2N-barb) * v), v = (mu,...,mu) */
    for (int j = 0; j < N; ++j) { tv[j] = mu; acc[j] = 0; }
    if (barb != 0) poly_mul_by_xai(acc + N, Nx2 - barb, tv, N);
    else memcpy(acc + N, tv, (size_t)N * sizeof(int32_t));
    /* step 4: blind rotation */
    for (int i = 0; i < n; ++i) {
        const int abar = oracle_mod_switch_from_torus32(in[i], Nx2);
        if (abar == 0) continue;
        cmux_step(K, mode, s, i, abar, acc);
    }
    /* step 5: sample extract at index 0 */
    out[0] = acc[0];
    for (int j = 1; j < N; ++j) out[j] = (int32_t)(0u - (uint32_t)acc[N - j]);
    out[N] = acc[N];
    free(acc); free(tv);
}

void oracle_bootstrap_woks(const oracle_keys* K, int mode, const int32_t* in, int32_t* out) {
    scratch* s = scratch_new(&K->p);
    bootstrap_woks((oracle_keys*)K, mode, s, in, out);
    scratch_free(s);
}

/* §8c step 6 [UPSTREAM-RECALL] lweKeySwitch. */
void oracle_keyswitch(const oracle_keys* K, const int32_t* in, int32_t* out) {
    const oracle_params* p = &K->p;
    const int n = p->n, N = p->N, t = p->ks_t, basebit = p->ks_basebit, base = 1 << basebit;
    const uint32_t prec_offset = UINT32_C(1) << (32 - (1 + basebit * t));
    const uint32_t mask = (uint32_t)base - 1;
    uint32_t* r = (uint32_t*)malloc((size_t)(n + 1) * sizeof(uint32_t));
    memset(r, 0, (size_t)n * sizeof(uint32_t));
    r[n] = (uint32_t)in[N];
    for (int i = 0; i < N; ++i) {
        const uint32_t aibar = (uint32_t)in[i] + prec_offset;
        for (int j = 0; j < t; ++j) {
            const uint32_t aij = (aibar >> (32 - (j + 1) * basebit)) & mask;
            if (aij != 0) {
                const uint32_t* row = (const uint32_t*)K->ksk + (((size_t)i * t + j) * base + aij) * (size_t)(n + 1);
                for (int q = 0; q <= n; ++q) r[q] -= row[q];
            }
        }
    }
    memcpy(out, r, (size_t)(n + 1) * sizeof(uint32_t));
    free(r);
}

static int gate_one(oracle_keys* K, int mode, scratch* s, int op, const int32_t* ca, const int32_t* cb,
                    const int32_t* cc, int32_t* out) {
    const oracle_params* p = &K->p;
    const int n = p->n, N = p->N;
    int32_t* t = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t));
    int32_t* u = (int32_t*)malloc((size_t)(N + 1) * sizeof(int32_t));
    int rc = 0;
    if (op == ORACLE_OP_MUX) {
        /* §8c step 7: u1 = PBS(-1/8 + a + b), u2 = PBS(-1/8 - a + c), KS((0,1/8) + u1 + u2). */
        int32_t* u2 = (int32_t*)malloc((size_t)(N + 1) * sizeof(int32_t));
        lincomb(p, -1, 1, ca, 1, cb, t);
        bootstrap_woks(K, mode, s, t, u);
        lincomb(p, -1, -1, ca, 1, cc, t);
        bootstrap_woks(K, mode, s, t, u2);
        for (int j = 0; j < N; ++j) u[j] = (int32_t)((uint32_t)u[j] + (uint32_t)u2[j]);
        u[N] = (int32_t)((uint32_t)oracle_mod_switch_to_torus32(1, 8) + (uint32_t)u[N] + (uint32_t)u2[N]);
        oracle_keyswitch(K, u, out);
        free(u2);
    } else {
        int c0, sa, sb;
        if (gate_consts(op, &c0, &sa, &sb)) rc = -1;
        else {
            lincomb(p, c0, sa, ca, sb, cb, t);
            bootstrap_woks(K, mode, s, t, u);
            oracle_keyswitch(K, u, out);
        }
    }
    free(t); free(u);
    return rc;
}

int oracle_gate(const oracle_keys* K, int mode, int op, const int32_t* ca, const int32_t* cb, const int32_t* cc,
                int32_t* out) {
    scratch* s = scratch_new(&K->p);
    int rc = gate_one((oracle_keys*)K, mode, s, op, ca, cb, cc, out);
    scratch_free(s);
    return rc;
}

typedef struct {
    oracle_keys* K; int mode, op; const int32_t *ca, *cb, *cc; int32_t* out; int64_t begin, end; int rc;
} batch_job;

static void* batch_worker(void* arg) {
    batch_job* J = (batch_job*)arg;
    const size_t w = (size_t)J->K->p.n + 1;
    scratch* s = scratch_new(&J->K->p);
    for (int64_t g = J->begin; g < J->end; ++g) {
        int rc = gate_one(J->K, J->mode, s, J->op, J->ca + (size_t)g * w, J->cb + (size_t)g * w,
                          J->cc ? J->cc + (size_t)g * w : NULL, J->out + (size_t)g * w);
        if (rc) J->rc = rc;
    }
    scratch_free(s);
    return NULL;
}

int oracle_gate_batch(const oracle_keys* K, int mode, int op, const int32_t* ca, const int32_t* cb,
                      const int32_t* cc, int32_t* out, int64_t count, int threads) {
    if (threads < 1) threads = 1;
    if (threads > count) threads = count > 0 ? (int)count : 1;
    if (mode == ORACLE_FFT) ensure_bkfft((oracle_keys*)K);
    batch_job* jobs = (batch_job*)calloc((size_t)threads, sizeof(batch_job));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    int rc = 0;
    for (int t = 0; t < threads; ++t) {
        jobs[t].K = (oracle_keys*)K; jobs[t].mode = mode; jobs[t].op = op;
        jobs[t].ca = ca; jobs[t].cb = cb; jobs[t].cc = cc; jobs[t].out = out;
        jobs[t].begin = count * t / threads; jobs[t].end = count * (t + 1) / threads;
        if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    for (int t = 0; t < threads; ++t) if (jobs[t].rc) rc = jobs[t].rc;
    free(jobs); free(th);
    return rc;
}
