/* blugen.c -- deterministic synthetic simplex-style bases for tests and bench.
 *
 * One generator feeds both the CUDA path and the CPU oracle (SURVEY.md 8d,
 * BASELINE.md section 5).  RNG: xoshiro256** seeded through splitmix64.
 * Values are uniform in [-1,-0.1] U [0.1,1]; one entry per column lies on a
 * random perfect matching with magnitude in [1,2] so the matrix has full
 * structural rank.  Row indices inside a column are distinct and ascending.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { uint64_t s[4]; } rng_t;

static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed) {
    for (int k = 0; k < 4; k++) r->s[k] = splitmix64(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static double rng_unif(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static int64_t rng_below(rng_t *r, int64_t n) { return (int64_t)(rng_unif(r) * (double)n); }
/* uniform in [-1,-0.1] U [0.1,1] */
static double rng_value(rng_t *r) {
    double v = 0.1 + 0.9 * rng_unif(r);
    return (rng_next(r) & 1) ? v : -v;
}
static int64_t rng_poisson(rng_t *r, double mean) {
    double l = exp(-mean), p = 1.0;
    int64_t k = 0;
    do { k++; p *= rng_unif(r); } while (p > l);
    return k - 1;
}
static void rng_perm(rng_t *r, int64_t n, int64_t *p) {
    for (int64_t i = 0; i < n; i++) p[i] = i;
    for (int64_t i = n - 1; i > 0; i--) {
        int64_t j = rng_below(r, i + 1);
        int64_t t = p[i]; p[i] = p[j]; p[j] = t;
    }
}
static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* Simplex-style basis: nslack unit columns e_{match(j)} (value 1.0) at random
 * column positions, the other columns have 1+Poisson(pmean) entries (capped at
 * cap).  colptr has m+1 entries.  Pass rowidx == NULL to only count.
 * Returns nnz. */
int64_t blugen_basis(uint64_t seed, int64_t m, int64_t nslack, double pmean, int64_t cap,
                     int64_t *colptr, int64_t *rowidx, double *value) {
    rng_t r;
    rng_seed(&r, seed);
    int64_t *match = malloc((size_t)m * sizeof(int64_t));
    int64_t *cpos = malloc((size_t)m * sizeof(int64_t));
    char *isslack = calloc((size_t)m, 1);
    int64_t *rows = malloc((size_t)(cap + 1) * sizeof(int64_t));
    rng_perm(&r, m, match);
    rng_perm(&r, m, cpos);
    for (int64_t k = 0; k < nslack; k++) isslack[cpos[k]] = 1;
    int64_t nnz = 0;
    for (int64_t j = 0; j < m; j++) {
        if (colptr) colptr[j] = nnz;
        if (isslack[j]) {
            if (rowidx) { rowidx[nnz] = match[j]; value[nnz] = 1.0; }
            nnz++;
            continue;
        }
        int64_t cnt = 1 + rng_poisson(&r, pmean);
        if (cnt > cap) cnt = cap;
        if (cnt > m) cnt = m;
        rows[0] = match[j];
        int64_t n = 1;
        while (n < cnt) {
            int64_t i = rng_below(&r, m);
            int dup = 0;
            for (int64_t q = 0; q < n; q++) if (rows[q] == i) { dup = 1; break; }
            if (!dup) rows[n++] = i;
        }
        qsort(rows, (size_t)cnt, sizeof(int64_t), cmp_i64);
        for (int64_t q = 0; q < cnt; q++) {
            double v;
            if (rows[q] == match[j]) {
                v = 1.0 + rng_unif(&r);
                if (rng_next(&r) & 1) v = -v;
            } else {
                v = rng_value(&r);
            }
            if (rowidx) { rowidx[nnz] = rows[q]; value[nnz] = v; }
            nnz++;
        }
    }
    if (colptr) colptr[m] = nnz;
    free(match); free(cpos); free(isslack); free(rows);
    return nnz;
}

/* dense right-hand side, uniform in [-1,1] */
void blugen_rhs(uint64_t seed, int64_t m, double *rhs) {
    rng_t r;
    rng_seed(&r, seed);
    for (int64_t i = 0; i < m; i++) rhs[i] = 2.0 * rng_unif(&r) - 1.0;
}

/* sparse right-hand side: nz distinct indices, values uniform in [-1,1] */
void blugen_sparse_rhs(uint64_t seed, int64_t m, int64_t nz, int64_t *idx, double *val) {
    rng_t r;
    rng_seed(&r, seed);
    int64_t n = 0;
    /* rejection through a marker array keeps this O(nz) for nz << m */
    char *seen = calloc((size_t)m, 1);
    while (n < nz) {
        int64_t i = rng_below(&r, m);
        if (seen[i]) continue;
        seen[i] = 1;
        idx[n] = i;
        val[n] = 2.0 * rng_unif(&r) - 1.0;
        n++;
    }
    free(seen);
}
