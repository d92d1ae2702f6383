/* CPU oracle for the log-uniform candidate sampler (TEST INFRASTRUCTURE ONLY).
 *
 * Restates /root/reference/U2GNN_pytorch/log_uniform/Log_Uniform_Sampler.cpp:
 *   ctor            :10-16   prob[i] = (log(i+2)-log(i+1)) / log(N+1), engine seeded 1111
 *   probability     :18-21
 *   expected_count  :23-32   -expm1(tries * log1p(-prob[id]))  (float result)
 *   sample          :57-71   draw x~U[0,1), v = lround(exp(x*log N)) - 1, until `size` distinct
 *   sample_unique   :73-88   same, skipping ids in `labels`
 * and the libstdc++ pieces the reference relies on (third-party, GCC 13 <random>):
 *   std::default_random_engine = minstd_rand0: x <- 16807*x mod (2^31-1), min 1, max 2^31-2
 *   std::generate_canonical<double,53>: k = 2 draws, sum = (g1-1) + (g2-1)*R, R = 2147483646,
 *   result sum/(R*R) (nextafter(1,0) if it rounds to 1); uniform_real_distribution(0,1).
 * Pinned against the reference class compiled from its own sources (oracle/_ref, see Makefile)
 * and the try counts in SURVEY.md §4 (929 / 1093 / 117) by tests/test_sampler_oracle.py.
 * The reference returns the ids in libstdc++ bucket order; only the SET is semantically used
 * (the loss sums over it), so this oracle returns first-insertion order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;
    uint32_t x; /* minstd_rand0 state */
} logu_oracle;

static uint32_t lcg_next(logu_oracle *s) {
    s->x = (uint32_t)(((uint64_t)s->x * 16807ull) % 2147483647ull);
    return s->x;
}

static double canonical(logu_oracle *s) {
    const double R = 2147483646.0;
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k) {
        sum += (double)(lcg_next(s) - 1u) * tmp;
        tmp *= R;
    }
    double r = sum / tmp;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

logu_oracle *logu_oracle_new(int n) {
    logu_oracle *s = (logu_oracle *)malloc(sizeof *s);
    s->n = n;
    s->x = 1111u % 2147483647u;
    return s;
}
void logu_oracle_free(logu_oracle *s) { free(s); }
void logu_oracle_reseed(logu_oracle *s) { s->x = 1111u; }
uint32_t logu_oracle_state(const logu_oracle *s) { return s->x; }

float logu_oracle_probability(const logu_oracle *s, int idx) {
    return (float)((log((double)idx + 2) - log((double)idx + 1)) / log((double)s->n + 1));
}

void logu_oracle_expected_count(const logu_oracle *s, int tries, const int64_t *ids, int64_t n, float *out) {
    for (int64_t i = 0; i < n; ++i) {
        float p = logu_oracle_probability(s, (int)ids[i]);
        out[i] = (float)(-expm1(tries * log1p(-p)));
    }
}

/* open-addressing set of int64 keyed by value; cap is a power of two >= 4*size */
static int set_insert(int64_t *tab, int64_t cap, int64_t v) {
    uint64_t h = ((uint64_t)v * 0x9E3779B97F4A7C15ull) & (uint64_t)(cap - 1);
    while (tab[h] != -1) {
        if (tab[h] == v) return 0;
        h = (h + 1) & (uint64_t)(cap - 1);
    }
    tab[h] = v;
    return 1;
}
static int set_has(const int64_t *tab, int64_t cap, int64_t v) {
    uint64_t h = ((uint64_t)v * 0x9E3779B97F4A7C15ull) & (uint64_t)(cap - 1);
    while (tab[h] != -1) {
        if (tab[h] == v) return 1;
        h = (h + 1) & (uint64_t)(cap - 1);
    }
    return 0;
}

/* returns 0, or -1 if size > N (the reference would loop forever) */
int logu_oracle_sample(logu_oracle *s, int64_t size, int64_t *out_ids, int *num_tries) {
    if (size > s->n) return -1;
    int64_t cap = 16;
    while (cap < 4 * size) cap <<= 1;
    int64_t *tab = (int64_t *)malloc(sizeof(int64_t) * cap);
    memset(tab, 0xFF, sizeof(int64_t) * cap);
    const double log_n = log((double)s->n);
    int64_t got = 0;
    *num_tries = 0;
    while (got != size) {
        *num_tries += 1;
        double x = canonical(s);
        int64_t v = lround(exp(x * log_n)) - 1;
        if (set_insert(tab, cap, v)) out_ids[got++] = v;
    }
    free(tab);
    return 0;
}

int logu_oracle_sample_unique(logu_oracle *s, int64_t size, const int64_t *labels, int64_t n_labels,
                              int64_t *out_ids) {
    int64_t lcap = 16, cap = 16;
    while (lcap < 4 * n_labels) lcap <<= 1;
    while (cap < 4 * size) cap <<= 1;
    int64_t *ltab = (int64_t *)malloc(sizeof(int64_t) * lcap);
    int64_t *tab = (int64_t *)malloc(sizeof(int64_t) * cap);
    memset(ltab, 0xFF, sizeof(int64_t) * lcap);
    memset(tab, 0xFF, sizeof(int64_t) * cap);
    int64_t distinct = 0;
    for (int64_t i = 0; i < n_labels; ++i) distinct += set_insert(ltab, lcap, labels[i]);
    if (size > s->n - distinct) { free(ltab); free(tab); return -1; }
    const double log_n = log((double)s->n);
    int64_t got = 0;
    while (got != size) {
        double x = canonical(s);
        int64_t v = lround(exp(x * log_n)) - 1;
        if (!set_has(ltab, lcap, v) && set_insert(tab, cap, v)) out_ids[got++] = v;
    }
    free(ltab); free(tab);
    return 0;
}
