/*
 * TEST INFRASTRUCTURE — C restatement of the reference's choose + aggregate steps, used as the
 * bit-exact checker at sizes where the Python port (oracle/port.py) is too slow, and as the
 * optimistic ("c_port") CPU timing in bench.py. Never linked or loaded by the product.
 *
 * Follows /root/reference/src/layers.py:
 *   :260-262  num_sample = ceil(d * threshold)
 *   :657-658  distance = |center_score - neighbour_score| in fp32, ascending sort
 *   :662-672  keep the first num_sample if d > num_sample + 1, else keep all
 *   :675-691  positive target: append the int(num_sample * rho) nearest train positives
 *   :694      the result is the SET of those ids (union, duplicates merged)
 *   :593-624  mean over the set  (graphsage.py:224-231: divide by sqrt(n) instead)
 * with the canonical order (distance fp32, then position) — position == id for id-sorted rows.
 *
 * Work item w = r * B + i  (relation-major). sel_idx rows come out id-ascending.
 * Build: gcc -O2 -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}
static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}
static inline uint32_t dist_bits(float a, float b) {
    volatile float d = a - b; /* fp32 subtraction, no contraction */
    float f = fabsf(d);
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

/* per-row sizes: kept count k, oversample count o (before the union) */
static inline void row_counts(int64_t d, double thr, double rho, int positive, int P, int64_t* k, int64_t* o) {
    int64_t c = (int64_t)ceil((double)d * thr);
    *k = (d > c + 1) ? c : d;
    int64_t oo = positive ? (int64_t)((double)c * rho) : 0;
    if (oo > P) oo = P;
    if (oo < 0) oo = 0;
    *o = oo;
}

/* Upper bound of selected entries for the batch: sum over items of k + o. */
int64_t pcgo_bound(const int64_t* indptr, int64_t n_nodes, int n_rel, const int32_t* targets, const uint8_t* is_pos,
                   int B, const double* thresh, double rho, int P, int train) {
    int64_t tot = 0;
    for (int r = 0; r < n_rel; ++r)
        for (int i = 0; i < B; ++i) {
            int64_t row = (int64_t)r * n_nodes + targets[i];
            int64_t k, o;
            row_counts(indptr[row + 1] - indptr[row], thresh[r], rho, train && is_pos[i], P, &k, &o);
            tot += k + o;
        }
    return tot;
}

/*
 * entry_score: optional per-CSR-entry score (explicit neighbour scores, the IntraAgg.forward
 * calling convention); when NULL the neighbour score is score[indices[e]].
 * k_override: optional per-item kept count (sample_list); when NULL k comes from thresh.
 * Returns 0, or -1 if cap is too small.
 */
int pcgo_choose(const int64_t* indptr, const int32_t* indices, const float* score, const float* entry_score,
                const float* center_score, int64_t n_nodes, int n_rel, const int32_t* targets, const uint8_t* is_pos,
                int B, const double* thresh, const int32_t* k_override, double rho, const int32_t* pool,
                const float* pool_score, int P, int train, int64_t* sel_ptr, int32_t* sel_idx, int64_t cap) {
    int64_t items = (int64_t)n_rel * B;
    int64_t* bound = (int64_t*)malloc(sizeof(int64_t) * (items + 1));
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * items);
    bound[0] = 0;
    for (int64_t w = 0; w < items; ++w) {
        int r = (int)(w / B), i = (int)(w % B);
        int64_t row = (int64_t)r * n_nodes + targets[i];
        int64_t d = indptr[row + 1] - indptr[row], k, o;
        row_counts(d, thresh[r], rho, train && is_pos[i], P, &k, &o);
        if (k_override) {
            int64_t c = k_override[w];
            k = (d > c + 1) ? c : d;
            o = (train && is_pos[i]) ? (int64_t)((double)c * rho) : 0;
            if (o > P) o = P;
        }
        bound[w + 1] = bound[w] + k + o;
    }
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (bound[items] > 0 ? bound[items] : 1));
    int fail = 0;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t w = 0; w < items; ++w) {
        int r = (int)(w / B), i = (int)(w % B);
        int32_t v = targets[i];
        int64_t row = (int64_t)r * n_nodes + v;
        int64_t beg = indptr[row], d = indptr[row + 1] - beg;
        float sv = center_score ? center_score[i] : score[v];
        int64_t c = k_override ? k_override[w] : (int64_t)ceil((double)d * thresh[r]);
        int64_t k = (d > c + 1) ? c : d;
        int32_t* out = tmp + bound[w];
        int64_t n = 0;
        if (k < d) {
            uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * d);
            for (int64_t j = 0; j < d; ++j) {
                float su = entry_score ? entry_score[beg + j] : score[indices[beg + j]];
                keys[j] = ((uint64_t)dist_bits(sv, su) << 32) | (uint64_t)j;
            }
            qsort(keys, d, sizeof(uint64_t), cmp_u64);
            for (int64_t j = 0; j < k; ++j) out[n++] = indices[beg + (keys[j] & 0xffffffffu)];
            free(keys);
        } else {
            for (int64_t j = 0; j < d; ++j) out[n++] = indices[beg + j];
        }
        if (train && is_pos[i]) {
            int64_t o = (int64_t)((double)c * rho);
            if (o > P) o = P;
            if (o > 0) {
                uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * P);
                for (int64_t p = 0; p < P; ++p) {
                    float sp = pool_score ? pool_score[p] : score[pool[p]];
                    keys[p] = ((uint64_t)dist_bits(sv, sp) << 32) | (uint64_t)p;
                }
                qsort(keys, P, sizeof(uint64_t), cmp_u64);
                for (int64_t p = 0; p < o; ++p) out[n++] = pool[keys[p] & 0xffffffffu];
                free(keys);
            }
        }
        /* set semantics: sort ids, drop duplicates */
        qsort(out, n, sizeof(int32_t), cmp_i32);
        int64_t m = 0;
        for (int64_t j = 0; j < n; ++j)
            if (m == 0 || out[j] != out[m - 1]) out[m++] = out[j];
        cnt[w] = m;
    }
    sel_ptr[0] = 0;
    for (int64_t w = 0; w < items; ++w) sel_ptr[w + 1] = sel_ptr[w] + cnt[w];
    if (sel_ptr[items] > cap) fail = 1;
    if (!fail) {
#pragma omp parallel for schedule(static)
        for (int64_t w = 0; w < items; ++w) memcpy(sel_idx + sel_ptr[w], tmp + bound[w], sizeof(int32_t) * cnt[w]);
    }
    free(tmp);
    free(bound);
    free(cnt);
    return fail ? -1 : 0;
}

/*
 * Segment aggregation over id lists: out[w, :] = sum_{j in seg w} feat[j, :] * scale(n_w)
 * norm 0: 1/n (mean, layers.py:612-624), norm 1: 1/sqrt(n) (graphsage.py:224-231).
 * Accumulates in double and rounds once (order-independent reference value).
 */
int pcgo_aggregate(const float* feat, int F, int64_t ldf, const int64_t* sel_ptr, const int32_t* sel_idx, int64_t rows,
                   int norm, float* out) {
#pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * F);
#pragma omp for schedule(dynamic, 8)
        for (int64_t w = 0; w < rows; ++w) {
            int64_t b = sel_ptr[w], n = sel_ptr[w + 1] - b;
            for (int f = 0; f < F; ++f) acc[f] = 0.0;
            for (int64_t j = 0; j < n; ++j) {
                const float* x = feat + (int64_t)sel_idx[b + j] * ldf;
                for (int f = 0; f < F; ++f) acc[f] += x[f];
            }
            double s = n ? (norm == 1 ? 1.0 / sqrt((double)n) : 1.0 / (double)n) : 0.0;
            for (int f = 0; f < F; ++f) out[w * F + f] = (float)(acc[f] * s);
        }
        free(acc);
    }
    return 0;
}

/* Select-all variant for the GCN / SAGE baselines: row ∪ {self} (graphsage.py:210) or the row as is. */
int pcgo_select_all(const int64_t* indptr, const int32_t* indices, int64_t n_nodes, const int32_t* targets, int B,
                    int add_self, int64_t* sel_ptr, int32_t* sel_idx) {
    sel_ptr[0] = 0;
    for (int i = 0; i < B; ++i) {
        int32_t v = targets[i];
        int64_t beg = indptr[v], d = indptr[v + 1] - beg, n = sel_ptr[i];
        int seen = 0;
        for (int64_t j = 0; j < d; ++j) {
            int32_t u = indices[beg + j];
            if (add_self && !seen && u >= v) {
                if (u != v) sel_idx[n++] = v;
                seen = 1;
            }
            sel_idx[n++] = u;
        }
        if (add_self && !seen) sel_idx[n++] = v;
        sel_ptr[i + 1] = n;
    }
    return 0;
}

/* pick_step replay (utils.py:274-278 + CPython random.choices): index = bisect_right(cum, u*total, 0, n-1). */
int pcgo_pick_replay(const double* cum, int64_t n, const double* u, int64_t k, int64_t* out) {
    double total = cum[n - 1] + 0.0;
    for (int64_t t = 0; t < k; ++t) {
        double x = u[t] * total;
        int64_t lo = 0, hi = n - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi) / 2;
            if (x < cum[mid]) hi = mid; else lo = mid + 1;
        }
        out[t] = lo;
    }
    return 0;
}
