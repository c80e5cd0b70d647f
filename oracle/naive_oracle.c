/* Naive CPU restatement of exact top-k search.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference (dorenwick/CloudVectorDB) ships README.md:1-2
 * and no code, tests or golden vectors.  This double loop restates the FAISS
 * IndexFlat convention BASELINE.json's north_star names (k best per query,
 * ties -> lower id, -1 / +-inf padding) independently of oracle/flat_oracle.py
 * so the two can be checked against each other.  Scores accumulate in double.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* metric: 0 = inner product (larger better), 1 = squared L2 (smaller better).
 * self_ids (nq) / group_db (n) / group_q (nq) may be NULL. */
int naive_search(const float* xb, int64_t n, const float* xq, int64_t nq, int d,
                 int k, int metric, const int64_t* self_ids,
                 const int32_t* group_db, const int32_t* group_q,
                 float* D, int64_t* I)
{
    double* bv = (double*)malloc(sizeof(double) * (size_t)k);
    int64_t* bi = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    if (!bv || !bi) return -1;
    for (int64_t q = 0; q < nq; ++q) {
        int cnt = 0;
        const float* qv = xq + q * d;
        for (int64_t j = 0; j < n; ++j) {
            if (self_ids && self_ids[q] == j) continue;
            if (group_db && group_q[q] >= 0 && group_db[j] == group_q[q]) continue;
            const float* xv = xb + j * d;
            double s = 0.0;
            if (metric == 0) {
                for (int t = 0; t < d; ++t) s += (double)qv[t] * (double)xv[t];
            } else {
                for (int t = 0; t < d; ++t) {
                    double df = (double)qv[t] - (double)xv[t];
                    s -= df * df; /* negated: larger is better */
                }
            }
            /* insertion into a sorted list; strict > keeps the lower id on ties */
            if (cnt < k || s > bv[cnt - 1]) {
                int p = cnt < k ? cnt : k - 1;
                while (p > 0 && s > bv[p - 1]) { bv[p] = bv[p - 1]; bi[p] = bi[p - 1]; --p; }
                bv[p] = s; bi[p] = j;
                if (cnt < k) ++cnt;
            }
        }
        for (int r = 0; r < k; ++r) {
            if (r < cnt) {
                D[q * k + r] = (float)(metric == 0 ? bv[r] : -bv[r]);
                I[q * k + r] = bi[r];
            } else {
                D[q * k + r] = metric == 0 ? -INFINITY : INFINITY;
                I[q * k + r] = -1;
            }
        }
    }
    free(bv); free(bi);
    return 0;
}

/* One Lloyd assignment: argmin_c ||x - c||^2, lower id on ties. */
int naive_kmeans_assign(const float* x, int64_t n, const float* c, int K, int d,
                        int32_t* assign, float* dist)
{
    for (int64_t i = 0; i < n; ++i) {
        double best = INFINITY; int bi = -1;
        for (int j = 0; j < K; ++j) {
            double s = 0.0;
            for (int t = 0; t < d; ++t) {
                double df = (double)x[i * d + t] - (double)c[(int64_t)j * d + t];
                s += df * df;
            }
            if (s < best) { best = s; bi = j; }
        }
        assign[i] = bi; dist[i] = (float)best;
    }
    return 0;
}
