/*
 * TEST INFRASTRUCTURE ONLY (oracle) -- never linked into or called from the product path.
 *
 * CPU restatement of the linear-assignment solver the reference reaches through
 *     lap.lapjv(cost_matrix, extend_cost=True, cost_limit=thresh)
 * at /root/reference/demo_bottrack_onnx_tflite.py:1686  (package lap==0.4.0, pinned at
 * /root/reference/Dockerfile:26 and demo:7).  The lap sources are NOT under /root/reference
 * (third-party dependency, not vendored, not installable here: no network), so this file
 * restates the PUBLISHED algorithm lap implements:
 *
 *   R. Jonker, A. Volgenant, "A shortest augmenting path algorithm for dense and sparse
 *   linear assignment problems", Computing 38 (1987) 325-340  -- the dense LAPJV variant
 *   with its four phases: column reduction, reduction transfer, augmenting row reduction
 *   (run twice), and shortest-path augmentation.
 *
 * plus lap 0.4.0's documented treatment of `cost_limit` / `extend_cost`: the N x M problem
 * is embedded into an (N+M) x (N+M) square matrix filled with cost_limit/2, whose lower
 * right M x N block is 0 and whose upper left block is the cost; after solving, partners
 * with index >= M (rows) / >= N (columns) are reported as -1.
 *
 * Parity status: "parity unpinned by the reference" for tie-breaking (no lap sources, no
 * golden vectors in the reference).  The optimum is pinned by tests/test_oracle_lap.py:
 * brute force on small cases and scipy.optimize.linear_sum_assignment on the extended
 * matrix on larger ones.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define JV_BIG 1.0e300

/* ---- phase 1+2: column reduction and reduction transfer ------------------------------- */
static int jv_column_reduction(int n, const double *c, int *free_rows, int *x, int *y, double *v)
{
    unsigned char *single = (unsigned char *)malloc((size_t)n);
    for (int i = 0; i < n; ++i) { x[i] = -1; y[i] = 0; v[i] = JV_BIG; single[i] = 1; }
    /* v[j] = column minimum, y[j] = first row attaining it */
    for (int i = 0; i < n; ++i) {
        const double *ci = c + (size_t)i * n;
        for (int j = 0; j < n; ++j) {
            if (ci[j] < v[j]) { v[j] = ci[j]; y[j] = i; }
        }
    }
    /* scan columns from the last to the first; a row keeps the highest-index column only
       if no other column also chose it */
    for (int j = n - 1; j >= 0; --j) {
        int i = y[j];
        if (x[i] < 0) x[i] = j;
        else { single[i] = 0; y[j] = -1; }
    }
    int nfree = 0;
    for (int i = 0; i < n; ++i) {
        if (x[i] < 0) {
            free_rows[nfree++] = i;
        } else if (single[i]) {
            /* reduction transfer: lower v of the assigned column by the row's second-best slack */
            int j1 = x[i];
            const double *ci = c + (size_t)i * n;
            double mn = JV_BIG;
            for (int j = 0; j < n; ++j) {
                if (j == j1) continue;
                double s = ci[j] - v[j];
                if (s < mn) mn = s;
            }
            v[j1] -= mn;
        }
    }
    free(single);
    return nfree;
}

/* ---- phase 3: augmenting row reduction ------------------------------------------------ */
static int jv_augmenting_row_reduction(int n, const double *c, int nfree, int *free_rows,
                                       int *x, int *y, double *v)
{
    int cur = 0, next_free = 0;
    long long budget_cnt = 0;
    while (cur < nfree) {
        ++budget_cnt;
        int fi = free_rows[cur++];
        const double *ci = c + (size_t)fi * n;
        int j1 = 0, j2 = -1;
        double s1 = ci[0] - v[0], s2 = JV_BIG;
        for (int j = 1; j < n; ++j) {
            double s = ci[j] - v[j];
            if (s < s2) {
                if (s >= s1) { s2 = s; j2 = j; }
                else { s2 = s1; s1 = s; j2 = j1; j1 = j; }
            }
        }
        int i0 = y[j1];
        double v1_new = v[j1] - (s2 - s1);
        int lowers = v1_new < v[j1];
        if (budget_cnt < (long long)cur * n) {
            if (lowers) v[j1] = v1_new;
            else if (i0 >= 0 && j2 >= 0) { j1 = j2; i0 = y[j2]; }
            if (i0 >= 0) {
                if (lowers) free_rows[--cur] = i0;       /* re-process the displaced row at once */
                else free_rows[next_free++] = i0;
            }
        } else {
            if (i0 >= 0) free_rows[next_free++] = i0;
        }
        x[fi] = j1;
        y[j1] = fi;
    }
    return next_free;
}

/* ---- phase 4: shortest augmenting paths (Dijkstra on reduced costs) -------------------- */
static int jv_collect_minima(int n, int lo, const double *d, int *cols)
{
    int hi = lo + 1;
    double mind = d[cols[lo]];
    for (int k = hi; k < n; ++k) {
        int j = cols[k];
        if (d[j] <= mind) {
            if (d[j] < mind) { hi = lo; mind = d[j]; }
            cols[k] = cols[hi];
            cols[hi++] = j;
        }
    }
    return hi;
}

static int jv_scan(int n, const double *c, int *plo, int *phi, double *d, int *cols, int *pred,
                   const int *y, const double *v)
{
    int lo = *plo, hi = *phi;
    while (lo != hi) {
        int j = cols[lo++];
        int i = y[j];
        double mind = d[j];
        const double *ci = c + (size_t)i * n;
        double h = ci[j] - v[j] - mind;
        for (int k = hi; k < n; ++k) {
            int jj = cols[k];
            double cred = ci[jj] - v[jj] - h;
            if (cred < d[jj]) {
                d[jj] = cred;
                pred[jj] = i;
                if (cred == mind) {
                    if (y[jj] < 0) return jj;
                    cols[k] = cols[hi];
                    cols[hi++] = jj;
                }
            }
        }
    }
    *plo = lo; *phi = hi;
    return -1;
}

static int jv_find_path(int n, const double *c, int start, const int *y, double *v, int *pred,
                        int *cols, double *d)
{
    int lo = 0, hi = 0, final_j = -1, n_ready = 0;
    const double *cs = c + (size_t)start * n;
    for (int j = 0; j < n; ++j) { cols[j] = j; pred[j] = start; d[j] = cs[j] - v[j]; }
    while (final_j == -1) {
        if (lo == hi) {
            n_ready = lo;
            hi = jv_collect_minima(n, lo, d, cols);
            for (int k = lo; k < hi; ++k) {
                int j = cols[k];
                if (y[j] < 0) final_j = j;
            }
        }
        if (final_j == -1) final_j = jv_scan(n, c, &lo, &hi, d, cols, pred, y, v);
    }
    double mind = d[cols[lo]];
    for (int k = 0; k < n_ready; ++k) {
        int j = cols[k];
        v[j] += d[j] - mind;
    }
    return final_j;
}

static void jv_augment(int n, const double *c, int nfree, const int *free_rows, int *x, int *y, double *v)
{
    int *pred = (int *)malloc(sizeof(int) * (size_t)n);
    int *cols = (int *)malloc(sizeof(int) * (size_t)n);
    double *d = (double *)malloc(sizeof(double) * (size_t)n);
    for (int f = 0; f < nfree; ++f) {
        int start = free_rows[f];
        int j = jv_find_path(n, c, start, y, v, pred, cols, d);
        int i = -1;
        while (i != start) {
            i = pred[j];
            y[j] = i;
            int t = x[i]; x[i] = j; j = t;
        }
    }
    free(pred); free(cols); free(d);
}

/* Solve the square n x n problem. x[i] = column of row i, y[j] = row of column j. */
int oracle_lapjv_square(int n, const double *c, int *x, int *y)
{
    if (n <= 0) return 0;
    int *free_rows = (int *)malloc(sizeof(int) * (size_t)n);
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    int nfree = jv_column_reduction(n, c, free_rows, x, y, v);
    for (int pass = 0; nfree > 0 && pass < 2; ++pass)
        nfree = jv_augmenting_row_reduction(n, c, nfree, free_rows, x, y, v);
    if (nfree > 0) jv_augment(n, c, nfree, free_rows, x, y, v);
    free(free_rows); free(v);
    return 0;
}

/*
 * lap.lapjv(cost, extend_cost=True, cost_limit=limit) restated: cost is n_rows x n_cols row major.
 * x_out[n_rows], y_out[n_cols]: partner index or -1.  Returns 0, or -1 on allocation failure.
 */
int oracle_lapjv_extended(int n_rows, int n_cols, const double *cost, double limit, int *x_out, int *y_out)
{
    int n = n_rows + n_cols;
    if (n == 0) return 0;
    double *ext = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    int *x = (int *)malloc(sizeof(int) * (size_t)n);
    int *y = (int *)malloc(sizeof(int) * (size_t)n);
    if (!ext || !x || !y) { free(ext); free(x); free(y); return -1; }
    double half = limit / 2.0;
    for (int i = 0; i < n; ++i) {
        double *row = ext + (size_t)i * n;
        if (i < n_rows) {
            memcpy(row, cost + (size_t)i * n_cols, sizeof(double) * (size_t)n_cols);
            for (int j = n_cols; j < n; ++j) row[j] = half;
        } else {
            for (int j = 0; j < n_cols; ++j) row[j] = half;
            for (int j = n_cols; j < n; ++j) row[j] = 0.0;
        }
    }
    oracle_lapjv_square(n, ext, x, y);
    for (int i = 0; i < n_rows; ++i) x_out[i] = (x[i] >= n_cols) ? -1 : x[i];
    for (int j = 0; j < n_cols; ++j) y_out[j] = (y[j] >= n_rows) ? -1 : y[j];
    free(ext); free(x); free(y);
    return 0;
}
