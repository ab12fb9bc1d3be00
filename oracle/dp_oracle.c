/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the WarpSTR DP fill and
 * traceback (reference: /root/reference/src/caller/caller.py:198-245 and :247-301).
 * Used by tests/ as the bulk checker of the CUDA path and by bench.py as the timed
 * "port" CPU baseline.  Never linked into the product library.
 *
 * Pinning: tests/test_oracle_pinned.py compares this against the unmodified
 * reference (oracle/refshim.py) and against the golden vectors in tests/golden/.
 *
 * Everything is float64 with the reference's operation order: the skip candidate is
 * ((D[i-back][p] + |x[i-back+1]-v_p|) + ... + |x[i-1]-v_p|) + |x[i]-v_j|, candidates
 * are tried 'stay' first, then incoming states in list order, strict '<'.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* caller.py:198-245.  D is T*S row-major, fully overwritten. */
void wso_fill(const double *x, int T, const double *v, const int32_t *seq_idx,
              const int32_t *in_ptr, const int32_t *in_idx, int S,
              const uint8_t *mask, int mv, int flank_length, double *D)
{
    const double inf = INFINITY;
    for (long n = 0; n < (long)T * S; ++n) D[n] = inf;
    double first = fabs(x[0] - v[0]);
    D[0] = first;
    for (int c = 1; c <= mv; ++c) D[c] = first + fabs(x[c] - v[0]);   /* row 0, column c */

    int boundary = flank_length - 10;
    int after = seq_idx[S - 1] - boundary;
    int th1 = 6 * boundary, th2 = T - 6 * boundary;

    for (int i = mv; i < T; ++i) {
        int back = (mask && mask[i]) ? mv - 1 : mv;
        double xi = x[i];
        double *row = D + (long)i * S;
        const double *up = row - S;
        const double *far = D + (long)(i - back) * S;
        int banded = !(i < th1) && (i > th2);
        for (int j = 0; j < S; ++j) {
            if (banded && seq_idx[j] < after) continue;
            double emit = fabs(xi - v[j]);
            double best = inf;
            if (up[j] != inf) {
                double c = up[j] + emit;
                if (c < best) best = c;
            }
            for (int e = in_ptr[j]; e < in_ptr[j + 1]; ++e) {
                int p = in_idx[e];
                double c = far[p];
                if (c == inf) continue;
                double vp = v[p];
                for (int t = i - back + 1; t < i; ++t) c += fabs(x[t] - vp);
                c += emit;
                if (c < best) best = c;
            }
            row[j] = best;
        }
    }
}

/* caller.py:247-301.  Returns 0, or 1 for the reference's RuntimeError. trace has T entries. */
int wso_backtrack(const double *D, const double *x, int T, const double *v,
                  const int32_t *in_ptr, const int32_t *in_idx, int S, int endstate,
                  const uint8_t *mask, int mv, int32_t *trace)
{
    const double inf = INFINITY;
    int j = endstate, i = T - 1, chosen = -1;
    long n = T;           /* fill from the back; the path consumes exactly T samples */
    while (i != 0) {
        double here = D[(long)i * S + j];
        double d_stay = inf;
        if (D[(long)(i - 1) * S + j] != inf)
            d_stay = fabs((D[(long)(i - 1) * S + j] + fabs(x[i] - v[j])) - here);
        int back = (mask && mask[i]) ? mv - 1 : mv;
        double d_skip = inf;
        for (int e = in_ptr[j]; e < in_ptr[j + 1]; ++e) {
            int p = in_idx[e];
            if (i - back < 0) continue;            /* cannot happen on a finite path */
            double c = D[(long)(i - back) * S + p];
            if (c == inf) continue;
            for (int t = i - back + 1; t < i; ++t) c += fabs(x[t] - v[p]);
            c += fabs(x[i] - v[j]);
            double d = fabs(c - here);
            if (d < d_skip) { d_skip = d; chosen = p; }
        }
        if (n <= 0) return 2;
        trace[--n] = j;
        if (d_skip < d_stay) {
            if (chosen == -1) return 1;
            for (int r = 0; r < back - 1; ++r) { if (n <= 0) return 2; trace[--n] = chosen; }
            i -= back;
            j = chosen;
        } else {
            i -= 1;
        }
    }
    if (n != 1) return 2;
    trace[0] = j;
    return 0;
}

int wso_warp(const double *x, int T, const double *v, const int32_t *seq_idx,
             const int32_t *in_ptr, const int32_t *in_idx, int S, int endstate,
             const uint8_t *mask, int mv, int flank_length, int32_t *trace, double *end_cost)
{
    double *D = (double *)malloc(sizeof(double) * (size_t)T * S);
    if (!D) return 3;
    wso_fill(x, T, v, seq_idx, in_ptr, in_idx, S, mask, mv, flank_length, D);
    if (end_cost) *end_cost = D[(long)(T - 1) * S + endstate];
    int rc = wso_backtrack(D, x, T, v, in_ptr, in_idx, S, endstate, mask, mv, trace);
    free(D);
    return rc;
}

/* Reads first..last-1 against one automaton, serially; off[r]..off[r+1] delimit read r in
 * x / mask / trace.  The Python side runs several of these ranges on threads (ctypes drops
 * the GIL), so no OpenMP runtime is needed. */
int wso_warp_range(const double *x, const int64_t *off, int first, int last, const double *v,
                   const int32_t *seq_idx, const int32_t *in_ptr, const int32_t *in_idx, int S,
                   int endstate, const uint8_t *mask, int mv, int flank_length,
                   int32_t *trace, int32_t *status)
{
    for (int r = first; r < last; ++r) {
        int T = (int)(off[r + 1] - off[r]);
        status[r] = wso_warp(x + off[r], T, v, seq_idx, in_ptr, in_idx, S, endstate,
                             mask ? mask + off[r] : NULL, mv, flank_length, trace + off[r], NULL);
    }
    return 0;
}
