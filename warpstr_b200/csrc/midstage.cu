// What WarpSTR does with a pass's trace: run statistics, spline rescale, bad-repeat mask,
// state-wise cost and the decoded sequence.  One warp per read.
//
// Replaces, for the default rescaling configuration (reps_as_one False, method mean):
//   WarpResult.state_transitions / create_alignment   reference src/caller/caller.py:58-96
//   StateAlignment.good_enough / filter_alignment      :23-39, :316-318
//   rescale_signal (splrep s=m, splev)                 :304-313
//   mask_bad_repeats and helpers                       :330-421
//   the cost lines of WarpSTR.run                      :138-139
//   WarpSTR._get_sequence                              :178-187
//
// Bit-level fidelity.  Means and variances use numpy's pairwise summation order
// (numpy/_core/src/umath/loops_utils.h.src: n<8 sequential, <=128 eight accumulators, else
// halves).  splrep(x, y, s=m) is FITPACK curfit (scipy pins 1.6.3; same algorithm in 1.18):
// its first iteration is the least-squares cubic polynomial on the knots [xb]*4+[xe]*4 and it
// is accepted whenever its residual fp <= s, which holds for every read whose filtered pairs
// satisfy |y-x| <= threshold <= 1 (fp <= m*threshold^2).  That iteration (fpcurf's Givens
// sweep: fpbspl, fpgivs, fprota, fpback) and splev are restated here operation by operation;
// this file is compiled with -fmad=false, and the host test compares the coefficients with
// scipy bit for bit.  If fp > s the spline needs interior knots: the read is flagged
// WSTR_READ_SPLINE_KNOTS and the Python host evaluates it with scipy.
// One known last-bit difference: the reference squares a standard deviation with pow(x, 2)
// inside calc_ttest (caller.py:351); glibc's pow differs from x*x in the last bit for about
// 0.1 % of arguments, so a t statistic can differ by one ulp; that flips a decision only
// if the statistic is within one ulp of +-3 or of its predecessor.
#include <math.h>

#include "wstr_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ---- numpy pairwise summation over f(0..n-1) ---------------------------------------------------
// (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE).  The recursion of the original
// (halves, the left one rounded down to a multiple of 8) is unrolled onto an explicit stack so
// that the kernels have a static frame size.
template <class F>
__device__ __forceinline__ double pairwise_leaf(const F &f, int start, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += f(start + i);
        return res;
    }
    double r0 = f(start), r1 = f(start + 1), r2 = f(start + 2), r3 = f(start + 3), r4 = f(start + 4),
           r5 = f(start + 5), r6 = f(start + 6), r7 = f(start + 7);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 += f(start + i);
        r1 += f(start + i + 1);
        r2 += f(start + i + 2);
        r3 += f(start + i + 3);
        r4 += f(start + i + 4);
        r5 += f(start + i + 5);
        r6 += f(start + i + 6);
        r7 += f(start + i + 7);
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += f(start + i);
    return res;
}

// the same order for n <= 16 values already in registers (a run of one state: its dwell): n < 8 one after the
// other; else eight accumulators over the first eight, the next eight added if there are sixteen, the
// accumulators combined pairwise, the rest (n % 8 values) added one after the other
__device__ __forceinline__ double pairwise16(const double (&w)[16], int n) {
    if (n < 8) {
        double res = 0.0;
#pragma unroll
        for (int i = 0; i < 7; ++i)
            if (i < n) res += w[i];
        return res;
    }
    double r0 = w[0], r1 = w[1], r2 = w[2], r3 = w[3], r4 = w[4], r5 = w[5], r6 = w[6], r7 = w[7];
    if (n == 16) {
        r0 += w[8];
        r1 += w[9];
        r2 += w[10];
        r3 += w[11];
        r4 += w[12];
        r5 += w[13];
        r6 += w[14];
        r7 += w[15];
        return ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
#pragma unroll
    for (int i = 8; i < 15; ++i)
        if (i < n) res += w[i];
    return res;
}

template <class F>
__device__ double pairwise_sum(const F &f, int start0, int n0) {
    if (n0 <= 128) return pairwise_leaf(f, start0, n0);
    struct Frame {
        int start, n, stage;
        double left;
    };
    Frame st[28];
    int sp = 0;
    st[sp++] = Frame{start0, n0, 0, 0.0};
    double ret = 0.0;
    while (sp > 0) {
        Frame &fr = st[sp - 1];
        if (fr.n <= 128) {
            ret = pairwise_leaf(f, fr.start, fr.n);
            --sp;
            continue;
        }
        int n2 = fr.n / 2;
        n2 -= n2 % 8;
        if (fr.stage == 0) {
            fr.stage = 1;
            st[sp++] = Frame{fr.start, n2, 0, 0.0};
        } else if (fr.stage == 1) {
            fr.left = ret;
            fr.stage = 2;
            st[sp++] = Frame{fr.start + n2, fr.n - n2, 0, 0.0};
        } else {
            ret = fr.left + ret;
            --sp;
        }
    }
    return ret;
}

// (exact division by a divisor that is used many times: Divisor, make_divisor, div_fast in wstr_internal.h)
// GUARD: test the numerator (and the divisor's flag) per call; otherwise the caller has
// established that every numerator is 0 or within 2^+-600
template <bool GUARD>
__device__ __forceinline__ double div_exact(double a, const Divisor &dv) {
    if (!GUARD) return div_fast(a, dv);
    const unsigned e = (static_cast<unsigned>(__double2hiint(a)) >> 20) & 0x7ffu;
    if (dv.fast && e - (1023u - 400u) <= 800u) return div_fast(a, dv);
    return a / dv.d;
}
// 0, or 2^-120 <= |v| <= 2^120: sums, differences and triple products of such values (and of
// quotients by a `fast` divisor) stay far inside the range div_fast needs
__device__ __forceinline__ bool tame(double v) {
    const unsigned hi = static_cast<unsigned>(__double2hiint(v)) & 0x7fffffffu;
    return (hi >> 20) - (1023u - 120u) <= 240u || (hi | static_cast<unsigned>(__double2loint(v))) == 0u;
}

// ---- FITPACK pieces for the single-interval cubic ------------------------------------------------
// fpbspl with k=3 on knots [xb,xb,xb,xb,xe,xe,xe,xe], interval l=4; span = make_divisor(xe - xb)
template <bool GUARD>
__device__ __forceinline__ void bspl3(double xb, double xe, const Divisor &span, double x, double *h) {
    double hh[3];
    h[0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; ++i) {
            if (GUARD && xe == xb) {   // (an unguarded caller has a non-zero span)
                h[i] = 0.0;
                continue;
            }
            const double f = div_exact<GUARD>(hh[i - 1], span);
            h[i - 1] = h[i - 1] + f * (xe - x);
            h[i] = f * (x - xb);
        }
    }
}

__device__ __forceinline__ void givens(double piv, double &ww, double &c, double &s) {
    const double store = fabs(piv);
    double dd;
    if (store >= ww) {
        const double r = ww / piv;
        dd = store * sqrt(1.0 + r * r);
    } else {
        const double r = piv / ww;
        dd = ww * sqrt(1.0 + r * r);
    }
    c = ww / dd;
    s = piv / dd;
    ww = dd;
}

__device__ __forceinline__ void rotate(double c, double s, double &a, double &b) {
    const double s1 = a, s2 = b;
    b = c * s2 + s * s1;
    a = c * s1 - s * s2;
}

struct Cubic {
    double xb, xe, c[4], fp;
};

// least-squares cubic through (sx[i], sy[i]), i < m, sx ascending: fpcurf's first iteration
__device__ void lsq_cubic(const double *sx, const double *sy, int m, Cubic &out) {
    const double xb = sx[0], xe = sx[m - 1];
    const Divisor span = make_divisor(xe - xb);
    double a[4][4], z[4];
    for (int i = 0; i < 4; ++i) {
        z[i] = 0.0;
        for (int j = 0; j < 4; ++j) a[i][j] = 0.0;
    }
    double fp = 0.0;
    for (int it = 0; it < m; ++it) {
        const double xi = sx[it];
        double yi = sy[it];
        double h[4];
        bspl3<true>(xb, xe, span, xi, h);
        for (int i = 0; i < 4; ++i) {
            const double piv = h[i];
            if (piv == 0.0) continue;
            double c, s;
            givens(piv, a[i][0], c, s);
            rotate(c, s, yi, z[i]);
            if (i == 3) break;
            int i2 = 0;
            for (int i1 = i + 1; i1 < 4; ++i1) {
                ++i2;
                rotate(c, s, h[i1], a[i][i2]);
            }
        }
        fp = fp + yi * yi;
    }
    // fpback, n = 4, bandwidth 4
    double c[4];
    c[3] = z[3] / a[3][0];
    int i = 2;
    for (int j = 2; j <= 4; ++j) {
        double store = z[i];
        const int i1 = j <= 3 ? j - 1 : 3;
        int mm = i;
        for (int l = 1; l <= i1; ++l) {
            ++mm;
            store = store - c[mm] * a[i][l];
        }
        c[i] = store / a[i][0];
        --i;
    }
    out.xb = xb;
    out.xe = xe;
    out.fp = fp;
    for (int q = 0; q < 4; ++q) out.c[q] = c[q];
}

template <bool GUARD>
__device__ __forceinline__ double splev3(const Cubic &cu, const Divisor &span, double x) {
    double h[4];
    bspl3<GUARD>(cu.xb, cu.xe, span, x, h);
    double sp = 0.0;
    for (int j = 0; j < 4; ++j) sp = sp + cu.c[j] * h[j];
    return sp;
}

// ---- sliding two-sample statistic (caller.py:347-354), windows x[c-3:c] and x[c:c+3] -----------
template <bool GUARD>
__device__ __forceinline__ void mean_sd3(const double *w, const Divisor &three, double &mean, double &sd) {
    mean = div_exact<GUARD>((w[0] + w[1]) + w[2], three);
    const double d0 = w[0] - mean, d1 = w[1] - mean, d2 = w[2] - mean;
    const double var = div_exact<GUARD>((d0 * d0 + d1 * d1) + d2 * d2, three);
    sd = sqrt(var);
}

template <bool GUARD>
__device__ __forceinline__ double tstat_of(double ma, double sa, double mb, double sb, const Divisor &three) {
    double sd = sqrt(div_exact<GUARD>(sa * sa + sb * sb, three));
    if (sd == 0.0) sd = sd + 0.0000001;
    return (ma - mb) / sd;
}

// number of detected segment borders minus one in t(c0..c1) (caller.py:357-378).  t(c) compares
// x[c-3:c] with x[c:c+3]; the right window of position c is the left window of c+3, so its mean
// and deviation are kept for three positions instead of being computed twice.
// GUARD = false: every sample of the read is `tame`, the divisions by 3 need no range test.
//
// `ties` counts the decisions that sit within `tie_ulps` units in the last place of flipping: |t| against
// 3 and, beyond 3, t against its predecessor.  The reference squares the two deviations with pow(x, 2)
// (np.float64 ** 2, caller.py:351), and a libm pow that is not correctly rounded (glibc >= 2.28: < 1 ULP)
// returns the neighbour of x*x for ~0.1 % of arguments; that moves t by at most ~3 ulp.  A read without
// ties is therefore decided identically whatever the libm; a read with ties is reported (d_ttest_ties)
// and the Python layer re-evaluates it with the host's own pow.
template <bool GUARD>
__device__ int count_segments(const double *__restrict__ x, int c0, int c1, const long long tie_ulps, int &ties) {
    const Divisor three = make_divisor(3.0);
    int borders = 0;
    bool rising = false;
    double prev = 0.0;
    double m0 = 0.0, s0 = 0.0, m1 = 0.0, s1 = 0.0, m2 = 0.0, s2 = 0.0;   // right windows of c-3, c-2, c-1
    if (c0 > c1) return -1;
    // the right window slides by one sample per position: two of its three samples are the previous
    // position's, the third is fetched a position ahead of its use
    double w[3] = {x[c0], x[c0 + 1], x[c0 + 2]};
    double nxt = c0 < c1 ? x[c0 + 3] : 0.0;
    for (int c = c0; c <= c1; ++c) {
        double ma, sa, mb, sb;
        if (c - c0 >= 3) {
            ma = m0;
            sa = s0;
        } else {
            mean_sd3<GUARD>(x + c - 3, three, ma, sa);
        }
        mean_sd3<GUARD>(w, three, mb, sb);
        w[0] = w[1];
        w[1] = w[2];
        w[2] = nxt;
        if (c + 1 < c1) nxt = x[c + 4];
        m0 = m1; s0 = s1;
        m1 = m2; s1 = s2;
        m2 = mb; s2 = sb;
        const double t = tstat_of<GUARD>(ma, sa, mb, sb, three);
        if (c == c0) prev = t;
        {
            const long long at = __double_as_longlong(fabs(t)), three_bits = 0x4008000000000000LL;
            long long d3 = at - three_bits;
            d3 = d3 < 0 ? -d3 : d3;
            bool tie = d3 <= tie_ulps;
            if (!tie && c != c0 && at >= three_bits - tie_ulps && ((t > 0.0) == (prev > 0.0))) {
                long long dp = at - __double_as_longlong(fabs(prev));
                dp = dp < 0 ? -dp : dp;
                tie = dp <= tie_ulps;
            }
            ties += tie ? 1 : 0;
        }
        if (t > 3.0 || t < -3.0) {
            if ((t > 3.0 && t >= prev) || (t < -3.0 && t <= prev)) {
                rising = true;
            } else {
                if (rising) ++borders;
                rising = false;
            }
        } else if (rising) {
            ++borders;
            rising = false;
        }
        prev = t;
    }
    return borders - 1;
}

__device__ __forceinline__ int warp_min(int v) {
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

__device__ __forceinline__ uint8_t complement(uint8_t b) {
    switch (b) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        default: return b;
    }
}

// np.median of a run (caller.py:323-325, rescaling.method 'median'): the middle order statistic,
// or the mean of the two middle ones; runs are short (a state's dwell), so ranks are counted
__device__ double run_median(const double *xs, int n) {
    const int k_lo = (n - 1) / 2, k_hi = n / 2;
    double v_lo = 0.0, v_hi = 0.0;
    for (int i = 0; i < n; ++i) {
        const double xi = xs[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double xj = xs[j];
            rank += (xj < xi) || (xj == xi && j < i);
        }
        if (rank == k_lo) v_lo = xi;
        if (rank == k_hi) v_hi = xi;
    }
    return (n & 1) ? v_hi : (v_lo + v_hi) / 2.0;
}

// k-th smallest (0-based) of xs[0..n) without a copy: most-significant-bit-first selection on the
// order-preserving integer image of the doubles, 64 counting passes (for the long lists of
// reps_as_one, where a state's samples come from every visit of it)
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ double select_kth(const double *xs, int n, int k) {
    unsigned long long prefix = 0ull;
    for (int bit = 63; bit >= 0; --bit) {
        const unsigned long long above = bit == 63 ? 0ull : ~((2ull << bit) - 1ull);   // bits already decided
        int zeros = 0;
        for (int i = 0; i < n; ++i) {
            const unsigned long long key = order_key(xs[i]);
            zeros += ((key & above) == prefix && !((key >> bit) & 1ull)) ? 1 : 0;
        }
        if (k >= zeros) {
            k -= zeros;
            prefix |= 1ull << bit;
        }
    }
    const unsigned long long b = (prefix >> 63) ? (prefix & 0x7fffffffffffffffull) : ~prefix;
    return __longlong_as_double(static_cast<long long>(b));
}
__device__ double list_median(const double *xs, int n) {
    if (n <= 48) return run_median(xs, n);
    if (n & 1) return select_kth(xs, n, n / 2);
    return (select_kth(xs, n, n / 2 - 1) + select_kth(xs, n, n / 2)) / 2.0;
}

struct Scratch {
    int32_t *run_start, *run_state;
    double *sv, *px, *py, *sx, *sy;
    int32_t *sidx;       // sort scratch for long reads; reps_as_one: where each run's samples go in gvals
    int32_t *astate;     // reps_as_one: state of each alignment entry
    int32_t *cnt, *base; // reps_as_one: samples per state, first slot of each state in gvals ([S], [S+1])
    double *gvals;       // reps_as_one: the read's samples grouped by state, time order inside a state ([T])
};

__host__ __device__ inline size_t scratch_core_bytes(int R) {
    size_t b = ((8 * (size_t)R + 4 + 7) / 8) * 8;   // run_start (R+1) + run_state (R), int32
    b += 5 * 8 * (size_t)R;                         // sv, px, py, sx, sy
    b += 4 * (size_t)R;                             // sidx
    b += 4 * (size_t)R;                             // astate
    return (b + 7) / 8 * 8;
}

__device__ __forceinline__ Scratch carve(unsigned char *ws, int R, int T, int s_max) {
    Scratch s;
    s.run_start = reinterpret_cast<int32_t *>(ws);   // R+1
    s.run_state = s.run_start + (R + 1);             // R
    s.sv = reinterpret_cast<double *>(ws + ((8 * (size_t)R + 4 + 7) / 8) * 8);
    s.px = s.sv + R;
    s.py = s.px + R;
    s.sx = s.py + R;
    s.sy = s.sx + R;
    s.sidx = reinterpret_cast<int32_t *>(s.sy + R);
    s.astate = s.sidx + R;
    unsigned char *ext = ws + scratch_core_bytes(R);          // only there when the call has reps_as_one
    s.gvals = reinterpret_cast<double *>(ext);
    s.cnt = reinterpret_cast<int32_t *>(s.gvals + T);
    s.base = s.cnt + s_max;
    return s;
}

// Stable ascending sort of (key[i], idx[i]), i < m, by one warp: a bitonic network in the form
// whose compare-exchanges all point the same way (first step of every merge mirrors the block,
// the others are butterflies), so the positions m..P-1 can be left out as virtual +inf.  Ties
// are broken by idx, which makes the order the stable one np.argsort(kind='stable') /
// sorted() produce.  key/idx may live in shared or global memory.
__device__ __forceinline__ void sort_cx(double *key, int32_t *idx, int i, int l) {
    const double ka = key[i], kb = key[l];
    const int32_t ia = idx[i], ib = idx[l];
    if (kb < ka || (kb == ka && ib < ia)) {
        key[i] = kb;
        key[l] = ka;
        idx[i] = ib;
        idx[l] = ia;
    }
}

__device__ __forceinline__ void warp_sort_pairs(double *key, int32_t *idx, int m, int lane) {
    int P = 1;
    while (P < m) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int flip = j == (k >> 1) ? k - 1 : j;      // partner = i ^ flip
            for (int t = lane; t < (P >> 1); t += 32) {
                // t-th pair of this step: i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i ^ flip;
                if (l < m) sort_cx(key, idx, i, l);             // i < l always; beyond m: +inf, nothing to do
            }
            __syncwarp();
        }
    }
}

// The same network for lists too long for shared memory (a thousand-repeat read has ~6 000 pairs): the
// steps whose partners are at least CH apart run on the global arrays, every run of steps with closer
// partners is done chunk by chunk in shared memory (one load and one store of the chunk for up to
// log2(CH) steps) -- 10 global passes instead of 91 for 8 192 positions.
template <int CH>
__device__ void warp_sort_pairs_tiled(double *gkey, int32_t *gidx, int m, int lane, double *skey, int32_t *sidx) {
    int P = 1;
    while (P < m) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        int j = k >> 1;
        for (; j >= CH; j >>= 1) {                            // partners in different chunks
            const int flip = j == (k >> 1) ? k - 1 : j;
            for (int t0 = 0; t0 < (P >> 1); t0 += 128) {      // four pairs per lane in flight
                int pi[4], pl[4];
                double ka[4], kb[4];
                int32_t ia[4], ib[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + 32 * u + lane;
                    pi[u] = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    pl[u] = pi[u] ^ flip;
                    const bool on = t < (P >> 1) && pl[u] < m;
                    if (!on) pl[u] = -1;
                    ka[u] = on ? gkey[pi[u]] : 0.0;
                    kb[u] = on ? gkey[pl[u]] : 0.0;
                    ia[u] = on ? gidx[pi[u]] : 0;
                    ib[u] = on ? gidx[pl[u]] : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (pl[u] >= 0 && (kb[u] < ka[u] || (kb[u] == ka[u] && ib[u] < ia[u]))) {
                        gkey[pi[u]] = kb[u];
                        gkey[pl[u]] = ka[u];
                        gidx[pi[u]] = ib[u];
                        gidx[pl[u]] = ia[u];
                    }
                }
            }
            __syncwarp();
        }
        if (j == 0) continue;
        for (int c0 = 0; c0 < m; c0 += CH) {                  // partners inside one CH-aligned chunk
            const int n = min(CH, m - c0);
            for (int e0 = 0; e0 < n; e0 += 128) {            // four loads of each array in flight per lane
                double kk[4];
                int32_t ii[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + 32 * u + lane;
                    kk[u] = e < n ? gkey[c0 + e] : 0.0;
                    ii[u] = e < n ? gidx[c0 + e] : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + 32 * u + lane;
                    if (e < n) {
                        skey[e] = kk[u];
                        sidx[e] = ii[u];
                    }
                }
            }
            __syncwarp();
            for (int jj = j; jj > 0; jj >>= 1) {
                const int flip = jj == (k >> 1) ? k - 1 : jj;   // (k <= CH: the mirror step is inside the chunk)
                for (int t = lane; t < (CH >> 1); t += 32) {
                    const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
                    const int l = i ^ flip;
                    if (l < n) sort_cx(skey, sidx, i, l);
                }
                __syncwarp();
            }
            for (int e = lane; e < n; e += 32) {
                gkey[c0 + e] = skey[e];
                gidx[c0 + e] = sidx[e];
            }
            __syncwarp();
        }
    }
}

constexpr int SORT_SMEM = 512;   // pairs per warp sorted in shared memory; longer lists use the scratch

// Kernel A (one warp per read): run-length view, per-run statistics, filtered pairs sorted by
// state value, repeat-region borders.
template <bool SECOND>
__global__ void __launch_bounds__(128) mid_stats_kernel(const MidParams p) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(p.queue, 1);
        q = __shfl_sync(FULL, q, 0);
        if (q >= p.n) break;
        const MidRead rd = p.reads[q];
        if (p.status[rd.read] != WSTR_READ_OK) continue;
        const MidAutomaton A = p.auts[rd.aut];
        const int T = rd.T;
        const double *x = p.x + rd.sig_off;
        const int32_t *trace = p.trace + rd.sig_off;
        const int R = rd.run_cap;
        const Scratch sc = carve(p.scratch + rd.ws_off, R, T, p.s_max);
        int32_t *run_start = sc.run_start, *run_state = sc.run_state;
        int fail = 0;

        // ---- run-length view of the trace (caller.py:58-60) -----------------------------------
        // (four trace words per lane are in flight before the first is looked at; the state of the
        // sample before comes from the neighbouring lane, not from a second load)
        int n_runs = 0;
        {
            const int32_t *__restrict__ tr = trace;
            int before = -1;                       // state of the sample in front of the current 32
            for (int t0 = 0; t0 < T; t0 += 128) {
                int sv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + 32 * u + lane;
                    sv[u] = t < T ? tr[t] : -1;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + 32 * u + lane;
                    const int st = sv[u];
                    int left = __shfl_up_sync(FULL, st, 1);
                    if (lane == 0) left = before;
                    before = __shfl_sync(FULL, st, 31);
                    const bool head = t < T && (t == 0 || left != st);
                    const unsigned bal = __ballot_sync(FULL, head);
                    if (head) {
                        const int r = n_runs + __popc(bal & ((1u << lane) - 1u));
                        if (r < R) {
                            run_start[r] = t;
                            run_state[r] = st;
                        }
                    }
                    n_runs += __popc(bal);
                }
            }
        }
        if (n_runs > R) fail = WSTR_READ_SEGMENT;   // cannot happen: a run has >= mv-1 samples
        if (!fail && lane == 0) run_start[n_runs] = T;
        __syncwarp();

        // ---- alignment list (caller.py:65-96) and its usable pairs (:23-39) -----------------------
        int m = 0;         // good pairs
        int n_align = 0;   // entries of the alignment list
        if (!fail && !p.reps) {
            n_align = n_runs;
            for (int r0 = 0; r0 < n_runs; r0 += 32) {
                const int r = r0 + lane;
                bool g = false;
                double mean = 0.0, expect = 0.0;
                if (r < n_runs) {
                    const int a = run_start[r], n = run_start[r + 1] - a;
                    const double *xs = x + a;
                    expect = A.values[run_state[r]];
                    if (n <= 16) {
                        // the usual case: the run's samples fetched once, all loads in flight together, both
                        // sums taken from registers
                        double w[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) w[u] = u < n ? xs[u] : 0.0;
                        mean = pairwise16(w, n) / (double)n;
                        const double avg = mean;
                        if (p.method == 1) mean = run_median(xs, n);
                        sc.sv[r] = mean;
                        if (n >= p.mv) {
#pragma unroll
                            for (int u = 0; u < 16; ++u) {
                                const double d = w[u] - avg;
                                w[u] = d * d;
                            }
                            const double sd = sqrt(pairwise16(w, n) / (double)n);
                            g = sd < p.max_std && fabs(expect - mean) <= p.threshold;
                        }
                    } else {
                        const double sum = pairwise_sum([xs](int i) { return xs[i]; }, 0, n);
                        mean = sum / (double)n;
                        const double avg = mean;                       // np.std is taken about the mean either way
                        if (p.method == 1) mean = run_median(xs, n);   // 'mean' is the run's state value from here on
                        sc.sv[r] = mean;
                        if (n >= p.mv) {
                            const double ss = pairwise_sum(
                                [xs, avg](int i) {
                                    const double d = xs[i] - avg;
                                    return d * d;
                                },
                                0, n);
                            const double sd = sqrt(ss / (double)n);
                            g = sd < p.max_std && fabs(expect - mean) <= p.threshold;
                        }
                    }
                }
                const unsigned bal = __ballot_sync(FULL, g);
                if (g) {
                    const int k = m + __popc(bal & ((1u << lane) - 1u));
                    sc.px[k] = mean;
                    sc.py[k] = expect;
                }
                m += __popc(bal);
            }
            __syncwarp();
            if (m <= 3) fail = WSTR_READ_SPLINE;   // splrep: 'm > k must hold'
        }
        if (!fail && p.reps) {
            // rescaling.reps_as_one (caller.py:69-79): one entry per distinct state of the path, ascending
            // state index (np.unique), over ALL samples mapped to the state in time order
            // (np.take(signal, np.where(trace == state)))
            const int S = A.n_states;
            int32_t *cnt = sc.cnt, *base = sc.base, *dest = sc.sidx;
            for (int j = lane; j < S; j += 32) cnt[j] = 0;
            __syncwarp();
            for (int r = lane; r < n_runs; r += 32) atomicAdd(&cnt[run_state[r]], run_start[r + 1] - run_start[r]);
            __syncwarp();
            if (lane == 0) {
                int acc = 0;
                for (int j = 0; j < S; ++j) {
                    base[j] = acc;
                    acc += cnt[j];
                }
                base[S] = acc;
                // every run's place inside its state's group: runs in time order, one after the other
                for (int j = 0; j < S; ++j) cnt[j] = base[j];
                for (int r = 0; r < n_runs; ++r) {
                    const int st = run_state[r];
                    dest[r] = cnt[st];
                    cnt[st] += run_start[r + 1] - run_start[r];
                }
            }
            __syncwarp();
            for (int r = lane; r < n_runs; r += 32) {
                const int a = run_start[r], n = run_start[r + 1] - a, d = dest[r];
                for (int k = 0; k < n; ++k) sc.gvals[d + k] = x[a + k];
            }
            __syncwarp();
            for (int j0 = 0; j0 < S; j0 += 32) {
                const int j = j0 + lane;
                const int n = j < S ? base[j + 1] - base[j] : 0;
                bool g = false;
                double value = 0.0, expect = 0.0;
                if (n > 0) {
                    const double *xs = sc.gvals + base[j];
                    const double avg = pairwise_sum([xs](int i) { return xs[i]; }, 0, n) / (double)n;
                    value = p.method == 1 ? list_median(xs, n) : avg;
                    expect = A.values[j];
                    if (n >= p.mv) {
                        const double ss = pairwise_sum(
                            [xs, avg](int i) {
                                const double d = xs[i] - avg;
                                return d * d;
                            },
                            0, n);
                        const double sd = sqrt(ss / (double)n);
                        g = sd < p.max_std && fabs(expect - value) <= p.threshold;
                    }
                }
                const unsigned present = __ballot_sync(FULL, n > 0);
                if (n > 0) {
                    const int k = n_align + __popc(present & ((1u << lane) - 1u));
                    sc.sv[k] = value;
                    sc.astate[k] = j;
                }
                n_align += __popc(present);
                const unsigned bal = __ballot_sync(FULL, g);
                if (g) {
                    const int k = m + __popc(bal & ((1u << lane) - 1u));
                    sc.px[k] = value;
                    sc.py[k] = expect;
                }
                m += __popc(bal);
            }
            __syncwarp();
            if (m <= 3) fail = WSTR_READ_SPLINE;
        }

        if (!SECOND && !fail) {
            // ---- stable sort by state value (caller.py:306) -----------------------------------------
            __shared__ double s_key[4][SORT_SMEM];
            __shared__ int32_t s_idx[4][SORT_SMEM];
            const bool in_smem = m <= SORT_SMEM;
            if (in_smem) {
                // (the shared arrays by name, so that the network's loads and stores are LDS/STS, not generic)
                double *key = s_key[threadIdx.x >> 5];
                int32_t *idx = s_idx[threadIdx.x >> 5];
                for (int i = lane; i < m; i += 32) {
                    key[i] = sc.px[i];
                    idx[i] = i;
                }
                __syncwarp();
                warp_sort_pairs(key, idx, m, lane);
                for (int i = lane; i < m; i += 32) {
                    sc.sx[i] = key[i];
                    sc.sy[i] = sc.py[idx[i]];
                }
            } else {
                double *key = sc.sx;
                int32_t *idx = sc.sidx;
                for (int i = lane; i < m; i += 32) {
                    key[i] = sc.px[i];
                    idx[i] = i;
                }
                __syncwarp();
                warp_sort_pairs_tiled<SORT_SMEM>(key, idx, m, lane, s_key[threadIdx.x >> 5], s_idx[threadIdx.x >> 5]);
                for (int i = lane; i < m; i += 32) sc.sy[i] = sc.py[idx[i]];
            }
            __syncwarp();
        }

        // ---- repeat-region borders (caller.py:381-406) -------------------------------------------
        int start = -1, end = -1, ra = 0, rb = 0, ncuts = 0;
        if (!fail) {
            int first = 1 << 30, last = -1;
            for (int r = lane; r < n_runs; r += 32) {
                if (A.rep_mask[run_state[r]]) {
                    first = min(first, r);
                    last = max(last, r);
                }
            }
            first = warp_min(first);
            last = warp_max(last);
            if (last < 0) {
                fail = WSTR_READ_NO_REPEAT_STATE;
            } else {
                start = first;
                end = last;
                const int s_state = run_state[start];
                int fa = 1 << 30;
                for (int r = lane; r < n_runs; r += 32)
                    if (run_state[r] == s_state) fa = min(fa, r);
                ra = warp_min(fa);
                for (int pass = 0; pass < 2 && !fail; ++pass) {
                    if (end >= n_runs) {
                        fail = WSTR_READ_SEGMENT;   // state_transitions[end]
                        break;
                    }
                    const int e_state = run_state[end];
                    int lb = -1;
                    for (int r = lane; r < n_runs; r += 32)
                        if (run_state[r] == e_state) lb = max(lb, r);
                    rb = warp_max(lb);
                    ncuts = rb > ra ? rb - ra : 0;
                    if (pass == 1) break;
                    const int sis = p.sis;
                    const int extra = (((ncuts - 1) % sis) + sis) % sis;
                    if (extra == 0) break;
                    end = end + (sis - extra);
                }
            }
        }
        int nb = 0;
        if (!fail) {
            nb = (ncuts + p.sis - 1) / p.sis;
            if (nb == 0) fail = WSTR_READ_SEGMENT;   // bounds[0]
        }
        // bounds[n] = run_start[ra + n*sis + 1] - 1; a window is x[b_n-3 : b_{n+1}+3] and needs
        // at least one full 6-sample stretch (caller.py:336,358-361)
        if (!fail && nb > 1) {
            const int b0 = run_start[ra + 1] - 1;
            const int bpen = run_start[ra + (nb - 2) * p.sis + 1] - 1;
            if (b0 - 3 < 0 || bpen > T - 3) fail = WSTR_READ_SEGMENT;
        }
        if (lane == 0) {
            MidState st;
            st.n_runs = n_runs;
            st.m = m;
            st.start = start;
            st.end = end;
            st.ra = ra;
            st.nb = nb;
            st.n_align = n_align;
            st.pad_ = 0;
            p.state[q] = st;
            if (fail) p.status[rd.read] = fail;
        }
        __syncwarp();
    }
}

// Kernel B (first pass only): the smoothing spline's accepted first iteration, one thread per read.
// Strictly sequential per read (4 m dependent Givens rotations), so reads are spread over threads: the
// form for large batches, where a hundred thousand reads keep every lane busy.
__global__ void __launch_bounds__(128) mid_fit_kernel(const MidParams p) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.n) return;
    const MidRead rd = p.reads[q];
    if (p.status[rd.read] != WSTR_READ_OK) return;
    const MidState st = p.state[q];
    const Scratch sc = carve(p.scratch + rd.ws_off, rd.run_cap, rd.T, p.s_max);
    Cubic cu;
    lsq_cubic(sc.sx, sc.sy, st.m, cu);
    const double s = (double)st.m, acc = 0.001 * s;
    const double fpms = cu.fp - s;
    // fpcurf: accept if |fp-s| < acc or fp < s; otherwise knots would be added
    if ((!(fabs(fpms) < acc) && !(fpms < 0.0)) || !(cu.fp == cu.fp)) p.status[rd.read] = WSTR_READ_SPLINE_KNOTS;
    double *o = p.cubic + (size_t)q * 8;
    o[0] = cu.xb;
    o[1] = cu.xe;
    o[2] = cu.c[0];
    o[3] = cu.c[1];
    o[4] = cu.c[2];
    o[5] = cu.c[3];
    o[6] = cu.fp;
}

// Kernel B for small batches of long reads (a few thousand reads of thousands of pairs each: the sweep's
// latency is then all there is).  The Givens sweep is sequential per matrix element (row `it` rotates
// into the triangle left by rows 0..it-1), but its four columns form a pipeline: while row `it` is
// rotated into column i, row `it-1` can be rotated into column i+1.  Four lanes work on one read, each
// owning one column of the triangle (pivot, off-diagonals, right-hand side) and handing the rest of the
// rotated row to the next lane by shuffle; eight reads share a warp.  A lane always sees its pivot as
// the first element of what it receives and passes on what is right of it, so all lanes run the same
// instructions (a column with fewer off-diagonals rotates zeros).  The rows' B-spline bases (fpbspl) do
// not depend on one another: the warp evaluates them for all rows first, 32 at a time, into the read's
// free scratch (the unsorted-pair arrays and the not yet written rescaled signal).  Every matrix
// element sees the same rotations in the same order as in FITPACK's loop: the coefficients are bit-equal
// to the thread-per-read kernel's and to scipy's.  A read costs m + 3 pipeline steps instead of 4 m
// dependent rotations.
constexpr int FIT_STAGES = 4;
constexpr int FIT_READS_PER_WARP = 32 / FIT_STAGES;   // 8
constexpr int FIT_PIPE_MAX_READS = 16384;             // larger batches fill the GPU with one thread per read

__device__ __forceinline__ double shfl_up_d(double v, int delta) { return __shfl_up_sync(FULL, v, delta); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

__global__ void __launch_bounds__(128) mid_fit_pipe_kernel(const MidParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int slot = lane / FIT_STAGES, stage = lane - slot * FIT_STAGES;

    // ---- the basis of every row of the warp's reads, all lanes on one read at a time ----------------
    for (int g = 0; g < FIT_READS_PER_WARP; ++g) {
        const int qg = warp * FIT_READS_PER_WARP + g;
        if (qg >= p.n) break;
        const MidRead rd = p.reads[qg];
        if (p.status[rd.read] != WSTR_READ_OK) continue;
        const Scratch sc = carve(p.scratch + rd.ws_off, rd.run_cap, rd.T, p.s_max);
        const int mg = p.state[qg].m;
        const double xb = sc.sx[0], xe = sc.sx[mg - 1];
        const Divisor span = make_divisor(xe - xb);
        double *b2 = p.rescaled + rd.sig_off, *b3 = b2 + rd.run_cap;     // 2 * run_cap <= T, checked by the launcher
        for (int it = lane; it < mg; it += 32) {
            double h[4];
            bspl3<true>(xb, xe, span, sc.sx[it], h);
            sc.px[it] = h[0];
            sc.py[it] = h[1];
            b2[it] = h[2];
            b3[it] = h[3];
        }
    }
    __syncwarp();

    const int q = warp * FIT_READS_PER_WARP + slot;
    const bool active = q < p.n && p.status[p.reads[q < p.n ? q : 0].read] == WSTR_READ_OK;
    int m = 0;
    const double *sy = nullptr, *b0 = nullptr, *b1 = nullptr, *b2 = nullptr, *b3 = nullptr;
    double xb = 0.0, xe = 0.0;
    if (active) {
        const MidRead rd = p.reads[q];
        const Scratch sc = carve(p.scratch + rd.ws_off, rd.run_cap, rd.T, p.s_max);
        m = p.state[q].m;
        sy = sc.sy;
        b0 = sc.px;
        b1 = sc.py;
        b2 = p.rescaled + rd.sig_off;
        b3 = b2 + rd.run_cap;
        xb = sc.sx[0];
        xe = sc.sx[m - 1];
    }
    int steps = m + FIT_STAGES - 1;
    for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, o));

    // this lane's column of the triangle (stage = column): a0 the pivot, a1..a3 the row's tail
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, z = 0.0, fp = 0.0;
    // what this lane hands to the next stage: the row right of this stage's column, after its rotation
    double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0, oy = 0.0;
    int ovalid = 0;
    // the first column reads its rows from memory, one step ahead of their use
    const bool feeds = active && stage == 0;
    double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0, ny = 0.0;
    if (feeds && m > 0) {
        n0 = b0[0];
        n1 = b1[0];
        n2 = b2[0];
        n3 = b3[0];
        ny = sy[0];
    }

    for (int tau = 0; tau < steps; ++tau) {
        // the previous stage's output of the step before
        double t0 = shfl_up_d(o0, 1), t1 = shfl_up_d(o1, 1), t2 = shfl_up_d(o2, 1), t3 = shfl_up_d(o3, 1);
        double yi = shfl_up_d(oy, 1);
        int valid = __shfl_up_sync(FULL, ovalid, 1);
        if (stage == 0) {
            t0 = n0;
            t1 = n1;
            t2 = n2;
            t3 = n3;
            yi = ny;
            valid = feeds && tau < m;
            if (feeds && tau + 1 < m) {
                n0 = b0[tau + 1];
                n1 = b1[tau + 1];
                n2 = b2[tau + 1];
                n3 = b3[tau + 1];
                ny = sy[tau + 1];
            }
        }
        ovalid = active && valid;
        if (ovalid) {
            if (t0 != 0.0) {                       // fpcurf skips a zero pivot
                double c, sn;
                givens(t0, a0, c, sn);
                rotate(c, sn, yi, z);
                rotate(c, sn, t1, a1);
                rotate(c, sn, t2, a2);
                rotate(c, sn, t3, a3);
            }
            fp = fp + yi * yi;                     // (the last column's is the residual sum)
            o0 = t1;
            o1 = t2;
            o2 = t3;
            o3 = 0.0;
            oy = yi;
        }
    }

    // fpback (n = 4, bandwidth 4) on the first lane of the read; the other columns come over by shuffle
    const int l0 = slot * FIT_STAGES;
    const double A00 = shfl_d(a0, l0), A01 = shfl_d(a1, l0), A02 = shfl_d(a2, l0), A03 = shfl_d(a3, l0), Z0 = shfl_d(z, l0);
    const double A10 = shfl_d(a0, l0 + 1), A11 = shfl_d(a1, l0 + 1), A12 = shfl_d(a2, l0 + 1), Z1 = shfl_d(z, l0 + 1);
    const double A20 = shfl_d(a0, l0 + 2), A21 = shfl_d(a1, l0 + 2), Z2 = shfl_d(z, l0 + 2);
    const double A30 = shfl_d(a0, l0 + 3), Z3 = shfl_d(z, l0 + 3), FP = shfl_d(fp, l0 + 3);
    if (active && stage == 0) {
        double c[4];
        c[3] = Z3 / A30;
        {   // i = 2: store = z[2] - c[3]*a[2][1]
            double store = Z2;
            store = store - c[3] * A21;
            c[2] = store / A20;
        }
        {   // i = 1: store = z[1] - c[2]*a[1][1] - c[3]*a[1][2]
            double store = Z1;
            store = store - c[2] * A11;
            store = store - c[3] * A12;
            c[1] = store / A10;
        }
        {   // i = 0: store = z[0] - c[1]*a[0][1] - c[2]*a[0][2] - c[3]*a[0][3]
            double store = Z0;
            store = store - c[1] * A01;
            store = store - c[2] * A02;
            store = store - c[3] * A03;
            c[0] = store / A00;
        }
        const MidRead rd = p.reads[q];
        const double s = (double)m, acc = 0.001 * s;
        const double fpms = FP - s;
        // fpcurf: accept if |fp-s| < acc or fp < s; otherwise knots would be added
        if ((!(fabs(fpms) < acc) && !(fpms < 0.0)) || !(FP == FP)) p.status[rd.read] = WSTR_READ_SPLINE_KNOTS;
        double *o = p.cubic + (size_t)q * 8;
        o[0] = xb;
        o[1] = xe;
        o[2] = c[0];
        o[3] = c[1];
        o[4] = c[2];
        o[5] = c[3];
        o[6] = FP;
    }
}

// Kernel C (one warp per read): rescaled signal + bad-repeat mask (first pass), cost, sequence.
template <bool SECOND>
__global__ void __launch_bounds__(128) mid_finish_kernel(const MidParams p) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(p.queue + 1, 1);
        q = __shfl_sync(FULL, q, 0);
        if (q >= p.n) break;
        const MidRead rd = p.reads[q];
        if (p.status[rd.read] != WSTR_READ_OK) continue;
        const MidAutomaton A = p.auts[rd.aut];
        const MidState st = p.state[q];
        const int T = rd.T;
        const double *x = p.x + rd.sig_off;
        const Scratch sc = carve(p.scratch + rd.ws_off, rd.run_cap, rd.T, p.s_max);
        const int32_t *run_start = sc.run_start, *run_state = sc.run_state;
        const int n_runs = st.n_runs, ra = st.ra, nb = st.nb;
        int fail = 0;

        if (!SECOND) {
            Cubic cu;
            const double *ci = p.cubic + (size_t)q * 8;
            cu.xb = ci[0];
            cu.xe = ci[1];
            cu.c[0] = ci[2];
            cu.c[1] = ci[3];
            cu.c[2] = ci[4];
            cu.c[3] = ci[5];
            // ---- rescaled signal (caller.py:312) ------------------------------------------------
            // four samples per lane are fetched before the first is evaluated; a sample whose distances
            // to both knots are tame takes the unguarded divisions (all of them, in practice)
            double *__restrict__ out = p.rescaled + rd.sig_off;
            const double *__restrict__ xr = x;
            const Divisor span = make_divisor(cu.xe - cu.xb);
            bool all_tame = true;
            for (int t0 = 0; t0 < T; t0 += 128) {
                double xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + 32 * u + lane;
                    xv[u] = t < T ? xr[t] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + 32 * u + lane;
                    if (t < T) {
                        const double v = xv[u];
                        all_tame = all_tame && tame(v);
                        if (span.fast && tame(cu.xe - v) && tame(v - cu.xb)) out[t] = splev3<false>(cu, span, v);
                        else out[t] = splev3<true>(cu, span, v);
                    }
                }
            }
            const bool tt_fast = __all_sync(FULL, all_tame);
            // ---- bad-repeat mask (caller.py:336-344, 409-421) -----------------------------------
            uint32_t *mw = p.maskbits + rd.mask_off;
            const int nwords = (T + 31) >> 5;
            for (int w = lane; w < nwords; w += 32) mw[w] = 0u;
            __syncwarp();
            int ties = 0;
            for (int n0 = 0; n0 < nb - 1; n0 += 32) {
                const int n = n0 + lane;
                if (n < nb - 1) {
                    const int b_lo = run_start[ra + n * p.sis + 1] - 1;
                    const int b_hi = run_start[ra + (n + 1) * p.sis + 1] - 1;
                    const int c1 = min(b_hi, T - 3);
                    const int segs = tt_fast ? count_segments<false>(x, b_lo, c1, p.tie_ulps, ties)
                                             : count_segments<true>(x, b_lo, c1, p.tie_ulps, ties);
                    if (segs >= p.sis + 1) {
                        for (int w = b_lo >> 5; w <= (b_hi - 1) >> 5; ++w) {
                            const int lo = max(b_lo, w << 5), hi = min(b_hi, (w + 1) << 5);   // [lo, hi)
                            if (hi <= lo) continue;
                            const int cnt = hi - lo;
                            const uint32_t bits = (cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u)) << (lo & 31);
                            atomicOr(mw + w, bits);
                        }
                    }
                }
            }
            if (p.ties) {
                for (int o = 16; o > 0; o >>= 1) ties += __shfl_xor_sync(FULL, ties, o);
                if (lane == 0) p.ties[rd.read] = ties;
            }
        }

        // ---- state-wise cost (caller.py:138-139) and decoded sequence (:178-187) ----------------
        {
            // alignment[start:end] (start, end are run indices even when the list is per state: the
            // reference applies them as they are, caller.py:138-139)
            const int start = st.start;
            const int e_clip = min(st.end, st.n_align);
            const int cnt = e_clip > start ? e_clip - start : 0;
            const int32_t *a_state = p.reps ? sc.astate : run_state;
            for (int r = start + lane; r < start + cnt; r += 32)
                sc.px[r - start] = fabs(sc.sv[r] - A.values[a_state[r]]);
            __syncwarp();
            if (lane == 0) {
                double cost;
                if (cnt == 0) {
                    cost = __longlong_as_double(0x7ff8000000000000LL);   // np.mean([]) is nan
                } else {
                    const double *cp = sc.px;
                    cost = pairwise_sum([cp](int i) { return cp[i]; }, 0, cnt) / (double)cnt;
                }
                p.cost[rd.read] = cost;
            }
            // python: seq[F - offset : -F] over one character per run
            const int F = A.flank_length;
            const int offset = A.seq_idx[run_state[0]];
            int a = F - offset, b = F == 0 ? 0 : n_runs - F;
            if (a < 0) a = max(n_runs + a, 0);
            if (a > n_runs) a = n_runs;
            if (b < 0) b = 0;
            const int len = b > a ? b - a : 0;
            if (lane == 0) p.len[rd.read] = len;
            if (p.seq && rd.seq_off >= 0) {
                uint8_t *so = p.seq + rd.seq_off;
                for (int i = lane; i < len; i += 32) {
                    if (rd.reverse) so[i] = complement(A.last_base[run_state[b - 1 - i]]);
                    else so[i] = A.last_base[run_state[a + i]];
                }
            }
        }
        if (fail && lane == 0) p.status[rd.read] = fail;
        __syncwarp();
    }
}

}  // namespace

// p.queue must point at two zeroed counters
int wstr_launch_midstage(const MidParams &p, bool second, cudaStream_t s) {
    if (p.n <= 0) return WSTR_OK;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        WSTR_CUDA(cudaGetDevice(&dev));
        WSTR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = (p.n + 3) / 4;
    if (grid > sms * 8) grid = sms * 8;
    if (second) {
        mid_stats_kernel<true><<<grid, 128, 0, s>>>(p);
        mid_finish_kernel<true><<<grid, 128, 0, s>>>(p);
    } else {
        mid_stats_kernel<false><<<grid, 128, 0, s>>>(p);
        if (p.n <= FIT_PIPE_MAX_READS && p.pipe_ok)
            mid_fit_pipe_kernel<<<(p.n + 4 * FIT_READS_PER_WARP - 1) / (4 * FIT_READS_PER_WARP), 128, 0, s>>>(p);
        else
            mid_fit_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
        mid_finish_kernel<false><<<grid, 128, 0, s>>>(p);
    }
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}

int64_t wstr_mid_scratch_bytes(int T, int mv, int reps, int s_max) {
    const int64_t R = T / (mv > 2 ? mv - 1 : 1) + 16;
    int64_t b = (int64_t)scratch_core_bytes((int)R);
    if (reps) b += 8 * (int64_t)T + 4 * (2 * (int64_t)s_max + 2);   // gvals, cnt, base
    return (b + 255) / 256 * 256;
}
