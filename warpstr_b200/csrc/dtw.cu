// DTW state-automaton fill + traceback for sm_100a.
//
// Replaces WarpSTR._calc_dtw_astates (reference src/caller/caller.py:198-245) and
// WarpSTR._backtracking (:247-301).
//
// One warp owns one read; the automaton's states live in registers, K = KC + KG per lane:
//   * KC "chain" slots: lane l holds KC consecutive states of a pure chain (every state has
//     exactly one incoming edge, from the state before it: the flanks).  Its skip candidate
//     comes from the neighbouring register; slot 0 reads the previous lane's tail.
//   * KG "generic" slots: one arbitrary state per lane and slot (the repeat region: loop
//     back-edges, optional-group skips, IUPAC fan-in, context copies).  Up to DEG incoming
//     edges each, kept in the reference's incoming order, gathered from a small published
//     row in shared memory.
// Row i of the DP needs only row i-1 (stay) and row i-back (skip through an incoming state),
// so per state the warp carries D[i-1] and the running partial sums
//     P_m[t] = D[t] + |x[t+1]-v| + ... + |x[t+m]-v|        m = 1..mv-1
// in the reference's own left-to-right addition order; the skip candidate into state j over
// incoming state p is P_{back-1}[i-back][p] + |x[i]-v_j|, bit-identical to the reference's
// nested loop.  Every row each lane publishes the values other lanes may need (its generic
// states and its chain tail, KG+1 stores), one __syncwarp, then everything is branch-free:
// lanes with fewer edges read a cell that always holds +inf.  All arithmetic is FP64
// add/abs/compare: no multiply, so no FMA contraction can occur.
//
// The read's signal is streamed through shared memory in 1 KB tiles with 1-D bulk async
// copies (cp.async.bulk + mbarrier, TMA engine) two tiles ahead of the row loop.
// Output: one 4-bit direction code per cell (0 stay, c>0 = c-th incoming edge), packed 8 per
// 32-bit word, stored row-major [row][word][lane] so that every row is one coalesced
// 128-byte line per word.  D itself never leaves the SM.
#include "wstr_internal.h"

namespace {

constexpr int CH = WSTR_SIG_CHUNK;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one elected lane: announce `bytes` and start the bulk copy global -> shared
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

template <int K, int MV>
struct LaneState {
    double v[K];
    double D[K];
    double P[MV - 1][K];   // P[m-1] = P_m
};

// the pipeline value a state offers to its successors in this row
template <int K, int MV, bool SHORT>
__device__ __forceinline__ double offer(const LaneState<K, MV> &s, int k) {
    if (SHORT) {
        if (MV >= 3) return s.P[MV >= 3 ? MV - 3 : 0][k];
        return s.D[k];
    }
    return s.P[MV - 2][k];
}

template <int KG>
struct LaneConsts {
    uint32_t band_bits;     // slot u of this lane is inside the end band's skipped prefix
    uint32_t src0;          // byte offset (within a published row) of chain slot 0's predecessor
    uint32_t gsrc[KG];      // generic slot g: 4 x 8-bit published-row indices, incoming order
};

// published row: [generic g][lane] ... [chain tail][lane], [+inf cell]
template <int KG>
__host__ __device__ constexpr int q_row_len() { return 32 * KG + 40; }

// One DP row.  SHORT: this row allows dwell mv-1 (masked row of the second pass).
// BAND: the end band is active (caller.py:223-224).
template <int KC, int KG, int DEG, int MV, bool SHORT, bool BAND>
__device__ __forceinline__ void dp_row(LaneState<KC + KG, MV> &s, const LaneConsts<KG> &lc, const double x,
                                       double *__restrict__ Q, const int lane,
                                       uint32_t *__restrict__ dir_row) {
    constexpr int K = KC + KG;
    constexpr int W = (K + 7) / 8;
    const double INF = dinf();

    // what other lanes may read this row: my generic states and my chain tail
#pragma unroll
    for (int g = 0; g < KG; ++g) Q[g * 32 + lane] = offer<K, MV, SHORT>(s, KC + g);
    Q[KG * 32 + lane] = offer<K, MV, SHORT>(s, KC - 1);
    __syncwarp();

    uint32_t codes[W];
#pragma unroll
    for (int w = 0; w < W; ++w) codes[w] = 0u;

    // ---- chain slots: stay or the single incoming edge -----------------------------------
    double qprev = *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(Q) + lc.src0);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const double ae = fabs(x - s.v[k]);
        const double qhere = offer<K, MV, SHORT>(s, k);
        const double stay = s.D[k] + ae;
        const double ch = qprev + ae;
        const bool take = ch < stay;
        double best = take ? ch : stay;
        uint32_t code = take ? 1u : 0u;
        if (BAND) {
            if ((lc.band_bits >> k) & 1u) {
                best = INF;
                code = 0u;
            }
        }
#pragma unroll
        for (int m = MV - 2; m >= 1; --m) s.P[m][k] = s.P[m - 1][k] + ae;
        s.P[0][k] = stay;
        s.D[k] = best;
        qprev = qhere;
        codes[k >> 3] |= code << (4 * (k & 7));
    }

    // ---- generic slots: stay, then up to DEG incoming edges in list order --------------------
#pragma unroll
    for (int g = 0; g < KG; ++g) {
        const int k = KC + g;
        const double ae = fabs(x - s.v[k]);
        const double stay = s.D[k] + ae;
        double best = stay;
        uint32_t code = 0u;
#pragma unroll
        for (int r = 0; r < DEG; ++r) {
            const uint32_t idx = (lc.gsrc[g] >> (8 * r)) & 0xffu;
            const double c = Q[idx] + ae;
            if (c < best) {
                best = c;
                code = static_cast<uint32_t>(r + 1);
            }
        }
        if (BAND) {
            if ((lc.band_bits >> k) & 1u) {
                best = INF;
                code = 0u;
            }
        }
#pragma unroll
        for (int m = MV - 2; m >= 1; --m) s.P[m][k] = s.P[m - 1][k] + ae;
        s.P[0][k] = stay;
        s.D[k] = best;
        codes[k >> 3] |= code << (4 * (k & 7));
    }
#pragma unroll
    for (int w = 0; w < W; ++w) dir_row[w * 32 + lane] = codes[w];
}

// per-warp shared memory: [sig 2 x CH f64][published rows 2 x q_row_len f64][2 mbarriers]
template <int KG>
struct alignas(16) FillSmem {
    double sig[2][CH];
    double Q[2][q_row_len<KG>()];
    uint64_t bar[2];
};

template <int KC, int KG, int DEG, int MV>
__global__ void __launch_bounds__(32 * WSTR_WARPS_PER_CTA, (KC + KG <= 8 ? 4 : (KC + KG <= 12 ? 3 : 2)))
dtw_fill_kernel(const FillParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int K = KC + KG;
    constexpr int W = (K + 7) / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FillSmem<KG> &sm = reinterpret_cast<FillSmem<KG> *>(smem_raw)[warp];
    const double INF = dinf();

    if (lane == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        fence_barrier_init();
    }
    for (int e = lane; e < 2 * q_row_len<KG>(); e += 32) (&sm.Q[0][0])[e] = INF;   // incl. the +inf cell
    __syncwarp();
    uint32_t uses0 = 0, uses1 = 0;   // completed phases of the two tile barriers

    LaneState<K, MV> s;
    LaneConsts<KG> lc;
    int cached_aut = -1;
    double v0 = 0.0;

    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(p.queue, 1);
        r = __shfl_sync(FULL, r, 0);
        if (r >= p.n) break;
        const ReadMeta m = p.meta[p.order[r]];
        const DevAutomaton *A = p.auts + m.aut;
        const int T = m.T;
        if (T <= MV) {
            if (lane == 0) p.status[m.read] = WSTR_READ_TOO_SHORT;
            continue;
        }

        if (m.aut != cached_aut) {   // (re)load the automaton into registers
#pragma unroll
            for (int k = 0; k < K; ++k) s.v[k] = __ldg(A->v_pos + lane * K + k);
            lc.band_bits = __ldg(A->lane_tab + lane * WSTR_LANE_TAB_STRIDE + 0);
            lc.src0 = __ldg(A->lane_tab + lane * WSTR_LANE_TAB_STRIDE + 1) * 8u;
#pragma unroll
            for (int g = 0; g < KG; ++g) lc.gsrc[g] = __ldg(A->lane_tab + lane * WSTR_LANE_TAB_STRIDE + 2 + g);
            v0 = __ldg(A->v_pos + A->init_pos[0]);
            cached_aut = m.aut;
        }

        // ---- signal tiles: two in flight --------------------------------------------------
        const double *gsig = p.signal + m.sig_off;
        const int nchunks = (T + CH - 1) / CH;
        auto issue = [&](int c) {
            if (lane == 0) {
                int n = T - c * CH;
                n = n > CH ? CH : n;
                n = (n + 1) & ~1;   // 16-byte granules
                bulk_load(sm.sig[c & 1], gsig + static_cast<int64_t>(c) * CH, static_cast<uint32_t>(n) * 8u,
                          &sm.bar[c & 1]);
            }
        };
        issue(0);
        if (nchunks > 1) issue(1);
        while (!mbar_try_wait(&sm.bar[0], uses0 & 1u)) {
        }
        ++uses0;

        // ---- row 0 (caller.py:201-208) and the empty rows 1..mv-1 -----------------------
        {
            const double *x0 = sm.sig[0];
            const double first = fabs(x0[0] - v0);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                s.D[k] = INF;
#pragma unroll
                for (int mm = 0; mm < MV - 1; ++mm) s.P[mm][k] = INF;
            }
            for (int c = 0; c <= MV; ++c) {
                const int pos = A->init_pos[c];
                double d0 = first;
                if (c > 0) d0 = first + fabs(x0[c] - v0);   // row 0, column c
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (pos == lane * K + k) {
                        double acc = d0;
                        for (int t = 1; t <= MV - 1; ++t) acc = acc + fabs(x0[t] - s.v[k]);
                        s.P[MV - 2][k] = acc;   // P_{mv-1}[0]
                    }
                }
            }
        }

        const int band_start = max(A->th1, T - A->band6 + 1);
        const uint32_t *mw_ptr = p.maskbits ? p.maskbits + m.mask_off : nullptr;
        uint32_t mw = 0u, mw_next = 0u;
        const int nwords = (T + 31) >> 5;
        int mblock = MV >> 5;            // 32-row block whose mask word is in mw
        if (mw_ptr) {
            mw = __ldg(mw_ptr + mblock);
            if (mblock + 1 < nwords) mw_next = __ldg(mw_ptr + mblock + 1);
        }
        uint32_t *dir = p.dir + m.dir_off;

        for (int c = 0; c < nchunks; ++c) {
            if (c > 0) {
                if (c & 1) {
                    while (!mbar_try_wait(&sm.bar[1], uses1 & 1u)) {
                    }
                    ++uses1;
                } else {
                    while (!mbar_try_wait(&sm.bar[0], uses0 & 1u)) {
                    }
                    ++uses0;
                }
            }
            const double *xs = sm.sig[c & 1];
            const int i_begin = c == 0 ? MV : c * CH;
            const int i_end = min(T, (c + 1) * CH);
            for (int i = i_begin; i < i_end; ++i) {
                if (mw_ptr && (i >> 5) != mblock) {   // next mask word, fetched one block ahead
                    mblock = i >> 5;
                    mw = mw_next;
                    if (mblock + 1 < nwords) mw_next = __ldg(mw_ptr + mblock + 1);
                }
                const double x = xs[i & (CH - 1)];
                const bool shortrow = (mw >> (i & 31)) & 1u;
                double *Q = sm.Q[i & 1];
                uint32_t *dir_row = dir + static_cast<int64_t>(i) * (W * 32);
                if (i < band_start) {
                    if (shortrow)
                        dp_row<KC, KG, DEG, MV, true, false>(s, lc, x, Q, lane, dir_row);
                    else
                        dp_row<KC, KG, DEG, MV, false, false>(s, lc, x, Q, lane, dir_row);
                } else {
                    if (shortrow)
                        dp_row<KC, KG, DEG, MV, true, true>(s, lc, x, Q, lane, dir_row);
                    else
                        dp_row<KC, KG, DEG, MV, false, true>(s, lc, x, Q, lane, dir_row);
                }
            }
            __syncwarp();                          // every lane is done with this tile
            if (c + 2 < nchunks) issue(c + 2);     // refill it
        }

        if (p.end_cost) {
            const int ep = A->end_pos;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (ep == lane * K + k) p.end_cost[m.read] = s.D[k];
        }
        if (lane == 0) p.status[m.read] = WSTR_READ_OK;
    }
}

// ------------------------------------------------------------------------------------------
// Traceback: follow the stored direction codes from (T-1, endstate) to row 0.  The
// reference re-derives each step by recomputing the candidates and picking the one that
// reproduces the stored cost (caller.py:254-299); the winner of the fill reproduces it
// exactly, so following the fill's arg-min with the same priority (stay, then incoming in
// list order) visits the same cells.  One thread per read.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) traceback_kernel(const TraceParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n) return;
    const ReadMeta m = p.meta[idx];
    if (p.status[m.read] != WSTR_READ_OK) return;
    const DevAutomaton *A = p.auts + m.aut;
    const int K = A->K, W = A->W, mv = A->mv;
    const int16_t *sop = A->state_of_pos;
    const int16_t *pred = A->pred_tab;
    const uint32_t *dir = p.dir + m.dir_off;
    const uint32_t *mw = p.maskbits ? p.maskbits + m.mask_off : nullptr;
    int32_t *tr = p.trace + m.sig_off;

    int i = m.T - 1;
    int pos = A->end_pos;
    int st = sop[pos];
    while (i > 0) {
        const int lane = pos / K, k = pos - lane * K;
        uint32_t code = 0u;
        if (i >= mv) {   // rows 1..mv-1 are never filled: the reference keeps them at +inf
            const uint32_t word = dir[(static_cast<int64_t>(i) * W + (k >> 3)) * 32 + lane];
            code = (word >> (4 * (k & 7))) & 15u;
        }
        tr[i] = st;
        if (code == 0u) {
            i -= 1;
            continue;
        }
        int back = mv;
        if (mw && ((mw[i >> 5] >> (i & 31)) & 1u)) back = mv - 1;
        const int ppos = pred[pos * WSTR_PRED_STRIDE + static_cast<int>(code)];
        if (i < back || ppos < 0) {
            p.status[m.read] = WSTR_READ_BACKTRACK;
            return;
        }
        const int pst = sop[ppos];
        for (int r = 1; r < back; ++r) tr[i - r] = pst;
        i -= back;
        pos = ppos;
        st = pst;
    }
    tr[0] = st;
}

template <int KC, int KG, int DEG, int MV>
int launch_fill_t(const FillParams &p, cudaStream_t s) {
    static int grid_cap = 0;
    const int smem = static_cast<int>(sizeof(FillSmem<KG>)) * WSTR_WARPS_PER_CTA;
    if (grid_cap == 0) {
        WSTR_CUDA(cudaFuncSetAttribute(dtw_fill_kernel<KC, KG, DEG, MV>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int dev = 0, sms = 0, per_sm = 0;
        WSTR_CUDA(cudaGetDevice(&dev));
        WSTR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        WSTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_fill_kernel<KC, KG, DEG, MV>,
                                                                32 * WSTR_WARPS_PER_CTA, smem));
        if (per_sm < 1) per_sm = 1;
        grid_cap = sms * per_sm;
    }
    int grid = (p.n + WSTR_WARPS_PER_CTA - 1) / WSTR_WARPS_PER_CTA;
    if (grid > grid_cap) grid = grid_cap;
    if (grid < 1) return WSTR_OK;
    dtw_fill_kernel<KC, KG, DEG, MV><<<grid, 32 * WSTR_WARPS_PER_CTA, smem, s>>>(p);
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}

template <int KC, int KG, int MV>
int launch_fill_deg(int deg, const FillParams &p, cudaStream_t s) {
    if (deg <= 2) return launch_fill_t<KC, KG, 2, MV>(p, s);
    return launch_fill_t<KC, KG, 4, MV>(p, s);
}

}  // namespace

// the (chain, generic) slot splits the library is built with; keep in sync with kSplits in api.cu
int wstr_launch_fill(int kc, int kg, int deg, int mv, const FillParams &p, cudaStream_t s) {
    if (deg > 4) return WSTR_ERR_UNSUPPORTED;
#define WSTR_CASE(KC_, KG_, MV_) \
    if (kc == KC_ && kg == KG_ && mv == MV_) return launch_fill_deg<KC_, KG_, MV_>(deg, p, s);
    WSTR_CASE(6, 2, 4)
    WSTR_CASE(6, 2, 3)
    WSTR_CASE(6, 2, 5)
    WSTR_CASE(4, 4, 4)
    WSTR_CASE(8, 2, 4)
    WSTR_CASE(6, 4, 4)
    WSTR_CASE(8, 4, 4)
    WSTR_CASE(12, 4, 4)
#undef WSTR_CASE
    return WSTR_ERR_UNSUPPORTED;
}

int wstr_launch_traceback(const TraceParams &p, cudaStream_t s) {
    if (p.n <= 0) return WSTR_OK;
    const int threads = 128;
    traceback_kernel<<<(p.n + threads - 1) / threads, threads, 0, s>>>(p);
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}
