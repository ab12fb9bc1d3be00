// DTW state-automaton fill + traceback for sm_100a.
//
// Replaces WarpSTR._calc_dtw_astates (reference src/caller/caller.py:198-245) and
// WarpSTR._backtracking (:247-301).
//
// One warp owns one read.  The automaton's states live in registers, K consecutive
// positions per lane (position p = lane*K + slot); a position is "chained" when the edge
// (p-1 -> p) exists, which the host-side layout arranges for almost every state.  Row i
// of the DP needs only row i-1 (stay) and row i-back (skip through an incoming state), so
// per position the warp carries D[i-1] and the running partial sums
//     P_m[t] = D[t] + |x[t+1]-v| + ... + |x[t+m]-v|        m = 1..mv-1
// in the reference's own left-to-right addition order; the skip candidate into state j
// over incoming state p is P_{back-1}[i-back][p] + |x[i]-v_j|, bit-identical to the
// reference's nested loop.  Chained candidates come from the neighbouring register (one
// 64-bit shuffle per row at the lane boundary), the few remaining edges (loop back-edges,
// optional-group skips, IUPAC fan-in) through a double-buffered shared-memory row.
// All arithmetic is FP64 add/abs/compare: no multiply, so no FMA contraction can occur.
//
// The read's signal is streamed through shared memory in 2 KB tiles with 1-D bulk async
// copies (cp.async.bulk + mbarrier, TMA engine) two tiles ahead of the row loop.
// Output: one 4-bit direction code per cell (0 stay, 1 chain, 2+r = r-th extra edge of
// the position), packed 8 per 32-bit word, stored row-major [row][word][lane] so that every
// row is one coalesced 128-byte line per word.  D itself never leaves the SM.
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

// the pipeline value a position offers to its successors in this row
template <int K, int MV, bool SHORT>
__device__ __forceinline__ double offer(const LaneState<K, MV> &s, int k) {
    if (SHORT) {
        if (MV >= 3) return s.P[MV >= 3 ? MV - 3 : 0][k];
        return s.D[k];
    }
    return s.P[MV - 2][k];
}

struct LaneTables {
    uint32_t chain_bits, band_bits, src_bits;       // per lane, K bits each
    uint32_t allchain_slots, extra_slots, src_slots;  // warp-uniform
};

// One DP row.  SHORT: this row allows dwell mv-1 (masked row of the second pass).
// BAND: the end band is active (caller.py:223-224).
template <int K, int MV, bool SHORT, bool BAND>
__device__ __forceinline__ void dp_row(LaneState<K, MV> &s, const LaneTables &lt, const double x,
                                       double *__restrict__ Qs, const uint16_t *__restrict__ xtab,
                                       const uint8_t *__restrict__ xoff, const int lane,
                                       uint32_t *__restrict__ dir_row) {
    constexpr int W = (K + 7) / 8;
    const double INF = dinf();

    if (lt.src_slots) {   // publish the values that non-chained edges read
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if ((lt.src_slots >> k) & 1u) {
                if ((lt.src_bits >> k) & 1u) Qs[lane * K + k] = offer<K, MV, SHORT>(s, k);
            }
        }
        __syncwarp();
    }
    double qprev = __shfl_up_sync(FULL, offer<K, MV, SHORT>(s, K - 1), 1);

    uint32_t codes[W];
#pragma unroll
    for (int w = 0; w < W; ++w) codes[w] = 0u;

#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double ae = fabs(x - s.v[k]);
        const double qhere = offer<K, MV, SHORT>(s, k);
        const double stay = s.D[k] + ae;
        double ch = qprev + ae;
        if (!((lt.allchain_slots >> k) & 1u)) {
            if (!((lt.chain_bits >> k) & 1u)) ch = INF;
        }
        double best = stay;
        uint32_t code = 0u;
        if (ch < best) {
            best = ch;
            code = 1u;
        }
        if ((lt.extra_slots >> k) & 1u) {
            const int r0 = xoff[k], r1 = xoff[k + 1];
            for (int r = r0; r < r1; ++r) {
                const uint32_t ent = xtab[r * 32 + lane];
                if (ent != WSTR_NO_EDGE) {
                    const double c = Qs[ent & 0x7fffu] + ae;
                    // an edge listed before the chain edge in the reference's incoming order
                    // beats it on ties; everything else needs a strictly smaller cost
                    const bool tie = (ent >> 15) && code == 1u && c == best;
                    if (c < best || tie) {
                        best = c;
                        code = 2u + static_cast<uint32_t>(r - r0);
                    }
                }
            }
        }
        if (BAND) {
            if ((lt.band_bits >> k) & 1u) {
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
#pragma unroll
    for (int w = 0; w < W; ++w) dir_row[w * 32 + lane] = codes[w];
}

template <int K>
struct FillSmem {
    double sig[2][CH];
    double Qs[2][32 * K];
    uint16_t xtab[WSTR_XTAB_MAX_ROWS * 32];
    uint64_t bar[2];
    uint8_t xoff[32];
};

template <int K>
constexpr int fill_min_blocks() { return K <= 8 ? 4 : (K <= 12 ? 3 : 2); }

template <int K, int MV>
__global__ void __launch_bounds__(32 * WSTR_WARPS_PER_CTA, fill_min_blocks<K>())
dtw_fill_kernel(const FillParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int W = (K + 7) / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FillSmem<K> &sm = reinterpret_cast<FillSmem<K> *>(smem_raw)[warp];
    const double INF = dinf();

    if (lane == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        fence_barrier_init();
    }
    __syncwarp();
    uint32_t uses0 = 0, uses1 = 0;   // completed phases of the two tile barriers

    LaneState<K, MV> s;
    LaneTables lt;
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

        if (m.aut != cached_aut) {   // (re)load the automaton into registers / shared memory
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) s.v[k] = __ldg(A->v_pos + lane * K + k);
            lt.chain_bits = __ldg(A->lane_bits + lane);
            lt.band_bits = __ldg(A->lane_bits + 32 + lane);
            lt.src_bits = __ldg(A->lane_bits + 64 + lane);
            lt.allchain_slots = A->allchain_slots;
            lt.extra_slots = A->extra_slots;
            lt.src_slots = A->src_slots;
            const int nx = A->n_xrows * 32;
            for (int e = lane; e < nx; e += 32) sm.xtab[e] = __ldg(A->xtab + e);
            if (lane <= K) sm.xoff[lane] = A->xoff[lane];
            v0 = __ldg(A->v_pos + A->init_pos[0]);
            cached_aut = m.aut;
            __syncwarp();
        }

        // ---- signal tiles: two in flight --------------------------------------------------
        const double *gsig = p.signal + m.sig_off;
        const int nchunks = (T + CH - 1) / CH;
        auto issue = [&](int c) {
            if (lane == 0) {
                int n = T - c * CH;
                n = n > CH ? CH : n;
                n = (n + 1) & ~1;   // 16-byte granules
                bulk_load(sm.sig[c & 1], gsig + static_cast<int64_t>(c) * CH,
                          static_cast<uint32_t>(n) * 8u, &sm.bar[c & 1]);
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
                double *Qs = sm.Qs[i & 1];
                uint32_t *dir_row = dir + static_cast<int64_t>(i) * (W * 32);
                if (i < band_start) {
                    if (shortrow)
                        dp_row<K, MV, true, false>(s, lt, x, Qs, sm.xtab, sm.xoff, lane, dir_row);
                    else
                        dp_row<K, MV, false, false>(s, lt, x, Qs, sm.xtab, sm.xoff, lane, dir_row);
                } else {
                    if (shortrow)
                        dp_row<K, MV, true, true>(s, lt, x, Qs, sm.xtab, sm.xoff, lane, dir_row);
                    else
                        dp_row<K, MV, false, true>(s, lt, x, Qs, sm.xtab, sm.xoff, lane, dir_row);
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
        if (i < back) {
            p.status[m.read] = WSTR_READ_BACKTRACK;
            return;
        }
        const int ppos = code == 1u ? pos - 1
                                    : (A->xtab[(A->xoff[k] + static_cast<int>(code) - 2) * 32 + lane] & 0x7fff);
        const int pst = sop[ppos];
        for (int r = 1; r < back; ++r) tr[i - r] = pst;
        i -= back;
        pos = ppos;
        st = pst;
    }
    tr[0] = st;
}

template <int K, int MV>
int launch_fill_t(const FillParams &p, cudaStream_t s) {
    static int grid_cap = 0;
    const int smem = static_cast<int>(sizeof(FillSmem<K>)) * WSTR_WARPS_PER_CTA;
    if (grid_cap == 0) {
        WSTR_CUDA(cudaFuncSetAttribute(dtw_fill_kernel<K, MV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int dev = 0, sms = 0, per_sm = 0;
        WSTR_CUDA(cudaGetDevice(&dev));
        WSTR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        WSTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_fill_kernel<K, MV>,
                                                                32 * WSTR_WARPS_PER_CTA, smem));
        if (per_sm < 1) per_sm = 1;
        grid_cap = sms * per_sm;
    }
    int grid = (p.n + WSTR_WARPS_PER_CTA - 1) / WSTR_WARPS_PER_CTA;
    if (grid > grid_cap) grid = grid_cap;
    if (grid < 1) return WSTR_OK;
    dtw_fill_kernel<K, MV><<<grid, 32 * WSTR_WARPS_PER_CTA, smem, s>>>(p);
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}

template <int MV>
int launch_fill_k(int K, const FillParams &p, cudaStream_t s) {
    switch (K) {
        case 4: return launch_fill_t<4, MV>(p, s);
        case 8: return launch_fill_t<8, MV>(p, s);
        case 9: return launch_fill_t<9, MV>(p, s);
        case 10: return launch_fill_t<10, MV>(p, s);
        case 12: return launch_fill_t<12, MV>(p, s);
        case 16: return launch_fill_t<16, MV>(p, s);
        default: return WSTR_ERR_TOO_MANY_STATES;
    }
}

}  // namespace

int wstr_fill_smem_bytes(int K) {
    switch (K) {
        case 4: return sizeof(FillSmem<4>) * WSTR_WARPS_PER_CTA;
        case 8: return sizeof(FillSmem<8>) * WSTR_WARPS_PER_CTA;
        case 9: return sizeof(FillSmem<9>) * WSTR_WARPS_PER_CTA;
        case 10: return sizeof(FillSmem<10>) * WSTR_WARPS_PER_CTA;
        case 12: return sizeof(FillSmem<12>) * WSTR_WARPS_PER_CTA;
        case 16: return sizeof(FillSmem<16>) * WSTR_WARPS_PER_CTA;
        default: return -1;
    }
}

int wstr_launch_fill(int K, int mv, const FillParams &p, cudaStream_t s) {
    switch (mv) {
        case 4: return launch_fill_k<4>(K, p, s);
        case 3: return launch_fill_k<3>(K, p, s);
        case 5: return launch_fill_k<5>(K, p, s);
        default: return WSTR_ERR_UNSUPPORTED;
    }
}

int wstr_launch_traceback(const TraceParams &p, cudaStream_t s) {
    if (p.n <= 0) return WSTR_OK;
    const int threads = 128;
    traceback_kernel<<<(p.n + threads - 1) / threads, threads, 0, s>>>(p);
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}
