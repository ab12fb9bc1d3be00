// Catch-all DTW fill + traceback: every configuration the reference accepts.
//
// Same arithmetic as dtw.cu (WarpSTR._calc_dtw_astates, reference src/caller/caller.py:198-245,
// and _backtracking, :247-301), but nothing is a compile-time constant: any
// min_values_per_state > 1 (src/config.py:115), any in-degree (up to 254), any number of states
// the shared memory holds (mv * S <= ~20 000).  The specialised kernels keep a lane's states and
// their pipelines in registers, which fixes (chain slots, generic slots, in-degree, mv) per
// instantiation; automata or settings without an instantiation run here instead of being refused.
//
// One CTA of 128 threads per read, thread t owns states t, t+128, ...  Per state the CTA keeps in
// shared memory a ring of mv running sums: slot (t mod mv) holds
//     D[t][j] + |x[t+1]-v_j| + ... + |x[i-1]-v_j|
// for the mv most recent rows t, built left to right exactly like the reference's inner loop
// (caller.py:231-244), so that at row i the slot of row i-back is the skip candidate's
// P_{back-1}[i-back][p] and the slot of row i-1, once |x_i - v_j| is added, the stay candidate.
// A state's ring is only ever touched by its owner; what other states read (the slots of rows
// i-mv and i-mv+1 of every state) is copied to a double-buffered published row at the end of the
// previous row: one __syncthreads per DP row.
//
// Direction codes: one byte per cell (0 = stay, 1 + r = incoming edge r took the lead last), four
// consecutive rows of one state per 32-bit word, [word row][state] so that the CTA's stores are
// coalesced.  The traceback is done by warp 0 right after the fill: lane l looks at row i-l of
// the path's state, one ballot finds the first row that leaves it.
#include "wstr_internal.h"

namespace {

constexpr int NT = WSTR_ANY_THREADS;
constexpr int TILE = WSTR_ANY_TILE;
constexpr unsigned FULL = 0xffffffffu;
static_assert(TILE == NT, "every thread loads one sample of a signal tile");

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

struct AnySmem {
    double *ring;      // [mv][spad]
    double *pub;       // [2][2][spad]   (row parity, 0: slot of row i-mv / 1: slot of row i-mv+1)
    double *v;         // [spad]
    double *sx;        // [2][TILE]
    uint32_t *cw;      // [spad] direction bytes of the current word row
    uint8_t *band;     // [spad] state is inside the end band's skipped set
};

__host__ __device__ inline size_t any_smem_bytes(int mv, int spad) {
    return sizeof(double) * ((size_t)mv * spad + 4 * (size_t)spad + spad + 2 * TILE) + sizeof(uint32_t) * (size_t)spad +
           (size_t)spad + 16;
}

__device__ __forceinline__ AnySmem carve(unsigned char *raw, int mv, int spad) {
    AnySmem s;
    s.ring = reinterpret_cast<double *>(raw);
    s.pub = s.ring + (size_t)mv * spad;
    s.v = s.pub + 4 * (size_t)spad;
    s.sx = s.v + spad;
    s.cw = reinterpret_cast<uint32_t *>(s.sx + 2 * TILE);
    s.band = reinterpret_cast<uint8_t *>(s.cw + spad);
    return s;
}

__device__ void traceback_any(const DevAutomaton *A, const int T, const int mv, const uint32_t *dir,
                              const uint32_t *mw, int32_t *tr, int32_t *status_slot, const int lane) {
    const int spad = A->spad;
    int i = T - 1;
    int j = A->endstate;
    bool failed = false;
    while (i > 0) {
        const int row = i - lane;
        uint32_t code = 0u, mb = 0u;
        if (row >= mv) {      // rows below mv were never coded (and never leave their state)
            const uint32_t w = dir[(int64_t)(row >> 2) * spad + j];
            code = (w >> (8 * (row & 3))) & 0xffu;
            if (mw) mb = (__ldg(mw + (row >> 5)) >> (row & 31)) & 1u;
        }
        const unsigned moves = __ballot_sync(FULL, code != 0u);
        if (moves == 0u) {                       // 32 stays
            if (row >= 1) tr[row] = j;
            i -= 32;
            continue;
        }
        const int tm = __ffs(moves) - 1;         // lane of the first row that leaves the state
        const int rm = i - tm;
        const int codem = static_cast<int>(__shfl_sync(FULL, code, tm));
        const int back = mv - static_cast<int>(__shfl_sync(FULL, mb, tm));
        if (lane <= tm) tr[row] = j;             // the move row keeps the state it leaves
        const int p = __ldg(A->in_idx + __ldg(A->in_ptr + j) + codem - 1);
        if (rm - back < 0) {
            failed = true;
            break;
        }
        for (int q = lane; q < back - 1; q += 32) tr[rm - 1 - q] = p;   // the skipped rows belong to the predecessor
        i = rm - back;
        j = p;
    }
    if (lane == 0) {
        if (failed) *status_slot = WSTR_READ_BACKTRACK;
        else tr[0] = j;
    }
}

__global__ void __launch_bounds__(NT) dtw_fill_any_kernel(const FillParams p, const int mv, const int spad_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_next;
    const AnySmem sm = carve(smem_raw, mv, spad_max);
    const int tid = threadIdx.x;
    const double INF = dinf();
    int cached_aut = -1;

    for (;;) {
        __syncthreads();       // the previous read (its traceback included) is done with shared memory
        if (tid == 0) s_next = atomicAdd(p.queue, 1);
        __syncthreads();
        const int r = s_next;
        if (r >= p.n) break;
        const ReadMeta m = p.meta[p.order[r]];
        const DevAutomaton *A = p.auts + m.aut;
        const int T = m.T;
        if (p.respect_status && p.status[m.read] != WSTR_READ_OK) continue;   // failed earlier in the call
        if (T <= mv) {
            if (tid == 0) p.status[m.read] = WSTR_READ_TOO_SHORT;
            continue;
        }
        const int S = A->S;
        const int spad = A->spad;
        const int32_t *__restrict__ in_ptr = A->in_ptr;
        const int32_t *__restrict__ in_idx = A->in_idx;
        if (m.aut != cached_aut) {
            const int after = A->after;
            for (int j = tid; j < S; j += NT) {
                sm.v[j] = __ldg(A->values + j);
                sm.band[j] = __ldg(A->seq_idx + j) < after ? 1 : 0;
            }
            cached_aut = m.aut;
        }
        const double *__restrict__ gx = p.signal + m.sig_off;
        const uint32_t *mw = p.maskbits ? p.maskbits + m.mask_off : nullptr;
        uint32_t *dir = p.dir + m.dir_off;
        const int band_start = max(A->th1, T - A->band6 + 1);
        const int ntiles = (T + TILE - 1) / TILE;

        sm.sx[tid] = tid < T ? gx[tid] : 0.0;                      // tile 0 (TILE == NT)
        __syncthreads();

        // ---- row 0 (caller.py:201-208) and the empty rows 1..mv-1 --------------------------------
        {
            const double v0 = sm.v[0];
            const double first = fabs(gx[0] - v0);
            const int par = mv & 1;
            for (int j = tid; j < S; j += NT) {
                double acc = INF;
                if (j == 0) acc = first;
                else if (j <= mv) acc = first + fabs(gx[j] - v0);     // row 0, column j
                const double vj = sm.v[j];
                for (int t = 1; t <= mv - 1; ++t) acc = acc + fabs(gx[t] - vj);
                sm.ring[j] = acc;                                     // slot of row 0
                for (int t = 1; t < mv; ++t) sm.ring[(size_t)t * spad_max + j] = INF;
                sm.pub[(size_t)(par * 2) * spad_max + j] = acc;       // row mv reads rows 0 and 1
                sm.pub[(size_t)(par * 2 + 1) * spad_max + j] = INF;
                sm.cw[j] = 0u;
            }
        }
        __syncthreads();

        int slot = 0;            // i mod mv (row mv: 0)
        for (int c = 0; c < ntiles; ++c) {
            const int i_begin = c == 0 ? mv : c * TILE;
            const int i_end = min(T, (c + 1) * TILE);
            const int nt = (c + 1) * TILE + tid;
            const double nx = (c + 1 < ntiles && nt < T) ? gx[nt] : 0.0;   // next tile, a whole tile ahead
            const double *xs = sm.sx + (c & 1) * TILE - c * TILE;
            for (int i = i_begin; i < i_end; ++i) {
                const double x = xs[i];
                const uint32_t mb = mw ? (__ldg(mw + (i >> 5)) >> (i & 31)) & 1u : 0u;
                const bool banded = i >= band_start;
                const int slot_prev = slot == 0 ? mv - 1 : slot - 1;      // row i-1
                const int slot_n1 = slot + 1 == mv ? 0 : slot + 1;        // row i+1-mv
                const int slot_n2 = slot_n1 + 1 == mv ? 0 : slot_n1 + 1;  // row i+2-mv
                const double *pb = sm.pub + (size_t)((i & 1) * 2 + mb) * spad_max;
                double *pn = sm.pub + (size_t)(((i + 1) & 1) * 2) * spad_max;
                const int sh = 8 * (i & 3);
                const bool flush = (i & 3) == 3 || i == T - 1;
                for (int j = tid; j < S; j += NT) {
                    const double e = fabs(x - sm.v[j]);
                    // every running sum of this state takes the row's emission, the oldest is replaced below
                    for (int s = 0; s < mv; ++s)
                        if (s != slot) sm.ring[(size_t)s * spad_max + j] += e;
                    double best = INF;
                    uint32_t code = 0u;
                    if (!(banded && sm.band[j])) {                        // end band: the cell stays +inf (:223-224)
                        const double stay = sm.ring[(size_t)slot_prev * spad_max + j];
                        if (stay < best) best = stay;
                        const int e0 = __ldg(in_ptr + j), e1 = __ldg(in_ptr + j + 1);
                        for (int q = e0; q < e1; ++q) {
                            const double cnd = pb[__ldg(in_idx + q)] + e;
                            if (cnd < best) {
                                best = cnd;
                                code = static_cast<uint32_t>(q - e0 + 1);
                            }
                        }
                    }
                    sm.ring[(size_t)slot * spad_max + j] = best;
                    pn[j] = sm.ring[(size_t)slot_n1 * spad_max + j];
                    if (mw) pn[spad_max + j] = sm.ring[(size_t)slot_n2 * spad_max + j];
                    uint32_t w = sm.cw[j] | (code << sh);
                    if (flush) {
                        dir[(int64_t)(i >> 2) * spad + j] = w;
                        w = 0u;
                    }
                    sm.cw[j] = w;
                }
                slot = slot_n1;
                __syncthreads();
            }
            sm.sx[((c + 1) & 1) * TILE + tid] = nx;
            __syncthreads();
        }

        if (tid == 0) {
            if (p.end_cost) {
                const int last = (T - 1) % mv;
                p.end_cost[m.read] = sm.ring[(size_t)last * spad_max + A->endstate];
            }
            p.status[m.read] = WSTR_READ_OK;
        }
        __syncthreads();     // direction words of all threads are visible to warp 0
        if (tid < 32) traceback_any(A, T, mv, dir, mw, p.trace + m.sig_off, p.status + m.read, tid);
    }
}

}  // namespace

size_t wstr_any_smem_bytes(int mv, int spad_max) { return any_smem_bytes(mv, spad_max); }

int wstr_launch_fill_any(int mv, int spad_max, const FillParams &p, cudaStream_t s) {
    static int sms = 0;
    static unsigned long long attr_done = 0ull;      // devices the function attributes are set on
    int dev = 0;
    WSTR_CUDA(cudaGetDevice(&dev));
    if (sms == 0 || !((attr_done >> (dev & 63)) & 1ull)) {
        WSTR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        WSTR_CUDA(cudaFuncSetAttribute(dtw_fill_any_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       WSTR_ANY_SMEM_MAX));
        WSTR_CUDA(cudaFuncSetAttribute(dtw_fill_any_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));
        attr_done |= 1ull << (dev & 63);
    }
    const size_t smem = any_smem_bytes(mv, spad_max);
    if (smem > WSTR_ANY_SMEM_MAX) return WSTR_ERR_TOO_MANY_STATES;
    int per_sm = 0;
    WSTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_fill_any_kernel, NT, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = p.n < sms * per_sm ? p.n : sms * per_sm;
    if (grid < 1) return WSTR_OK;
    dtw_fill_any_kernel<<<grid, NT, smem, s>>>(p, mv, spad_max);
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}
