// Kernels (1) expected-signal lookup and (2) per-read normalisation, plus the FP64-add
// rate probe used as the measured roofline denominator.
//
// (1) replaces PoreModel.get_value / Squiggler._generate_signal
//     (reference src/squiggler/pore_model.py:45-47, src/squiggler/Squiggler.py:20-28).
// (2) replaces Fast5.remove_spikes + normalize_signal_mad + the window slice
//     (reference src/schemas/fast5.py:45-57, 68-77, 90-114).  Exact: the percentiles and
//     the MAD are order statistics, obtained from a histogram of the int16 samples rather
//     than from a sort, and the float formulas follow numpy's (numpy 2.3
//     lib/_function_base_impl.py: 'linear' virtual index (n-1)*q, _get_indexes, _lerp,
//     median = mean of the middle one/two).  Compiled with -fmad=false.
#include <limits.h>
#include <math.h>

#include <string.h>

#include "wstr_internal.h"
#include <type_traits>

// ------------------------------------------------------------------------------------------
// (1) pore-model lookup
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t base_code(uint8_t c, bool &ok) {
    switch (c) {
        case 'A': return 0u;
        case 'C': return 1u;
        case 'G': return 2u;
        case 'T': return 3u;
        default: ok = false; return 0u;
    }
}

// Every thread produces VPT consecutive levels from one rolling 2-bit-packed k-mer index
// (VPT + k - 1 byte loads instead of VPT * k) and writes them with 16-byte stores when the
// output is aligned for it.
constexpr int PORE_VPT = 4;
__global__ void pore_lookup_kernel(const uint8_t *__restrict__ seq, int64_t n_out, const double *__restrict__ table,
                                   int k, double *__restrict__ out, int32_t *bad, int vec_ok) {
    const uint32_t mask = k >= 16 ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * PORE_VPT; i < n_out;
         i += (int64_t)gridDim.x * blockDim.x * PORE_VPT) {
        uint32_t idx = 0;
        int valid = 0;                       // consecutive valid bases ending at the current one
        for (int p = 0; p < k - 1; ++p) {
            bool ok = true;
            const uint32_t c = base_code(seq[i + p], ok);
            idx = (idx << 2) | c;
            valid = ok ? valid + 1 : 0;
        }
        double v[PORE_VPT];
        int n_bad = 0;
#pragma unroll
        for (int j = 0; j < PORE_VPT; ++j) {
            v[j] = nan;
            if (i + j < n_out) {
                bool ok = true;
                const uint32_t c = base_code(seq[i + j + k - 1], ok);
                idx = ((idx << 2) | c) & mask;
                valid = ok ? valid + 1 : 0;
                if (valid >= k) v[j] = __ldg(table + idx);
                else ++n_bad;
            }
        }
        if (vec_ok && i + PORE_VPT <= n_out) {
            reinterpret_cast<double2 *>(out + i)[0] = make_double2(v[0], v[1]);
            reinterpret_cast<double2 *>(out + i)[1] = make_double2(v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < PORE_VPT; ++j)
                if (i + j < n_out) out[i + j] = v[j];
        }
        if (n_bad && bad) atomicAdd(bad, n_bad);
    }
}

// (2 bits, valid) of an upper-case base without branches: A 0, C 1, G 2, T 3
__device__ __forceinline__ uint32_t base_code2(uint32_t c, bool &ok) {
    ok = (c >> 5) == 2u && ((0x0010008Au >> (c & 31u)) & 1u);   // bits 1, 3, 7, 20 = 'A','C','G','T' - 64
    const uint32_t t = (c >> 1) & 3u;                            // A 0, C 1, G 3, T 2
    return t ^ (t >> 1);
}

// Long sequences (whole-panel expected signals): the table sits in shared memory -- a gather from L1
// costs a tag look-up per distinct line, up to 32 per warp instruction for random k-mers, which is what
// bounded the kernel above at 0.4 of the HBM rate -- every thread produces 8 consecutive levels from
// 8 + k - 1 bases fetched with two 8-byte loads, and writes them with four 16-byte stores.
// Needs k <= 6 (4^k doubles in shared memory), seq 8-byte and out 16-byte aligned.
constexpr int PORE_VPT2 = 8;
__global__ void __launch_bounds__(256) pore_lookup_smem_kernel(const uint8_t *__restrict__ seq, int64_t n_bases,
                                                               int64_t n_out, const double *__restrict__ table, int k,
                                                               int n_entries, double *__restrict__ out, int32_t *bad) {
    extern __shared__ double pore_tab[];
    for (int e = threadIdx.x; e < n_entries; e += blockDim.x) pore_tab[e] = table[e];
    __syncthreads();
    const uint32_t mask = (1u << (2 * k)) - 1u;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    int n_bad = 0;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * PORE_VPT2; i < n_out;
         i += (int64_t)gridDim.x * blockDim.x * PORE_VPT2) {
        // bases i .. i+15: 8 of this thread's own and up to k-1 <= 7 of look-ahead
        uint64_t w0 = 0, w1 = 0;
        if (i + 16 <= n_bases) {
            w0 = *reinterpret_cast<const uint64_t *>(seq + i);
            w1 = *reinterpret_cast<const uint64_t *>(seq + i + 8);
        } else {
            for (int b = 0; b < 8; ++b) {
                if (i + b < n_bases) w0 |= (uint64_t)seq[i + b] << (8 * b);
                if (i + 8 + b < n_bases) w1 |= (uint64_t)seq[i + 8 + b] << (8 * b);
            }
        }
        uint32_t idx = 0;
        int valid = 0;                       // consecutive valid bases ending at the current one
#pragma unroll
        for (int p = 0; p < 7; ++p) {
            if (p < k - 1) {
                bool ok;
                const uint32_t c = base_code2((uint32_t)(w0 >> (8 * p)) & 0xffu, ok);
                idx = (idx << 2) | c;
                valid = ok ? valid + 1 : 0;
            }
        }
        double v[PORE_VPT2];
#pragma unroll
        for (int j = 0; j < PORE_VPT2; ++j) {
            const int q = j + k - 1;         // position of the k-mer's last base relative to i (< 15)
            const uint32_t ch = q < 8 ? (uint32_t)(w0 >> (8 * q)) & 0xffu : (uint32_t)(w1 >> (8 * (q - 8))) & 0xffu;
            bool ok;
            const uint32_t c = base_code2(ch, ok);
            idx = ((idx << 2) | c) & mask;
            valid = ok ? valid + 1 : 0;
            v[j] = nan;
            if (i + j < n_out) {
                if (valid >= k) v[j] = pore_tab[idx];
                else ++n_bad;
            }
        }
        const int64_t wbase = i - (int64_t)(threadIdx.x & 31) * PORE_VPT2;   // first output of this warp's 256
        if (wbase + 32 * PORE_VPT2 <= n_out) {
            // The warp's 256 levels go through shared memory so that every store instruction writes 512
            // contiguous bytes: 16-byte stores straight from the registers land 64 bytes apart, half a sector
            // each.  Rows of 8 levels padded to 9 keep both the 8-byte writes and the reads off each other's banks.
            double *stg = pore_tab + n_entries + (threadIdx.x >> 5) * (32 * (PORE_VPT2 + 1));
            const int lane = threadIdx.x & 31;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < PORE_VPT2; ++j) stg[lane * (PORE_VPT2 + 1) + j] = v[j];
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < PORE_VPT2 / 2; ++jj) {
                const int e = jj * 64 + lane * 2;                 // levels e, e + 1 of the warp's 256
                const double *src = stg + (e >> 3) * (PORE_VPT2 + 1) + (e & 7);
                reinterpret_cast<double2 *>(out + wbase)[jj * 32 + lane] = make_double2(src[0], src[1]);
            }
        } else if (i + PORE_VPT2 <= n_out) {
#pragma unroll
            for (int j = 0; j < PORE_VPT2; j += 2)
                reinterpret_cast<double2 *>(out + i)[j / 2] = make_double2(v[j], v[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < PORE_VPT2; ++j)
                if (i + j < n_out) out[i + j] = v[j];
        }
    }
    if (n_bad && bad) atomicAdd(bad, n_bad);
}

extern "C" int wstr_pore_lookup(const uint8_t *d_seq, int64_t n, const double *d_table, int32_t k, double *d_out,
                                int32_t *d_bad, void *stream) {
    if (!d_seq || !d_table || !d_out || k < 1 || k > 15) return WSTR_ERR_INVALID_ARGUMENT;
    const int64_t n_out = n - k + 1;
    if (n_out <= 0) return WSTR_OK;
    const int threads = 256;
    if (n_out >= (1 << 18) && k <= 6 && (reinterpret_cast<uintptr_t>(d_seq) & 7u) == 0 &&
        (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0) {
        const int n_entries = 1 << (2 * k);
        int64_t blocks2 = (n_out + (int64_t)threads * PORE_VPT2 - 1) / ((int64_t)threads * PORE_VPT2);
        if (blocks2 > 148 * 4) blocks2 = 148 * 4;
        const size_t smem = (n_entries + (threads / 32) * 32 * (PORE_VPT2 + 1)) * sizeof(double);   // table + a staging tile per warp
        static unsigned long long attr_done = 0ull;      // devices the function attribute is set on
        int dev = 0;
        WSTR_CUDA(cudaGetDevice(&dev));
        if (!((attr_done >> (dev & 63)) & 1ull)) {
            WSTR_CUDA(cudaFuncSetAttribute(pore_lookup_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)((4096 + (threads / 32) * 32 * (PORE_VPT2 + 1)) * sizeof(double))));
            attr_done |= 1ull << (dev & 63);
        }
        wstr_prof_begin(3, static_cast<cudaStream_t>(stream));
        pore_lookup_smem_kernel<<<(int)blocks2, threads, smem, static_cast<cudaStream_t>(stream)>>>(
            d_seq, n, n_out, d_table, k, n_entries, d_out, d_bad);
        wstr_prof_end(static_cast<cudaStream_t>(stream));
        WSTR_CUDA(cudaGetLastError());
        return WSTR_OK;
    }
    int64_t blocks = (n_out + (int64_t)threads * PORE_VPT - 1) / ((int64_t)threads * PORE_VPT);
    if (blocks > 148 * 16) blocks = 148 * 16;
    const int vec_ok = (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
    wstr_prof_begin(3, static_cast<cudaStream_t>(stream));
    pore_lookup_kernel<<<(int)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(d_seq, n_out, d_table, k, d_out,
                                                                                       d_bad, vec_ok);
    wstr_prof_end(static_cast<cudaStream_t>(stream));
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}

// ------------------------------------------------------------------------------------------
// (2) normalisation
// ------------------------------------------------------------------------------------------
#ifdef WSTR_NORM_TIMING
__device__ unsigned long long g_norm_phase[8];
#define WSTR_NORM_T(i)                                                      \
    do {                                                                    \
        if (threadIdx.x == 0) {                                             \
            const long long now_ = clock64();                               \
            if ((i) > 0) atomicAdd(&g_norm_phase[i], (unsigned long long)(now_ - t_phase_)); \
            else if (t_phase_) atomicAdd(&g_norm_phase[7], (unsigned long long)(now_ - t_phase_)); \
            t_phase_ = now_;                                                \
        }                                                                   \
    } while (0)
extern "C" int wstr_debug_norm_phases(unsigned long long *out8, int reset) {
    if (reset) {
        unsigned long long z[8] = {0};
        return (int)cudaMemcpyToSymbol(g_norm_phase, z, sizeof(z));
    }
    return (int)cudaMemcpyFromSymbol(out8, g_norm_phase, 8 * sizeof(unsigned long long));
}
#else
#define WSTR_NORM_T(i)
#endif
namespace {

constexpr int NT = 256;              // threads per CTA
constexpr int PER = 8;               // samples per thread per tile
constexpr int TILE = NT * PER;       // 2048
constexpr int HBINS = 8192;          // shared histogram (general path) covers values [0, HBINS)
constexpr int WBINS = 512;           // window histogram (Brute / None): values [wlo, wlo + WBINS), one column per lane
constexpr int GBINS = 65536;         // global fallback histogram covers every int16
#ifndef WSTR_NORM_DEPTH
#define WSTR_NORM_DEPTH 3            // stages in flight per CTA
#endif
constexpr int NSTAGE = WSTR_NORM_DEPTH + 1;   // stages of shared memory the window scan's bulk copies rotate through
constexpr int WSTEP = 2;                      // tiles per stage (one wait and one release per WSTEP tiles)
constexpr int WCOLS = 16;                     // words per bin of the window histogram: a 16-bit counter per lane
constexpr int WIN_MAX_N = (65535 - 64) * 32;  // longest read whose per-lane counts cannot overflow them
constexpr int SPIKE_CAP = 256;       // out-of-range samples of one read the barrier-free path can hold

struct NormParams {
    const int16_t *raw;
    const int64_t *raw_off;          // device copy, n+1
    const int32_t *win_lo, *win_hi;  // device copies
    const int64_t *out_off;          // device copy
    double *out;
    double *shift_scale;             // may be NULL
    uint32_t *ghist;                 // gridDim.x * GBINS, zero on entry and on exit
    int32_t *queue;
    int32_t n_reads;
    int32_t spike_mode;
};

struct NormSmem {
    // general path: hist[v], v in [0, HBINS).  Window path: a 16-bit counter per bin and lane, hist[(v - wlo) *
    // WCOLS + lane / 2], halves by lane parity: no two lanes of a warp ever meet in a bank at different words.
    // corr[bin]: the samples Brute moved into (+) or out of (-) a bin, kept apart because a counter must not
    // borrow from its neighbour.
    alignas(16) uint32_t hist[WBINS * WCOLS];
    int32_t corr[WBINS];
    // Window path: NSTAGE tiles of raw samples, filled by bulk asynchronous copies (one thread issues,
    // every thread reads its eight samples back with one 16-byte load).  Tile path and the final
    // conversion: tile[HALO + t] = sample t of the current tile; tile[HALO-2], tile[HALO-1] = the two
    // (patched) samples before it.  HALO = 8 keeps the tile 16-byte aligned for vector access.
    union {
        alignas(128) int16_t stage[NSTAGE * WSTEP * TILE];
        alignas(16) int16_t tile[TILE + 16];
    };
    alignas(8) uint64_t full[NSTAGE];
    uint32_t readers[NSTAGE];        // warps that have taken their samples out of a stage
    uint32_t spike_bits[TILE / 32];  // Brute, tile path: out-of-range samples of the current tile (all zero between tiles)
    // Brute, barrier-free path: the out-of-range samples of the read, (index << 32 | slot) for sorting, and
    // the raw samples i-2 .. i+2 around each, fetched by the thread that found it
    unsigned long long spike_key[SPIKE_CAP];
    int16_t spike_w[SPIKE_CAP][5];
    int16_t spike_nv[SPIKE_CAP];     // window path: the samples' values after patching
    int32_t n_spikes;
    // the read being worked on and the next one, prepared by the last warp while warp 0 patches and ranks:
    // its queue slot, offsets, window, and (Brute / None) where its histogram window goes
    struct Meta {
        int64_t raw_off, out_off;
        int32_t r, N, lo, hi, wlo;
    } meta[2];
    int32_t below, above;            // window path: samples under / over the window
    int32_t win_ok, old_dirty;
    int32_t vmin, vmax;
    int32_t ghist_dirty;
    double shift, scale;
};
constexpr int HALO = 8;
#ifndef WSTR_NORM_BLOCKS
#define WSTR_NORM_BLOCKS 3          // resident CTAs per SM (75 KB of shared memory each)
#endif

__device__ __forceinline__ uint32_t hcount(const NormSmem &sm, const uint32_t *gh, int v) {
    if (v >= 0 && v < HBINS) return sm.hist[v];
    if (v < -32768 || v > 32767) return 0u;
    return sm.ghist_dirty ? gh[v + 32768] : 0u;
}

__device__ __forceinline__ int16_t median_small(int16_t *w, int n) {
    // insertion sort of <= 5 values, then numpy's median (mean of the middle two for even n,
    // truncated toward zero when stored back into the int16 array: fast5.py:100)
    for (int a = 1; a < n; ++a) {
        int16_t key = w[a];
        int b = a - 1;
        while (b >= 0 && w[b] > key) {
            w[b + 1] = w[b];
            --b;
        }
        w[b + 1] = key;
    }
    if (n & 1) return w[n / 2];
    const double mid = ((double)w[n / 2 - 1] + (double)w[n / 2]) / 2.0;
    return (int16_t)mid;
}

// warp 0: smallest value v >= vstart with  (count of samples <= v) > rank, scanning 32 bins a step
__device__ int value_at_rank(const NormSmem &sm, const uint32_t *gh, int64_t rank, int lane) {
    int64_t cum = 0;
    for (int base = sm.vmin; base <= sm.vmax; base += 32) {
        const int v = base + lane;
        uint32_t c = v <= sm.vmax ? hcount(sm, gh, v) : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, cum + inc > rank);
        if (hit) return base + (__ffs(hit) - 1);
        cum += __shfl_sync(0xffffffffu, inc, 31);
    }
    return sm.vmax;
}

// warp 0: the rank-th smallest |v - shift| over all samples (float64, as numpy computes it)
__device__ double absdev_at_rank(const NormSmem &sm, const uint32_t *gh, double shift, int64_t rank, int lane) {
    const int fl = (int)floor(shift);
    int64_t cum = 0;
    const int span = max(fl - sm.vmin, sm.vmax - (fl + 1)) + 1;
    for (int base = 0; base < span; base += 32) {
        const int m = base + lane;
        const int lo = fl - m, hi = fl + 1 + m;
        const uint32_t cl = hcount(sm, gh, lo), chh = hcount(sm, gh, hi);
        uint32_t inc = cl + chh;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, cum + inc > rank);
        if (hit) {
            const int src = __ffs(hit) - 1;
            const int64_t before = cum + __shfl_sync(0xffffffffu, (int64_t)inc - (cl + chh), src);
            const uint32_t scl = __shfl_sync(0xffffffffu, cl, src), sch = __shfl_sync(0xffffffffu, chh, src);
            const int mm = base + src;
            const double dl = fabs((double)(fl - mm) - shift), du = fabs((double)(fl + 1 + mm) - shift);
            // inside the pair the nearer value comes first
            const int64_t within = rank - before;
            if (dl <= du) return within < (int64_t)scl ? dl : du;
            return within < (int64_t)sch ? du : dl;
        }
        cum += __shfl_sync(0xffffffffu, inc, 31);
    }
    return 0.0;
}

// ---- order statistics off the prefix-summed histogram (no sample outside [0, HBINS)) ------------------
// After the scan the CTA turns hist[vmin..vmax] into inclusive prefix sums in place; every rank query is
// then a binary search by one thread instead of a walk over the bins by a warp.
__device__ __forceinline__ int64_t cum_le(const NormSmem &sm, int v) {     // samples <= v
    if (v < sm.vmin) return 0;
    if (v >= sm.vmax) return sm.hist[sm.vmax];
    return sm.hist[v];
}
__device__ int value_at_rank_cum(const NormSmem &sm, int64_t rank) {       // smallest v with cum(v) > rank
    int a = sm.vmin, b = sm.vmax;
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (cum_le(sm, mid) > rank) b = mid;
        else a = mid + 1;
    }
    return a;
}
__device__ double absdev_at_rank_cum(const NormSmem &sm, double shift, int64_t rank) {
    // values by |v - shift|: the pairs (fl - m, fl + 1 + m), m = 0, 1, ...; the first m pairs hold the
    // samples in [fl - m + 1, fl + m]
    const int fl = (int)floor(shift);
    const int span = max(fl - sm.vmin, sm.vmax - (fl + 1)) + 1;
    int a = 0, b = span;                                                   // smallest m with count(m + 1) > rank
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (cum_le(sm, fl + mid + 1) - cum_le(sm, fl - mid - 1) > rank) b = mid;
        else a = mid + 1;
    }
    const int mm = a;
    const int64_t before = cum_le(sm, fl + mm) - cum_le(sm, fl - mm);
    const int64_t scl = cum_le(sm, fl - mm) - cum_le(sm, fl - mm - 1);
    const int64_t sch = cum_le(sm, fl + 1 + mm) - cum_le(sm, fl + mm);
    const double dl = fabs((double)(fl - mm) - shift), du = fabs((double)(fl + 1 + mm) - shift);
    const int64_t within = rank - before;
    if (dl <= du) return within < scl ? dl : du;                           // inside the pair the nearer value first
    return within < sch ? du : dl;
}

__device__ __forceinline__ uint32_t nsmem_u32(const void *q) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(q));
}
__device__ __forceinline__ void nbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nsmem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool nbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(nsmem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// one thread: announce `bytes` on `bar` and start the bulk copy global -> shared (TMA engine; no register or L1
// staging: with three 75 KB CTAs on an SM the L1 left over is too small to hold a deep queue of 16-byte loads)
__device__ __forceinline__ void nbulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nsmem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     nsmem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(nsmem_u32(bar))
                 : "memory");
}

struct NormCtx {
    const int16_t *raw;
    int16_t *stash;
    uint32_t *gh;
    int N, lo, hi, mis, wlo, spike_mode, tid;
};

// One thread's eight samples of tile k of the barrier-free scan, no assumption made: the tile may hang over either
// end of the read or meet the output window, samples may be Brute's to patch (noted in the spike list with their
// neighbourhoods) or lie outside the histogram.  WIN: the per-lane window histogram, returns (samples under the
// window) | (samples over it) << 8; otherwise the value-indexed histogram with the global one behind it, returns
// (min & 0xffff) | max << 16 of the valid samples.
template <bool WIN>
__device__ __noinline__ int norm_step_general(NormSmem &sm, const NormCtx &c, const uint4 loaded, const int g0) {
    union {
        uint4 q;
        int16_t h[PER];
    } cur;
    cur.q = loaded;                                               // samples g0 .. g0 + 7 of the read
    const int N = c.N, lo = c.lo, hi = c.hi, wlo = c.wlo;
    const bool in_window = g0 <= hi && g0 + PER > lo;
    int tmin = 32767, tmax = -32768;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int g = g0 + u;
        if (g >= 0 && g < N) {
            tmin = min(tmin, (int)cur.h[u]);
            tmax = max(tmax, (int)cur.h[u]);
        }
    }
    if (c.spike_mode == 1 && (tmin < 250 || tmax > 1000)) {       // note where Brute will patch
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int g = g0 + u;
            if (g >= 0 && g < N && (cur.h[u] > 1000 || cur.h[u] < 250)) {
                const int pos = atomicAdd(&sm.n_spikes, 1);
                if (pos < SPIKE_CAP) {
                    if (WIN) {                                    // (the neighbourhoods are fetched after the scan)
                        reinterpret_cast<uint32_t *>(sm.spike_key)[pos] = (uint32_t)g;
                    } else {
                        sm.spike_key[pos] = ((unsigned long long)(unsigned)g << 32) | (unsigned)pos;
#pragma unroll
                        for (int d = 0; d < 5; ++d) {
                            const int idx = g - 2 + d;
                            sm.spike_w[pos][d] = (idx >= 0 && idx < N) ? c.raw[idx] : (int16_t)0;
                        }
                    }
                }
            }
        }
    }
    int ret;
    if (WIN) {
        uint32_t *const col = sm.hist + ((c.tid & 31) >> 1) - wlo * WCOLS;
        const uint32_t one = 1u << ((c.tid & 1) * 16);
        int below = 0, above = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int g = g0 + u;
            if (g < 0 || g >= N) continue;
            const int vv = cur.h[u];
            if (vv < wlo) ++below;
            else if (vv >= wlo + WBINS) ++above;
            else atomicAdd(col + vv * WCOLS, one);
        }
        ret = below | above << 8;
    } else {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int g = g0 + u;
            if (g < 0 || g >= N) continue;
            const int vv = cur.h[u];
            if (vv >= 0 && vv < HBINS) {
                atomicAdd(&sm.hist[vv], 1u);
            } else {
                atomicAdd(&c.gh[vv + 32768], 1u);
                sm.ghist_dirty = 1;
            }
        }
        ret = (tmin & 0xffff) | tmax << 16;
    }
    if (!WIN && in_window) {                                      // (the window form converts straight from the read)
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int g = g0 + u;
            if (g >= lo && g <= hi && g < N) c.stash[g - lo] = cur.h[u];
        }
    }
    return ret;
}

// One CTA per read (persistent: the CTAs draw reads from a queue), one pass over the samples.  The read is
// walked from the 16-byte boundary at or before its first sample (samples in front of / behind the read are
// masked).  Indices inside a read are 32-bit.
//
// Brute / None (the reference's default and its off switch) -- the window path.  A histogram is additive, so
// every thread counts its RAW samples and only notes where the ones Brute will patch are; the order statistics
// (two percentiles around the median, then the median absolute deviation) only ever look at values near the
// median, so the histogram covers a WBINS-value window placed around the median of 32 samples spread over the
// read and the samples outside it are merely counted (under / over).  What the scan costs is set by three things,
// each measured (profiles/r02_summary.md, clock64 phase counters under -DWSTR_NORM_TIMING):
//   * bytes in flight: 16-byte loads into registers are bounded by what is left of the L1 beside 3 x 72 KB of
//     shared memory, and a rotating register queue (a = b; b = load) waits for the newest load every step; the
//     samples therefore come by bulk asynchronous copy (TMA) into NSTAGE shared-memory stages of two tiles, a
//     stage refilled by whichever warp is last to have taken its samples out of it;
//   * instructions per sample: the common step (eight samples all inside the window, none to patch) is min/max
//     two samples at a time, two compares, eight `red.shared` on precomputed 32-bit addresses; everything
//     unusual is out of line or behind a rarely taken branch;
//   * what happens between scans, when one warp works and seven wait: the noted samples are sorted in registers,
//     their neighbourhoods fetched all at once, and patched lane-parallel (fast5.py:90-101: out[i] =
//     median(out[i-2:i+3]) in index order, so only samples within two of each other depend on one another);
//     the other warps meanwhile sum the 32 per-lane counters of every bin and the last one fetches the next
//     read's offsets and window estimate; then the prefix sums, and the rank queries answered by warp 0 with two
//     ballots each; the output window is asked for before the statistics and divided out of registers after them.
// A read whose statistics fall outside the window (or with more than SPIKE_CAP noted samples, or too long for
// the 16-bit counters) is redone on the general path: the value-indexed histogram of [0, HBINS) with a global
// one behind it.  The median-filter modes (every sample changes) take the tile loop: the tile in shared memory,
// one barrier per tile.
__global__ void __launch_bounds__(NT, WSTR_NORM_BLOCKS) normalize_kernel(const NormParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NormSmem &sm = *reinterpret_cast<NormSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *gh = p.ghist + (size_t)blockIdx.x * GBINS;
    uint32_t *const cum = reinterpret_cast<uint32_t *>(sm.tile);   // window path: prefix sums of the window's bins

    // the window path finds its histogram zero and leaves it zero
    for (int b = tid; b < WBINS * WCOLS; b += NT) sm.hist[b] = 0u;
    for (int b = tid; b < WBINS; b += NT) sm.corr[b] = 0;
    if (tid == 0) {
        sm.old_dirty = 0;
        for (int b = 0; b < NSTAGE; ++b) {
            nbar_init(&sm.full[b], 1);
            sm.readers[b] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t jstep = 0;              // window-scan steps so far (all reads): stage = jstep % NSTAGE, use = jstep / NSTAGE

    long long t_phase_ = 0;
    (void)t_phase_;
    const bool mode_window = p.spike_mode <= 1;
    // (last warp) the read in queue slot `slot` into meta[which]: offsets, window and -- for the window path --
    // the median of 32 samples spread over the read (each lane ranks its own), WBINS / 2 under which the
    // histogram window starts
    auto prepare = [&](int slot, int which) {
        NormSmem::Meta &m = sm.meta[which];
        if (lane == 0) m.r = slot;
        if (slot >= p.n_reads) return;
        const int64_t ro = p.raw_off[slot];
        // (a read has fewer than 2^31 - 2*TILE samples, checked by the host: tile-relative indices fit an int)
        const int n = (int)(p.raw_off[slot + 1] - ro);
        if (lane == 0) {
            m.raw_off = ro;
            m.out_off = p.out_off[slot];
            m.N = n;
            m.lo = p.win_lo[slot];
            m.hi = p.win_hi[slot];
        }
        if (mode_window) {
            const int mine = p.raw[ro + (((int64_t)lane * n) >> 5)];
            int rank = 0;
#pragma unroll
            for (int o = 0; o < 32; ++o) {
                const int other = __shfl_sync(0xffffffffu, mine, o);
                rank += (other < mine || (other == mine && o < lane)) ? 1 : 0;
            }
            if (rank == 16) m.wlo = mine - WBINS / 2;
        }
    };
    constexpr int PREP_WARP = NT / 32 - 1;
    if (warp == PREP_WARP) {
        int slot = lane == 0 ? atomicAdd(p.queue, 1) : 0;
        slot = __shfl_sync(0xffffffffu, slot, 0);
        prepare(slot, 0);
    }
    __syncthreads();
    for (int it = 0;; ++it) {
        const NormSmem::Meta &meta = sm.meta[it & 1];
        const int r = meta.r;
        WSTR_NORM_T(0);
        if (r >= p.n_reads) break;
        // the next read's queue slot: asked for now, looked at after the scan
        int next_slot = (tid == PREP_WARP * 32) ? atomicAdd(p.queue, 1) : 0;
        bool prepared = false;
        auto prepare_next = [&]() {                                  // (all of the last warp, once per read)
            if (warp == PREP_WARP && !prepared) {
                next_slot = __shfl_sync(0xffffffffu, next_slot, 0);
                prepare(next_slot, (it + 1) & 1);
                prepared = true;
            }
        };
        const int16_t *raw = p.raw + meta.raw_off;
        const int N = meta.N;
        const int lo = meta.lo, hi = meta.hi;
        const int Tw = hi >= lo ? max(min(hi, N - 1) - lo + 1, 0) : 0;   // numpy slice semantics
        double *out = p.out + meta.out_off;
        int16_t *stash = reinterpret_cast<int16_t *>(out) + 3 * (int64_t)Tw;   // tail of the output window

        const bool try_window = mode_window && N <= WIN_MAX_N;
        if (try_window && sm.old_dirty)                                  // (block-uniform: written before a barrier)
            for (int b = tid; b < HBINS; b += NT) sm.hist[b] = 0u;
        if (tid == 0) {
            sm.n_spikes = 0;
            sm.below = 0;
            sm.above = 0;
            sm.win_ok = 0;
            sm.ghist_dirty = 0;
        }
        __syncthreads();
        if (tid == 0 && try_window) sm.old_dirty = 0;

        // sample index g (0-based in the read) of tile k, thread tid, element u: k*TILE + tid*PER + u - mis
        const int mis = (int)((reinterpret_cast<uintptr_t>(raw) >> 1) & 7);
        const uint4 *vec = reinterpret_cast<const uint4 *>(raw - mis);
        const int n_tiles = (N + mis + TILE - 1) / TILE;
        const int n_vec = (N + mis + PER - 1) / PER;
        auto fetch = [&](int k) {
            const int v = k * NT + tid;
            return v < n_vec ? __ldg(vec + v) : make_uint4(0u, 0u, 0u, 0u);
        };
        int lmin = 32767, lmax = -32768;
        // np.percentile(data, q) (method 'linear') from a rank -> value function
        auto percentile_from = [&](auto &&rank_value, double q) {
            const double vidx = (double)(N - 1) * q;
            int64_t i0 = (int64_t)floor(vidx), i1 = i0 + 1;
            if (vidx >= (double)(N - 1)) i0 = i1 = N - 1;
            const double gamma = vidx - (double)i0;
            const int a = rank_value(i0);
            const int b = rank_value(i1);
            const int16_t diff16 = (int16_t)(b - a);              // numpy subtracts in int16
            const double diff = (double)diff16;
            double res = (double)a + diff * gamma;
            if (gamma >= 0.5) res = (double)b - diff * (1.0 - gamma);
            return res;
        };
        // ---- barrier-free scans (Brute / None) ------------------------------------------------------------
        NormCtx ctx;
        ctx.raw = raw;
        ctx.stash = stash;
        ctx.gh = gh;
        ctx.N = N;
        ctx.lo = lo;
        ctx.hi = hi;
        ctx.mis = mis;
        ctx.wlo = 0;
        ctx.spike_mode = p.spike_mode;
        ctx.tid = tid;
        int n_below = 0, n_above = 0;
        // Window form.  The loop body is what a step is for nearly every tile of nearly every read -- the
        // thread's eight samples all inside the window (and inside [250, 1000] for Brute), the tile whole and
        // not meeting the output window: min/max two samples per instruction, eight increments -- and is kept
        // that small on purpose; everything else is one out-of-line call.
        // tile k of this read is step j of the kernel: stage j % NSTAGE, its (j / NSTAGE)-th use
        const int n_steps = (n_tiles + WSTEP - 1) / WSTEP;       // a stage holds WSTEP tiles
        auto issue = [&](int k, uint32_t st) {                    // (one thread) step k of this read into stage st
            constexpr int SBYTES = WSTEP * TILE * 2;
            const int bytes = min(SBYTES, n_vec * 16 - k * SBYTES);
            nbulk_load(&sm.stage[st * (WSTEP * TILE)], reinterpret_cast<const char *>(vec) + (size_t)k * SBYTES,
                       (uint32_t)bytes, &sm.full[st]);
        };
        auto scan_window_prologue = [&]() {                       // (thread 0; every stage is free between reads)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stages were last written by threads
            for (int d = 0; d < NSTAGE; ++d)
                if (d < n_steps) issue(d, (jstep + d) % NSTAGE);
        };
        auto scan_window = [&]() {
            const int wlo = meta.wlo;
            ctx.wlo = wlo;
            // the word this lane's 16-bit counter of value v lives in is at col + v * (4 * WCOLS), `one` its unit
            const uint32_t col = nsmem_u32(sm.hist + (lane >> 1)) - (uint32_t)(wlo * (4 * WCOLS));
            const uint32_t one = 1u << ((lane & 1) * 16);
            // all eight samples inside the window and none of them Brute's to patch: the common step
            const int ok_lo = p.spike_mode == 1 ? max(wlo, 250) : wlo;
            const int ok_hi = p.spike_mode == 1 ? min(wlo + WBINS - 1, 1000) : wlo + WBINS - 1;
            for (int k = 0; k < n_steps; ++k) {
                const uint32_t j = jstep + k, st = j % NSTAGE, use = j / NSTAGE;
                while (!nbar_try_wait(&sm.full[st], use & 1u)) {
                }
                union {
                    uint4 q;
                    int16_t h[PER];
                } cur[WSTEP];
#pragma unroll
                for (int w = 0; w < WSTEP; ++w)
                    cur[w].q = *reinterpret_cast<const uint4 *>(&sm.stage[(st * WSTEP + w) * TILE + tid * PER]);
                // The last warp to have its samples in registers sends the stage off for step k + NSTAGE: no stage
                // waits for a particular thread to come by, NSTAGE - 1 copies are in flight whenever a warp waits.
                asm volatile("" ::"r"(cur[0].q.x), "r"(cur[WSTEP - 1].q.x));   // (the loads have landed)
                __syncwarp();
                if (lane == 0) {
                    if (atomicAdd(&sm.readers[st], 1u) == NT / 32 - 1) {
                        sm.readers[st] = 0u;
                        if (k + NSTAGE < n_steps) issue(k + NSTAGE, st);
                    }
                }
#pragma unroll
                for (int w = 0; w < WSTEP; ++w) {
                    const int t_base = (k * WSTEP + w) * TILE - mis;
                    if (t_base >= N) continue;                     // (the stage's second tile lies beyond the read)
                    const uint4 q = cur[w].q;
                    const unsigned mn = __vmins2(__vmins2(q.x, q.y), __vmins2(q.z, q.w));
                    const unsigned mx = __vmaxs2(__vmaxs2(q.x, q.y), __vmaxs2(q.z, q.w));
                    const int tmin = min((int)(int16_t)(mn & 0xffffu), (int)(int16_t)(mn >> 16));
                    const int tmax = max((int)(int16_t)(mx & 0xffffu), (int)(int16_t)(mx >> 16));
                    if (t_base >= 0 && t_base + TILE <= N) {
                        if (tmin >= ok_lo && tmax <= ok_hi) {
#pragma unroll
                            for (int u = 0; u < PER; ++u)
                                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(col + (uint32_t)((int)cur[w].h[u] * (4 * WCOLS))),
                                             "r"(one)
                                             : "memory");
                        } else {
#pragma unroll
                            for (int u = 0; u < PER; ++u) {
                                const int vv = cur[w].h[u];
                                if ((unsigned)(vv - wlo) < (unsigned)WBINS)
                                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(col + (uint32_t)(vv * (4 * WCOLS))), "r"(one)
                                                 : "memory");
                                else if (vv < wlo) ++n_below;
                                else ++n_above;
                            }
                            if (p.spike_mode == 1 && (tmin < 250 || tmax > 1000)) {   // note where Brute will patch
                                const int g0 = t_base + tid * PER;
#pragma unroll
                                for (int u = 0; u < PER; ++u) {
                                    if (cur[w].h[u] > 1000 || cur[w].h[u] < 250) {
                                        const int pos = atomicAdd(&sm.n_spikes, 1);
                                        if (pos < SPIKE_CAP)
                                            reinterpret_cast<uint32_t *>(sm.spike_key)[pos] = (uint32_t)(g0 + u);
                                    }
                                }
                            }
                        }
                    } else {                                       // a tile that hangs over an end of the read
                        const int ba = norm_step_general<true>(sm, ctx, q, t_base + tid * PER);
                        n_below += ba & 0xff;
                        n_above += ba >> 8;
                    }
                }
            }
            jstep += (uint32_t)n_steps;
        };
        // Value-indexed form (the reads the window did not hold): every step out of line.
        auto scan_fast = [&]() {
            uint4 nxt = fetch(0);
            for (int k = 0; k < n_tiles; ++k) {
                const uint4 q = nxt;
                nxt = fetch(k + 1);
                const int mm = norm_step_general<false>(sm, ctx, q, k * TILE - mis + tid * PER);
                lmin = min(lmin, (int)(int16_t)(mm & 0xffff));
                lmax = max(lmax, mm >> 16);
            }
        };
        // warp 0, after a barrier: the noted out-of-range samples in ascending order (bitonic network over the
        // list, +inf beyond ns), then fast5.py:90-101 one sample after the other by one lane: out[i] =
        // median(out[i-2:i+3]) sees the samples before i as already patched, the ones after it raw.
        // move(old, new) takes the sample from one histogram bin to another.
        auto patch_spikes = [&](int ns, auto &&move) {
            int P = 1;
            while (P < ns) P <<= 1;
            for (int kk = 2; kk <= P; kk <<= 1) {
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const int flip = j == (kk >> 1) ? kk - 1 : j;
                    for (int t = lane; t < (P >> 1); t += 32) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const int l = i ^ flip;
                        if (l < ns) {
                            const unsigned long long a = sm.spike_key[i], b = sm.spike_key[l];
                            if (b < a) {
                                sm.spike_key[i] = b;
                                sm.spike_key[l] = a;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) {
                int p1i = -8, p2i = -8;
                int16_t p1v = 0, p2v = 0;
                for (int k = 0; k < ns; ++k) {
                    const unsigned long long key = sm.spike_key[k];
                    const int i = (int)(key >> 32), slot = (int)(key & 0xffffffffu);
                    if (i <= 2) continue;
                    int16_t w5[5];
                    const int n = min(5, N - (i - 2));
                    for (int u = 0; u < n; ++u) {
                        const int idx = i - 2 + u;
                        int16_t vv = sm.spike_w[slot][u];
                        if (idx == p1i) vv = p1v;
                        else if (idx == p2i) vv = p2v;
                        w5[u] = vv;
                    }
                    const int16_t nv = median_small(w5, n);
                    const int16_t old = sm.spike_w[slot][2];
                    if (nv != old) move((int)old, (int)nv);
                    if (i >= lo && i <= hi) stash[i - lo] = nv;
                    p2i = p1i;
                    p2v = p1v;
                    p1i = i;
                    p1v = nv;
                }
            }
        };
        // one lane: move a sample between histogram bins
        auto hist_move = [&](int from, int to) {
            if (from >= 0 && from < HBINS) sm.hist[from] -= 1u;
            else {
                gh[from + 32768] -= 1u;
                sm.ghist_dirty = 1;
            }
            if (to >= 0 && to < HBINS) sm.hist[to] += 1u;
            else {
                gh[to + 32768] += 1u;
                sm.ghist_dirty = 1;
            }
        };
        // ---- tile scan (median filters; Brute with very many out-of-range samples) -----------------------
        auto scan_tiles = [&]() {
        uint4 nxt = fetch(0);
        for (int k = 0; k < n_tiles; ++k) {
            union {
                uint4 q;
                int16_t h[PER];
            } cur;
            cur.q = nxt;
            if (k + 1 < n_tiles) nxt = fetch(k + 1);
            const int t_base = k * TILE - mis;                           // read index of tile element 0
            const int g0 = t_base + tid * PER;                           // ... of this thread's first sample
            // block-uniform: every sample of the tile belongs to the read / the tile meets the window
            const bool interior = t_base >= 0 && t_base + TILE <= N;
            const bool in_window = t_base <= hi && t_base + TILE > lo;
            int tmin, tmax;                                              // over this thread's valid samples
            if (interior) {                                              // two samples per instruction
                const unsigned mn = __vmins2(__vmins2(cur.q.x, cur.q.y), __vmins2(cur.q.z, cur.q.w));
                const unsigned mx = __vmaxs2(__vmaxs2(cur.q.x, cur.q.y), __vmaxs2(cur.q.z, cur.q.w));
                tmin = min((int)(int16_t)(mn & 0xffffu), (int)(int16_t)(mn >> 16));
                tmax = max((int)(int16_t)(mx & 0xffffu), (int)(int16_t)(mx >> 16));
            } else {
                tmin = 32767;
                tmax = -32768;
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int g = g0 + u;
                    if (g >= 0 && g < N) {
                        tmin = min(tmin, (int)cur.h[u]);
                        tmax = max(tmax, (int)cur.h[u]);
                    } else {
                        cur.h[u] = 0;                                    // zero padding (scipy medfilt's, too)
                    }
                }
            }
            const bool spike = tmin < 250 || tmax > 1000;               // Brute's test on the original samples
            *reinterpret_cast<uint4 *>(&sm.tile[HALO + tid * PER]) = cur.q;
            const bool slow = p.spike_mode == 1 ? __syncthreads_or(spike) != 0 : (__syncthreads(), p.spike_mode != 0);

            if (slow && p.spike_mode == 1) {
                // Brute (fast5.py:90-101).  Only the threads that hold an out-of-range sample do
                // anything: they mark their samples in a bitmap of the tile (a thread's 8 samples share
                // one word) ...
                if (spike) {
                    uint32_t m = 0u;
#pragma unroll
                    for (int u = 0; u < PER; ++u) {
                        const int g = g0 + u;
                        if (g >= 0 && g < N && (cur.h[u] > 1000 || cur.h[u] < 250)) m |= 1u << u;
                    }
                    if (m) atomicOr(&sm.spike_bits[tid >> 2], m << ((tid & 3) * PER));
                }
                __syncthreads();
                // ... and warp 0 walks the marked samples in ascending order, one lane patching them
                // one after the other: later medians see earlier fixes
                if (warp == 0) {
                    for (int w0 = 0; w0 < TILE / 32; w0 += 32) {
                        uint32_t word = sm.spike_bits[w0 + lane];
                        if (word) sm.spike_bits[w0 + lane] = 0u;        // leave the bitmap clean
                        unsigned live = __ballot_sync(0xffffffffu, word != 0u);
                        while (live) {
                            const int src = __ffs(live) - 1;
                            live &= live - 1;
                            uint32_t bits = __shfl_sync(0xffffffffu, word, src);
                            if (lane == 0) {
                                while (bits) {
                                    const int t = (w0 + src) * 32 + (__ffs(bits) - 1);
                                    bits &= bits - 1;
                                    const int g = t_base + t;
                                    if (g > 2) {
                                        int16_t w5[5];
                                        const int n = min(5, N - (g - 2));
                                        for (int u = 0; u < n; ++u) {                // samples g-2 .. g-2+n-1
                                            const int e = t - 2 + u;                 // tile element (-2, -1 = carried)
                                            w5[u] = e < TILE ? sm.tile[HALO + e] : raw[g - 2 + u];   // not yet patched: raw
                                        }
                                        sm.tile[HALO + t] = median_small(w5, n);
                                    }
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                cur.q = *reinterpret_cast<const uint4 *>(&sm.tile[HALO + tid * PER]);
                // (a patched value is a median of original values: tmin/tmax still bound it)
            } else if (slow) {
                // scipy.signal.medfilt: zero-padded running median of the raw samples (fast5.py:72-75)
                const int h = p.spike_mode / 2;
                int16_t res[PER];
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    int16_t w5[5];
                    for (int d = -h; d <= h; ++d) {
                        const int e = tid * PER + u + d;
                        const int g = t_base + e;
                        int16_t vv = 0;
                        if (g >= 0 && g < N) vv = e < TILE ? sm.tile[HALO + e] : raw[g];
                        w5[d + h] = vv;
                    }
                    res[u] = median_small(w5, 2 * h + 1);
                }
                __syncthreads();                                        // everyone has read its neighbours
                if (tid == NT - 1) {                                    // raw carry for the next tile
                    sm.tile[HALO - 2] = cur.h[PER - 2];
                    sm.tile[HALO - 1] = cur.h[PER - 1];
                }
                tmin = 32767;                                           // the zero padding can put values
                tmax = -32768;                                          // below the raw minimum
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    cur.h[u] = res[u];
                    const int g = g0 + u;
                    if (g >= 0 && g < N) {
                        tmin = min(tmin, (int)res[u]);
                        tmax = max(tmax, (int)res[u]);
                    }
                }
            }
            lmin = min(lmin, tmin);
            lmax = max(lmax, tmax);

            // histogram
            if (interior && tmin >= 0 && tmax < HBINS) {
#pragma unroll
                for (int u = 0; u < PER; ++u) atomicAdd(&sm.hist[cur.h[u]], 1u);
            } else {
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int g = g0 + u;
                    if (g < 0 || g >= N) continue;
                    const int vv = cur.h[u];
                    if (vv >= 0 && vv < HBINS) {
                        atomicAdd(&sm.hist[vv], 1u);
                    } else {
                        atomicAdd(&gh[vv + 32768], 1u);
                        sm.ghist_dirty = 1;
                    }
                }
            }
            // the window's samples are parked in the tail of their own output slot
            if (in_window) {
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int g = g0 + u;
                    if (g >= lo && g <= hi && g < N) stash[g - lo] = cur.h[u];
                }
            }
            if (p.spike_mode == 1 && tid == NT - 1) {                   // patched carry for the next tile
                sm.tile[HALO - 2] = cur.h[PER - 2];
                sm.tile[HALO - 1] = cur.h[PER - 1];
            }
            // the next iteration's tile store cannot pass this tile's readers: in the slow paths they
            // are fenced by the barriers above, in the fast path nobody reads the tile
        }
        };

        // ---- window path (Brute / None) -------------------------------------------------------------------
        bool done = false;
        int win_ns = 0;
        constexpr int WREG = 8;                                   // output windows up to 4096 samples are held in registers
        uint32_t wreg[WREG];
        if (try_window) {
            if (tid == 0) scan_window_prologue();
        WSTR_NORM_T(1);
            scan_window();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                n_below += __shfl_xor_sync(0xffffffffu, n_below, o);
                n_above += __shfl_xor_sync(0xffffffffu, n_above, o);
            }
            if (lane == 0 && (n_below | n_above)) {
                atomicAdd(&sm.below, n_below);
                atomicAdd(&sm.above, n_above);
            }
            __syncthreads();
        WSTR_NORM_T(2);
            const int wlo = meta.wlo;
            const int ns = sm.n_spikes;
            win_ns = ns;
            // the output window's samples, two per register, asked for now and used after the statistics
            if (Tw <= NT * 2 * WREG) {
#pragma unroll
                for (int i = 0; i < WREG; ++i) {
                    const int t0 = tid + (2 * i) * NT, t1 = t0 + NT;
                    const uint32_t a = t0 < Tw ? (uint16_t)raw[lo + t0] : 0u;
                    const uint32_t b2 = t1 < Tw ? (uint16_t)raw[lo + t1] : 0u;
                    wreg[i] = a | b2 << 16;
                }
            }
            uint32_t *const sidx = reinterpret_cast<uint32_t *>(sm.spike_key);   // the noted samples' indices
            if (warp != 0) {
                // warps 1-7, while warp 0 patches: every bin's 32 counters summed into cum[bin] and left zero.
                // Neighbouring threads' rows start 16 banks apart, and with the 16-byte pieces taken in an order
                // skewed by bin / 2 a quarter warp's eight accesses cover the 32 banks once.
                static_assert(WCOLS == 16, "window reduction layout");
                for (int bin = tid - 32; bin < WBINS; bin += NT - 32) {
                    uint4 *row = reinterpret_cast<uint4 *>(sm.hist + bin * WCOLS);
                    uint32_t acc = 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = (j + (bin >> 1)) & 3;
                        const uint4 x = row[c];
                        row[c] = make_uint4(0u, 0u, 0u, 0u);
                        acc += (x.x & 0xffffu) + (x.x >> 16) + (x.y & 0xffffu) + (x.y >> 16) + (x.z & 0xffffu) + (x.z >> 16) +
                               (x.w & 0xffffu) + (x.w >> 16);
                    }
                    cum[bin] = acc;
                }
            }
            prepare_next();                                       // (the last warp, meanwhile)
            if (ns > 0 && ns <= SPIKE_CAP && warp == 0) {
                // Brute (fast5.py:90-101) on the noted samples.  Ascending order (bitonic network, +inf beyond ns) ...
                if (ns <= 32) {                                    // one index per lane, sorted across the warp
                    uint32_t key = lane < ns ? sidx[lane] : 0xffffffffu;
#pragma unroll
                    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                        for (int j = kk >> 1; j > 0; j >>= 1) {
                            const uint32_t other = __shfl_xor_sync(0xffffffffu, key, j);
                            const bool up = (lane & kk) == 0, lower = (lane & j) == 0;
                            key = (lower == up) ? min(key, other) : max(key, other);
                        }
                    }
                    if (lane < ns) sidx[lane] = key;
                    __syncwarp();
                }
                int P = 1;
                while (P < ns) P <<= 1;
                for (int kk = 2; kk <= P && ns > 32; kk <<= 1) {
                    for (int j = kk >> 1; j > 0; j >>= 1) {
                        const int flip = j == (kk >> 1) ? kk - 1 : j;
                        for (int t = lane; t < (P >> 1); t += 32) {
                            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const int l = i ^ flip;
                            if (l < ns) {
                                const uint32_t a = sidx[i], b2 = sidx[l];
                                if (b2 < a) {
                                    sidx[i] = b2;
                                    sidx[l] = a;
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                // ... the raw samples i-2 .. i+2 around each, all loads in flight at once ...
                for (int t = lane; t < ns; t += 32) {
                    const int i = (int)sidx[t];
#pragma unroll
                    for (int d = 0; d < 5; ++d) {
                        const int idx = i - 2 + d;
                        sm.spike_w[t][d] = (idx >= 0 && idx < N) ? raw[idx] : (int16_t)0;
                    }
                }
                __syncwarp();
                // ... and out[i] = median(out[i-2:i+3]) in index order: it sees the samples before i as already
                // patched, the ones after it raw, so only samples within two of each other depend on one another.
                // A lane takes a run of such samples (nearly always a single one) from its first sample on.
                for (int t0 = lane; t0 < ns; t0 += 32) {
                    if (t0 > 0 && sidx[t0] - sidx[t0 - 1] <= 2u) continue;
                    int p1i = -8, p2i = -8, p1v = 0, p2v = 0;
                    for (int t = t0; t < ns && (t == t0 || sidx[t] - sidx[t - 1] <= 2u); ++t) {
                        const int i = (int)sidx[t];
                        const int old = sm.spike_w[t][2];
                        int nv = old;
                        if (i > 2) {                               // (fast5.py:93: the first three samples are left alone)
                            const int n = min(5, N - (i - 2));     // 3, 4 or 5 samples; the rest +inf
                            int v[5];
#pragma unroll
                            for (int u = 0; u < 5; ++u) {
                                const int idx = i - 2 + u;
                                int vv = sm.spike_w[t][u];
                                if (idx == p1i) vv = p1v;
                                else if (idx == p2i) vv = p2v;
                                v[u] = u < n ? vv : INT_MAX;
                            }
#define WSTR_CSWAP(a_, b_)                     \
    {                                          \
        const int lo_ = min(v[a_], v[b_]);     \
        v[b_] = max(v[a_], v[b_]);             \
        v[a_] = lo_;                           \
    }
                            WSTR_CSWAP(0, 1) WSTR_CSWAP(3, 4) WSTR_CSWAP(2, 4) WSTR_CSWAP(2, 3) WSTR_CSWAP(0, 3)
                            WSTR_CSWAP(0, 2) WSTR_CSWAP(1, 4) WSTR_CSWAP(1, 3) WSTR_CSWAP(1, 2)
#undef WSTR_CSWAP
                            // numpy's median; the mean of the middle two is truncated toward zero when stored back
                            // into the int16 array (fast5.py:100)
                            nv = n == 5 ? v[2] : n == 4 ? (v[1] + v[2]) / 2 : v[1];
                            if (nv != old) {                       // from one histogram bin to another
                                if (old < wlo) atomicAdd(&sm.below, -1);
                                else if (old >= wlo + WBINS) atomicAdd(&sm.above, -1);
                                else atomicAdd(&sm.corr[old - wlo], -1);
                                if (nv < wlo) atomicAdd(&sm.below, 1);
                                else if (nv >= wlo + WBINS) atomicAdd(&sm.above, 1);
                                else atomicAdd(&sm.corr[nv - wlo], 1);
                            }
                            p2i = p1i;
                            p2v = p1v;
                            p1i = i;
                            p1v = nv;
                        }
                        sm.spike_nv[t] = (int16_t)nv;
                    }
                }
            }
            __syncthreads();
        WSTR_NORM_T(3);
            // Brute's corrections added, inclusive prefix sums over the bins: two adjacent bins per thread
            static_assert(WBINS == 2 * NT, "window prefix layout");
            const uint32_t v0 = cum[2 * tid] + (uint32_t)sm.corr[2 * tid];
            const uint32_t v1 = cum[2 * tid + 1] + (uint32_t)sm.corr[2 * tid + 1];
            sm.corr[2 * tid] = 0;
            sm.corr[2 * tid + 1] = 0;
            uint32_t inc = v0 + v1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) sm.spike_bits[warp] = inc;            // (the bitmap is free between reads: 8 warp totals)
            __syncthreads();
            {
                uint32_t base = inc - (v0 + v1);
                for (int w = 0; w < warp; ++w) base += sm.spike_bits[w];
                cum[2 * tid] = base + v0;
                cum[2 * tid + 1] = base + v0 + v1;
            }
            __syncthreads();
        WSTR_NORM_T(4);
            if (tid < TILE / 32) sm.spike_bits[tid] = 0u;         // leave the bitmap clean
            if (warp == 0 && ns <= SPIKE_CAP) {
                // Order statistics by the whole warp: a monotone predicate over the bins is located with two
                // ballots (the lanes' block ends, then the block's bins).  Every rank query has to fall inside the
                // window; one that does not sends the read to the general path below.
                // (N <= WIN_MAX_N: every count and rank fits 32 bits)
                const int below = sm.below;
                const int inside = (int)cum[WBINS - 1];
                bool ok = true;                                   // (warp-uniform throughout)
                auto wcum = [&](int v) -> int {                   // samples <= v, for v in [wlo - 1, wlo + WBINS)
                    return v < wlo ? below : below + (int)cum[v - wlo];
                };
                auto value_at = [&](int rank) -> int {            // smallest v with (samples <= v) > rank
                    if (rank < below || rank >= below + inside) {
                        ok = false;
                        return 0;
                    }
                    constexpr int B = WBINS / 32;
                    const bool e1 = below + (int)cum[lane * B + B - 1] > rank;
                    const int L = __ffs(__ballot_sync(0xffffffffu, e1)) - 1;
                    const bool e2 = lane < B && below + (int)cum[L * B + min(lane, B - 1)] > rank;
                    const int q = __ffs(__ballot_sync(0xffffffffu, e2)) - 1;
                    return wlo + L * B + q;
                };
                auto percentile32 = [&](double q) {               // percentile_from with 32-bit indices: same doubles
                    const double vidx = (double)(N - 1) * q;
                    int i0 = (int)floor(vidx), i1 = i0 + 1;
                    if (vidx >= (double)(N - 1)) i0 = i1 = N - 1;
                    const double gamma = vidx - (double)i0;
                    const int a = value_at(i0);
                    const int b2 = value_at(i1);
                    const double diff = (double)(int16_t)(b2 - a);   // numpy subtracts in int16
                    double res = (double)a + diff * gamma;
                    if (gamma >= 0.5) res = (double)b2 - diff * (1.0 - gamma);
                    return res;
                };
                const double p0 = percentile32(46.5 / 100.0);
                const double p1 = percentile32(53.5 / 100.0);
                const double shift = (p0 + p1) / 2.0;
                const int fl = (int)floor(shift);
                // values by |v - shift|: the pairs (fl - m, fl + 1 + m), m = 0, 1, ...; the first m pairs hold the
                // samples in [fl - m + 1, fl + m].  m_max: the last pair inside the window
                const int m_max = ok ? min(fl - wlo, wlo + WBINS - 2 - fl) : -1;
                auto pairs = [&](int m) -> int { return wcum(fl + m + 1) - wcum(fl - m - 1); };
                auto absdev_at = [&](int rank) -> double {
                    if (m_max < 0 || pairs(m_max) <= rank) {
                        ok = false;
                        return 0.0;
                    }
                    const int B = (m_max + 32) / 32;              // smallest m with pairs(m) > rank
                    const bool e1 = pairs(min(lane * B + B - 1, m_max)) > rank;
                    const int L = __ffs(__ballot_sync(0xffffffffu, e1)) - 1;
                    const int mq = L * B + lane;
                    const bool e2 = lane < B && mq <= m_max && pairs(min(mq, m_max)) > rank;
                    const int mm = L * B + __ffs(__ballot_sync(0xffffffffu, e2)) - 1;
                    const int before = wcum(fl + mm) - wcum(fl - mm);
                    const int scl = wcum(fl - mm) - wcum(fl - mm - 1);
                    const int sch = wcum(fl + 1 + mm) - wcum(fl + mm);
                    const double dl = fabs((double)(fl - mm) - shift), du = fabs((double)(fl + 1 + mm) - shift);
                    const int within = rank - before;
                    if (dl <= du) return within < scl ? dl : du;  // inside the pair the nearer value first
                    return within < sch ? du : dl;
                };
                double scale = 0.0;
                if (N & 1) {
                    scale = absdev_at(N / 2);
                } else {
                    const double m1 = absdev_at(N / 2 - 1);
                    const double m2 = absdev_at(N / 2);
                    scale = (m1 + m2) / 2.0;
                }
                if (ok && lane == 0) {
                    sm.shift = shift;
                    sm.scale = scale;
                    if (p.shift_scale) {
                        p.shift_scale[2 * r] = shift;
                        p.shift_scale[2 * r + 1] = scale;
                    }
                    sm.win_ok = 1;
                }
            }
            __syncthreads();
            done = sm.win_ok != 0;
        }

        // ---- general path: the median filters, and the reads the window did not hold ----------------------
        if (!done) {
            for (int b = tid; b < HBINS; b += NT) sm.hist[b] = 0u;
            if (tid == 0) {
                sm.vmin = 32767;
                sm.vmax = -32768;
                sm.n_spikes = 0;
                sm.old_dirty = 1;
            }
            if (tid < HALO) sm.tile[tid] = 0;
            if (tid < TILE / 32) sm.spike_bits[tid] = 0u;
            __syncthreads();
            if (p.spike_mode <= 1) {
                scan_fast();
                __syncthreads();
                const int ns = sm.n_spikes;
                if (ns > SPIKE_CAP) {
                    // too many for the list: start over on the tile path
                    __syncthreads();
                    for (int b = tid; b < HBINS; b += NT) sm.hist[b] = 0u;
                    if (sm.ghist_dirty)
                        for (int b = tid; b < GBINS; b += NT) gh[b] = 0u;
                    lmin = 32767;
                    lmax = -32768;
                    __syncthreads();
                    if (tid == 0) sm.ghist_dirty = 0;
                    __syncthreads();
                    scan_tiles();
                } else if (ns > 0 && warp == 0) {
                    patch_spikes(ns, hist_move);
                }
            } else {
                scan_tiles();
            }
            atomicMin(&sm.vmin, lmin);
            atomicMax(&sm.vmax, lmax);
            __threadfence_block();
            __syncthreads();

            // np.percentile(data, (46.5, 53.5)), their mean = shift, np.median(|data - shift|) = scale
            const bool by_prefix = !sm.ghist_dirty && N > 0;           // block-uniform (set before the barrier above)
            if (by_prefix) {
                // in-place inclusive prefix sums of hist[vmin..vmax]: a contiguous run of bins per thread, the
                // threads' totals scanned through shared memory
                const int v0 = sm.vmin, R = sm.vmax - sm.vmin + 1;
                const int per = (R + NT - 1) / NT;
                const int b0 = v0 + tid * per, b1 = min(b0 + per, v0 + R);
                uint32_t sum = 0u;
                for (int b = b0; b < b1; ++b) sum += sm.hist[b];
                uint32_t inc = sum;
    #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) sm.spike_bits[warp] = inc;            // (free between reads; 8 warp totals)
                __syncthreads();
                uint32_t base = inc - sum;
                for (int w = 0; w < warp; ++w) base += sm.spike_bits[w];
                for (int b = b0; b < b1; ++b) {
                    base += sm.hist[b];
                    sm.hist[b] = base;
                }
                __syncthreads();
                if (tid < TILE / 32) sm.spike_bits[tid] = 0u;         // leave the bitmap clean
                if (tid == 0) {
                    const double p0 = percentile_from([&](int64_t rk) { return value_at_rank_cum(sm, rk); }, 46.5 / 100.0);
                    const double p1 = percentile_from([&](int64_t rk) { return value_at_rank_cum(sm, rk); }, 53.5 / 100.0);
                    const double shift = (p0 + p1) / 2.0;
                    double scale;
                    if (N & 1) {
                        scale = absdev_at_rank_cum(sm, shift, N / 2);
                    } else {
                        const double m1 = absdev_at_rank_cum(sm, shift, N / 2 - 1);
                        const double m2 = absdev_at_rank_cum(sm, shift, N / 2);
                        scale = (m1 + m2) / 2.0;
                    }
                    sm.shift = shift;
                    sm.scale = scale;
                    if (p.shift_scale) {
                        p.shift_scale[2 * r] = shift;
                        p.shift_scale[2 * r + 1] = scale;
                    }
                }
            } else if (warp == 0 && N > 0) {
                // samples outside [0, HBINS) went to the global histogram: walk the bins
                const double p0 = percentile_from([&](int64_t rk) { return value_at_rank(sm, gh, rk, lane); }, 46.5 / 100.0);
                const double p1 = percentile_from([&](int64_t rk) { return value_at_rank(sm, gh, rk, lane); }, 53.5 / 100.0);
                const double shift = (p0 + p1) / 2.0;
                // np.median(np.abs(data - shift))
                double scale;
                if (N & 1) {
                    scale = absdev_at_rank(sm, gh, shift, N / 2, lane);
                } else {
                    const double m1 = absdev_at_rank(sm, gh, shift, N / 2 - 1, lane);
                    const double m2 = absdev_at_rank(sm, gh, shift, N / 2, lane);
                    scale = (m1 + m2) / 2.0;
                }
                if (lane == 0) {
                    sm.shift = shift;
                    sm.scale = scale;
                    if (p.shift_scale) {
                        p.shift_scale[2 * r] = shift;
                        p.shift_scale[2 * r + 1] = scale;
                    }
                }
            }
        }
        prepare_next();
        __syncthreads();
        WSTR_NORM_T(5);
        const double shift = sm.shift, scale = sm.scale;
        // (x - shift) / scale, the reference's expression (fast5.py:113).  The divisor is the same for the whole
        // read: the five-instruction exact division of wstr_internal.h (the very bits of a / d) where its
        // ranges hold -- the numerator is 0 or a multiple of ulp(shift) below 2^17 -- and the plain one otherwise
        // (scale 0, a constant read).
        const Divisor dv = make_divisor(scale);
        auto normalised = [&](int v) {
            const double a = (double)v - shift;
            return dv.fast ? div_fast(a, dv) : a / scale;
        };

        if (done) {
            // window path: straight from the read, then the patched samples the window holds
            if (Tw <= NT * 2 * WREG) {
#pragma unroll
                for (int i = 0; i < WREG; ++i) {                 // (the quotients unconditionally: independent chains)
                    const int t0 = tid + (2 * i) * NT, t1 = t0 + NT;
                    const double q0 = normalised((int16_t)(wreg[i] & 0xffffu));
                    const double q1 = normalised((int16_t)(wreg[i] >> 16));
                    if (t0 < Tw) out[t0] = q0;
                    if (t1 < Tw) out[t1] = q1;
                }
            } else {
                for (int t = tid; t < Tw; t += NT) out[t] = normalised(raw[lo + t]);
            }
            if (win_ns > 0) {
                __syncthreads();
                const uint32_t *const sidx = reinterpret_cast<const uint32_t *>(sm.spike_key);
                for (int t = tid; t < win_ns; t += NT) {
                    const int i = (int)sidx[t];
                    if (i >= lo && i - lo < Tw) out[i - lo] = normalised(sm.spike_nv[t]);
                }
            }
        } else {
            // convert the stashed int16 window in place, front to back, one tile at a time
            for (int64_t t0 = 0; t0 < Tw; t0 += TILE) {
                const int len = (int)min((int64_t)TILE, Tw - t0);
                for (int t = tid; t < len; t += NT) sm.tile[HALO + t] = stash[t0 + t];
                __syncthreads();
                for (int t = tid; t < len; t += NT) out[t0 + t] = normalised(sm.tile[HALO + t]);
                __syncthreads();
            }
        }
        // leave the global fallback histogram clean for the next read
        if (sm.ghist_dirty) {
            for (int b = tid; b < GBINS; b += NT) gh[b] = 0u;
        }
        __syncthreads();
        WSTR_NORM_T(6);
    }
}

int norm_grid(int n_reads) {
    int g = 148 * WSTR_NORM_BLOCKS;   // persistent: one CTA per resident slot
    return n_reads < g ? (n_reads < 1 ? 1 : n_reads) : g;
}

}  // namespace

extern "C" int64_t wstr_normalize_workspace_bytes(int32_t n_reads) {
    if (n_reads < 0) return WSTR_ERR_INVALID_ARGUMENT;
    const int64_t meta = 256 + ((int64_t)n_reads + 1) * 8 * 2 + (int64_t)n_reads * 4 * 2 + 1024;
    return meta + (int64_t)norm_grid(n_reads) * GBINS * 4;
}

extern "C" int wstr_normalize_batch(const int16_t *d_raw, const int64_t *raw_off, const int32_t *win_lo,
                                    const int32_t *win_hi, int32_t n_reads, int32_t spike_mode, double *d_out,
                                    const int64_t *out_off, double *d_shift_scale, void *d_workspace,
                                    int64_t workspace_bytes, void *stream) {
    if (!d_raw || !raw_off || !win_lo || !win_hi || !d_out || !out_off || !d_workspace || n_reads < 0)
        return WSTR_ERR_INVALID_ARGUMENT;
    if (spike_mode != 0 && spike_mode != 1 && spike_mode != 3 && spike_mode != 5) return WSTR_ERR_INVALID_ARGUMENT;
    if (n_reads == 0) return WSTR_OK;
    if (workspace_bytes < wstr_normalize_workspace_bytes(n_reads)) return WSTR_ERR_WORKSPACE_TOO_SMALL;
    for (int r = 0; r < n_reads; ++r)   // (the kernel indexes a read's samples with an int)
        if (raw_off[r + 1] <= raw_off[r] || raw_off[r + 1] - raw_off[r] > (int64_t)INT32_MAX - 2 * TILE || win_lo[r] < 0)
            return WSTR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(d_workspace);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = (off + bytes + 255) / 256 * 256;
        return o;
    };
    // [queue][raw_off][out_off][win_lo][win_hi]: one block, staged through pinned mapped memory and a copy
    // kernel (a pipelining caller keeps the DMA engines busy with the next chunk's samples)
    const size_t o_q = take(256), o_ro = take(sizeof(int64_t) * (n_reads + 1)), o_oo = take(sizeof(int64_t) * n_reads),
                 o_lo = take(sizeof(int32_t) * n_reads), o_hi = take(sizeof(int32_t) * n_reads);
    const size_t block_bytes = off;
    const int grid = norm_grid(n_reads);
    const size_t o_gh = take(0);
    if ((int64_t)(o_gh + (size_t)grid * GBINS * 4) > workspace_bytes) return WSTR_ERR_WORKSPACE_TOO_SMALL;
    {
        void *h = nullptr, *token = nullptr;
        int rc = wstr_stage_begin(block_bytes, &h, &token);
        if (rc != WSTR_OK) return rc;
        unsigned char *hb = static_cast<unsigned char *>(h);
        memset(hb + o_q, 0, 256);
        memcpy(hb + o_ro, raw_off, sizeof(int64_t) * (n_reads + 1));
        memcpy(hb + o_oo, out_off, sizeof(int64_t) * n_reads);
        memcpy(hb + o_lo, win_lo, sizeof(int32_t) * n_reads);
        memcpy(hb + o_hi, win_hi, sizeof(int32_t) * n_reads);
        rc = wstr_stage_commit(token, ws, block_bytes, s);
        if (rc != WSTR_OK) return rc;
        rc = wstr_zero_async(ws + o_gh, (size_t)grid * GBINS * 4, s);
        if (rc != WSTR_OK) return rc;
    }
    NormParams p;
    p.raw = d_raw;
    p.raw_off = reinterpret_cast<const int64_t *>(ws + o_ro);
    p.out_off = reinterpret_cast<const int64_t *>(ws + o_oo);
    p.win_lo = reinterpret_cast<const int32_t *>(ws + o_lo);
    p.win_hi = reinterpret_cast<const int32_t *>(ws + o_hi);
    p.out = d_out;
    p.shift_scale = d_shift_scale;
    p.ghist = reinterpret_cast<uint32_t *>(ws + o_gh);
    p.queue = reinterpret_cast<int32_t *>(ws + o_q);
    p.n_reads = n_reads;
    p.spike_mode = spike_mode;
    static unsigned long long attr_done = 0ull;      // devices the function attribute is set on
    int dev = 0;
    WSTR_CUDA(cudaGetDevice(&dev));
    if (!((attr_done >> (dev & 63)) & 1ull)) {
        WSTR_CUDA(cudaFuncSetAttribute(normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(NormSmem)));
        attr_done |= 1ull << (dev & 63);
    }
    wstr_prof_begin(2, s);
    normalize_kernel<<<grid, NT, sizeof(NormSmem), s>>>(p);
    wstr_prof_end(s);
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}

// ------------------------------------------------------------------------------------------
// (2b) already-normalised reads shipped as what determines their bits: the int16 samples of the
// window (after spike removal) and the read's {shift, scale}.  out[t] = (raw[t] - shift) / scale is the
// reference's own expression (schemas/fast5.py:113: one float64 subtraction and one division per sample),
// so the float64 window is reproduced bit for bit from a quarter of the bytes.
// ------------------------------------------------------------------------------------------
namespace {

struct DeqRead {
    int64_t raw_off, out_off;
    int32_t n, pad_;
};

__global__ void __launch_bounds__(256) dequantize_kernel(const int16_t *__restrict__ raw, const DeqRead *__restrict__ reads,
                                                         const double *__restrict__ shift_scale, double *__restrict__ out,
                                                         int n_reads) {
    // one CTA per read (grid-stride); eight samples per thread and step when the read is 16-byte aligned
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const DeqRead rd = reads[r];
        const double shift = shift_scale[2 * r], scale = shift_scale[2 * r + 1];
        const int16_t *__restrict__ src = raw + rd.raw_off;
        double *__restrict__ dst = out + rd.out_off;
        const int n = rd.n;
        int done = 0;
        if (((rd.raw_off & 7) == 0) && ((rd.out_off & 1) == 0)) {
            const int n8 = n & ~7;
            for (int t = threadIdx.x * 8; t < n8; t += blockDim.x * 8) {
                const uint4 q = *reinterpret_cast<const uint4 *>(src + t);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double a = (double)(int16_t)(w[k] & 0xffffu), b = (double)(int16_t)(w[k] >> 16);
                    double2 v;
                    v.x = (a - shift) / scale;
                    v.y = (b - shift) / scale;
                    *reinterpret_cast<double2 *>(dst + t + 2 * k) = v;
                }
            }
            done = n8;
        }
        for (int t = done + threadIdx.x; t < n; t += blockDim.x) dst[t] = ((double)src[t] - shift) / scale;
    }
}

}  // namespace

extern "C" int64_t wstr_dequantize_workspace_bytes(int32_t n_reads) {
    if (n_reads < 0) return WSTR_ERR_INVALID_ARGUMENT;
    return (int64_t)sizeof(DeqRead) * n_reads + 256;
}

extern "C" int wstr_dequantize_batch(const int16_t *d_raw, const int64_t *raw_off, const int32_t *lengths,
                                     const double *d_shift_scale, int32_t n_reads, double *d_out,
                                     const int64_t *out_off, void *d_workspace, int64_t workspace_bytes, void *stream) {
    if (!d_raw || !raw_off || !lengths || !d_shift_scale || !d_out || !out_off || !d_workspace || n_reads < 0)
        return WSTR_ERR_INVALID_ARGUMENT;
    if (n_reads == 0) return WSTR_OK;
    if (workspace_bytes < wstr_dequantize_workspace_bytes(n_reads)) return WSTR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int r = 0; r < n_reads; ++r)
        if (lengths[r] < 0 || raw_off[r] < 0 || out_off[r] < 0) return WSTR_ERR_INVALID_ARGUMENT;
    void *h = nullptr, *token = nullptr;
    const size_t bytes = sizeof(DeqRead) * (size_t)n_reads;
    int rc = wstr_stage_begin(bytes, &h, &token);
    if (rc != WSTR_OK) return rc;
    DeqRead *hr = static_cast<DeqRead *>(h);
    for (int r = 0; r < n_reads; ++r) {
        hr[r].raw_off = raw_off[r];
        hr[r].out_off = out_off[r];
        hr[r].n = lengths[r];
        hr[r].pad_ = 0;
    }
    rc = wstr_stage_commit(token, d_workspace, bytes, s);
    if (rc != WSTR_OK) return rc;
    const int grid = n_reads < 148 * 8 ? n_reads : 148 * 8;
    wstr_prof_begin(2, s);
    dequantize_kernel<<<grid, 256, 0, s>>>(d_raw, static_cast<const DeqRead *>(d_workspace), d_shift_scale, d_out,
                                           n_reads);
    wstr_prof_end(s);
    WSTR_CUDA(cudaGetLastError());
    return WSTR_OK;
}

// ------------------------------------------------------------------------------------------
// FP64 add-rate probe
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_add_probe(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double inc = seed * 0.5;
    for (int i = 0; i < iters; ++i) {
        a0 += inc; a1 += inc; a2 += inc; a3 += inc; a4 += inc; a5 += inc; a6 += inc; a7 += inc;
        a0 += a1;  a2 += a3;  a4 += a5;  a6 += a7;  a1 += inc; a3 += inc; a5 += inc; a7 += inc;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int wstr_measure_fp64_add_rate(double *tera_adds_per_s, void *stream) {
    if (!tera_adds_per_s) return WSTR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    WSTR_CUDA(cudaGetDevice(&dev));
    WSTR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double *d = nullptr;
    WSTR_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, s);
        fp64_add_probe<<<blocks, threads, 0, s>>>(d, iters, 1.0 + rep);
        cudaEventRecord(e1, s);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(d);
            return wstr_set_cuda_error(e, "fp64 probe");
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double adds = (double)blocks * threads * iters * 16.0;
        const double rate = adds / (ms * 1e-3) / 1e12;
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tera_adds_per_s = best;
    return WSTR_OK;
}
