// Internal structures shared by the host side of the C ABI and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "warpstr_b200.h"

#define WSTR_MAX_K 16            // states per lane in the widest kernel (32*16 = 512 positions)
#define WSTR_MAX_MV 8            // largest min_values_per_state of the specialised kernels (the catch-all has no limit)
#define WSTR_ANY_THREADS 128     // CTA of the catch-all kernel
#define WSTR_ANY_TILE 128        // samples per signal tile of the catch-all kernel
#define WSTR_ANY_SMEM_MAX (200 * 1024)
#define WSTR_SIG_CHUNK 126       // samples per bulk-copied signal tile (1008 B; a multiple of the 3-row cycle)
// warps per CTA of the fill kernel.  Warps never talk to each other, so a CTA is one warp: every
// shared-memory address is then a compile-time constant (the 3-row cycle of the (7,1) kernel is
// 262 instructions instead of 297) and a finished warp frees its slot at once -- 7 % faster on
// the bench batch than 4-warp CTAs.
#ifndef WSTR_WARPS_PER_CTA
#define WSTR_WARPS_PER_CTA 1
#endif
#define WSTR_MAX_DEG 4           // incoming edges of a generic-slot state
#define WSTR_LANE_TAB_STRIDE 8   // u32 per lane: band bits, slot-0 source, gsrc[0..5]
#define WSTR_PRED_STRIDE 8       // i32 per position: predecessor (state << 16 | position) by direction code

// Device view of one automaton laid out for the warp-per-read kernel (see dtw.cu).
// Position p = lane*K + u; u < KC is a chain slot, u >= KC the generic slot u-KC.
// Published-row index of a value other lanes can read: generic (g, lane) -> g*32 + lane,
// chain tail of lane -> KG*32 + lane, the constant +inf cell -> KG*32 + 32.
struct DevAutomaton {
    const double *v_pos;             // [32*K] level per position (0 for padding)
    const int16_t *state_of_pos;     // [32*K] state index, -1 = padding
    const uint32_t *lane_tab;        // [32][WSTR_LANE_TAB_STRIDE]
    const int32_t *pred_tab;         // [32*K][WSTR_PRED_STRIDE]: (state << 16) | position of the predecessor for code c (-1 none)
    int32_t K, KC, KG, DEG;
    int32_t NB, RPW;                 // direction bits per lane and row; rows per 32-bit word (32/NB)
    int32_t S;                       // states
    int32_t end_pos;                 // position of the end state
    int32_t mv;                      // min_values_per_state
    int32_t th1;                     // 6*(flank_length-10)
    int32_t band6;                   // 6*(flank_length-10) (second threshold = T - band6)
    int32_t band_closed;             // no edge leads from outside the end band's skipped set into it
    int32_t init_pos[WSTR_MAX_MV + 1];  // positions of states 0..mv (row-0 initialisation)
    // ---- catch-all kernel (dtw_any.cu): the automaton as the reference holds it -------------------
    // any != 0: no specialised layout (K = 0); the tables below are valid for every automaton
    const double *values;            // [S]
    const int32_t *seq_idx;          // [S]
    const int32_t *in_ptr;           // [S+1]
    const int32_t *in_idx;           // [E] incoming states, list order
    int32_t any;
    int32_t endstate;
    int32_t after;                   // end band skips states with seq_idx < after (caller.py:211-212)
    int32_t spad;                    // S rounded up to 32: row stride of the catch-all direction words
};

// Per-read record of one wave.
struct ReadMeta {
    int64_t sig_off;     // element offset into the signal buffer
    int64_t dir_off;     // word offset into the direction-bit workspace
    int64_t mask_off;    // word offset into maskbits (ignored if maskbits == NULL)
    int32_t T;
    int32_t aut;
    int32_t read;        // index in the caller's arrays (status / end_cost)
    int32_t pad_;
};

struct FillParams {
    const DevAutomaton *auts;
    const ReadMeta *meta;        // this wave's reads
    const int32_t *order;        // indices into meta, longest first, for this K class
    int32_t n;                   // entries in order
    int32_t *queue;              // work counter (zeroed before launch)
    const double *signal;
    const uint32_t *maskbits;    // may be NULL
    uint32_t *dir;               // direction-bit workspace
    int32_t *trace;              // state index per sample, same offsets as the signal
    double *end_cost;            // may be NULL; indexed by ReadMeta.read
    int32_t *status;             // indexed by ReadMeta.read
    int32_t respect_status;      // skip reads whose status is already an error (second pass)
};

// ---- mid-stage (midstage.cu) ------------------------------------------------------------------
struct MidAutomaton {
    const double *values;        // [S]
    const int32_t *seq_idx;      // [S]
    const uint8_t *rep_mask;     // [S]
    const uint8_t *last_base;    // [S]
    int32_t flank_length;
    int32_t n_states;
};

struct MidRead {
    int64_t sig_off;     // element offset into signal / trace / rescaled
    int64_t mask_off;    // word offset into maskbits
    int64_t ws_off;      // byte offset into the scratch region
    int64_t seq_off;     // byte offset into the sequence output, -1 = none
    int32_t T, aut, read, reverse;
    int32_t run_cap;     // capacity of the run arrays
    int32_t pad_;
};

struct MidState {
    int32_t n_runs, m, start, end, ra, nb;
    int32_t n_align, pad_;       // entries of the alignment list: runs, or distinct states with reps_as_one
};

struct MidParams {
    const MidAutomaton *auts;
    const MidRead *reads;
    int32_t n;
    int32_t *queue;              // two zeroed counters
    const double *x;             // the signal this pass aligned
    const int32_t *trace;
    double *rescaled;            // first pass: output
    uint32_t *maskbits;          // first pass: output
    unsigned char *scratch;
    MidState *state;             // [n]
    double *cubic;               // [n][8]
    int32_t *len;
    double *cost;
    uint8_t *seq;                // may be NULL
    int32_t *status;
    int32_t mv, sis;
    int32_t method, reps;        // 0 mean, 1 median (state value of a run); reps: rescaling.reps_as_one
    int32_t s_max, pipe_ok;      // most states of any automaton of the call (sizes the reps_as_one scratch);
                                 // pipe_ok: every read has 2 * run_cap <= T (the pipelined fit parks its basis there)
    double threshold, max_std;
    int32_t *ties;               // may be NULL: t-test decisions within tie_ulps of flipping, per read (first pass)
    long long tie_ulps;
};

int wstr_launch_midstage(const MidParams &p, bool second, cudaStream_t s);
int64_t wstr_mid_scratch_bytes(int T, int mv, int reps, int s_max);

struct wstr_automaton {
    DevAutomaton dev;            // pointers into d_blob
    void *d_blob;
    int32_t n_edges, n_generic, n_chain_lanes;
    int32_t flank_length;
    int32_t *h_state_of_pos;     // host copy, 32*K
    // tables kept for the mid-stage kernels (device, inside d_blob)
    const double *d_values;      // [S]
    const int32_t *d_seq_idx;    // [S]
    const uint8_t *d_rep_mask;   // [S]
    const uint8_t *d_last_base;  // [S]
};

int wstr_set_cuda_error(cudaError_t e, const char *where);
// Host arrays of a call travel through a pinned, device-mapped staging slot and a copy kernel on the
// caller's stream, never through the DMA engines (api.cu): write `bytes` at *h_ptr, then commit.
int wstr_stage_begin(size_t bytes, void **h_ptr, void **token);
int wstr_stage_commit(void *token, void *d_dst, size_t bytes, cudaStream_t s);
int wstr_zero_async(void *d, size_t bytes, cudaStream_t s);   // by a kernel; d 16-byte aligned, bytes % 16 == 0
void wstr_prof_begin(int category, cudaStream_t s);
void wstr_prof_end(cudaStream_t s);

#define WSTR_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return wstr_set_cuda_error(e_, #call);     \
    } while (0)

// kernels / launchers (dtw.cu)
int wstr_launch_fill(int kc, int kg, int deg, int mv, const FillParams &p, cudaStream_t s);
// catch-all (dtw_any.cu): any min_values_per_state, any in-degree; spad_max = widest automaton of the launch
int wstr_launch_fill_any(int mv, int spad_max, const FillParams &p, cudaStream_t s);
size_t wstr_any_smem_bytes(int mv, int spad_max);

#ifdef __CUDACC__
// ---- exact division by a divisor that is used many times -------------------------------------------
// The spline basis divides by (xe - xb) six times per evaluated sample, the t statistic by 3.0 five times per
// position, the normalisation by the read's scale once per window sample; a double division is ~15 FP64-pipe instructions on this GPU.  With
// y = RN(1/d) taken once (one IEEE division),
//     q0 = RN(a*y);  r0 = a - q0*d (exact, one FMA);  q1 = RN(q0 + r0*y);
//     r1 = a - q1*d (exact);                          q  = RN(q1 + r1*y)
// is the correctly rounded a/d (Markstein 1990: one such step on a faithful q with a correctly
// rounded reciprocal rounds correctly; the first step makes q1 faithful), i.e. the very bits
// `a / d` gives, in 5 instructions.  Used only where nothing can over- or underflow: |d| within
// 2^+-60 and the numerator zero or within 2^+-600, established per call (GUARD), per sample or per
// read (`tame`); every other operand takes the plain division.  The identity is also checked on
// the host against a/d (oracle/div_identity.c: 4e9 operand pairs over those ranges, the normalisation's operands and adversarial
// significands included -- divisor all ones / a power of two / 1.5 -- no mismatch).
struct Divisor {
    double d, y;
    bool fast;     // 2^-60 <= |d| <= 2^60
};
__device__ __forceinline__ Divisor make_divisor(double d) {
    Divisor r;
    r.d = d;
    r.y = 1.0 / d;
    const unsigned e = (static_cast<unsigned>(__double2hiint(d)) >> 20) & 0x7ffu;
    r.fast = e - (1023u - 60u) <= 120u;
    return r;
}
// the five-instruction form; the caller vouches for the operand ranges
__device__ __forceinline__ double div_fast(double a, const Divisor &dv) {
    const double q0 = __dmul_rn(a, dv.y);
    const double r0 = __fma_rn(-q0, dv.d, a);
    const double q1 = __fma_rn(r0, dv.y, q0);
    const double r1 = __fma_rn(-q1, dv.d, a);
    return __fma_rn(r1, dv.y, q1);
}
#endif
