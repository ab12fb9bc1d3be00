// Host side of the C ABI: automaton layout + upload, wave scheduling of the DP passes.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#include "wstr_internal.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_cuda_err[512] = "";

int wstr_set_cuda_error(cudaError_t e, const char *where) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return WSTR_ERR_NO_DEVICE;
    return WSTR_ERR_CUDA;
}

extern "C" int wstr_version(void) { return 100; }

extern "C" const char *wstr_last_cuda_error(void) { return g_cuda_err; }

extern "C" const char *wstr_error_string(int code) {
    switch (code) {
        case WSTR_OK: return "ok";
        case WSTR_ERR_INVALID_ARGUMENT: return "invalid argument";
        case WSTR_ERR_CUDA: return "CUDA error (see wstr_last_cuda_error)";
        case WSTR_ERR_TOO_MANY_STATES: return "automaton has more states than the widest kernel supports (512)";
        case WSTR_ERR_UNSUPPORTED: return "unsupported configuration";
        case WSTR_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case WSTR_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown error";
    }
}

// ------------------------------------------------------------------------------------------
// automaton layout
//
// The kernel serves the edge (position p-1 -> position p) from registers; every other edge
// costs shared-memory traffic and, worse, instructions in every lane of the warp for the
// slot it lands in.  So: cover the automaton with as few vertex-disjoint paths as possible
// (greedy, forward edges only), lay the paths end to end, and everything that is not a
// path edge becomes an "extra".
// ------------------------------------------------------------------------------------------
namespace {

const int kAllowedK[] = {4, 8, 9, 10, 12, 16};

struct Extra {
    int src_state;
    bool before_chain;
};

struct Layout {
    int K = 0;
    std::vector<int> state_of_pos;              // 32*K, -1 padding
    std::vector<int> pos_of_state;              // S
    std::vector<char> chained;                  // per position
    std::vector<std::vector<Extra>> extras;     // per position, in incoming order
};

int pick_k(int n_pos) {
    for (int k : kAllowedK)
        if (32 * k >= n_pos) return k;
    return -1;
}

bool build_layout(int S, const int32_t *in_ptr, const int32_t *in_idx, Layout &L) {
    // greedy path cover: state j extends the path ending in its closest unextended predecessor
    std::vector<int> next_in_path(S, -1), prev_in_path(S, -1);
    for (int j = 0; j < S; ++j) {
        int best = -1;
        for (int e = in_ptr[j]; e < in_ptr[j + 1]; ++e) {
            int p = in_idx[e];
            if (p < j && next_in_path[p] == -1 && p > best) best = p;
        }
        if (best >= 0) {
            next_in_path[best] = j;
            prev_in_path[j] = best;
        }
    }
    std::vector<int> seq;
    seq.reserve(S);
    for (int h = 0; h < S; ++h) {
        if (prev_in_path[h] != -1) continue;
        for (int j = h; j != -1; j = next_in_path[j]) seq.push_back(j);
    }
    if ((int)seq.size() != S) return false;
    L.K = pick_k(S);
    if (L.K < 0) return false;
    const int NP = 32 * L.K;
    L.state_of_pos.assign(NP, -1);
    L.pos_of_state.assign(S, -1);
    L.chained.assign(NP, 0);
    L.extras.assign(NP, {});
    for (int p = 0; p < S; ++p) {
        L.state_of_pos[p] = seq[p];
        L.pos_of_state[seq[p]] = p;
    }
    for (int p = 0; p < S; ++p) {
        const int j = seq[p];
        const int chain_src = (p > 0 && prev_in_path[j] == seq[p - 1]) ? seq[p - 1] : -1;
        int chain_rank = -1;
        if (chain_src >= 0) {
            for (int e = in_ptr[j]; e < in_ptr[j + 1]; ++e)
                if (in_idx[e] == chain_src) {
                    chain_rank = e - in_ptr[j];
                    break;
                }
        }
        L.chained[p] = chain_rank >= 0;
        for (int e = in_ptr[j]; e < in_ptr[j + 1]; ++e) {
            const int rank = e - in_ptr[j];
            if (rank == chain_rank) continue;
            L.extras[p].push_back({in_idx[e], chain_rank >= 0 && rank < chain_rank});
        }
    }
    return true;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

extern "C" int wstr_automaton_create(const double *values, const int32_t *seq_idx, const int32_t *in_ptr,
                                     const int32_t *in_idx, const uint8_t *rep_mask, const uint8_t *last_base,
                                     int32_t S, int32_t endstate, int32_t flank_length, int32_t mv,
                                     wstr_automaton **out) {
    if (!values || !seq_idx || !in_ptr || !out || S <= 0) return WSTR_ERR_INVALID_ARGUMENT;
    if (mv < 2 || mv > WSTR_MAX_MV) return WSTR_ERR_UNSUPPORTED;
    if (S <= mv) return WSTR_ERR_INVALID_ARGUMENT;          // reference: IndexError at caller.py:208
    if (endstate < 0 || endstate >= S) return WSTR_ERR_INVALID_ARGUMENT;
    if (S > 32 * WSTR_MAX_K) return WSTR_ERR_TOO_MANY_STATES;
    const int E = in_ptr[S];
    if (E > 0 && !in_idx) return WSTR_ERR_INVALID_ARGUMENT;
    for (int e = 0; e < E; ++e)
        if (in_idx[e] < 0 || in_idx[e] >= S) return WSTR_ERR_INVALID_ARGUMENT;

    Layout L;
    if (!build_layout(S, in_ptr, in_idx, L)) return WSTR_ERR_TOO_MANY_STATES;
    const int K = L.K, NP = 32 * K;

    // per-slot rows of extra edges
    std::vector<int> rows_of_slot(K, 0);
    int n_extra = 0;
    for (int p = 0; p < NP; ++p) {
        const int k = p % K;
        rows_of_slot[k] = std::max(rows_of_slot[k], (int)L.extras[p].size());
        n_extra += (int)L.extras[p].size();
        if (L.extras[p].size() > 14) return WSTR_ERR_UNSUPPORTED;   // 4-bit direction codes
    }
    DevAutomaton d;
    memset(&d, 0, sizeof(d));
    int n_xrows = 0;
    for (int k = 0; k < K; ++k) {
        d.xoff[k] = (uint8_t)n_xrows;
        n_xrows += rows_of_slot[k];
    }
    for (int k = K; k <= WSTR_MAX_K; ++k) d.xoff[k] = (uint8_t)n_xrows;
    if (n_xrows > WSTR_XTAB_MAX_ROWS) return WSTR_ERR_UNSUPPORTED;

    std::vector<uint16_t> xtab((size_t)std::max(n_xrows, 1) * 32, WSTR_NO_EDGE);
    std::vector<uint32_t> lane_bits(32 * 4, 0u);
    std::vector<double> v_pos(NP, 0.0);
    std::vector<int16_t> sop(NP, -1);
    const int boundary = flank_length - 10;
    const int after = seq_idx[S - 1] - boundary;     // caller.py:211-212
    uint32_t extra_slots = 0, src_slots = 0, broken = 0;
    for (int p = 0; p < NP; ++p) {
        const int lane = p / K, k = p % K;
        const int j = L.state_of_pos[p];
        sop[p] = (int16_t)j;
        if (j < 0) continue;
        v_pos[p] = values[j];
        lane_bits[96 + lane] |= 1u << k;
        if (L.chained[p]) lane_bits[lane] |= 1u << k;
        else broken |= 1u << k;
        if (seq_idx[j] < after) lane_bits[32 + lane] |= 1u << k;
        for (size_t r = 0; r < L.extras[p].size(); ++r) {
            const int sp = L.pos_of_state[L.extras[p][r].src_state];
            xtab[(size_t)(d.xoff[k] + r) * 32 + lane] = (uint16_t)(sp | (L.extras[p][r].before_chain ? 0x8000 : 0));
            lane_bits[64 + sp / K] |= 1u << (sp % K);
            src_slots |= 1u << (sp % K);
            extra_slots |= 1u << k;
        }
    }
    // padding positions never hold a finite cost; their chain bit stays 0
    for (int p = 0; p < NP; ++p)
        if (L.state_of_pos[p] < 0) broken |= 1u << (p % K);

    // one device blob
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    const size_t o_v = take(sizeof(double) * NP), o_sop = take(sizeof(int16_t) * NP),
                 o_bits = take(sizeof(uint32_t) * 128), o_x = take(sizeof(uint16_t) * xtab.size()),
                 o_val = take(sizeof(double) * S), o_sq = take(sizeof(int32_t) * S), o_rm = take(S), o_lb = take(S);
    std::vector<unsigned char> blob(off, 0);
    memcpy(blob.data() + o_v, v_pos.data(), sizeof(double) * NP);
    memcpy(blob.data() + o_sop, sop.data(), sizeof(int16_t) * NP);
    memcpy(blob.data() + o_bits, lane_bits.data(), sizeof(uint32_t) * 128);
    memcpy(blob.data() + o_x, xtab.data(), sizeof(uint16_t) * xtab.size());
    memcpy(blob.data() + o_val, values, sizeof(double) * S);
    memcpy(blob.data() + o_sq, seq_idx, sizeof(int32_t) * S);
    if (rep_mask) memcpy(blob.data() + o_rm, rep_mask, S);
    if (last_base) memcpy(blob.data() + o_lb, last_base, S);

    void *d_blob = nullptr;
    WSTR_CUDA(cudaMalloc(&d_blob, off));
    cudaError_t ce = cudaMemcpy(d_blob, blob.data(), off, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        cudaFree(d_blob);
        return wstr_set_cuda_error(ce, "cudaMemcpy(automaton)");
    }
    unsigned char *base = static_cast<unsigned char *>(d_blob);
    d.v_pos = reinterpret_cast<const double *>(base + o_v);
    d.state_of_pos = reinterpret_cast<const int16_t *>(base + o_sop);
    d.lane_bits = reinterpret_cast<const uint32_t *>(base + o_bits);
    d.xtab = reinterpret_cast<const uint16_t *>(base + o_x);
    d.K = K;
    d.W = (K + 7) / 8;
    d.S = S;
    d.n_xrows = n_xrows;
    d.end_pos = L.pos_of_state[endstate];
    d.mv = mv;
    d.th1 = 6 * boundary;
    d.band6 = 6 * boundary;
    for (int c = 0; c <= mv; ++c) d.init_pos[c] = L.pos_of_state[c];
    d.allchain_slots = ~broken;
    d.extra_slots = extra_slots;
    d.src_slots = src_slots;

    wstr_automaton *a = new wstr_automaton();
    a->dev = d;
    a->d_blob = d_blob;
    a->n_edges = E;
    a->n_extra = n_extra;
    a->n_extra_slots = __builtin_popcount(extra_slots);
    a->n_broken_slots = __builtin_popcount(broken & ((1u << K) - 1));
    a->flank_length = flank_length;
    a->h_state_of_pos = new int32_t[NP];
    for (int p = 0; p < NP; ++p) a->h_state_of_pos[p] = L.state_of_pos[p];
    a->d_values = reinterpret_cast<const double *>(base + o_val);
    a->d_seq_idx = reinterpret_cast<const int32_t *>(base + o_sq);
    a->d_rep_mask = base + o_rm;
    a->d_last_base = base + o_lb;
    *out = a;
    return WSTR_OK;
}

extern "C" int wstr_automaton_destroy(wstr_automaton *a) {
    if (!a) return WSTR_OK;
    cudaFree(a->d_blob);
    delete[] a->h_state_of_pos;
    delete a;
    return WSTR_OK;
}

extern "C" int wstr_automaton_info(const wstr_automaton *a, int32_t *info, int32_t n_info) {
    if (!a || !info) return WSTR_ERR_INVALID_ARGUMENT;
    const int32_t vals[7] = {a->dev.K, a->dev.W, a->n_extra, a->n_extra_slots, a->dev.S, a->n_edges,
                             a->n_broken_slots};
    for (int i = 0; i < n_info && i < 7; ++i) info[i] = vals[i];
    return WSTR_OK;
}

extern "C" int wstr_automaton_layout(const wstr_automaton *a, int32_t *state_of_pos, int32_t n_pos) {
    if (!a || !state_of_pos) return WSTR_ERR_INVALID_ARGUMENT;
    for (int p = 0; p < n_pos && p < 32 * a->dev.K; ++p) state_of_pos[p] = a->h_state_of_pos[p];
    return WSTR_OK;
}

// ------------------------------------------------------------------------------------------
// DP passes over a batch, in waves that fit the workspace
// ------------------------------------------------------------------------------------------
namespace {

struct WsPlan {
    size_t o_queue, o_auts, o_meta, o_order, o_dir, fixed_end;
};

WsPlan plan_workspace(int n_automata, int n_reads) {
    WsPlan w;
    size_t off = 0;
    w.o_queue = off;
    off = align_up(off + 64 * sizeof(int32_t), 256);
    w.o_auts = off;
    off = align_up(off + sizeof(DevAutomaton) * (size_t)n_automata, 256);
    w.o_meta = off;
    off = align_up(off + sizeof(ReadMeta) * (size_t)n_reads, 256);
    w.o_order = off;
    off = align_up(off + sizeof(int32_t) * (size_t)n_reads, 256);
    w.o_dir = off;
    w.fixed_end = off;
    return w;
}

inline int64_t dir_words(const wstr_automaton *a, int T) { return (int64_t)T * a->dev.W * 32; }

}  // namespace

extern "C" int64_t wstr_warp_workspace_bytes(wstr_automaton *const *automata, int32_t n_automata,
                                             const int32_t *read_automaton, const int32_t *lengths,
                                             int32_t n_reads) {
    if (!automata || n_automata <= 0 || n_reads < 0) return WSTR_ERR_INVALID_ARGUMENT;
    WsPlan w = plan_workspace(n_automata, n_reads);
    int64_t words = 0;
    for (int r = 0; r < n_reads; ++r) {
        const int a = read_automaton ? read_automaton[r] : 0;
        if (a < 0 || a >= n_automata) return WSTR_ERR_INVALID_ARGUMENT;
        words += dir_words(automata[a], lengths[r]);
    }
    return (int64_t)w.fixed_end + words * 4 + 256;
}

extern "C" int wstr_warp_batch(wstr_automaton *const *automata, int32_t n_automata,
                               const int32_t *read_automaton, const double *d_signal, const int64_t *sig_off,
                               const int32_t *lengths, const uint32_t *d_maskbits, const int64_t *mask_off,
                               int32_t n_reads, void *d_workspace, int64_t workspace_bytes, int32_t *d_trace,
                               double *d_end_cost, int32_t *d_status, void *stream) {
    if (!automata || n_automata <= 0 || n_reads < 0 || !d_signal || !sig_off || !lengths || !d_workspace ||
        !d_trace || !d_status)
        return WSTR_ERR_INVALID_ARGUMENT;
    if (d_maskbits && !mask_off) return WSTR_ERR_INVALID_ARGUMENT;
    if (n_reads == 0) return WSTR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int mv = automata[0]->dev.mv;
    for (int a = 0; a < n_automata; ++a)
        if (!automata[a] || automata[a]->dev.mv != mv) return WSTR_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < n_reads; ++r) {
        const int a = read_automaton ? read_automaton[r] : 0;
        if (a < 0 || a >= n_automata || lengths[r] < 0 || (sig_off[r] & 1)) return WSTR_ERR_INVALID_ARGUMENT;
    }

    const WsPlan w = plan_workspace(n_automata, n_reads);
    if ((int64_t)w.fixed_end >= workspace_bytes) return WSTR_ERR_WORKSPACE_TOO_SMALL;
    const int64_t dir_capacity = (workspace_bytes - (int64_t)w.o_dir) / 4;
    unsigned char *ws = static_cast<unsigned char *>(d_workspace);
    int32_t *d_queue = reinterpret_cast<int32_t *>(ws + w.o_queue);
    DevAutomaton *d_auts = reinterpret_cast<DevAutomaton *>(ws + w.o_auts);
    ReadMeta *d_meta = reinterpret_cast<ReadMeta *>(ws + w.o_meta);
    int32_t *d_order = reinterpret_cast<int32_t *>(ws + w.o_order);
    uint32_t *d_dir = reinterpret_cast<uint32_t *>(ws + w.o_dir);

    std::vector<DevAutomaton> h_auts(n_automata);
    for (int a = 0; a < n_automata; ++a) h_auts[a] = automata[a]->dev;
    WSTR_CUDA(cudaMemcpyAsync(d_auts, h_auts.data(), sizeof(DevAutomaton) * n_automata, cudaMemcpyHostToDevice, s));

    // longest reads first, grouped by automaton so that a warp rarely reloads its tables
    std::vector<int> idx(n_reads);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lengths[x] > lengths[y]; });

    std::vector<ReadMeta> meta;
    std::vector<int32_t> order;
    size_t cursor = 0;
    while (cursor < (size_t)n_reads) {
        meta.clear();
        int64_t used = 0;
        while (cursor < (size_t)n_reads) {
            const int r = idx[cursor];
            const int a = read_automaton ? read_automaton[r] : 0;
            const int64_t need = dir_words(automata[a], lengths[r]);
            if (used + need > dir_capacity) break;
            ReadMeta m;
            m.sig_off = sig_off[r];
            m.dir_off = used;
            m.mask_off = mask_off ? mask_off[r] : 0;
            m.T = lengths[r];
            m.aut = a;
            m.read = r;
            m.pad_ = 0;
            meta.push_back(m);
            used += need;
            ++cursor;
        }
        if (meta.empty()) return WSTR_ERR_WORKSPACE_TOO_SMALL;   // one read does not fit
        const int nw = (int)meta.size();
        WSTR_CUDA(cudaMemcpyAsync(d_meta, meta.data(), sizeof(ReadMeta) * nw, cudaMemcpyHostToDevice, s));
        WSTR_CUDA(cudaMemsetAsync(d_queue, 0, 64 * sizeof(int32_t), s));

        // one launch per kernel width present in this wave
        order.assign(nw, 0);
        int filled = 0, cls = 0;
        for (int k : kAllowedK) {
            const int begin = filled;
            for (int i = 0; i < nw; ++i)
                if (automata[meta[i].aut]->dev.K == k) order[filled++] = i;
            if (filled == begin) continue;
            // within a width: by automaton among equal lengths is already implied by the stable sort
            WSTR_CUDA(cudaMemcpyAsync(d_order + begin, order.data() + begin, sizeof(int32_t) * (filled - begin),
                                      cudaMemcpyHostToDevice, s));
            FillParams fp;
            fp.auts = d_auts;
            fp.meta = d_meta;
            fp.order = d_order + begin;
            fp.n = filled - begin;
            fp.queue = d_queue + cls;
            fp.signal = d_signal;
            fp.maskbits = d_maskbits;
            fp.dir = d_dir;
            fp.end_cost = d_end_cost;
            fp.status = d_status;
            int rc = wstr_launch_fill(k, mv, fp, s);
            if (rc != WSTR_OK) return rc;
            ++cls;
        }
        TraceParams tp;
        tp.auts = d_auts;
        tp.meta = d_meta;
        tp.n = nw;
        tp.maskbits = d_maskbits;
        tp.dir = d_dir;
        tp.trace = d_trace;
        tp.status = d_status;
        int rc = wstr_launch_traceback(tp, s);
        if (rc != WSTR_OK) return rc;
        if (cursor < (size_t)n_reads) {
            // the next wave overwrites meta/order staging on the host side; the device copies are
            // stream-ordered, but the pageable source vectors are reused, so drain first
            WSTR_CUDA(cudaStreamSynchronize(s));
        }
    }
    return WSTR_OK;
}
