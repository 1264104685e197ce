// libmvosr.so -- host side of the C ABI declared in include/mvosr.h, plus the small kernels
// (stand-alone stage 1, temporal filter).  Built for sm_100a only; there is no CPU fallback:
// every entry point either runs the CUDA kernels or returns an error code.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "handle.h"
#include "frame_kernel.cuh"
#include "aux_kernels.cuh"
#include "bucket_kernel.cuh"

using namespace mvosr;

thread_local char g_cuda_err[256] = "";

// -------------------------------------------------------------------------------------------------
// small kernels
// -------------------------------------------------------------------------------------------------
namespace mvosr {

// Stage 1 alone: one CTA per frame (grid-stride), order-preserving compaction of the surviving features.
__global__ void __launch_bounds__(256) triangulate_kernel(int n_frames, const int32_t *offsets,
        const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v, const uint8_t *e_mask,
        const double *poses, mvosr_config cfg, float *x, float *y, float *z, float *u, float *v, int32_t *n_out) {
    __shared__ int wcnt[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const int base = offsets[f], n = offsets[f + 1] - base;
        Pose pose;
#pragma unroll
        for (int i = 0; i < 9; ++i) pose.R[i] = poses[12 * f + (i / 3) * 4 + (i % 3)];
        pose.t[0] = poses[12 * f + 3]; pose.t[1] = poses[12 * f + 7]; pose.t[2] = poses[12 * f + 11];
        int total = 0;
        for (int c0 = 0; c0 < n; c0 += 256) {
            int i = c0 + tid;
            bool ok = false; double X = 0, Y = 0, Z = 0, uu = 0, vv = 0;
            if (i < n) {
                ok = triangulate_point(cur_u[base + i], cur_v[base + i], ref_u[base + i], ref_v[base + i], pose,
                                       cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.triangulation_max_depth, X, Y, Z, uu, vv);
                if (e_mask) ok = ok && e_mask[base + i] != 0;
            }
            unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
            if (lane == 0) wcnt[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { int c = wcnt[w]; if (w < warp) woff += c; tot += c; }
            if (ok) {
                // destination index <= source index and earlier chunks are already consumed: in-place safe
                int pos = base + total + woff + __popc(bal & ((1u << lane) - 1u));
                x[pos] = (float)X; y[pos] = (float)Y; z[pos] = (float)Z; u[pos] = (float)uu; v[pos] = (float)vv;
            }
            total += tot;
            __syncthreads();
        }
        if (tid == 0) n_out[f] = total;
    }
}

// median of w[0..n) (n <= 32) by rank counting -- np.median semantics: mean of the two middle values for even n
__device__ __forceinline__ double median_window(const double *w, int n) {
    const int k0 = (n - 1) >> 1, k1 = n >> 1;
    double a = 0, b = 0;
    for (int i = 0; i < n; ++i) {
        const double v = w[i];
        int less = 0, leq = 0;
        for (int j = 0; j < n; ++j) { less += w[j] < v; leq += w[j] <= v; }
        if (less <= k0 && k0 < leq) a = v;
        if (less <= k1 && k1 < leq) b = v;
    }
    return (n & 1) ? a : (a + b) / 2.0;
}

// Stage 6: driver gating (main_offline.py:57-88) + slew limiter + median of the last window_size states
// (rescale.py:168-178), then filter(data, 10) of script/evaluate_scale.py:25-29.  One CTA per sequence, frames staged in shared
// memory in chunks of FCH.  Only the slew limiter is a true recurrence (state_i = f(state_{i-1}, raw_i), unbounded memory); it is
// run SPECULATIVELY by one warp: every lane advances FL consecutive frames from a guessed start state -- the raw scale of the last
// updated frame before its segment, which is the exact state whenever that frame did not hit the +-0.3 limit --, then the lanes
// compare their start with the end state of the lane before and the wrong ones recompute, lowest first, until none changes (the
// arithmetic of every step is the sequential one, so the result is bit-identical to the loop; a slewing stretch costs one round
// per lane it crosses).  Everything else is a scan or a window and runs on all threads: the push counter of the deque, the
// medians, "repeat the last output" (a last-valid-index scan), filter_10.
constexpr int FCH = 1024;            // frames per chunk
constexpr int FHIST = 32;            // history carried between chunks (window_size <= 31, filter_10 needs 9)
constexpr int FT = 256;              // threads; each owns FCH / FT consecutive frames in the scans
constexpr int FL = 8;                // frames per lane and round of the speculative recurrence

// inclusive block scan over FCH elements, 4 consecutive per thread: OP = 0 sum, 1 max.  v[4]: in/out.
template <int OP>
__device__ __forceinline__ void filter_scan4(int v[4], int *wtmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 1; k < 4; ++k) v[k] = OP ? max(v[k], v[k - 1]) : v[k] + v[k - 1];
    int tot = v[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, tot, o); if (lane >= o) tot = OP ? max(tot, u) : tot + u; }
    if (lane == 31) wtmp[warp] = tot;
    __syncthreads();
    int excl = __shfl_up_sync(0xFFFFFFFFu, tot, 1);
    if (lane == 0) excl = OP ? -0x40000000 : 0;
    for (int w = 0; w < warp; ++w) excl = OP ? max(excl, wtmp[w]) : excl + wtmp[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = OP ? max(v[k], excl) : v[k] + excl;
    __syncthreads();
}

__global__ void __launch_bounds__(FT) filter_kernel(int n_seq, const int32_t *seq_offsets, const double *raw, const uint8_t *status,
                              const uint8_t *move, const int32_t *n_features, const mvosr_frame_record *rec, const int32_t *slot,
                              mvosr_config cfg, double *out, double *out10) {
    __shared__ double P[FHIST + FCH];       // pushed states: [0,FHIST) = tail of the earlier chunks
    __shared__ double O[FHIST + FCH];       // outputs, same layout
    __shared__ double R[FCH];               // raw scales
    __shared__ double S[FCH];               // slew-limited state after every frame, then the deque medians
    __shared__ int pidx[FCH];               // number of states pushed up to and including this frame (chunk-local, + FHIST)
    __shared__ uint8_t kind[FCH], flags[FCH];
    __shared__ int wtmp[FT / 32];
    __shared__ double s_scale, s_last; __shared__ int s_nhist, s_ohist;
    static_assert(FCH == 4 * FT, "the scans give every thread four consecutive frames");
    const int tid = threadIdx.x, lane = tid & 31;
    const int win = cfg.window_size < 1 ? 1 : (cfg.window_size > 31 ? 31 : cfg.window_size);
    const double lim = cfg.slew_limit;
    for (int s = blockIdx.x; s < n_seq; s += gridDim.x) {
        const int f0 = seq_offsets[s], f1 = seq_offsets[s + 1];
        if (tid == 0) { s_scale = 1.0; s_last = 0.0; s_nhist = 0; s_ohist = 0; }     // self.scale = 1 (rescale.py:27), scales = [0] (main_offline.py:45)
        __syncthreads();
        for (int c0 = f0; c0 < f1; c0 += FCH) {
            const int n = min(FCH, f1 - c0);
            // ---- 1. load; kind: 0 not moving -> 0 (:64-68), 1 too few features -> repeat (:73,84-86), 2 estimator called
            for (int i = tid; i < FCH; i += FT) {
                int kd = 1, fl = 0; double r = 0.0;          // padding frames repeat: they change nothing
                if (i < n) {
                    const int f = c0 + i;
                    int nfeat = 0x7FFFFFFF;
                    if (rec) {                                                            // gathered records, consumed in place
                        const mvosr_frame_record q = rec[slot ? slot[f] : f];
                        r = q.raw_scale; fl = (q.status & MVOSR_ST_UPDATED) ? 1 : 0; nfeat = q.n_features;
                    } else {
                        r = raw[f]; fl = (status[f] & MVOSR_ST_UPDATED) ? 1 : 0;
                        if (n_features) nfeat = n_features[f];
                    }
                    kd = (move && !move[f]) ? 0 : (nfeat <= cfg.min_features ? 1 : 2);
                }
                R[i] = r; kind[i] = (uint8_t)kd; flags[i] = (uint8_t)(kd == 2 && fl);
            }
            __syncthreads();
            // ---- 2. push counter of the deque (one push per estimator call)
            int v4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v4[k] = kind[4 * tid + k] == 2;
            filter_scan4<0>(v4, wtmp);
#pragma unroll
            for (int k = 0; k < 4; ++k) pidx[4 * tid + k] = FHIST + v4[k];
            // ---- 3. the slew limiter, speculatively (warp 0)
            if (tid < 32) {
                double carry = s_scale;
                for (int base = 0; base < n; base += 32 * FL) {
                    const int a = min(base + lane * FL, n), b = min(a + FL, n);
                    // guess: the raw scale of the last updated frame before this lane's segment, else the carried state
                    double lastv = 0.0; bool has = false;
                    for (int i = a; i < b; ++i) if (flags[i]) { lastv = R[i]; has = true; }
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double uv = __shfl_up_sync(0xFFFFFFFFu, lastv, o); const bool uh = __shfl_up_sync(0xFFFFFFFFu, (int)has, o) != 0;
                        if (lane >= o && !has) { lastv = uv; has = uh; }
                    }
                    double gv = __shfl_up_sync(0xFFFFFFFFu, lastv, 1); const bool gh = __shfl_up_sync(0xFFFFFFFFu, (int)has, 1) != 0;
                    double start = (lane == 0 || !gh) ? carry : gv, end = start;
                    bool run = true;
                    for (;;) {
                        if (run) {
                            double sc = start;
                            for (int i = a; i < b; ++i) {
                                if (flags[i]) {                                        // rescale.py:168-174
                                    const double r = R[i];
                                    if (r - sc > lim) sc += lim;
                                    else if (r - sc < -lim) sc -= lim;
                                    else sc = r;
                                }
                                S[i] = sc;
                            }
                            end = sc;
                        }
                        double prev = __shfl_up_sync(0xFFFFFFFFu, end, 1);
                        if (lane == 0) prev = carry;
                        run = __double_as_longlong(prev) != __double_as_longlong(start);
                        if (!__any_sync(0xFFFFFFFFu, run)) break;
                        if (run) start = prev;
                    }
                    carry = __shfl_sync(0xFFFFFFFFu, end, 31);
                }
                if (lane == 0) s_scale = carry;
            }
            __syncthreads();
            // ---- 4. the pushed states, in push order
            const int nhist = s_nhist;
            for (int i = tid; i < n; i += FT) if (kind[i] == 2) P[pidx[i] - 1] = S[i];
            __syncthreads();
            // ---- 5. median of the last `win` pushed states (rescale.py:175-178); S becomes the per-frame value 0 / median
            for (int i = tid; i < n; i += FT) {
                double m = 0.0;
                if (kind[i] == 2) {
                    const int e = pidx[i];
                    int b = e - win; if (b < FHIST - nhist) b = FHIST - nhist;
                    m = median_window(P + b, e - b);
                }
                S[i] = m;
            }
            // ---- 6. outputs: "repeat the last output" = the value at the last frame that was not a repeat
#pragma unroll
            for (int k = 0; k < 4; ++k) v4[k] = kind[4 * tid + k] != 1 ? 4 * tid + k : -1;
            __syncthreads();
            filter_scan4<1>(v4, wtmp);
            const double last = s_last;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int i = 4 * tid + k; if (i < n) O[FHIST + i] = v4[k] >= 0 ? S[v4[k]] : last; }
            __syncthreads();
            const int ohist = s_ohist;
            for (int i = tid; i < n; i += FT) {
                out[c0 + i] = O[FHIST + i];
                if (out10) {                                             // causal running median over the last 10 outputs
                    int b = FHIST + i - 9; if (b < FHIST - ohist) b = FHIST - ohist;
                    out10[c0 + i] = median_window(O + b, FHIST + i + 1 - b);
                }
            }
            __syncthreads();
            // ---- carry the tails into the history slots
            const int np = pidx[n - 1], keepP = min(FHIST, nhist + (np - FHIST)), keepO = min(FHIST, ohist + n);
            double tp = 0, to = 0;
            if (tid < FHIST) {
                if (tid >= FHIST - keepP) tp = P[np - FHIST + tid];
                if (tid >= FHIST - keepO) to = O[n + tid];
            }
            const double newlast = O[FHIST + n - 1];
            __syncthreads();
            if (tid < FHIST) { P[tid] = tp; O[tid] = to; }
            if (tid == 0) { s_nhist = keepP; s_ohist = keepO; s_last = newlast; }
            __syncthreads();
        }
    }
}

}  // namespace mvosr

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
extern "C" {

int mvosr_version(void) { return MVOSR_VERSION; }

const char *mvosr_error_string(int code) {
    switch (code) {
        case MVOSR_OK: return "ok";
        case MVOSR_E_INVALID: return "invalid argument";
        case MVOSR_E_CUDA: return "CUDA error";
        case MVOSR_E_NOMEM: return "out of memory";
        case MVOSR_E_CAPACITY: return "frame exceeds kernel capacity";
        case MVOSR_E_NO_DEVICE: return "no CUDA device";
        default: return "unknown error";
    }
}

const char *mvosr_last_cuda_error(void) { return g_cuda_err; }

static double sin_threshold(double deg) {
    // largest double s with asin(s)*180/pi < deg (deg < 0), by bisection on the monotone libm asin
    double lo = -1.0, hi = 0.0;          // lo satisfies, hi does not
    for (int i = 0; i < 200; ++i) {
        double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        if (asin(mid) * 180.0 / M_PI < deg) lo = mid; else hi = mid;
    }
    return lo;
}

static uint32_t default_pass_mask(void) {
    // graph.py:6-17,134-145 with edge potential [[3,1],[2,2],[2,2],[0,4]]: bit (idx*3+k) <=> p_k(idx) > 0.6
    const double ep[4][2] = { {3, 1}, {2, 2}, {2, 2}, {0, 4} };
    double tp[8][8];
    for (int row = 0; row < 8; ++row)
        for (int col = 0; col < 8; ++col) {
            int r0 = (row >> 2) & 1, r1 = (row >> 1) & 1, r2 = row & 1, c0 = (col >> 2) & 1, c1 = (col >> 1) & 1, c2 = col & 1;
            tp[row][col] = ep[r0 * 2 + r1][c0] * ep[r1 * 2 + r2][c1] * ep[r0 * 2 + r2][c2];
        }
    uint32_t m = 0;
    for (int idx = 0; idx < 8; ++idx) {
        double z = 0, pk[3] = { 0, 0, 0 };
        for (int row = 0; row < 8; ++row) {
            z += tp[row][idx];
            if (row & 4) pk[0] += tp[row][idx];
            if (row & 2) pk[1] += tp[row][idx];
            if (row & 1) pk[2] += tp[row][idx];
        }
        for (int k = 0; k < 3; ++k) if (pk[k] / z > 0.6) m |= 1u << (idx * 3 + k);
    }
    return m;
}

int mvosr_default_config(mvosr_config *c) {
    if (!c) return MVOSR_E_INVALID;
    memset(c, 0, sizeof(*c));
    c->absolute_reference = 1.75;        // src/param.py:36
    c->fx = 718.856; c->fy = 718.856; c->cx = 607.1928; c->cy = 185.2157;
    c->vanish = 185.0f;
    c->min_features = 100; c->min_kept = 10; c->min_selected = 12;
    c->sin_loose = sin_threshold(-80.0);
    c->sin_tight = sin_threshold(-85.0);
    c->height_level_factor = 0.9;
    c->ransac_iterations = 100; c->ransac_stop_at_goal = 1;
    c->ransac_threshold = 0.005; c->ransac_goal_fraction = 0.8;
    c->graph_pass_mask = default_pass_mask();
    c->slew_limit = 0.3; c->window_size = 5;
    c->triangulation_max_depth = 100.0;
    return MVOSR_OK;
}

int mvosr_create(const mvosr_config *cfg, int device, mvosr_handle **out) {
    if (!out) return MVOSR_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MVOSR_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return MVOSR_E_INVALID;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    CK(cudaSetDevice(device));
    mvosr_handle *h = new (std::nothrow) mvosr_handle();
    if (!h) return MVOSR_E_NOMEM;
    memset(h, 0, sizeof(*h));
    if (cfg) h->cfg = *cfg; else mvosr_default_config(&h->cfg);
    h->device = device;
    // any failure below releases what was acquired and leaves *out untouched
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) {
        h->num_sms = prop.multiProcessorCount;
        h->smem_optin = (int)prop.sharedMemPerBlockOptin;
        int cap = 256;
        while (make_plan(cap + 64).total + (int)sizeof(Ctl) + 1024 <= h->smem_optin) cap += 64;
        h->cap_max = cap;
        e = cudaMalloc(&h->work_counter, NCOUNTERS * sizeof(int));
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(frame_kernel<SRC_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin - (int)sizeof(Ctl) - 512);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(frame_kernel<SRC_CORR>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin - (int)sizeof(Ctl) - 512);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(frame_kernel<SRC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin - (int)sizeof(Ctl) - 512);
    if (e != cudaSuccess) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "%s in mvosr_create", cudaGetErrorString(e));
        if (h->work_counter) cudaFree(h->work_counter);
        delete h;
        cudaSetDevice(prev_dev);
        return MVOSR_E_CUDA;
    }
    cudaSetDevice(prev_dev);
    *out = h;
    return MVOSR_OK;
}

int mvosr_destroy(mvosr_handle *h) {
    if (!h) return MVOSR_OK;
    cudaSetDevice(h->device);
    if (h->work_counter) cudaFree(h->work_counter);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_frame) cudaFree(h->d_frame);
    if (h->h_frame) cudaFreeHost(h->h_frame);
    for (int k = 0; k < 2; ++k) { if (h->d_ws[k]) cudaFree(h->d_ws[k]); if (h->ev_ws_ready[k]) cudaEventDestroy(h->ev_ws[k]); }
    if (h->streams_ready) {
        cudaStreamDestroy(h->s_copy); cudaStreamDestroy(h->s_comp[0]); cudaStreamDestroy(h->s_comp[1]);
        for (int i = 0; i < 8; ++i) cudaEventDestroy(h->ev_copy[i]);
    }
    delete h;
    return MVOSR_OK;
}

int mvosr_get_config(const mvosr_handle *h, mvosr_config *cfg) {
    if (!h || !cfg) return MVOSR_E_INVALID;
    *cfg = h->cfg;
    return MVOSR_OK;
}

int64_t mvosr_launch_count(const mvosr_handle *h) { return h ? h->launches : 0; }

int mvosr_set_phase_timing(mvosr_handle *h, int64_t *phase_cycles_device) {
    if (!h) return MVOSR_E_INVALID;
    h->phase_cycles = (long long *)phase_cycles_device;
    return MVOSR_OK;
}

}  // extern "C"

// largest capacity the 16-bit indices of the frame kernel allow (triangle blocks are addressed below 2*cap < 65536)
static const int CAP_LIMIT = 32704;

template <int SRC>
static int launch_frames(mvosr_handle *h, FrameParams &P, int max_features, cudaStream_t st) {
    if (P.n_frames <= 0) return MVOSR_OK;
    int cap = (max_features + 63) / 64 * 64;
    if (cap < 256) cap = 256;
    if (cap > CAP_LIMIT) return MVOSR_E_CAPACITY;
    P.cap = cap;
    P.cfg = h->cfg;
    h->counter_slot = (h->counter_slot + 1) % NCOUNTERS;      // launches in flight on different streams must not share a counter
    P.work_counter = h->work_counter + h->counter_slot;
    P.phase_cycles = h->phase_cycles;
    SmemPlan pl = make_plan(P.cap);
    int grid = P.n_frames < h->num_sms ? P.n_frames : h->num_sms;
    size_t dyn = (size_t)pl.total;
    P.workspace = nullptr; P.ws_stride = 0;
    int ws = -1;
    if (cap > h->cap_max) {
        // large-frame mode: the staging lives in global memory, one slab per CTA
        ws = h->ws_slot; h->ws_slot ^= 1;
        size_t stride = ((size_t)pl.total + 255) & ~(size_t)255, need = stride * (size_t)grid;
        if (need > h->ws_bytes[ws]) {
            CK(cudaDeviceSynchronize());                     // earlier launches on any stream may still use the old buffer
            if (h->d_ws[ws]) cudaFree(h->d_ws[ws]);
            h->d_ws[ws] = nullptr; h->ws_bytes[ws] = 0;
            if (cudaMalloc(&h->d_ws[ws], need) != cudaSuccess) { cudaGetLastError(); return MVOSR_E_NOMEM; }
            h->ws_bytes[ws] = need;
        }
        P.workspace = (unsigned char *)h->d_ws[ws]; P.ws_stride = stride;
        dyn = 0;
        if (!h->ev_ws_ready[ws]) { CK(cudaEventCreateWithFlags(&h->ev_ws[ws], cudaEventDisableTiming)); h->ev_ws_ready[ws] = 1; }
        else CK(cudaStreamWaitEvent(st, h->ev_ws[ws], 0));
    }
    CK(cudaMemsetAsync(P.work_counter, 0, sizeof(int), st));
    frame_kernel<SRC><<<grid, NT, dyn, st>>>(P);
    CK(cudaGetLastError());
    if (ws >= 0) CK(cudaEventRecord(h->ev_ws[ws], st));
    h->launches += 1;
    return MVOSR_OK;
}

extern "C" {

int mvosr_triangulate_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                             const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                             const uint8_t *e_mask, const double *poses,
                             float *x, float *y, float *z, float *u, float *v, int32_t *n_out, void *stream) {
    if (!h || n_frames < 0 || !offsets || !cur_u || !cur_v || !ref_u || !ref_v || !poses || !x || !y || !z || !u || !v || !n_out)
        return MVOSR_E_INVALID;
    if (n_frames == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    int grid = n_frames < 8 * h->num_sms ? n_frames : 8 * h->num_sms;
    triangulate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_frames, offsets, cur_u, cur_v, ref_u, ref_v, e_mask, poses,
                                                                  h->cfg, x, y, z, u, v, n_out);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_scale_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const int32_t *counts,
                       const float *x, const float *y, const float *z, const float *u, const float *v,
                       int32_t max_features, int32_t frame_index0, int32_t seq_id, uint64_t seed,
                       double *raw_scale, uint8_t *status, mvosr_frame_stats *stats,
                       const mvosr_debug_buffers *debug, void *stream) {
    if (!h || n_frames < 0 || !offsets || !x || !y || !z || !u || !v || !raw_scale || !status || max_features < 0)
        return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    FrameParams P; memset(&P, 0, sizeof(P));
    P.n_frames = n_frames; P.offsets = offsets; P.counts = counts;
    P.x = x; P.y = y; P.z = z; P.u = u; P.v = v;
    P.mode = MODE_FULL; P.gate = 0;
    P.frame_index0 = frame_index0; P.seq_id = seq_id; P.seed = seed;
    P.raw_scale = raw_scale; P.status = status; P.stats = stats;
    if (debug) { P.dbg = *debug; P.has_dbg = 1; }
    return launch_frames<SRC_F32>(h, P, max_features, (cudaStream_t)stream);
}

int mvosr_scale_frames_f64(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const double *feature3d, const double *feature2d,
                           int32_t max_features, int32_t frame_index0, int32_t seq_id, uint64_t seed,
                           double *raw_scale, uint8_t *status, mvosr_frame_stats *stats, const mvosr_debug_buffers *debug, void *stream) {
    if (!h || n_frames < 0 || !offsets || !feature3d || !feature2d || !raw_scale || !status || max_features < 0) return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    FrameParams P; memset(&P, 0, sizeof(P));
    P.n_frames = n_frames; P.offsets = offsets; P.f3d = feature3d; P.f2d = feature2d;
    P.mode = MODE_FULL; P.gate = 0;
    P.frame_index0 = frame_index0; P.seq_id = seq_id; P.seed = seed;
    P.raw_scale = raw_scale; P.status = status; P.stats = stats;
    if (debug) { P.dbg = *debug; P.has_dbg = 1; }
    return launch_frames<SRC_F64>(h, P, max_features, (cudaStream_t)stream);
}

// One frame from host memory, the call behind the per-frame drop-in (compat/rescale.py): the two numpy arrays go to the device
// as they are (one copy each through a pinned staging buffer), one launch, and ONE small copy back (stats + record).
int mvosr_scale_frame_host_f64(mvosr_handle *h, int32_t n_features, const double *feature3d_host, const double *feature2d_host,
                               int32_t frame_index, int32_t seq_id, uint64_t seed,
                               mvosr_frame_record *record_out_host, mvosr_frame_stats *stats_out_host) {
    if (!h || n_features < 0 || (n_features > 0 && (!feature3d_host || !feature2d_host)) || !record_out_host) return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)n_features;
    // device staging: [offsets 2 x int32 | pad][record 16][stats][f3 3n doubles][f2 2n doubles]; the same layout pinned on the host
    const size_t o_rec = 16, o_stats = 32, o_f3 = 32 + ((sizeof(mvosr_frame_stats) + 15) & ~(size_t)15), o_f2 = o_f3 + 24 * n, total = o_f2 + 16 * n;
    if (total > h->frame_bytes) {
        CK(cudaDeviceSynchronize());
        if (h->d_frame) cudaFree(h->d_frame);
        if (h->h_frame) cudaFreeHost(h->h_frame);
        h->d_frame = nullptr; h->h_frame = nullptr; h->frame_bytes = 0;
        const size_t cap = total + total / 2 + 4096;
        if (cudaMalloc(&h->d_frame, cap) != cudaSuccess) { cudaGetLastError(); return MVOSR_E_NOMEM; }
        if (cudaMallocHost(&h->h_frame, cap) != cudaSuccess) { cudaGetLastError(); cudaFree(h->d_frame); h->d_frame = nullptr; return MVOSR_E_NOMEM; }
        h->frame_bytes = cap;
    }
    if (!h->s_frame_ready) { CK(cudaStreamCreateWithFlags(&h->s_frame, cudaStreamNonBlocking)); h->s_frame_ready = 1; }
    char *hp = (char *)h->h_frame, *dp = (char *)h->d_frame;
    int32_t *off = (int32_t *)hp; off[0] = 0; off[1] = n_features;
    memcpy(hp + o_f3, feature3d_host, 24 * n);
    memcpy(hp + o_f2, feature2d_host, 16 * n);
    cudaStream_t st = h->s_frame;
    CK(cudaMemcpyAsync(dp, hp, 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dp + o_f3, hp + o_f3, 40 * n, cudaMemcpyHostToDevice, st));
    FrameParams P; memset(&P, 0, sizeof(P));
    P.n_frames = 1; P.offsets = (const int32_t *)dp; P.f3d = (const double *)(dp + o_f3); P.f2d = (const double *)(dp + o_f2);
    P.mode = MODE_FULL; P.gate = 0;
    P.frame_index0 = frame_index; P.seq_id = seq_id; P.seed = seed;
    P.records = (mvosr_frame_record *)(dp + o_rec); P.stats = (mvosr_frame_stats *)(dp + o_stats);
    int rc = launch_frames<SRC_F64>(h, P, n_features > 0 ? n_features : 1, st);
    if (rc != MVOSR_OK) return rc;
    CK(cudaMemcpyAsync(hp + o_rec, dp + o_rec, o_f3 - o_rec, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(record_out_host, hp + o_rec, sizeof(mvosr_frame_record));
    if (stats_out_host) memcpy(stats_out_host, hp + o_stats, sizeof(mvosr_frame_stats));
    return MVOSR_OK;
}

int mvosr_scale_frames_from_correspondences(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                       const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                       const uint8_t *e_mask, const double *poses,
                       int32_t max_features, int32_t frame_index0, int32_t seq_id, uint64_t seed,
                       double *raw_scale, uint8_t *status, int32_t *n_features, mvosr_frame_stats *stats, void *stream) {
    if (!h || n_frames < 0 || !offsets || !cur_u || !cur_v || !ref_u || !ref_v || !poses || !raw_scale || !status || max_features < 0)
        return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    FrameParams P; memset(&P, 0, sizeof(P));
    P.n_frames = n_frames; P.offsets = offsets;
    P.cur_u = cur_u; P.cur_v = cur_v; P.ref_u = ref_u; P.ref_v = ref_v; P.e_mask = e_mask; P.poses = poses;
    P.mode = MODE_FULL; P.gate = 1;
    P.frame_index0 = frame_index0; P.seq_id = seq_id; P.seed = seed;
    P.raw_scale = raw_scale; P.status = status; P.n_features = n_features; P.stats = stats;
    return launch_frames<SRC_CORR>(h, P, max_features, (cudaStream_t)stream);
}

int mvosr_delaunay_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const float *u, const float *v,
                          int32_t max_features, int32_t *tri, int32_t *n_tri, uint8_t *status, void *stream) {
    if (!h || n_frames < 0 || !offsets || !u || !v || !tri || !n_tri || max_features < 0) return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    FrameParams P; memset(&P, 0, sizeof(P));
    P.n_frames = n_frames; P.offsets = offsets; P.u = u; P.v = v;
    P.mode = MODE_DT_ONLY;
    P.tri_out = tri; P.n_tri_out = n_tri; P.status = status;
    return launch_frames<SRC_F32>(h, P, max_features, (cudaStream_t)stream);
}

int mvosr_filter_sequences(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets,
                           const double *raw_scale, const uint8_t *status, const uint8_t *move_flags,
                           const int32_t *n_features, double *scale_out, double *filter10_out, void *stream) {
    if (!h || n_sequences < 0 || !seq_offsets || !raw_scale || !status || !scale_out) return MVOSR_E_INVALID;
    if (n_sequences == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = n_sequences < 4 * h->num_sms ? n_sequences : 4 * h->num_sms;
    filter_kernel<<<grid, 256, 0, st>>>(n_sequences, seq_offsets, raw_scale, status, move_flags, n_features, nullptr, nullptr, h->cfg, scale_out, filter10_out);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_filter_records(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets,
                         const mvosr_frame_record *records, const int32_t *slot, const uint8_t *move_flags,
                         double *scale_out, double *filter10_out, void *stream) {
    if (!h || n_sequences < 0 || !seq_offsets || !records || !scale_out) return MVOSR_E_INVALID;
    if (n_sequences == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    const int grid = n_sequences < 4 * h->num_sms ? n_sequences : 4 * h->num_sms;
    filter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_sequences, seq_offsets, nullptr, nullptr, move_flags, nullptr, records, slot, h->cfg,
                                                         scale_out, filter10_out);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_scale_shard_from_correspondences(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                        const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                        const uint8_t *e_mask, const double *poses, int32_t max_features,
                        const int32_t *frame_seq, const int32_t *frame_index, int32_t frame_index0, int32_t seq_id,
                        const int32_t *order, uint64_t seed, mvosr_frame_record *records, void *stream) {
    if (!h || n_frames < 0 || !offsets || !cur_u || !cur_v || !ref_u || !ref_v || !poses || !records || max_features < 0)
        return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    FrameParams P; memset(&P, 0, sizeof(P));
    P.n_frames = n_frames; P.offsets = offsets;
    P.cur_u = cur_u; P.cur_v = cur_v; P.ref_u = ref_u; P.ref_v = ref_v; P.e_mask = e_mask; P.poses = poses;
    P.mode = MODE_FULL; P.gate = 1;
    P.frame_seq = frame_seq; P.frame_index = frame_index; P.frame_index0 = frame_index0; P.seq_id = seq_id; P.order = order; P.seed = seed;
    P.records = records;
    return launch_frames<SRC_CORR>(h, P, max_features, (cudaStream_t)stream);
}

// Host-buffer pipeline shared by the two _host entry points: S sequences (frame ranges seq_off[0..S], Philox sequence ids
// seq_id0 + s, frame counters restarting at every sequence) packed in one CSR batch.  The batch is cut into about eight chunks
// regardless of the sequence boundaries (per-frame Philox tables carry them): chunk k+1 is copied while chunk k is processed, and
// consecutive chunks run on two compute streams so that the tail of one launch overlaps the head of the next.
static int recover_host(mvosr_handle *h, int32_t n_frames, const int32_t *offsets_host,
                        const float *cur_u_host, const float *cur_v_host, const float *ref_u_host, const float *ref_v_host,
                        const double *poses_host, const uint8_t *move_flags_host, int32_t max_features,
                        int32_t n_sequences, const int32_t *seq_off, int32_t seq_id0, uint64_t seed,
                        double *scale_out_host, double *raw_scale_out_host, uint8_t *status_out_host) {
    // the sized copies below trust offsets_host: check it first (a wrong dtype or a non-monotone table would read out of bounds)
    if (offsets_host[0] != 0) return MVOSR_E_INVALID;
    for (int f = 0; f < n_frames; ++f) if (offsets_host[f + 1] < offsets_host[f]) return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    const size_t M = (size_t)offsets_host[n_frames];
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const bool tables = n_sequences > 1;
    size_t o_off = 0, o_cu = o_off + al(4 * (size_t)(n_frames + 1)), o_cv = o_cu + al(4 * M), o_ru = o_cv + al(4 * M),
           o_rv = o_ru + al(4 * M), o_pose = o_rv + al(4 * M), o_move = o_pose + al(96 * (size_t)n_frames),
           o_raw = o_move + al((size_t)n_frames), o_st = o_raw + al(8 * (size_t)n_frames), o_nf = o_st + al((size_t)n_frames),
           o_out = o_nf + al(4 * (size_t)n_frames), o_seq = o_out + al(8 * (size_t)n_frames), o_tab = o_seq + al(4 * (size_t)(n_sequences + 1)),
           total = o_tab + (tables ? al(8 * (size_t)n_frames) : 0);
    if (total > h->stage_bytes) {
        CK(cudaDeviceSynchronize());
        if (h->d_stage) cudaFree(h->d_stage);
        h->d_stage = nullptr; h->stage_bytes = 0;
        if (cudaMalloc(&h->d_stage, total) != cudaSuccess) { cudaGetLastError(); return MVOSR_E_NOMEM; }
        h->stage_bytes = total;
    }
    char *d = (char *)h->d_stage;
    if (!h->streams_ready) {
        CK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->s_comp[0], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->s_comp[1], cudaStreamNonBlocking));
        for (int i = 0; i < 8; ++i) CK(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
        h->streams_ready = 1;
    }
    cudaStream_t sc = h->s_copy;
    CK(cudaMemcpyAsync(d + o_off, offsets_host, 4 * (size_t)(n_frames + 1), cudaMemcpyHostToDevice, sc));
    CK(cudaMemcpyAsync(d + o_pose, poses_host, 96 * (size_t)n_frames, cudaMemcpyHostToDevice, sc));
    if (move_flags_host) CK(cudaMemcpyAsync(d + o_move, move_flags_host, (size_t)n_frames, cudaMemcpyHostToDevice, sc));
    CK(cudaMemcpyAsync(d + o_seq, seq_off, 4 * (size_t)(n_sequences + 1), cudaMemcpyHostToDevice, sc));
    if (tables) {
        // per-frame (sequence id, frame index inside the sequence): the Philox stream of a frame does not depend on the chunking
        int32_t *tab = (int32_t *)malloc(8 * (size_t)n_frames);
        if (!tab) return MVOSR_E_NOMEM;
        for (int s = 0; s < n_sequences; ++s)
            for (int f = seq_off[s]; f < seq_off[s + 1]; ++f) { tab[f] = seq_id0 + s; tab[n_frames + f] = f - seq_off[s]; }
        // (pageable source: the call returns once the data is staged, so the table can be freed right away)
        cudaError_t e = cudaMemcpyAsync(d + o_tab, tab, 8 * (size_t)n_frames, cudaMemcpyHostToDevice, sc);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sc);
        free(tab);
        CK(e);
    }
    // about 8 chunks over the whole call, the first one half-sized so that the GPU starts early; at least one frame per SM and chunk
    int n_chunks = 8;
    if (n_chunks > n_frames / h->num_sms) n_chunks = n_frames / h->num_sms > 0 ? n_frames / h->num_sms : 1;
    for (int c = 0; c < n_chunks; ++c) {
        int f0, f1;
        if (n_chunks > 1) {
            f0 = c == 0 ? 0 : (int)((long long)n_frames * (2 * c - 1) / (2 * n_chunks - 1));
            f1 = (int)((long long)n_frames * (2 * c + 1) / (2 * n_chunks - 1));
        } else { f0 = 0; f1 = n_frames; }
        if (f1 > n_frames) f1 = n_frames;
        if (f1 <= f0) continue;
        const size_t a0 = (size_t)offsets_host[f0], a1 = (size_t)offsets_host[f1];
        CK(cudaMemcpyAsync(d + o_cu + 4 * a0, cur_u_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
        CK(cudaMemcpyAsync(d + o_cv + 4 * a0, cur_v_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
        CK(cudaMemcpyAsync(d + o_ru + 4 * a0, ref_u_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
        CK(cudaMemcpyAsync(d + o_rv + 4 * a0, ref_v_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
        cudaEvent_t ev = h->ev_copy[c & 7];
        CK(cudaEventRecord(ev, sc));
        cudaStream_t st = h->s_comp[c & 1];
        CK(cudaStreamWaitEvent(st, ev, 0));
        FrameParams P; memset(&P, 0, sizeof(P));
        P.n_frames = f1 - f0; P.offsets = (const int32_t *)(d + o_off) + f0;
        P.cur_u = (const float *)(d + o_cu); P.cur_v = (const float *)(d + o_cv); P.ref_u = (const float *)(d + o_ru); P.ref_v = (const float *)(d + o_rv);
        P.poses = (const double *)(d + o_pose) + 12 * (size_t)f0;
        P.mode = MODE_FULL; P.gate = 1; P.seed = seed;
        if (tables) { P.frame_seq = (const int32_t *)(d + o_tab) + f0; P.frame_index = (const int32_t *)(d + o_tab) + n_frames + f0; }
        else { P.frame_index0 = f0; P.seq_id = seq_id0; }
        P.raw_scale = (double *)(d + o_raw) + f0; P.status = (uint8_t *)(d + o_st) + f0; P.n_features = (int32_t *)(d + o_nf) + f0;
        int rc = launch_frames<SRC_CORR>(h, P, max_features, st);
        if (rc != MVOSR_OK) return rc;
    }
    // join: the filter runs on compute stream 0 after both compute streams
    CK(cudaEventRecord(h->ev_copy[0], h->s_comp[1]));
    CK(cudaStreamWaitEvent(h->s_comp[0], h->ev_copy[0], 0));
    cudaStream_t st = h->s_comp[0];
    int rc = mvosr_filter_sequences(h, n_sequences, (const int32_t *)(d + o_seq), (const double *)(d + o_raw), (const uint8_t *)(d + o_st),
                                move_flags_host ? (const uint8_t *)(d + o_move) : nullptr, (const int32_t *)(d + o_nf),
                                (double *)(d + o_out), nullptr, st);
    if (rc != MVOSR_OK) return rc;
    CK(cudaMemcpyAsync(scale_out_host, d + o_out, 8 * (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    if (raw_scale_out_host) CK(cudaMemcpyAsync(raw_scale_out_host, d + o_raw, 8 * (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    if (status_out_host) CK(cudaMemcpyAsync(status_out_host, d + o_st, (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return MVOSR_OK;
}

int mvosr_recover_scales_host(mvosr_handle *h, int32_t n_frames, const int32_t *offsets_host,
                       const float *cur_u_host, const float *cur_v_host, const float *ref_u_host, const float *ref_v_host,
                       const double *poses_host, const uint8_t *move_flags_host,
                       int32_t max_features, int32_t seq_id, uint64_t seed,
                       double *scale_out_host, double *raw_scale_out_host, uint8_t *status_out_host) {
    if (!h || n_frames < 0 || !offsets_host || !cur_u_host || !cur_v_host || !ref_u_host || !ref_v_host || !poses_host || !scale_out_host)
        return MVOSR_E_INVALID;
    if (n_frames == 0) return MVOSR_OK;
    const int32_t seq_off[2] = { 0, n_frames };
    return recover_host(h, n_frames, offsets_host, cur_u_host, cur_v_host, ref_u_host, ref_v_host, poses_host, move_flags_host, max_features,
                        1, seq_off, seq_id, seed, scale_out_host, raw_scale_out_host, status_out_host);
}

int mvosr_recover_fleet_host(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets_host, const int32_t *offsets_host,
                       const float *cur_u_host, const float *cur_v_host, const float *ref_u_host, const float *ref_v_host,
                       const double *poses_host, const uint8_t *move_flags_host,
                       int32_t max_features, int32_t seq_id0, uint64_t seed,
                       double *scale_out_host, double *raw_scale_out_host, uint8_t *status_out_host) {
    if (!h || n_sequences < 0 || !seq_offsets_host || !offsets_host || !cur_u_host || !cur_v_host || !ref_u_host || !ref_v_host || !poses_host || !scale_out_host)
        return MVOSR_E_INVALID;
    if (n_sequences == 0) return MVOSR_OK;
    if (seq_offsets_host[0] != 0) return MVOSR_E_INVALID;
    for (int s = 0; s < n_sequences; ++s) if (seq_offsets_host[s + 1] < seq_offsets_host[s]) return MVOSR_E_INVALID;
    const int32_t n_frames = seq_offsets_host[n_sequences];
    if (n_frames == 0) return MVOSR_OK;
    return recover_host(h, n_frames, offsets_host, cur_u_host, cur_v_host, ref_u_host, ref_v_host, poses_host, move_flags_host, max_features,
                        n_sequences, seq_offsets_host, seq_id0, seed, scale_out_host, raw_scale_out_host, status_out_host);
}

// ---- stand-alone primitives for callers that bring their own triangles / point lists / motions (aux_kernels.cuh) ----
int mvosr_triangle_planes(mvosr_handle *h, int32_t n_tri, const int32_t *tri, const double *xyz,
                          double *normal, double *height, double *mean_y, void *stream) {
    if (!h || n_tri < 0 || !tri || !xyz) return MVOSR_E_INVALID;
    if (n_tri == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    const int grid = min((n_tri + 255) / 256, 8 * h->num_sms);
    triangle_planes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_tri, tri, xyz, normal, height, mean_y);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_triangle_votes(mvosr_handle *h, int32_t n_tri, const int32_t *tri, const double *v, const double *d,
                         int32_t n_points, int32_t *flagged, int32_t *incident, void *stream) {
    if (!h || n_tri < 0 || n_points < 0 || !tri || !v || !d || !flagged) return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(flagged, 0, sizeof(int32_t) * (size_t)n_points, st));
    if (incident) CK(cudaMemsetAsync(incident, 0, sizeof(int32_t) * (size_t)n_points, st));
    if (n_tri == 0) return MVOSR_OK;
    const int grid = min((n_tri + 255) / 256, 8 * h->num_sms);
    triangle_votes_kernel<<<grid, 256, 0, st>>>(n_tri, tri, v, d, flagged, incident);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_ransac_planes(mvosr_handle *h, int32_t n_sets, const int32_t *offsets, const double *xyz,
                        int32_t iterations, double threshold, double goal_fraction, int32_t stop_at_goal,
                        uint64_t seed, const int32_t *frame_index, int32_t seq_id,
                        double *model, int32_t *ic, int32_t *best_hyp, int32_t *hyps_used, void *stream) {
    if (!h || n_sets < 0 || !offsets || !xyz || !model || iterations < 0) return MVOSR_E_INVALID;
    if (n_sets == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    const int grid = min(n_sets, 8 * h->num_sms);
    ransac_planes_kernel<<<grid, 32 * RW, 0, (cudaStream_t)stream>>>(n_sets, offsets, xyz, iterations, threshold, goal_fraction, stop_at_goal,
                                                                      seed, frame_index, seq_id, model, ic, best_hyp, hyps_used);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_integrate_paths(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets, const double *motions,
                          const double *scales, double *poses_out, void *stream) {
    if (!h || n_sequences < 0 || !seq_offsets || !motions || !poses_out) return MVOSR_E_INVALID;
    if (n_sequences == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    const int grid = min(n_sequences, 4 * h->num_sms);
    integrate_paths_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(n_sequences, seq_offsets, motions, scales, poses_out);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_depth_from_mesh(mvosr_handle *h, int32_t width, int32_t height, double fx, double fy, double cx, double cy,
                          int32_t n_tri, const int32_t *tri, const double *uv, const double *datas,
                          double *depth, int32_t *tri_id, void *stream) {
    if (!h || width <= 0 || height <= 0 || n_tri < 0 || !tri || !uv || !datas || !depth || !tri_id) return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = (size_t)width * (size_t)height;
    CK(cudaMemsetAsync(tri_id, 0x7F, sizeof(int32_t) * npix, st));          // 0x7F7F7F7F: larger than any triangle index
    if (n_tri > 0) {
        raster_mesh_kernel<<<min((n_tri + 7) / 8, 8 * h->num_sms), 256, 0, st>>>(n_tri, tri, uv, width, height, tri_id);
        CK(cudaGetLastError());
    }
    mesh_depth_kernel<<<min((int)((npix + 255) / 256), 8 * h->num_sms), 256, 0, st>>>(width, height, fx, fy, cx, cy, datas, tri_id, depth);
    CK(cudaGetLastError());
    h->launches += n_tri > 0 ? 2 : 1;
    return MVOSR_OK;
}

int mvosr_recover_pose_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                              const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                              const uint8_t *e_mask, const double *essential, double *poses_out, int32_t *n_good, void *stream) {
    if (!h || n_frames < 0 || !offsets || !cur_u || !cur_v || !ref_u || !ref_v || !essential || !poses_out) return MVOSR_E_INVALID;
    if (n_frames == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    const int grid = min(n_frames, 8 * h->num_sms);
    recover_pose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_frames, offsets, cur_u, cur_v, ref_u, ref_v, e_mask, essential, h->cfg, poses_out, n_good);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_bucket_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const float *u, const float *v,
                        int32_t bucket_size, int32_t density, uint64_t seed, const int32_t *frame_index, int32_t seq_id,
                        int32_t *out_index, int32_t *n_out, uint8_t *status, void *stream) {
    if (!h || n_frames < 0 || !offsets || !u || !v || !out_index || !n_out || bucket_size < 1 || density < 1) return MVOSR_E_INVALID;
    if (n_frames == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    bucket_kernel<<<min(n_frames, 8 * h->num_sms), BUCKET_THREADS, 0, (cudaStream_t)stream>>>(n_frames, offsets, u, v, bucket_size, density, seed,
                                                                                            frame_index, seq_id, out_index, n_out, status);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

int mvosr_pose_mask_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                           const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                           const uint8_t *e_mask, const double *poses, uint8_t *mask_out, void *stream) {
    if (!h || n_frames < 0 || !offsets || !cur_u || !cur_v || !ref_u || !ref_v || !poses || !mask_out) return MVOSR_E_INVALID;
    if (n_frames == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    pose_mask_kernel<<<min(n_frames, 8 * h->num_sms), 256, 0, (cudaStream_t)stream>>>(n_frames, offsets, cur_u, cur_v, ref_u, ref_v, e_mask, poses, h->cfg, mask_out);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

}  // extern "C"
