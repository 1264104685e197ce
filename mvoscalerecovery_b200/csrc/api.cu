// libmvosr.so -- host side of the C ABI declared in include/mvosr.h, plus the small kernels
// (stand-alone stage 1, temporal filter).  Built for sm_100a only; there is no CPU fallback:
// every entry point either runs the CUDA kernels or returns an error code.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "frame_kernel.cuh"
#include "aux_kernels.cuh"
#include "five_point_kernel.cuh"
#include "bucket_kernel.cuh"

using namespace mvosr;

static thread_local char g_cuda_err[256] = "";

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return MVOSR_E_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

static const int NCOUNTERS = 16;

struct mvosr_handle {
    mvosr_config cfg;
    int device;
    int num_sms;
    int smem_optin;
    int cap_max;
    int *work_counter;           // device: NCOUNTERS dynamic-scheduler counters (one per in-flight launch)
    int counter_slot;            // round-robin
    cudaStream_t s_copy, s_comp[2]; cudaEvent_t ev_copy[8]; int streams_ready;    // host-buffer pipeline
    int64_t launches;
    // host-API staging (grown on demand)
    void *d_stage; size_t stage_bytes;
    void *d_ws; size_t ws_bytes;   // large-frame staging (frames beyond the shared-memory capacity)
    cudaEvent_t ev_ws; int ev_ws_ready;   // the staging is one buffer: launches that use it are chained through this event
    long long *phase_cycles;     // optional profiling sink (device), set by mvosr_set_phase_timing
};

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
// (rescale.py:168-178), then filter(data, 10) of script/evaluate_scale.py:25-29.  One CTA per sequence, frames in chunks
// staged in shared memory: thread 0 runs the strictly sequential recurrences (slew limiter, "repeat the last output"),
// everything else -- the windowed medians -- is computed by all threads in parallel.
constexpr int FCH = 1024;            // frames per chunk
constexpr int FHIST = 32;            // history carried between chunks (window_size <= 31, filter_10 needs 9)
__global__ void __launch_bounds__(256) filter_kernel(int n_seq, const int32_t *seq_offsets, const double *raw, const uint8_t *status,
                              const uint8_t *move, const int32_t *n_features, mvosr_config cfg, double *out, double *out10) {
    __shared__ double P[FHIST + FCH];       // pushed states: [0,FHIST) = tail of the earlier chunks
    __shared__ double O[FHIST + FCH];       // outputs, same layout
    __shared__ double R[FCH];               // raw scales, then the medians
    __shared__ int pidx[FCH];               // number of states pushed up to and including this frame (chunk-local, + FHIST)
    __shared__ uint8_t kind[FCH], flags[FCH];
    __shared__ double s_scale, s_last; __shared__ int s_npush, s_nhist, s_ohist;
    const int tid = threadIdx.x;
    const int win = cfg.window_size < 1 ? 1 : (cfg.window_size > 31 ? 31 : cfg.window_size);
    for (int s = blockIdx.x; s < n_seq; s += gridDim.x) {
        const int f0 = seq_offsets[s], f1 = seq_offsets[s + 1];
        if (tid == 0) { s_scale = 1.0; s_last = 0.0; s_nhist = 0; s_ohist = 0; }     // self.scale = 1 (rescale.py:27), scales = [0] (main_offline.py:45)
        __syncthreads();
        for (int c0 = f0; c0 < f1; c0 += FCH) {
            const int n = min(FCH, f1 - c0);
            for (int i = tid; i < n; i += blockDim.x) {
                const int f = c0 + i;
                R[i] = raw[f];
                int fl = (status[f] & MVOSR_ST_UPDATED) ? 1 : 0;
                if (move && !move[f]) fl |= 2;                                        // not moving -> 0 (:64-68)
                else if (n_features && n_features[f] <= cfg.min_features) fl |= 4;    // too few features -> repeat (:73,84-86)
                flags[i] = (uint8_t)fl;
            }
            __syncthreads();
            if (tid == 0) {
                // the slew limiter is a strict recurrence: inputs are fetched eight frames ahead so that only its own
                // arithmetic is on the dependent path
                double scale = s_scale; int np = FHIST;
                for (int i0 = 0; i0 < n; i0 += 8) {
                    double r8[8]; int f8[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const int i = min(i0 + k, n - 1); r8[k] = R[i]; f8[k] = flags[i]; }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int i = i0 + k;
                        if (i >= n) break;
                        const int fl = f8[k];
                        int kd = 2;
                        if (fl & 2) kd = 0;
                        else if (fl & 4) kd = 1;
                        else {
                            if (fl & 1) {
                                const double r = r8[k];
                                if (r - scale > cfg.slew_limit) scale += cfg.slew_limit;
                                else if (r - scale < -cfg.slew_limit) scale -= cfg.slew_limit;
                                else scale = r;
                            }
                            P[np++] = scale;
                        }
                        kind[i] = (uint8_t)kd; pidx[i] = np;
                    }
                }
                s_scale = scale; s_npush = np;
            }
            __syncthreads();
            const int nhist = s_nhist;
            for (int i = tid; i < n; i += blockDim.x) {
                if (kind[i] != 2) continue;
                const int e = pidx[i];                                   // window = the last `win` pushed states
                int b = e - win; if (b < FHIST - nhist) b = FHIST - nhist;
                R[i] = median_window(P + b, e - b);
            }
            __syncthreads();
            if (tid == 0) {
                double last = s_last;
                for (int i0 = 0; i0 < n; i0 += 8) {
                    double r8[8]; int k8[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const int i = min(i0 + k, n - 1); r8[k] = R[i]; k8[k] = kind[i]; }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int i = i0 + k;
                        if (i >= n) break;
                        const double o = k8[k] == 0 ? 0.0 : (k8[k] == 1 ? last : r8[k]);
                        O[FHIST + i] = o; last = o;
                    }
                }
                s_last = last;
            }
            __syncthreads();
            const int ohist = s_ohist;
            for (int i = tid; i < n; i += blockDim.x) {
                out[c0 + i] = O[FHIST + i];
                if (out10) {                                             // causal running median over the last 10 outputs
                    int b = FHIST + i - 9; if (b < FHIST - ohist) b = FHIST - ohist;
                    out10[c0 + i] = median_window(O + b, FHIST + i + 1 - b);
                }
            }
            __syncthreads();
            // carry the tails into the history slots
            const int np = s_npush, keepP = min(FHIST, nhist + (np - FHIST)), keepO = min(FHIST, ohist + n);
            double tp = 0, to = 0;
            if (tid < FHIST) {
                if (tid >= FHIST - keepP) tp = P[np - FHIST + tid];
                if (tid >= FHIST - keepO) to = O[n + tid];
            }
            __syncthreads();
            if (tid < FHIST) { P[tid] = tp; O[tid] = to; }
            if (tid == 0) { s_nhist = keepP; s_ohist = keepO; }
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
    CK(cudaSetDevice(device));
    mvosr_handle *h = new (std::nothrow) mvosr_handle();
    if (!h) return MVOSR_E_NOMEM;
    memset(h, 0, sizeof(*h));
    if (cfg) h->cfg = *cfg; else mvosr_default_config(&h->cfg);
    h->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    int cap = 256;
    while (make_plan(cap + 64).total + (int)sizeof(Ctl) + 1024 <= h->smem_optin) cap += 64;
    h->cap_max = cap;
    CK(cudaMalloc(&h->work_counter, NCOUNTERS * sizeof(int)));
    CK(cudaFuncSetAttribute(frame_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin - (int)sizeof(Ctl) - 512));
    CK(cudaFuncSetAttribute(frame_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin - (int)sizeof(Ctl) - 512));
    *out = h;
    return MVOSR_OK;
}

int mvosr_destroy(mvosr_handle *h) {
    if (!h) return MVOSR_OK;
    cudaSetDevice(h->device);
    if (h->work_counter) cudaFree(h->work_counter);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_ws) cudaFree(h->d_ws);
    if (h->ev_ws_ready) cudaEventDestroy(h->ev_ws);
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

template <bool FROM_CORR>
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
    if (cap > h->cap_max) {
        // large-frame mode: the staging lives in global memory, one slab per CTA
        size_t stride = ((size_t)pl.total + 255) & ~(size_t)255, need = stride * (size_t)grid;
        if (need > h->ws_bytes) {
            CK(cudaDeviceSynchronize());                     // earlier launches on any stream may still use the old buffer
            if (h->d_ws) cudaFree(h->d_ws);
            h->d_ws = nullptr; h->ws_bytes = 0;
            if (cudaMalloc(&h->d_ws, need) != cudaSuccess) { cudaGetLastError(); return MVOSR_E_NOMEM; }
            h->ws_bytes = need;
        }
        P.workspace = (unsigned char *)h->d_ws; P.ws_stride = stride;
        dyn = 0;
        // two launches in flight (e.g. the two compute streams of the host pipeline) must not share the slabs
        if (!h->ev_ws_ready) { CK(cudaEventCreateWithFlags(&h->ev_ws, cudaEventDisableTiming)); h->ev_ws_ready = 1; }
        else CK(cudaStreamWaitEvent(st, h->ev_ws, 0));
    }
    CK(cudaMemsetAsync(P.work_counter, 0, sizeof(int), st));
    frame_kernel<FROM_CORR><<<grid, NT, dyn, st>>>(P);
    CK(cudaGetLastError());
    if (P.workspace) CK(cudaEventRecord(h->ev_ws, st));
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
    return launch_frames<false>(h, P, max_features, (cudaStream_t)stream);
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
    return launch_frames<true>(h, P, max_features, (cudaStream_t)stream);
}

int mvosr_delaunay_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const float *u, const float *v,
                          int32_t max_features, int32_t *tri, int32_t *n_tri, uint8_t *status, void *stream) {
    if (!h || n_frames < 0 || !offsets || !u || !v || !tri || !n_tri || max_features < 0) return MVOSR_E_INVALID;
    CK(cudaSetDevice(h->device));
    FrameParams P; memset(&P, 0, sizeof(P));
    P.n_frames = n_frames; P.offsets = offsets; P.u = u; P.v = v;
    P.mode = MODE_DT_ONLY;
    P.tri_out = tri; P.n_tri_out = n_tri; P.status = status;
    return launch_frames<false>(h, P, max_features, (cudaStream_t)stream);
}

int mvosr_filter_sequences(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets,
                           const double *raw_scale, const uint8_t *status, const uint8_t *move_flags,
                           const int32_t *n_features, double *scale_out, double *filter10_out, void *stream) {
    if (!h || n_sequences < 0 || !seq_offsets || !raw_scale || !status || !scale_out) return MVOSR_E_INVALID;
    if (n_sequences == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = n_sequences < 4 * h->num_sms ? n_sequences : 4 * h->num_sms;
    filter_kernel<<<grid, 256, 0, st>>>(n_sequences, seq_offsets, raw_scale, status, move_flags, n_features, h->cfg, scale_out, filter10_out);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

// Host-buffer pipeline shared by the two _host entry points: S sequences (frame ranges seq_off[0..S], Philox sequence ids
// seq_id0 + s, frame counters restarting at every sequence) packed in one CSR batch.
static int recover_host(mvosr_handle *h, int32_t n_frames, const int32_t *offsets_host,
                        const float *cur_u_host, const float *cur_v_host, const float *ref_u_host, const float *ref_v_host,
                        const double *poses_host, const uint8_t *move_flags_host, int32_t max_features,
                        int32_t n_sequences, const int32_t *seq_off, int32_t seq_id0, uint64_t seed,
                        double *scale_out_host, double *raw_scale_out_host, uint8_t *status_out_host) {
    CK(cudaSetDevice(h->device));
    const size_t M = (size_t)offsets_host[n_frames];
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t o_off = 0, o_cu = o_off + al(4 * (size_t)(n_frames + 1)), o_cv = o_cu + al(4 * M), o_ru = o_cv + al(4 * M),
           o_rv = o_ru + al(4 * M), o_pose = o_rv + al(4 * M), o_move = o_pose + al(96 * (size_t)n_frames),
           o_raw = o_move + al((size_t)n_frames), o_st = o_raw + al(8 * (size_t)n_frames), o_nf = o_st + al((size_t)n_frames),
           o_out = o_nf + al(4 * (size_t)n_frames), o_seq = o_out + al(8 * (size_t)n_frames), total = o_seq + al(4 * (size_t)(n_sequences + 1));
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
    // Pipeline: every sequence is cut into chunks; chunk k+1 is copied while chunk k is processed, and consecutive chunks
    // run on two compute streams so that the tail of one launch overlaps the head of the next.
    cudaStream_t sc = h->s_copy;
    CK(cudaMemcpyAsync(d + o_off, offsets_host, 4 * (size_t)(n_frames + 1), cudaMemcpyHostToDevice, sc));
    CK(cudaMemcpyAsync(d + o_pose, poses_host, 96 * (size_t)n_frames, cudaMemcpyHostToDevice, sc));
    if (move_flags_host) CK(cudaMemcpyAsync(d + o_move, move_flags_host, (size_t)n_frames, cudaMemcpyHostToDevice, sc));
    CK(cudaMemcpyAsync(d + o_seq, seq_off, 4 * (size_t)(n_sequences + 1), cudaMemcpyHostToDevice, sc));
    int launch = 0;
    for (int s = 0; s < n_sequences; ++s) {
        const int s0 = seq_off[s], sn = seq_off[s + 1] - s0;
        // about 8 chunks over the whole call, the first chunk of the call small so that the GPU starts early
        int n_chunks = (int)((long long)8 * sn / (n_frames > 0 ? n_frames : 1));
        if (sn >= 2 * h->num_sms && n_chunks < 2) n_chunks = 2;
        if (sn < 2 * h->num_sms || n_chunks < 1) n_chunks = 1;
        if (n_chunks > 8) n_chunks = 8;
        if (n_chunks > sn / h->num_sms) n_chunks = sn / h->num_sms > 0 ? sn / h->num_sms : 1;       // at least one frame per SM and chunk
        for (int c = 0; c < n_chunks; ++c) {
            int f0, f1;
            if (s == 0 && n_chunks > 1) {
                f0 = c == 0 ? 0 : (int)((long long)sn * (2 * c - 1) / (2 * n_chunks - 1));
                f1 = (int)((long long)sn * (2 * c + 1) / (2 * n_chunks - 1));
            } else { f0 = (int)((long long)sn * c / n_chunks); f1 = (int)((long long)sn * (c + 1) / n_chunks); }
            if (f1 > sn) f1 = sn;
            if (f1 <= f0) continue;
            const size_t a0 = (size_t)offsets_host[s0 + f0], a1 = (size_t)offsets_host[s0 + f1];
            CK(cudaMemcpyAsync(d + o_cu + 4 * a0, cur_u_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
            CK(cudaMemcpyAsync(d + o_cv + 4 * a0, cur_v_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
            CK(cudaMemcpyAsync(d + o_ru + 4 * a0, ref_u_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
            CK(cudaMemcpyAsync(d + o_rv + 4 * a0, ref_v_host + a0, 4 * (a1 - a0), cudaMemcpyHostToDevice, sc));
            cudaEvent_t ev = h->ev_copy[launch & 7];
            // an event slot is reused after eight launches: by then its waiter (two compute streams, in order) has long consumed it
            CK(cudaEventRecord(ev, sc));
            cudaStream_t st = h->s_comp[launch & 1];
            CK(cudaStreamWaitEvent(st, ev, 0));
            const int g0 = s0 + f0;
            int rc = mvosr_scale_frames_from_correspondences(h, f1 - f0, (const int32_t *)(d + o_off) + g0, (const float *)(d + o_cu),
                        (const float *)(d + o_cv), (const float *)(d + o_ru), (const float *)(d + o_rv), nullptr, (const double *)(d + o_pose) + 12 * (size_t)g0,
                        max_features, f0, seq_id0 + s, seed, (double *)(d + o_raw) + g0, (uint8_t *)(d + o_st) + g0, (int32_t *)(d + o_nf) + g0, nullptr, st);
            if (rc != MVOSR_OK) return rc;
            ++launch;
        }
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

int mvosr_find_essential_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                                const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                                int32_t hypotheses, double threshold_px, double confidence, uint64_t seed, const int32_t *frame_index, int32_t seq_id,
                                double *essential, uint8_t *e_mask_out, int32_t *n_inliers, int32_t *best_hyp, int32_t *hyps_used, void *stream) {
    if (!h || n_frames < 0 || !offsets || !cur_u || !cur_v || !ref_u || !ref_v || !essential || hypotheses < 1 || hypotheses > (1 << 24) ||
        !(threshold_px > 0.0) || !(confidence >= 0.0))
        return MVOSR_E_INVALID;
    if (n_frames == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    const int grid = min(n_frames, 16 * h->num_sms);
    find_essential_kernel<<<grid, FP5_THREADS, 0, (cudaStream_t)stream>>>(n_frames, offsets, cur_u, cur_v, ref_u, ref_v,
        h->cfg.fx, h->cfg.fy, h->cfg.cx, h->cfg.cy, hypotheses, threshold_px, confidence, seed, frame_index, seq_id, essential, e_mask_out, n_inliers, best_hyp, hyps_used);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

}  // extern "C"
