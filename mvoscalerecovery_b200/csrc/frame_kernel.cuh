// The fused per-frame kernel: one CTA stages one frame in shared memory and runs stages 1-5.
//
// Reference path replaced (file:line into the reference tree):
//   stage 1  cv2.recoverPose triangulation + dehomogenise     thirdparty/MonocularVO/visual_odometry.py:129-147, main.py:102-104
//   ROI cut  feature2d[:,1] > 185                             rescale.py:115-117
//   DT #1    Delaunay(feature2d).simplices                    rescale.py:124-125
//   graph    GraphChecker.find_inliers                        graph.py:18-36,124-145
//   DT #2    re-triangulate survivors if > 10 kept            rescale.py:133-137
//   planes   flat_selection                                   rescale.py:75-102
//   RANSAC   get_pitch_ransac / run_ransac                    estimate_road_norm.py:8-18,66-70, thirdparty/Ransac/ransac.py:3-23
//   scale    height = h_bar/|n|, scale = ref/height           rescale.py:156-167
//
// Shared-memory plan for capacity `cap` ROI features:
//   X,Y,Z 12cap (float32, feature order) | region T 12cap: U,V (float32 pixel coordinates, feature order) until the last
//   grid build of the frame, then the triangle list (uint16 x 3 x 2cap) | region H 16cap: the cell-sorted copy of the
//   points (x,y float32, orig uint16), cell_start, the deferred-star queue during the Delaunay passes, then the triangle
//   heights (float64 x 2cap) | scr 4cap (uint32: cell counters, later scans) | tbase 2cap | tcnt cap | pflag cap |
//   mult 2cap | tflags 2cap | ring pool 12cap + ring info 4cap (the stars of Delaunay #1, kept for Delaunay #2): 68 bytes
//   per feature.  Between the two Delaunay passes mult holds the old -> new feature index map and tflags the list of stars
//   that must be rebuilt.
#pragma once
#include <stdint.h>
#include <math_constants.h>
#include "../../include/mvosr.h"
#ifdef MVOSR_UNIFORM_GRID
#include "gstar.cuh"            // round-1 spatial index: uniform grid tuned to image-uniform features
#else
#include "gstrip.cuh"           // density-adaptive strips (round 2)
#endif
#include "philox.cuh"
#include "triangulate.cuh"

namespace mvosr {

enum { MODE_FULL = 0, MODE_DT_ONLY = 1 };
// where a frame's features come from: float32 structure-of-arrays (x,y,z,u,v), tracked correspondences + pose (stage 1 fused), or
// the float64 arrays of the reference's own hand-off -- feature3d (n,3) / feature2d (n,2) as numpy holds them (src/main.py:102-113)
enum { SRC_F32 = 0, SRC_CORR = 1, SRC_F64 = 2 };

struct FrameParams {
    int n_frames;
    const int32_t *offsets;
    const int32_t *counts;          // optional
    // triangulated features (FROM_CORR == false)
    const float *x, *y, *z, *u, *v;
    // float64 features, array-of-structures (SRC_F64): f3d[M][3], f2d[M][2].  The Delaunay runs on the float32 roundings of the
    // pixel coordinates; the ROI cut, the depth-order votes (ties of the roundings resolved in float64), the planes, the gates
    // and the RANSAC evaluate the float64 values, as the reference does.
    const double *f3d, *f2d;
    // correspondences (SRC_CORR)
    const float *cur_u, *cur_v, *ref_u, *ref_v;
    const uint8_t *e_mask;
    const double *poses;
    int cap;                        // ROI feature capacity of the staging
    int mode;
    int gate;                       // 1: skip frames with n_features <= min_features (main_offline.py:73)
    int frame_index0, seq_id;
    // shard mode (a frame range that spans sequences, fleet.py): per-frame Philox indices, processing order, packed outputs
    const int32_t *frame_seq, *frame_index;      // optional [F]: sequence id / frame index inside its sequence (else seq_id, frame_index0 + f)
    const int32_t *order;                        // optional [F]: work item i processes frame order[i] (largest frames first)
    mvosr_frame_record *records;                 // optional [F]: (raw_scale, n_features, status) in one 16-byte record per frame
    uint64_t seed;
    mvosr_config cfg;
    double *raw_scale;
    uint8_t *status;
    int32_t *n_features;
    mvosr_frame_stats *stats;
    mvosr_debug_buffers dbg;
    int has_dbg;
    // DT-only mode outputs
    int32_t *tri_out; int32_t *n_tri_out;
    int *work_counter;              // dynamic frame scheduler
    unsigned char *workspace;       // large-frame mode: per-CTA staging in global memory (NULL: dynamic shared memory)
    unsigned long long ws_stride;
    long long *phase_cycles;        // optional [F][16] per-phase SM cycles (profiling aid)
};

struct SmemPlan {
    int cap, off_X, off_Y, off_Z, off_T, off_H, off_scr, off_tbase, off_tcnt, off_pflag, off_mult, off_tflags, off_rpool, off_rinfo, total;
    // inside region T / H
    int t_U, t_V, h_sx, h_sy, h_sorig, h_cell_start, h_defer;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline SmemPlan make_plan(int cap) {   // cap: multiple of 64
    SmemPlan p; p.cap = cap; int o = 0;
    p.off_X = o; o += 4 * cap;
    p.off_Y = o; o += 4 * cap;
    p.off_Z = o; o += 4 * cap;
    p.off_T = o; o += 12 * cap;
    p.off_H = o; o += 16 * cap;
    p.off_scr = o; o += 4 * cap + 16;
    p.off_tbase = o; o += 2 * cap;
    p.off_tcnt = o; o += cap;
    p.off_pflag = o; o += cap;
    p.off_mult = o; o += 2 * cap + 16;
    p.off_tflags = o; o += 2 * cap;
    p.off_rinfo = o; o += 4 * cap;
    p.off_rpool = o; o += 12 * cap;
    p.total = align16(o);
    p.t_U = p.off_T; p.t_V = p.off_T + 4 * cap;
    p.h_sx = p.off_H; p.h_sy = p.off_H + 4 * cap; p.h_sorig = p.off_H + 8 * cap;
    p.h_cell_start = p.off_H + 10 * cap;           // at most cap - 1 cells: cap uint16 entries
    p.h_defer = p.off_H + 12 * cap;                // two queues of 2cap bytes each: ends at 16cap
    return p;
}

struct Ctl {                         // static shared control block
    StarCtl sc;
    int n_roi, n_feat, status, bad, singular, n_dup, n_dup1, n_kept, T, n_exact, n_deferred_total, rcount, n_todo;
    int n_loose, n_tight, n_valid, best_hyp, best_ic, hyps_used, n_degenerate;
    int warp_cnt[NWARP], warp_cnt2[NWARP];
    float red[4][NWARP];
    double height_level;
    unsigned hist[256];
    unsigned long long sel_prefix; int sel_k;
    int frame;
    long long tphase[16];
    int round_ic[NWARP];
    double round_model[NWARP][5];
    SortedSet ps;
};

// ---------------------------------------------------------------------------------------------
// small block-level helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += t; }
    return v;
}

// per-warp counts already stored in cnt[0..NWARP): this warp's exclusive offset and the block total
__device__ __forceinline__ void warp_offsets(const int *cnt, int warp, int lane, int &woff, int &tot) {
    int c = lane < NWARP ? cnt[lane] : 0;                 // NWARP <= 32
    int inc = warp_incl_scan(c, lane);
    tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
    woff = __shfl_sync(0xFFFFFFFFu, inc - c, warp);
}

// exclusive scan of a[0..n) (uint32) in shared memory, in place; a[n] receives the total. Block-wide.
__device__ __forceinline__ void block_excl_scan(uint32_t *a, int n, int *warp_tmp) {
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int per = (n + NT - 1) / NT;
    int b = tid * per, e = min(n, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += (int)a[i];
    int inc = warp_incl_scan(sum, lane);
    if (lane == 31) warp_tmp[warp] = inc;
    __syncthreads();
    int woff, tot;
    warp_offsets(warp_tmp, warp, lane, woff, tot);
    int run = woff + inc - sum;
    for (int i = b; i < e; ++i) { int t = (int)a[i]; a[i] = (uint32_t)run; run += t; }
    if (tid == 0) a[n] = (uint32_t)tot;
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// grid build: cell-sorted copy (sx, sy, sorig, cell_start) of the n points in U,V
// ---------------------------------------------------------------------------------------------
struct GridArrays { float *sx, *sy; uint16_t *sorig, *cell_start; uint32_t *scr; };

#ifdef MVOSR_UNIFORM_GRID
__device__ __forceinline__ void build_grid(int n, const float *U, const float *V, uint8_t *pflag, GridArrays ga, int cap, Ctl *ctl, float density, uint16_t *, int, float) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float xmn = CUDART_INF_F, xmx = -CUDART_INF_F, ymn = CUDART_INF_F, ymx = -CUDART_INF_F;
    for (int i = tid; i < n; i += NT) {
        float a = U[i], b = V[i];
        xmn = fminf(xmn, a); xmx = fmaxf(xmx, a); ymn = fminf(ymn, b); ymx = fmaxf(ymx, b);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        xmn = fminf(xmn, __shfl_xor_sync(0xFFFFFFFFu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xFFFFFFFFu, xmx, o));
        ymn = fminf(ymn, __shfl_xor_sync(0xFFFFFFFFu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xFFFFFFFFu, ymx, o));
    }
    if (lane == 0) { ctl->red[0][warp] = xmn; ctl->red[1][warp] = xmx; ctl->red[2][warp] = ymn; ctl->red[3][warp] = ymx; }
    __syncthreads();
    if (warp == 0) {
        const int wl = lane < NWARP ? lane : 0;
        xmn = ctl->red[0][wl]; xmx = ctl->red[1][wl]; ymn = ctl->red[2][wl]; ymx = ctl->red[3][wl];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            xmn = fminf(xmn, __shfl_xor_sync(0xFFFFFFFFu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xFFFFFFFFu, xmx, o));
            ymn = fminf(ymn, __shfl_xor_sync(0xFFFFFFFFu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xFFFFFFFFu, ymx, o));
        }
        if (lane == 0) {
            const int max_cells = cap - 1;
            float w = xmx - xmn, hgt = ymx - ymn, h;
            if (w > 0.f && hgt > 0.f) h = sqrtf(density * w * hgt / (float)n);
            else h = fmaxf(w, hgt) * 2.f / (float)n;
            if (!(h > 0.f)) h = 1.f;
            int gx, gy; float inv_h;
            for (;;) {
                inv_h = 1.f / h;
                gx = (int)((xmx - xmn) * inv_h) + 1; gy = (int)((ymx - ymn) * inv_h) + 1;     // same expression as cell_of
                if ((long long)gx * gy <= max_cells) break;
                h *= 1.25f;
            }
            SortedSet &ps = ctl->ps;
            ps.x = ga.sx; ps.y = ga.sy; ps.orig = ga.sorig; ps.cell_start = ga.cell_start; ps.n = n;
            ps.xmin = xmn; ps.ymin = ymn; ps.h = h; ps.inv_h = inv_h; ps.gx = gx; ps.gy = gy;
        }
    }
    __syncthreads();
    const float xmin = ctl->ps.xmin, ymin = ctl->ps.ymin, inv_h = ctl->ps.inv_h;
    const int gx = ctl->ps.gx, gy = ctl->ps.gy, ncell = gx * gy;
    for (int i = tid; i <= ncell; i += NT) ga.scr[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        int c = cell_of(U[i], xmin, inv_h, gx) + gx * cell_of(V[i], ymin, inv_h, gy);
        atomicAdd(&ga.scr[c], 1u);
    }
    __syncthreads();
    block_excl_scan(ga.scr, ncell, ctl->warp_cnt);
    for (int i = tid; i <= ncell; i += NT) ga.cell_start[i] = (uint16_t)ga.scr[i];
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        int c = cell_of(U[i], xmin, inv_h, gx) + gx * cell_of(V[i], ymin, inv_h, gy);
        unsigned pos = atomicAdd(&ga.scr[c], 1u);
        ga.sorig[pos] = (uint16_t)i;
    }
    __syncthreads();
    // per cell: sort by feature index (determinism; lowest index first), turn exact duplicates into holes at the
    // tail (the lowest index is kept; Qhull drops duplicates too, into .coplanar)
    int ndup = 0;
    for (int c = tid; c < ncell; c += NT) {
        const int b = ga.cell_start[c], e = ga.cell_start[c + 1];
        for (int i = b + 1; i < e; ++i) {
            uint16_t v = ga.sorig[i]; int j = i;
            while (j > b && ga.sorig[j - 1] > v) { ga.sorig[j] = ga.sorig[j - 1]; --j; }
            ga.sorig[j] = v;
        }
        int m = b;
        for (int i = b; i < e; ++i) {
            const int s = ga.sorig[i]; bool dup = false;
            for (int j = b; j < m; ++j) { int q = ga.sorig[j]; if (U[q] == U[s] && V[q] == V[s]) { dup = true; break; } }
            if (dup) { pflag[s] |= 1; ++ndup; }
            else ga.sorig[m++] = (uint16_t)s;
        }
        for (int i = m; i < e; ++i) ga.sorig[i] = INF16;
    }
    if (ndup) atomicAdd(&ctl->n_dup, ndup);
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const int o = ga.sorig[i];
        ga.sx[i] = o != INF16 ? U[o] : 0.f; ga.sy[i] = o != INF16 ? V[o] : 0.f;
    }
    __syncthreads();
}

#else
// Strip build (gstrip.cuh): y-histogram with the x-extent of every bin -> strips grown until count x height >= k x extent ->
// every strip cut into uniform sub-cells of about two points -> counting sort by (strip, sub-cell) -> rank sort by (x, feature
// index) inside each sub-cell -> exact duplicates become holes in place.
// tmp: cap uint16 of scratch (the feature list in sub-cell order).  scr: cap + 4 uint32.
__device__ __forceinline__ void build_grid(int n, const float *U, const float *V, uint8_t *pflag, GridArrays ga, int cap, Ctl *ctl, float density,
                                           uint16_t *tmp, int win_m, float wfac) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float xmn = CUDART_INF_F, xmx = -CUDART_INF_F, ymn = CUDART_INF_F, ymx = -CUDART_INF_F;
    for (int i = tid; i < n; i += NT) {
        float a = U[i], b = V[i];
        xmn = fminf(xmn, a); xmx = fmaxf(xmx, a); ymn = fminf(ymn, b); ymx = fmaxf(ymx, b);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        xmn = fminf(xmn, __shfl_xor_sync(0xFFFFFFFFu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xFFFFFFFFu, xmx, o));
        ymn = fminf(ymn, __shfl_xor_sync(0xFFFFFFFFu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xFFFFFFFFu, ymx, o));
    }
    if (lane == 0) { ctl->red[0][warp] = xmn; ctl->red[1][warp] = xmx; ctl->red[2][warp] = ymn; ctl->red[3][warp] = ymx; }
    __syncthreads();
    // tables kept with the sorted set (2 cap bytes at h_cell_start): row_start, row_bin, bin_row, row_cell, row_xi, cell_start
    const int NB = max(1, min(128, (cap - 32) / 18));
    uint16_t *row_start = ga.cell_start, *row_bin = row_start + (NB + 1), *bin_row = row_bin + (NB + 1), *rowoff = bin_row + NB;
    float2 *rowxi = (float2 *)(((uintptr_t)(rowoff + (NB + 1)) + 7) & ~(uintptr_t)7);
    uint16_t *cstart = (uint16_t *)(rowxi + NB);                  // <= cap / 2 + NB + 1 entries
    uint32_t *cnt = ga.scr, *bxmin = cnt + NB, *bxmax = bxmin + NB, *cc = bxmax + NB;      // cc: sub-cell counters / cursors
    if (warp == 0) {
        const int wl = lane < NWARP ? lane : 0;
        xmn = ctl->red[0][wl]; xmx = ctl->red[1][wl]; ymn = ctl->red[2][wl]; ymx = ctl->red[3][wl];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            xmn = fminf(xmn, __shfl_xor_sync(0xFFFFFFFFu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xFFFFFFFFu, xmx, o));
            ymn = fminf(ymn, __shfl_xor_sync(0xFFFFFFFFu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xFFFFFFFFu, ymx, o));
        }
        if (lane == 0) {
            SortedSet &ps = ctl->ps;
            float bh = (ymx - ymn) / (float)NB;
            if (!(bh > 0.f)) bh = 1.f;                             // all points on one horizontal line: one bin, one strip
            ps.x = ga.sx; ps.y = ga.sy; ps.orig = ga.sorig; ps.row_start = row_start; ps.row_bin = row_bin; ps.bin_row = bin_row;
            ps.row_cell = rowoff; ps.row_xi = rowxi; ps.cell_start = cstart;
            ps.n = n; ps.NB = NB; ps.R = 0; ps.win_m = win_m; ps.kdens = density; ps.wfac = wfac;
            ps.xmin = xmn; ps.xmax = xmx; ps.ymin = ymn; ps.ymax = ymx; ps.bh = bh; ps.inv_bh = 1.f / bh;
        }
    }
    for (int i = tid; i < NB; i += NT) { cnt[i] = 0; bxmin[i] = 0xFFFFFFFFu; bxmax[i] = 0u; }
    __syncthreads();
    const SortedSet &ps = ctl->ps;
    for (int i = tid; i < n; i += NT) {
        const int b = bin_of(ps, V[i]);
        const unsigned kx = fkey(U[i]);
        atomicAdd(&cnt[b], 1u); atomicMin(&bxmin[b], kx); atomicMax(&bxmax[b], kx);
    }
    __syncthreads();
    if (warp == 0) {
        // A strip is closed once a cell of `density` points would be at least as tall as it is wide: count x height >= k x extent.
        // One warp grows the strips: lane l looks at "the strip ends with bin b + l" (running count / x-extent by warp scans).
        const unsigned FULL = 0xFFFFFFFFu;
        const float Lmin = 0.2f * (ps.xmax - ps.xmin);
        int R = 0, b = 0, start = 0, ncell = 0;
        unsigned c0 = 0, lo0 = 0xFFFFFFFFu, hi0 = 0u;              // carried over when a strip is longer than 32 bins
        while (b < NB) {
            const int idx = b + lane; const bool in = idx < NB;
            unsigned c = in ? cnt[idx] : 0u, lo = in ? bxmin[idx] : 0xFFFFFFFFu, hi = in ? bxmax[idx] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned uc = __shfl_up_sync(FULL, c, o), ul = __shfl_up_sync(FULL, lo, o), uh = __shfl_up_sync(FULL, hi, o);
                if (lane >= o) { c += uc; lo = min(lo, ul); hi = max(hi, uh); }
            }
            c += c0; lo = min(lo, lo0); hi = max(hi, hi0);
            const float L = c ? fmaxf(funkey(hi) - funkey(lo), Lmin) : Lmin;
            const bool stop = in && (idx == NB - 1 || !(density * L - (float)c * ((float)(idx + 1 - start) * ps.bh) > 0.f));
            const unsigned sm = __ballot_sync(FULL, stop);
            if (!sm) { c0 = __shfl_sync(FULL, c, 31); lo0 = __shfl_sync(FULL, lo, 31); hi0 = __shfl_sync(FULL, hi, 31); b += 32; continue; }
            const int e = __ffs(sm) - 1, end = b + e;            // the strip is bins start .. end
            const unsigned cs = __shfl_sync(FULL, c, e), klo = __shfl_sync(FULL, lo, e), khi = __shfl_sync(FULL, hi, e);
            for (int q = start + lane; q <= end; q += 32) bin_row[q] = (uint16_t)R;
            if (lane == 0) {
                const int nc = (int)(cs >> 1) + 1;                 // sub-cells of about two points
                const float x0 = funkey(klo), ext = funkey(khi) - x0;
                row_bin[R] = (uint16_t)start; rowoff[R] = (uint16_t)ncell; rowxi[R] = make_float2(x0, ext > 0.f ? (float)nc / ext : 0.f);
            }
            ncell += (int)(cs >> 1) + 1;
            ++R; b = end + 1; start = b; c0 = 0; lo0 = 0xFFFFFFFFu; hi0 = 0u;
        }
        if (lane == 0) { row_bin[R] = (uint16_t)NB; rowoff[R] = (uint16_t)ncell; ctl->ps.R = R; }
    }
    __syncthreads();
    const int R = ps.R, NC = rowoff[R];
    // sub-cell of a point: strips in order, sub-cells left to right inside a strip
    auto subcell = [&](float x, float y) {
        const int r = bin_row[bin_of(ps, y)];
        return rowoff[r] + strip_cell(rowxi[r], rowoff[r + 1] - rowoff[r], x);
    };
    for (int i = tid; i <= NC; i += NT) cc[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) atomicAdd(&cc[subcell(U[i], V[i])], 1u);
    __syncthreads();
    block_excl_scan(cc, NC, ctl->warp_cnt);
    for (int i = tid; i <= NC; i += NT) cstart[i] = (uint16_t)cc[i];
    for (int r = tid; r <= R; r += NT) row_start[r] = (uint16_t)cc[rowoff[r]];
    __syncthreads();
    for (int i = tid; i < n; i += NT) tmp[atomicAdd(&cc[subcell(U[i], V[i])], 1u)] = (uint16_t)i;
    __syncthreads();
    for (int a = tid; a < n; a += NT) {
        const int i = tmp[a];
        const float xi = U[i], yi = V[i];
        const int sc = subcell(xi, yi), b0 = cstart[sc], e0 = cstart[sc + 1];
        int rank = 0;
        for (int j = b0; j < e0; ++j) { const int q = tmp[j]; const float xq = U[q]; rank += (xq < xi) || (xq == xi && q < i); }
        ga.sorig[b0 + rank] = (uint16_t)i; ga.sx[b0 + rank] = xi; ga.sy[b0 + rank] = yi;
    }
    __syncthreads();
    // exact duplicates (same strip, equal x: adjacent up to other points of that x): all but the lowest index become holes
    // (Qhull drops duplicates too, into .coplanar)
    for (int a = tid; a < n; a += NT) {
        const int b0 = row_start[bin_row[bin_of(ps, ga.sy[a])]];
        bool dup = false;
        for (int c = a - 1; c >= b0 && ga.sx[c] == ga.sx[a]; --c) if (ga.sy[c] == ga.sy[a]) { dup = true; break; }
        tmp[a] = dup;
    }
    __syncthreads();
    int ndup = 0;
    for (int a = tid; a < n; a += NT) if (tmp[a]) { pflag[ga.sorig[a]] |= 1; ga.sorig[a] = INF16; ++ndup; }
    if (ndup) atomicAdd(&ctl->n_dup, ndup);
    __syncthreads();
}
#endif

#ifndef MVOSR_UNIFORM_GRID
#ifndef MVOSR_FILTER_MIN_PCT
#define MVOSR_FILTER_MIN_PCT 60
#endif
constexpr int FILTER_MIN_PCT = MVOSR_FILTER_MIN_PCT;       // filter the index of Delaunay #1 when at least this share of the points survives, else build anew
// Strip index of a SUBSET of the indexed points (Delaunay #2 runs over the survivors of the graph check): instead of building it
// again from the feature arrays -- histogram, strips, counting sort, rank sort, duplicates: 31 k cycles per frame -- the sorted copy
// is filtered in place.  Strips, bins and sub-cells stay as they are (an order-preserving filter keeps every run sorted by (x,
// index): the new feature indices are monotone in the old ones); row_start / cell_start become the survivor counts before their old
// values.  The bounding box stays the old one (a superset: only used conservatively).  newidx[old feature] = new feature index or
// INF16; tmp: cap + 1 uint16 of scratch.  The caller uses this while most points survive (the strips were sized for the old density).
__device__ __forceinline__ void filter_grid(int n_old, int n_new, const uint16_t *newidx, GridArrays ga, Ctl *ctl, uint16_t *tmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SortedSet &ps = ctl->ps;
    uint16_t *row_start = (uint16_t *)ps.row_start, *cstart = (uint16_t *)ps.cell_start;
    // tmp[i] = survivors among the sorted positions before i
    int run = 0;
    for (int c0 = 0; c0 < n_old; c0 += NT) {
        const int i = c0 + tid;
        bool k = false;
        if (i < n_old) { const int o = ga.sorig[i]; k = o != INF16 && newidx[o] != INF16; }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, k);
        if (lane == 0) ctl->warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int woff, tot;
        warp_offsets(ctl->warp_cnt, warp, lane, woff, tot);
        if (i < n_old) tmp[i] = (uint16_t)(run + woff + __popc(bal & ((1u << lane) - 1u)));
        run += tot;
        __syncthreads();
    }
    if (tid == 0) tmp[n_old] = (uint16_t)run;
    __syncthreads();
    const int NC = ps.row_cell[ps.R];
    for (int c = tid; c <= NC; c += NT) cstart[c] = tmp[cstart[c]];
    for (int r = tid; r <= ps.R; r += NT) row_start[r] = tmp[row_start[r]];
    // the sorted copy, in place: every chunk is read before any of its (smaller or equal) target positions is written
    for (int c0 = 0; c0 < n_old; c0 += NT) {
        const int i = c0 + tid;
        float x = 0.f, y = 0.f; int no = INF16, pos = 0;
        if (i < n_old) {
            const int o = ga.sorig[i];
            if (o != INF16) no = newidx[o];
            if (no != INF16) { x = ga.sx[i]; y = ga.sy[i]; pos = tmp[i]; }
        }
        __syncthreads();
        if (no != INF16) { ga.sx[pos] = x; ga.sy[pos] = y; ga.sorig[pos] = (uint16_t)no; }
        __syncthreads();
    }
    if (tid == 0) ps.n = n_new;
    __syncthreads();
}
#endif

// write this frame's triangle list to global memory in canonical order
__device__ __forceinline__ void write_canonical(int n, const FrameView &fv, uint32_t *scr, int *warp_tmp,
                                                int32_t *out, int32_t *n_out) {
    int tid = threadIdx.x;
    for (int i = tid; i < n; i += NT) scr[i] = (fv.pflag[i] & 1) ? 0u : fv.tcnt[i];
    __syncthreads();
    block_excl_scan(scr, n, warp_tmp);
    for (int p = tid; p < n; p += NT) {
        if (fv.pflag[p] & 1) continue;
        int k = fv.tcnt[p], o = (int)scr[p], b = fv.tbase[p];
        for (int j = 0; j < k; ++j) {
            out[3 * (o + j) + 0] = fv.tri[3 * (b + j) + 0];
            out[3 * (o + j) + 1] = fv.tri[3 * (b + j) + 1];
            out[3 * (o + j) + 2] = fv.tri[3 * (b + j) + 2];
        }
    }
    if (tid == 0 && n_out) *n_out = (int)scr[n];
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// k-th smallest of the loose heights (radix select over float64 bit patterns, all positive)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_select(const double *h, const uint8_t *flags, int T, int k, Ctl *ctl) {
    // returns the key of rank k (0-based) among {h[t] : flags[t]&1}
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned long long prefix = 0;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += NT) ctl->hist[i] = 0;
        __syncthreads();
        const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int t = tid; t < T; t += NT) {
            if (!(flags[t] & 1)) continue;
            const unsigned long long key = (unsigned long long)__double_as_longlong(h[t]);
            if ((key & mask) == prefix) atomicAdd(&ctl->hist[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // warp 0: lane l owns bins 8l..8l+7; the bin holding rank k is where the running count first exceeds k
            unsigned c[8]; int sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = ctl->hist[8 * lane + i]; sum += (int)c[i]; }
            const int inc = warp_incl_scan(sum, lane);
            const unsigned before = __ballot_sync(0xFFFFFFFFu, inc <= k);      // lanes entirely below rank k (a prefix of the lanes)
            const int owner = __popc(before);
            if (lane == owner) {
                int acc = inc - sum, b = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { if (acc + (int)c[i] <= k) { acc += (int)c[i]; b = i + 1; } else break; }
                ctl->sel_k = k - acc;
                ctl->sel_prefix = prefix | ((unsigned long long)(8 * lane + b) << shift);
            }
        }
        __syncthreads();
        k = ctl->sel_k; prefix = ctl->sel_prefix;
        __syncthreads();
    }
    return prefix;
}

// given the key a of rank k, the key of rank k + 1: a itself if more than k + 1 values are <= a, else the smallest larger one
__device__ __forceinline__ unsigned long long block_select_next(const double *h, const uint8_t *flags, int T, int k, unsigned long long a, Ctl *ctl) {
    const int tid = threadIdx.x;
    if (tid == 0) { ctl->hist[0] = 0; ctl->sel_prefix = ~0ull; }
    __syncthreads();
    unsigned cnt = 0; unsigned long long mn = ~0ull;
    for (int t = tid; t < T; t += NT) {
        if (!(flags[t] & 1)) continue;
        const unsigned long long key = (unsigned long long)__double_as_longlong(h[t]);
        if (key <= a) ++cnt; else mn = key < mn ? key : mn;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        const unsigned long long m2 = __shfl_xor_sync(0xFFFFFFFFu, mn, o); mn = m2 < mn ? m2 : mn;
    }
    if ((tid & 31) == 0) { if (cnt) atomicAdd(&ctl->hist[0], cnt); if (mn != ~0ull) atomicMin(&ctl->sel_prefix, mn); }
    __syncthreads();
    const unsigned long long res = (int)ctl->hist[0] >= k + 2 ? a : ctl->sel_prefix;
    __syncthreads();
    return res;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int SRC>
__global__ void __launch_bounds__(NT, 1) frame_kernel(FrameParams P) {
    constexpr bool FROM_CORR = SRC == SRC_CORR, F64 = SRC == SRC_F64;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ Ctl ctl;
    // frames beyond the shared-memory capacity are staged in a per-CTA slab of global memory (L2-resident): same code
    unsigned char *smem = P.workspace ? P.workspace + (size_t)blockIdx.x * P.ws_stride : dyn_smem;
    const SmemPlan pl = make_plan(P.cap);
    const int cap = P.cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *X = (float *)(smem + pl.off_X), *Y = (float *)(smem + pl.off_Y), *Z = (float *)(smem + pl.off_Z);
    float *U = (float *)(smem + pl.t_U), *V = (float *)(smem + pl.t_V);
    uint8_t *pflag = smem + pl.off_pflag;
    GridArrays ga;
    ga.sx = (float *)(smem + pl.h_sx); ga.sy = (float *)(smem + pl.h_sy); ga.sorig = (uint16_t *)(smem + pl.h_sorig);
    ga.cell_start = (uint16_t *)(smem + pl.h_cell_start); ga.scr = (uint32_t *)(smem + pl.off_scr);
    uint16_t *defer = (uint16_t *)(smem + pl.h_defer), *defer2 = defer + cap;
    uint16_t *mult = (uint16_t *)(smem + pl.off_mult);
    double *theight = (double *)(smem + pl.off_H);           // aliases the sorted copy (dead by then)
    uint8_t *tflags = smem + pl.off_tflags;
    FrameView fv;
    fv.Z = Z; fv.pflag = pflag; fv.tri = (uint16_t *)(smem + pl.off_T);
    fv.tbase = (uint16_t *)(smem + pl.off_tbase); fv.tcnt = smem + pl.off_tcnt;
    fv.T = &ctl.T; fv.status = &ctl.status; fv.pass_mask = P.cfg.graph_pass_mask; fv.tri_cap = 2 * cap;
    fv.f3d = nullptr; fv.f2d = nullptr; fv.srcidx = nullptr;
    uint32_t *const SRCI = (uint32_t *)X;                        // SRC_F64: X's storage holds the feature's index in the frame's float64 arrays (X, Y unused)
    fv.rpool = nullptr; fv.rinfo = (uint32_t *)(smem + pl.off_rinfo); fv.rcount = &ctl.rcount; fv.rpool_cap = 6 * cap; fv.oldof = nullptr;
    uint16_t *const inv = (uint16_t *)(smem + pl.off_scr), *const oldof = inv + cap;     // new index -> sorted position / old index (scr is free between the grid build and the planes)
    uint16_t *const rpool = (uint16_t *)(smem + pl.off_rpool);
    uint16_t *const todo = (uint16_t *)(smem + pl.off_tflags);   // list of the stars Delaunay #2 must rebuild (tflags is free until the planes)
    const mvosr_config &cfg = P.cfg;
    const float density = cfg.reserved[0] > 0 ? 0.01f * (float)cfg.reserved[0] : GRID_DENSITY;     // tuning knob: mean points per grid cell x 100
#ifdef MVOSR_UNIFORM_GRID
    const int win_m = 0; const float wfac = 0.f;
#else
    const int win_m = WIN_M;
    const float wfac = cfg.reserved[1] > 0 ? 0.01f * (float)cfg.reserved[1] : WIN_FACTOR;            // tuning knob: half-width of the candidate window in cell sides x 100
#endif

    long long tlast = 0;
#define TMARK(k) do { if (tid == 0) { long long tn_ = clock64(); ctl.tphase[k] += tn_ - tlast; tlast = tn_; } } while (0)
    for (;;) {
        __syncthreads();
        if (tid == 0) ctl.frame = atomicAdd(P.work_counter, 1);
        __syncthreads();
        if (ctl.frame >= P.n_frames) break;
        const int f = P.order ? P.order[ctl.frame] : ctl.frame;
        if (tid == 0) {
            ctl.n_roi = ctl.n_feat = ctl.status = ctl.bad = ctl.singular = ctl.n_dup = ctl.n_kept = ctl.T = 0; ctl.n_dup1 = -1;
            ctl.n_exact = ctl.n_deferred_total = 0; ctl.rcount = 0; ctl.n_todo = 0;
            ctl.n_loose = ctl.n_tight = ctl.n_valid = 0; ctl.best_hyp = -1; ctl.best_ic = 0; ctl.hyps_used = 0;
            ctl.n_degenerate = 0; ctl.height_level = CUDART_NAN;
            for (int k = 0; k < 16; ++k) ctl.tphase[k] = 0;
            for (int k = 0; k < 8; ++k) ctl.sc.cnt[k] = 0;
            ctl.sc.n_wrap = 0;
            tlast = clock64();
        }
        __syncthreads();
        const int base = P.offsets[f];
        const int n_in = P.counts ? P.counts[f] : (P.offsets[f + 1] - base);
        if (F64) { fv.f3d = P.f3d; fv.f2d = P.f2d; fv.srcidx = SRCI; fv.fbase = base; }
        // feature i of the staged frame as float64 (x, y, z): the staged float32 values, or the caller's float64 ones
        auto point3 = [&](int i, double &px, double &py, double &pz) {
            if (F64) { const double *q = P.f3d + 3 * ((size_t)base + SRCI[i]); px = q[0]; py = q[1]; pz = q[2]; }
            else { px = X[i]; py = Y[i]; pz = Z[i]; }
        };

        // ---------------- load + (stage 1) + ROI cut, order preserving ----------------
        Pose pose;
        if (FROM_CORR) {
#pragma unroll
            for (int i = 0; i < 9; ++i) pose.R[i] = P.poses[12 * f + (i / 3) * 4 + (i % 3)];
            pose.t[0] = P.poses[12 * f + 3]; pose.t[1] = P.poses[12 * f + 7]; pose.t[2] = P.poses[12 * f + 11];
        }
        int total = 0, total_feat = 0;
        for (int c0 = 0; c0 < n_in; c0 += NT) {
            int i = c0 + tid;
            bool feat = false, ok = false;
            float fu = 0, fv_ = 0, fx3 = 0, fy3 = 0, fz3 = 0;
            if (i < n_in) {
                if (FROM_CORR) {
                    double X3, Y3, Z3, uu, vv;
                    feat = triangulate_point(P.cur_u[base + i], P.cur_v[base + i], P.ref_u[base + i], P.ref_v[base + i], pose,
                                             cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.triangulation_max_depth, X3, Y3, Z3, uu, vv);
                    if (P.e_mask) feat = feat && P.e_mask[base + i] != 0;
                    fx3 = (float)X3; fy3 = (float)Y3; fz3 = (float)Z3; fu = (float)uu; fv_ = (float)vv;
                } else if (F64) {
                    feat = true;
                    const double v64 = P.f2d[2 * (size_t)(base + i) + 1];
                    ok = v64 > (double)cfg.vanish;                  // the ROI cut on the float64 value (rescale.py:115)
                    fv_ = (float)v64;
                    if (ok) {
                        fu = (float)P.f2d[2 * (size_t)(base + i)];
                        const double *p3 = P.f3d + 3 * (size_t)(base + i);
                        fz3 = (float)p3[2];
                        if (!isfinite(p3[0]) || !isfinite(p3[1]) || !isfinite(p3[2])) ctl.bad = 1;
                        fx3 = __uint_as_float((unsigned)i);          // -> SRCI
                    }
                } else {
                    feat = true;
                    fv_ = P.v[base + i];
                }
                if (!F64) ok = feat && (P.mode == MODE_DT_ONLY || fv_ > cfg.vanish);
                if (ok && SRC == SRC_F32) {
                    fu = P.u[base + i];
                    if (P.mode != MODE_DT_ONLY) { fx3 = P.x[base + i]; fy3 = P.y[base + i]; fz3 = P.z[base + i]; }
                }
            }
            unsigned bal = __ballot_sync(0xFFFFFFFFu, ok), balf = __ballot_sync(0xFFFFFFFFu, feat);
            if (lane == 0) { ctl.warp_cnt[warp] = __popc(bal); ctl.warp_cnt2[warp] = __popc(balf); }
            __syncthreads();
            int woff, tot, woff2, totf;
            warp_offsets(ctl.warp_cnt, warp, lane, woff, tot);
            warp_offsets(ctl.warp_cnt2, warp, lane, woff2, totf);
            if (ok) {
                int pos = total + woff + __popc(bal & ((1u << lane) - 1u));
                if (!(fabsf(fu) < 4096.f) || !(fabsf(fv_) < 4096.f)) ctl.bad = 1;
                if (!F64 && P.mode != MODE_DT_ONLY && !(isfinite(fx3) && isfinite(fy3) && isfinite(fz3))) ctl.bad = 1;   // NaN / inf depth
                if (fabsf(fu) < 7.62939453125e-06f) fu = 0.f;       // 2^-17: keeps every difference exact in float64
                if (fabsf(fv_) < 7.62939453125e-06f) fv_ = 0.f;
                if (pos < cap) { U[pos] = fu; V[pos] = fv_; X[pos] = fx3; Y[pos] = fy3; Z[pos] = fz3; pflag[pos] = 0; }
            }
            total += tot; total_feat += totf;
            __syncthreads();
        }
        TMARK(0);
        const int n_feat = total_feat;
        int n = total;
        int status = 0;
        if (P.gate && n_feat <= cfg.min_features) status |= MVOSR_ST_SKIPPED;
        if (n > cap) status |= MVOSR_ST_OVERFLOW;
        if (ctl.bad) status |= MVOSR_ST_BAD_INPUT;
        if (n < 3) status |= MVOSR_ST_FEW_ROI;

        bool second = false;
        const int n1 = n;
        int n_exact = 0;
        if (!status) {
            // ---------------- Delaunay #1 -> graph vote (or triangles in DT-only mode) ----------------
            build_grid(n, U, V, pflag, ga, cap, &ctl, density, defer, win_m, wfac);
            TMARK(1);
            if (P.mode == MODE_DT_ONLY) {
                for (int i = tid; i < n; i += NT) { fv.tcnt[i] = 0; fv.tbase[i] = 0; }
                __syncthreads();
                int nd = run_stars<true>(ctl.ps, fv, &ctl.sc, defer, defer2, n_exact, &ctl.tphase[3]);
                if (tid == 0) ctl.n_deferred_total += nd;
            } else {
                for (int i = tid; i < n; i += NT) fv.rinfo[i] = 0;
                __syncthreads();
                fv.rpool = rpool;                                  // the vote pass also stores every finished star
                int nd = run_stars<false>(ctl.ps, fv, &ctl.sc, defer, defer2, n_exact, &ctl.tphase[3]);
                fv.rpool = nullptr;
                if (tid == 0) ctl.n_deferred_total += nd;
                __syncthreads();
#ifndef MVOSR_UNIFORM_GRID
                votes_from_rings(n, V, fv, rpool);                 // the graph votes of the stored stars, one star per thread
#endif
            }
            __syncthreads();
            status |= ctl.status;
            TMARK(2);
        }
        if (!status && P.mode == MODE_DT_ONLY) {
            write_canonical(n, fv, ga.scr, ctl.warp_cnt, P.tri_out + 3 * (size_t)(2 * base), P.n_tri_out + f);
            if (ctl.T == 0) status |= MVOSR_ST_FEW_ROI;
        }
        if (!status && P.mode == MODE_FULL) {
            // ---------------- keep mask, survivor rule (rescale.py:131-137) ----------------
            int cnt = 0;
            for (int i = tid; i < n; i += NT) cnt += (pflag[i] >> 1) & 1;
#pragma unroll
            for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (lane == 0 && cnt) atomicAdd(&ctl.n_kept, cnt);
            __syncthreads();
            const int n_kept = ctl.n_kept;
            if (P.has_dbg && P.dbg.keep) for (int i = tid; i < n; i += NT) P.dbg.keep[base + i] = (pflag[i] >> 1) & 1;
            second = n_kept > cfg.min_kept;
            if (P.has_dbg && P.dbg.tri1) {
                // parity probe: also materialise Delaunay #1 (re-runs the stars in emit mode; the triangle list
                // overwrites U,V, which are restored from the sorted copy afterwards)
                for (int i = tid; i < n; i += NT) { fv.tcnt[i] = 0; fv.tbase[i] = 0; }
                __syncthreads();
                run_stars<true>(ctl.ps, fv, &ctl.sc, defer, defer2, n_exact, nullptr);
                write_canonical(n, fv, ga.scr, ctl.warp_cnt, P.dbg.tri1 + 3 * (size_t)(2 * base), P.dbg.n_tri1 ? P.dbg.n_tri1 + f : nullptr);
                if (tid == 0) ctl.T = 0;
                for (int i = tid; i < n; i += NT) { int o = ga.sorig[i]; if (o != INF16) { U[o] = ga.sx[i]; V[o] = ga.sy[i]; } }
                __syncthreads();
            }
            if (second) {
                // order-preserving in-place compaction by the keep flag, chunk by chunk
                int run = 0;
                for (int c0 = 0; c0 < n; c0 += NT) {
                    int i = c0 + tid;
                    bool k = i < n && (pflag[i] & 2);
                    float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
                    if (k) { a0 = U[i]; a1 = V[i]; a2 = X[i]; a3 = Y[i]; a4 = Z[i]; }
                    unsigned bal = __ballot_sync(0xFFFFFFFFu, k);
                    if (lane == 0) ctl.warp_cnt[warp] = __popc(bal);
                    __syncthreads();
                    int woff, tot;
                    warp_offsets(ctl.warp_cnt, warp, lane, woff, tot);
                    if (k) {
                        int pos = run + woff + __popc(bal & ((1u << lane) - 1u));
                        U[pos] = a0; V[pos] = a1; X[pos] = a2; Y[pos] = a3; Z[pos] = a4;
                        mult[i] = (uint16_t)pos;                   // old -> new feature index
                    } else if (i < n) mult[i] = INF16;
                    run += tot;
                    __syncthreads();
                }
                n = n_kept;
                for (int i = tid; i < n; i += NT) pflag[i] = 0;
                if (tid == 0) { ctl.n_dup1 = ctl.n_dup; ctl.n_dup = 0; }
                __syncthreads();
#ifndef MVOSR_UNIFORM_GRID
                if (100 * n_kept >= FILTER_MIN_PCT * n1) filter_grid(n1, n, mult, ga, &ctl, defer);          // most points survive: filter the index
                else
#endif
                build_grid(n, U, V, pflag, ga, cap, &ctl, density, defer, win_m, wfac);
            } else {
                for (int i = tid; i < n; i += NT) pflag[i] &= 1;
                __syncthreads();
            }
            // ---------------- Delaunay #2 (or #1 again) with triangle emission ----------------
            for (int i = tid; i < n; i += NT) { fv.tcnt[i] = 0; fv.tbase[i] = 0; }
            __syncthreads();
            if (second) {
                // A star of Delaunay #1 none of whose neighbours was dropped is a star of Delaunay #2 too (its triangles keep
                // their empty circles and still close the fan around the point; the symbolic tie-break depends on the relative
                // index order only, which the compaction preserves): emit its triangles from the stored ring, one thread per
                // star, and rebuild only the others.
                // The ring (at most RD entries; larger stars are rebuilt) is held in registers as NEW feature indices.
                // The rings of the stars to rebuild are rewritten in place as sorted positions of the new grid (dropped
                // neighbours marked): they seed the rebuild (stars_pair).
                constexpr int RD = 10;
                for (int i = tid; i < n; i += NT) { const int og = ctl.ps.orig[i]; if (og != INF16) inv[og] = (uint16_t)i; }
                __syncthreads();
                for (int c0 = 0; c0 < n1; c0 += NT) {
                    const int o = c0 + tid;
                    int np = INF16, d = 0, dfull = 0;
                    if (o < n1) np = mult[o];
                    uint32_t info = 0;
                    if (np != INF16) { info = fv.rinfo[o]; if (info & RING_PARTIAL) info = 0; dfull = d = (int)(info & 0xFFu); if (d > RD) d = 0; oldof[np] = (uint16_t)o; }
                    uint16_t *ring = rpool + (info >> 8);
                    int nb[RD];                                    // new index of ring entry j; INF16: hull gap; -1: dropped
                    bool clean = d > 0;
#pragma unroll
                    for (int j = 0; j < RD; ++j) {
                        nb[j] = INF16;
                        if (j < d) { const int q = ring[j]; if (q != INF16) { const int m = mult[q]; nb[j] = m == INF16 ? -1 : m; } }
                        clean = clean && nb[j] >= 0;
                    }
                    if (!clean) {
                        d = 0;
                        for (int j = 0; j < dfull; ++j) {
                            const int q = ring[j];
                            if (q != INF16) { const int m = mult[q]; ring[j] = m == INF16 ? RING_DROPPED : inv[m]; }
                        }
                    }
                    // owned triangles (np smaller than both other vertices): key = (min << 16) | max
                    unsigned key[RD]; int k = 0;
#pragma unroll
                    for (int j = 0; j < RD; ++j) {
                        int nx = nb[0];
#pragma unroll
                        for (int i = 1; i < RD; ++i) if (i == j + 1 && i < d) nx = nb[i];       // nb[(j + 1) % d]
                        const int a = nb[j];
                        const bool own = j < d && a != INF16 && nx != INF16 && np < a && np < nx;
                        key[j] = own ? (((unsigned)min(a, nx) << 16) | (unsigned)max(a, nx)) : 0xFFFFFFFFu;
                        k += own;
                    }
                    if (clean) pflag[np] |= 4;
                    // one atomic per warp for the triangle blocks of its stars
                    const int inc = warp_incl_scan(k, lane), wtot = __shfl_sync(0xFFFFFFFFu, inc, 31);
                    int tb = 0;
                    if (lane == 31 && wtot) tb = atomicAdd(&ctl.T, wtot);
                    tb = __shfl_sync(0xFFFFFFFFu, tb, 31) + inc - k;
                    if (!k) continue;
                    if (tb + k > fv.tri_cap) { atomicOr(&ctl.status, MVOSR_ST_OVERFLOW); continue; }
#pragma unroll
                    for (int j = 0; j < RD; ++j) {
                        if (key[j] == 0xFFFFFFFFu) continue;
                        int r = 0;                                 // rank inside the block: keys of one star are distinct
#pragma unroll
                        for (int i = 0; i < RD; ++i) r += key[i] < key[j];
                        uint16_t *t = fv.tri + 3 * (tb + r);
                        t[0] = (uint16_t)np; t[1] = (uint16_t)(key[j] >> 16); t[2] = (uint16_t)(key[j] & 0xFFFFu);
                    }
                    fv.tbase[np] = (uint16_t)tb; fv.tcnt[np] = (uint8_t)k;
                }
                __syncthreads();
                // the stars to rebuild, as sorted positions, hull rows first (see stars_pair)
                const SortedSet &ps = ctl.ps;
#ifdef MVOSR_UNIFORM_GRID
                const int rot = ps.cell_start[(ps.gy - 1) * ps.gx];
#else
                const int rot = ps.row_start[ps.R - 1];
#endif
                int run = 0;
                for (int c0 = 0; c0 < n; c0 += NT) {
                    const int j = c0 + tid;
                    int i = j + rot; if (i >= n) i -= n;
                    bool k = false;
                    if (j < n) { const int og = ps.orig[i]; k = og != INF16 && !(pflag[og] & 4); }
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, k);
                    if (lane == 0) ctl.warp_cnt[warp] = __popc(bal);
                    __syncthreads();
                    int woff, tot;
                    warp_offsets(ctl.warp_cnt, warp, lane, woff, tot);
                    if (k) todo[run + woff + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
                    run += tot;
                    __syncthreads();
                }
                if (tid == 0) ctl.n_todo = run;
                __syncthreads();
            }
            TMARK(6);
            if (second) { fv.rpool = rpool; fv.oldof = oldof; }
            int nd = run_stars<true>(ctl.ps, fv, &ctl.sc, defer, defer2, n_exact, &ctl.tphase[8], second ? todo : nullptr, ctl.n_todo);
            fv.rpool = nullptr; fv.oldof = nullptr;
            if (tid == 0) ctl.n_deferred_total += nd;
            __syncthreads();
            status |= ctl.status;
            TMARK(7);
        }
        const int T = ctl.T;
        if (!status && P.mode == MODE_FULL) {
            if (T == 0) status |= MVOSR_ST_FEW_ROI;
        }
        if (!status && P.mode == MODE_FULL) {
            // ---------------- per-triangle plane, pitch and height gates (rescale.py:75-96) ----------------
            int c_loose = 0, c_tight = 0;
            for (int t = tid; t < T; t += NT) {
                int i0 = fv.tri[3 * t], i1 = fv.tri[3 * t + 1], i2 = fv.tri[3 * t + 2];
                double p0x, p0y, p0z, p1x, p1y, p1z, p2x, p2y, p2z;
                point3(i0, p0x, p0y, p0z); point3(i1, p1x, p1y, p1z); point3(i2, p2x, p2y, p2z);
                double e1x = p1x - p0x, e1y = p1y - p0y, e1z = p1z - p0z;
                double e2x = p2x - p0x, e2y = p2y - p0y, e2z = p2z - p0z;
                // n = P^-1 . 1 = (e1 x e2) / (p0 . (e1 x e2))
                double cx = e1y * e2z - e1z * e2y, cy = e1z * e2x - e1x * e2z, cz = e1x * e2y - e1y * e2x;
                double det = p0x * cx + p0y * cy + p0z * cz;
                if (det == 0.0) ctl.singular = 1;                 // the reference's np.matrix(...).I raises LinAlgError (rescale.py:79)
                double clen = sqrt(cx * cx + cy * cy + cz * cz);
                double hgt = fabs(det) / clen;                     // 1/|n|
                double s = -(det < 0 ? -cy : cy) / clen;          // -n_y/|n| = sin(pitch)
                int fl = 0;
                if (s <= cfg.sin_loose) fl |= 1;
                if (s <= cfg.sin_tight) fl |= 2;
                theight[t] = hgt; tflags[t] = (uint8_t)fl;
                c_loose += fl & 1; c_tight += (fl >> 1) & 1;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                c_loose += __shfl_xor_sync(0xFFFFFFFFu, c_loose, o); c_tight += __shfl_xor_sync(0xFFFFFFFFu, c_tight, o);
            }
            if (lane == 0) { atomicAdd(&ctl.n_loose, c_loose); atomicAdd(&ctl.n_tight, c_tight); }
            __syncthreads();
            if (ctl.singular) status |= MVOSR_ST_SINGULAR;
            TMARK(10);
            const int n_loose = ctl.n_loose;
            // height_level = 0.9 * median(height[loose])  (np.median: mean of the two middle values for even counts)
            double level = CUDART_NAN;
            if (n_loose > 0) {
                unsigned long long ka = block_select(theight, tflags, T, (n_loose - 1) / 2, &ctl);
                double a = __longlong_as_double((long long)ka), b = a;
                if ((n_loose & 1) == 0) {
                    unsigned long long kb = block_select_next(theight, tflags, T, (n_loose - 1) / 2, ka, &ctl);
                    b = __longlong_as_double((long long)kb);
                }
                double med = (n_loose & 1) ? a : (a + b) / 2.0;
                level = cfg.height_level_factor * med;
            }
            if (tid == 0) ctl.height_level = level;
            TMARK(11);
            // ---------------- valid triangles -> canonical vertex list, multiplicities ----------------
            for (int i = tid; i < n + 2; i += NT) mult[i] = 0;
            __syncthreads();
            for (int p = tid; p < n; p += NT) {
                int k = fv.tcnt[p], b = fv.tbase[p], kv = 0;
                for (int j = 0; j < k; ++j) {
                    int t = b + j;
                    bool valid = (tflags[t] & 2) && (theight[t] > level);
                    if (valid) { tflags[t] |= 4; ++kv; }
                }
                ga.scr[p] = (uint32_t)kv;
            }
            __syncthreads();
            block_excl_scan(ga.scr, n, ctl.warp_cnt);
            const int n_valid = (int)ga.scr[n];
            for (int t = tid; t < T; t += NT) {
                if (!(tflags[t] & 4)) continue;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int vtx = fv.tri[3 * t + c];
                    atomicAdd((unsigned int *)mult + (vtx >> 1), (vtx & 1) ? 0x10000u : 1u);
                }
            }
            __syncthreads();
            if (tid == 0) ctl.n_valid = n_valid;
            const int n_sel = 3 * n_valid;
            TMARK(12);

            // ---------------- RANSAC over the vertex list (ransac.py:3-23) ----------------
            // One warp per hypothesis, NWARP hypotheses per round; after each round every thread replays the
            // sequential bookkeeping of run_ransac over that round's counts (keep the first strictly larger ic,
            // stop at the first ic > goal), so the result is the one the sequential loop returns.
            double m_a = CUDART_NAN, m_b = CUDART_NAN, m_c = CUDART_NAN, m_d = CUDART_NAN;
            if (n_sel >= cfg.min_selected && !(status & MVOSR_ST_SINGULAR)) {
                const double goal = (double)n_sel * cfg.ransac_goal_fraction;
                const int H = cfg.ransac_iterations;
                int h_done = 0, best = -1, best_ic = 0, ndeg = 0, used = 0;
                bool stop = false;
                double b_nx = 0, b_ny = 0, b_nz = 0, b_dd = 0, b_n4 = 1;
                while (h_done < H && !stop) {
                    const int h = h_done + warp;
                    double nx = 0, ny = 0, nz = 0, dd = 0, n4 = 1;
                    if (h < H) {
                        uint32_t posv[3];
                        sample3_positions(P.seed, (uint32_t)h, (uint32_t)(P.frame_index ? P.frame_index[f] : P.frame_index0 + f),
                                          (uint32_t)(P.frame_seq ? P.frame_seq[f] : P.seq_id), (uint32_t)n_sel,
                                          posv[0], posv[1], posv[2]);
                        int vtx[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            uint32_t r = posv[c] / 3u, corner = posv[c] % 3u;
                            // largest p with scr[p] <= r  (scr = exclusive scan of valid-triangle counts per emitting point)
                            int lo = 0, hi = n;        // scr[n] = n_valid > r
                            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ga.scr[mid] <= r) lo = mid; else hi = mid; }
                            int k = (int)(r - ga.scr[lo]);
                            int t = fv.tbase[lo];
                            for (;; ++t) { if (tflags[t] & 4) { if (k == 0) break; --k; } }
                            vtx[c] = fv.tri[3 * t + corner];
                        }
                        int ic = 0;
                        bool degenerate = (vtx[0] == vtx[1]) || (vtx[0] == vtx[2]) || (vtx[1] == vtx[2]);
                        double p0x, p0y, p0z, p1x, p1y, p1z, p2x, p2y, p2z;
                        point3(vtx[0], p0x, p0y, p0z); point3(vtx[1], p1x, p1y, p1z); point3(vtx[2], p2x, p2y, p2z);
                        double e1x = p1x - p0x, e1y = p1y - p0y, e1z = p1z - p0z;
                        double e2x = p2x - p0x, e2y = p2y - p0y, e2z = p2z - p0z;
                        // null vector of [p 1] (3x4) in closed form: (n, -n.p0), n = e1 x e2 (estimate_road_norm.py:13-15)
                        nx = e1y * e2z - e1z * e2y; ny = e1z * e2x - e1x * e2z; nz = e1x * e2y - e1y * e2x;
                        dd = -(nx * p0x + ny * p0y + nz * p0z);
                        n4 = sqrt(nx * nx + ny * ny + nz * nz + dd * dd);
                        if (!(nx * nx + ny * ny + nz * nz > 0)) degenerate = true;
                        if (!degenerate) {
                            const double thr = cfg.ransac_threshold * n4;   // |m.[x 1]| < thr for the unit-4-norm model m
                            // two points per lane and iteration, no data-dependent branch: the loads of both pipeline
                            for (int q = lane; q < n; q += 64) {
                                const int q2 = q + 32 < n ? q + 32 : q;
                                const int w0 = mult[q], w1 = q + 32 < n ? mult[q2] : 0;
                                double ax, ay, az, bx, by, bz;
                                point3(q, ax, ay, az); point3(q2, bx, by, bz);
                                const double r0 = nx * ax + ny * ay + nz * az + dd;
                                const double r1 = nx * bx + ny * by + nz * bz + dd;
                                ic += (fabs(r0) < thr ? w0 : 0) + (fabs(r1) < thr ? w1 : 0);
                            }
#pragma unroll
                            for (int o = 16; o; o >>= 1) ic += __shfl_xor_sync(0xFFFFFFFFu, ic, o);
                        }
                        if (lane == 0) {
                            ctl.round_ic[warp] = degenerate ? -1 : ic;
                            ctl.round_model[warp][0] = nx; ctl.round_model[warp][1] = ny; ctl.round_model[warp][2] = nz;
                            ctl.round_model[warp][3] = dd; ctl.round_model[warp][4] = n4;
                        }
                    }
                    __syncthreads();
                    // run_ransac's sequential bookkeeping over this round, lane = hypothesis (every warp does it redundantly):
                    // the loop stops at the first count above the goal (it is an update: earlier counts were <= goal, or an
                    // earlier round would have stopped); up to there the returned model is the FIRST maximum, if it beats the
                    // best so far.
                    const int hi_ = min(H, h_done + NWARP), cnt = hi_ - h_done;
                    const int icl = lane < cnt ? ctl.round_ic[lane] : -2;
                    const unsigned over = cfg.ransac_stop_at_goal ? __ballot_sync(0xFFFFFFFFu, lane < cnt && (double)icl > goal) : 0u;
                    const int L = over ? __ffs(over) : cnt;                    // hypotheses the sequential loop evaluates in this round
                    stop = over != 0u;
                    used = h_done + L;
                    ndeg += __popc(__ballot_sync(0xFFFFFFFFu, lane < L && icl == -1));
                    const int mx = __reduce_max_sync(0xFFFFFFFFu, lane < L ? icl : -2);
                    if (mx > best_ic) {
                        const int w = __ffs(__ballot_sync(0xFFFFFFFFu, lane < L && icl == mx)) - 1;
                        best_ic = mx; best = h_done + w;
                        b_nx = ctl.round_model[w][0]; b_ny = ctl.round_model[w][1]; b_nz = ctl.round_model[w][2];
                        b_dd = ctl.round_model[w][3]; b_n4 = ctl.round_model[w][4];
                    }
                    h_done = hi_;
                    __syncthreads();
                }
                if (best >= 0) {
                    // unit 4-norm, sign normalised so that b >= 0 (rescale.py:158-160 flips n and h_bar together)
                    double sg = b_ny < 0 ? -1.0 : 1.0;
                    m_a = sg * b_nx / b_n4; m_b = sg * b_ny / b_n4; m_c = sg * b_nz / b_n4; m_d = sg * b_dd / b_n4;
                    status |= MVOSR_ST_UPDATED;
                    if (P.has_dbg && P.dbg.inlier) {
                        const double thr = cfg.ransac_threshold * b_n4;
                        for (int q = tid; q < n; q += NT) {
                            double qx, qy, qz;
                            point3(q, qx, qy, qz);
                            double r = b_nx * qx + b_ny * qy + b_nz * qz + b_dd;
                            P.dbg.inlier[base + q] = (mult[q] && fabs(r) < thr) ? 1 : 0;
                        }
                    }
                } else {
                    status |= MVOSR_ST_NO_MODEL;
                }
                if (tid == 0) { ctl.best_hyp = best; ctl.best_ic = best_ic; ctl.hyps_used = used; ctl.n_degenerate = ndeg; }
            }
            TMARK(13);
            // ---------------- probes that need the final flags ----------------
            if (P.has_dbg) {
                if (P.dbg.data_id) {
                    for (int p = tid; p < n; p += NT) {
                        int k = fv.tcnt[p], b = fv.tbase[p], o = (int)ga.scr[p];
                        for (int j = 0; j < k; ++j) {
                            int t = b + j;
                            if (!(tflags[t] & 4)) continue;
                            int32_t *dst = P.dbg.data_id + 6 * (size_t)base + 3 * o;
                            dst[0] = fv.tri[3 * t]; dst[1] = fv.tri[3 * t + 1]; dst[2] = fv.tri[3 * t + 2];
                            ++o;
                        }
                    }
                }
                __syncthreads();
                if (P.dbg.tri2 || P.dbg.tri_flags || P.dbg.tri_height) {
                    // canonical offsets over ALL triangles: scan tcnt (scr is free again after the probes above)
                    for (int i = tid; i < n; i += NT) ga.scr[i] = (pflag[i] & 1) ? 0u : fv.tcnt[i];
                    __syncthreads();
                    block_excl_scan(ga.scr, n, ctl.warp_cnt);
                    for (int p = tid; p < n; p += NT) {
                        if (pflag[p] & 1) continue;
                        int k = fv.tcnt[p], b = fv.tbase[p], o = (int)ga.scr[p];
                        for (int j = 0; j < k; ++j) {
                            size_t row = (size_t)(2 * base) + o + j;
                            if (P.dbg.tri2) {
                                P.dbg.tri2[3 * row] = fv.tri[3 * (b + j)]; P.dbg.tri2[3 * row + 1] = fv.tri[3 * (b + j) + 1];
                                P.dbg.tri2[3 * row + 2] = fv.tri[3 * (b + j) + 2];
                            }
                            if (P.dbg.tri_flags) P.dbg.tri_flags[row] = tflags[b + j];
                            if (P.dbg.tri_height) P.dbg.tri_height[row] = theight[b + j];
                        }
                    }
                    __syncthreads();
                }
            }
            // ---------------- height and raw scale (rescale.py:156-167) ----------------
            if (tid == 0) {
                double height = CUDART_NAN, scale = CUDART_NAN;
                if (status & MVOSR_ST_UPDATED) {
                    double nn = sqrt(m_a * m_a + m_b * m_b + m_c * m_c);
                    height = (-m_d) / nn;                           // sign already normalised (b >= 0)
                    scale = cfg.absolute_reference / height;
                }
                if (P.raw_scale) P.raw_scale[f] = scale;
                if (P.records) P.records[f].raw_scale = scale;
                if (P.stats) {
                    mvosr_frame_stats &s = P.stats[f];
                    s.model[0] = m_a; s.model[1] = m_b; s.model[2] = m_c; s.model[3] = m_d; s.height = height;
                }
            }
        } else if (P.mode == MODE_FULL) {
            if (tid == 0) {
                if (P.raw_scale) P.raw_scale[f] = CUDART_NAN;
                if (P.records) P.records[f].raw_scale = CUDART_NAN;
                if (P.stats) { mvosr_frame_stats &s = P.stats[f]; s.model[0] = s.model[1] = s.model[2] = s.model[3] = CUDART_NAN; s.height = CUDART_NAN; }
            }
        }
        if (second) status |= MVOSR_ST_SECOND_DT;
        if (n_exact) atomicAdd(&ctl.n_exact, n_exact);
        __syncthreads();
        if (tid == 0) {
            if (P.status) P.status[f] = (uint8_t)status;
            if (P.n_features) P.n_features[f] = n_feat;
            if (P.records) { P.records[f].n_features = n_feat; P.records[f].status = (uint8_t)status; }
            if (P.mode == MODE_DT_ONLY && (status & ~MVOSR_ST_SECOND_DT) && P.n_tri_out) P.n_tri_out[f] = 0;
            if (P.stats) {
                mvosr_frame_stats &s = P.stats[f];
                s.n_features = n_feat; s.n_roi = n1; s.n_dup = ctl.n_dup1 >= 0 ? ctl.n_dup1 : ctl.n_dup; s.n_kept = ctl.n_kept; s.n_tri = ctl.T;
                s.n_loose = ctl.n_loose; s.n_tight = ctl.n_tight; s.n_valid = ctl.n_valid;
                s.best_hyp = ctl.best_hyp; s.best_ic = ctl.best_ic; s.hyps_used = ctl.hyps_used; s.n_degenerate = ctl.n_degenerate;
                s.n_deferred = ctl.n_deferred_total; s.n_exact = ctl.n_exact; s.height_level = ctl.height_level;
            }
#if defined(MVOSR_STAR_COUNTERS) || defined(MVOSR_WRAP_COUNTERS)
            ctl.tphase[4] = ctl.sc.cnt[0]; ctl.tphase[5] = ctl.sc.cnt[1]; ctl.tphase[11] = ctl.sc.cnt[2]; ctl.tphase[12] = ctl.sc.cnt[3]; ctl.tphase[14] = ctl.sc.cnt[4];
            ctl.tphase[15] = ctl.sc.cnt[5]; ctl.tphase[1] = ctl.sc.cnt[6]; ctl.tphase[10] = ctl.sc.cnt[7];
#endif
            ctl.tphase[9] = ctl.sc.n_wrap;                  // stars that left the pair path (both passes)
            if (P.phase_cycles) for (int k = 0; k < 16; ++k) P.phase_cycles[16 * (size_t)f + k] = ctl.tphase[k];
        }
    }
}

}  // namespace mvosr
