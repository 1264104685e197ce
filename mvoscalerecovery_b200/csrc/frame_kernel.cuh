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
// Shared-memory plan for capacity `cap` ROI features (bytes): px,py 16cap (float64 copies: no conversions in the hot loops) | X,Y,Z 12cap | pflag cap |
// cell_start/cell_n/cell_pts 6cap(+) | u32 scratch 4cap (cell counters, later canonical offsets) |
// tri list 12cap | tbase,tcnt 3cap | mult 2cap | region A = max(star storage, 16cap heights) | flags 2cap.
#pragma once
#include <stdint.h>
#include <math_constants.h>
#include "../../include/mvosr.h"
#include "star.cuh"
#include "gstar.cuh"
#include "philox.cuh"
#include "triangulate.cuh"

namespace mvosr {

enum { MODE_FULL = 0, MODE_DT_ONLY = 1 };

struct FrameParams {
    int n_frames;
    const int32_t *offsets;
    const int32_t *counts;          // optional
    // triangulated features (FROM_CORR == false)
    const float *x, *y, *z, *u, *v;
    // correspondences (FROM_CORR == true)
    const float *cur_u, *cur_v, *ref_u, *ref_v;
    const uint8_t *e_mask;
    const double *poses;
    int cap;                        // ROI feature capacity of the staging
    int mode;
    int gate;                       // 1: skip frames with n_features <= min_features (main_offline.py:73)
    int frame_index0, seq_id;
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
    long long *phase_cycles;        // optional [F][16] per-phase SM cycles (profiling aid)
};

struct SmemPlan {
    int cap, off_px, off_py, off_X, off_Y, off_Z, off_pflag, off_cell_start, off_cell_n, off_cell_pts, off_scr,
        off_tri, off_tbase, off_tcnt, off_mult, off_A, off_flags, off_defer, total;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline SmemPlan make_plan(int cap) {
    SmemPlan p; p.cap = cap; int o = 0;
    p.off_px = o; o = align16(o + 8 * cap);
    p.off_py = o; o = align16(o + 8 * cap);
    p.off_X = o; o = align16(o + 4 * cap);
    p.off_Y = o; o = align16(o + 4 * cap);
    p.off_Z = o; o = align16(o + 4 * cap);
    p.off_pflag = o; o = align16(o + cap);
    p.off_cell_start = o; o = align16(o + 2 * (cap + 2));
    p.off_cell_n = o; o = align16(o + 2 * (cap + 2));
    p.off_cell_pts = o; o = align16(o + 2 * cap);
    p.off_scr = o; o = align16(o + 4 * (cap + 2));
    p.off_tri = o; o = align16(o + 12 * cap);
    p.off_tbase = o; o = align16(o + 2 * cap);
    p.off_tcnt = o; o = align16(o + cap);
    p.off_mult = o; o = align16(o + 2 * cap + 4);
    p.off_defer = o; o = align16(o + 2 * cap);
    int star_bytes = NWARP * WARPSTAR_BYTES;
    int a_bytes = 16 * cap > star_bytes ? 16 * cap : star_bytes;
    p.off_A = o; o = align16(o + a_bytes);
    p.off_flags = o; o = align16(o + 2 * cap);
    p.total = o;
    return p;
}

struct Ctl {                         // static shared control block
    int next_pos;
    int n_roi, n_feat, n2, status, bad, n_dup, n_dup1, n_kept, T, n_defer, n_exact, n_deferred_total;
    int n_loose, n_tight, n_valid, best_hyp, best_ic, hyps_used, n_degenerate, err;
    int warp_cnt[NWARP], warp_cnt2[NWARP];
    float bb[4];
    float red[4][NWARP];
    double height_level;
    unsigned hist[256];
    unsigned long long sel_prefix; int sel_k; unsigned long long sel_val[2];
    int frame;
    long long tphase[16];
    int round_ic[NWARP];
    double round_model[NWARP][5];
};

// ---------------------------------------------------------------------------------------------
// small block-level helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += t; }
    return v;
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
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) { int c = warp_tmp[w]; if (w < warp) woff += c; tot += c; }
    int run = woff + inc - sum;
    for (int i = b; i < e; ++i) { int t = (int)a[i]; a[i] = (uint32_t)run; run += t; }
    if (tid == 0) a[n] = (uint32_t)tot;
    __syncthreads();
}

__device__ __forceinline__ bool edge_consistent(const double *py, const float *Z, int a, int b) {
    // check_triangle (graph.py:124-129): (v_a - v_b) * (d_a - d_b) < 0, float64 on float32-exact values
    return (py[a] - py[b]) * ((double)Z[a] - (double)Z[b]) < 0.0;
}

// graph vote of one star triangle (p,qa,qb) for vertex p under the canonical (ascending) vertex order
__device__ __forceinline__ int graph_vote(const double *py, const float *Z, int p, int qa, int qb, uint32_t pass_mask) {
    int i0 = p, i1 = qa, i2 = qb;
    if (i0 > i1) { int t = i0; i0 = i1; i1 = t; }
    if (i1 > i2) { int t = i1; i1 = i2; i2 = t; }
    if (i0 > i1) { int t = i0; i0 = i1; i1 = t; }
    int a = edge_consistent(py, Z, i0, i1), b = edge_consistent(py, Z, i1, i2), c = edge_consistent(py, Z, i0, i2);
    int idx = a * 4 + b * 2 + c;
    int k = (p == i0) ? 0 : (p == i1 ? 1 : 2);
    return (pass_mask >> (idx * 3 + k)) & 1u;
}

// ---------------------------------------------------------------------------------------------
// star consumers
// ---------------------------------------------------------------------------------------------
struct FrameView {
    double *px, *py;                 // 2-D pixel coordinates (float32-exact values)
    float *X, *Y, *Z;
    uint8_t *pflag;                  // bit0 duplicate, bit1 keep
    uint16_t *tri;                   // [T][3]
    uint16_t *tbase; uint8_t *tcnt;
    Ctl *ctl;
    uint32_t pass_mask;
};

template <class Star>
__device__ __forceinline__ void consume_vote(const Star &st, int d, int p, FrameView &fv) {
    int total = 0, pass = 0;
    for (int i = 0; i < d; ++i) {
        int qa = st.get(i), qb = st.get(i + 1 < d ? i + 1 : 0);
        if (qa == INF16 || qb == INF16) continue;
        ++total;
        pass += graph_vote(fv.py, fv.Z, p, qa, qb, fv.pass_mask);
    }
    // keep = (#incident triangles with p>0.6)/(#incident) > 0.5 (graph.py:33-35,131-132); 0/0 -> False
    if (2 * pass > total) fv.pflag[p] |= 2;
}

template <class Star>
__device__ __forceinline__ void consume_emit(const Star &st, int d, int p, FrameView &fv, int tri_cap) {
    int k = 0;
    for (int i = 0; i < d; ++i) {
        int qa = st.get(i), qb = st.get(i + 1 < d ? i + 1 : 0);
        if (qa == INF16 || qb == INF16) continue;
        if (qa > p && qb > p) ++k;
    }
    fv.tcnt[p] = (uint8_t)k;
    if (!k) { fv.tbase[p] = 0; return; }
    int base = atomicAdd(&fv.ctl->T, k);
    if (base + k > tri_cap) { atomicOr(&fv.ctl->status, MVOSR_ST_OVERFLOW); fv.tcnt[p] = 0; return; }
    fv.tbase[p] = (uint16_t)base;
    int w = 0;
    for (int i = 0; i < d; ++i) {
        int qa = st.get(i), qb = st.get(i + 1 < d ? i + 1 : 0);
        if (qa == INF16 || qb == INF16) continue;
        if (!(qa > p && qb > p)) continue;
        int a = min(qa, qb), b = max(qa, qb);
        // insertion sort by (a,b) inside this point's block
        int j = w++;
        uint16_t *t = fv.tri + 3 * base;
        while (j > 0 && (t[3 * (j - 1) + 1] > a || (t[3 * (j - 1) + 1] == a && t[3 * (j - 1) + 2] > b))) {
            t[3 * j + 1] = t[3 * (j - 1) + 1]; t[3 * j + 2] = t[3 * (j - 1) + 2]; --j;
        }
        t[3 * j + 0] = (uint16_t)p; t[3 * j + 1] = (uint16_t)a; t[3 * j + 2] = (uint16_t)b;
    }
    for (int j = 0; j < k; ++j) fv.tri[3 * (base + j)] = (uint16_t)p;
}

// ---------------------------------------------------------------------------------------------
// grid build over the n points currently staged in px/py
// ---------------------------------------------------------------------------------------------
struct GridArrays { uint16_t *cell_start, *cell_n, *cell_pts; uint32_t *scr; };

__device__ __forceinline__ bool build_grid(int n, const double *px, const double *py, uint8_t *pflag, GridArrays ga,
                                           int cap, Ctl *ctl, Grid &g) {
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float xmn = CUDART_INF_F, xmx = -CUDART_INF_F, ymn = CUDART_INF_F, ymx = -CUDART_INF_F;
    for (int i = tid; i < n; i += NT) {
        float a = (float)px[i], b = (float)py[i];      // exact: the values are float32
        xmn = fminf(xmn, a); xmx = fmaxf(xmx, a); ymn = fminf(ymn, b); ymx = fmaxf(ymx, b);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        xmn = fminf(xmn, __shfl_xor_sync(0xFFFFFFFFu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xFFFFFFFFu, xmx, o));
        ymn = fminf(ymn, __shfl_xor_sync(0xFFFFFFFFu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xFFFFFFFFu, ymx, o));
    }
    if (lane == 0) { ctl->red[0][warp] = xmn; ctl->red[1][warp] = xmx; ctl->red[2][warp] = ymn; ctl->red[3][warp] = ymx; }
    __syncthreads();
    xmn = ctl->red[0][0]; xmx = ctl->red[1][0]; ymn = ctl->red[2][0]; ymx = ctl->red[3][0];
#pragma unroll
    for (int w = 1; w < NWARP; ++w) {
        xmn = fminf(xmn, ctl->red[0][w]); xmx = fmaxf(xmx, ctl->red[1][w]);
        ymn = fminf(ymn, ctl->red[2][w]); ymx = fmaxf(ymx, ctl->red[3][w]);
    }
    double w = (double)xmx - (double)xmn, hgt = (double)ymx - (double)ymn;
    double h;
    if (w > 0 && hgt > 0) h = sqrt(2.0 * w * hgt / (double)n);
    else h = fmax(w, hgt) * 2.0 / (double)n;
    h = fmax(h, fmax(sqrt(w * hgt / (double)cap), fmax(w, hgt) / (double)cap));
    if (!(h > 0)) h = 1.0;
    int gx, gy;
    for (;;) {
        double gxd = floor(w / h) + 1.0, gyd = floor(hgt / h) + 1.0;
        if (gxd * gyd <= (double)cap) { gx = (int)gxd; gy = (int)gyd; break; }
        h *= 1.25;
    }
    g.xmin = (double)xmn; g.ymin = (double)ymn; g.h = h; g.inv_h = 1.0 / h; g.gx = gx; g.gy = gy;
    int ncell = gx * gy;
    for (int i = tid; i <= ncell; i += NT) ga.scr[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        int c = cell_coord(px[i], g.xmin, g.inv_h, gx) + gx * cell_coord(py[i], g.ymin, g.inv_h, gy);
        atomicAdd(&ga.scr[c], 1u);
    }
    __syncthreads();
    block_excl_scan(ga.scr, ncell, ctl->warp_cnt);
    for (int i = tid; i <= ncell; i += NT) ga.cell_start[i] = (uint16_t)ga.scr[i];
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        int c = cell_coord(px[i], g.xmin, g.inv_h, gx) + gx * cell_coord(py[i], g.ymin, g.inv_h, gy);
        unsigned pos = atomicAdd(&ga.scr[c], 1u);
        ga.cell_pts[pos] = (uint16_t)i;
    }
    __syncthreads();
    // per cell: sort ids ascending (determinism), drop exact duplicates (lowest index kept; Qhull drops them too)
    for (int c = tid; c < ncell; c += NT) {
        int b = ga.cell_start[c], e = ga.cell_start[c + 1];
        for (int i = b + 1; i < e; ++i) {
            uint16_t v = ga.cell_pts[i]; int j = i;
            while (j > b && ga.cell_pts[j - 1] > v) { ga.cell_pts[j] = ga.cell_pts[j - 1]; --j; }
            ga.cell_pts[j] = v;
        }
        int m = b;
        for (int i = b; i < e; ++i) {
            int s = ga.cell_pts[i]; bool dup = false;
            for (int j = b; j < m; ++j) { int q = ga.cell_pts[j]; if (px[q] == px[s] && py[q] == py[s]) { dup = true; break; } }
            if (dup) { pflag[s] |= 1; atomicAdd(&ctl->n_dup, 1); }
            else ga.cell_pts[m++] = (uint16_t)s;
        }
        ga.cell_n[c] = (uint16_t)(m - b);
        for (int i = m; i < e; ++i) ga.cell_pts[i] = INF16;
    }
    __syncthreads();
    return true;
}

// lane-parallel consumers of a register-resident group star (lane i owns star triangle i)
__device__ __forceinline__ void g_consume_vote(const GCtx &c, FrameView &fv) {
    bool fin = c.gl < c.d && c.sid != INF16 && c.nid != INF16;
    int v = fin ? graph_vote(fv.py, fv.Z, c.p, c.sid, c.nid, fv.pass_mask) : 0;
    unsigned mf = gballot(c, fin), mv = gballot(c, fin && v);
    // keep = (#incident triangles with p>0.6)/(#incident) > 0.5 (graph.py:33-35,131-132); 0/0 -> False
    if (c.gl == 0 && 2 * __popc(mv) > __popc(mf)) fv.pflag[c.p] |= 2;
}

__device__ __forceinline__ void g_consume_emit(const GCtx &c, FrameView &fv, int tri_cap) {
    bool fin = c.gl < c.d && c.sid != INF16 && c.nid != INF16;
    bool em = fin && c.sid > c.p && c.nid > c.p;
    int a = min(c.sid, c.nid), b = max(c.sid, c.nid);
    unsigned key = em ? (((unsigned)a << 16) | (unsigned)b) : 0xFFFFFFFFu;
    unsigned me = gballot(c, em);
    int k = __popc(me);
    if (!k) return;
    int base = 0;
    if (c.gl == 0) base = atomicAdd(&fv.ctl->T, k);
    base = gshfl(c, base, 0);
    if (base + k > tri_cap) { if (c.gl == 0) atomicOr(&fv.ctl->status, MVOSR_ST_OVERFLOW); return; }
    int rank = 0;
    unsigned m2 = me;
    while (m2) { int j = __ffs(m2) - 1; m2 &= m2 - 1; unsigned kj = gshfl(c, key, j); rank += (kj < key); }
    if (em) { uint16_t *t = fv.tri + 3 * (base + rank); t[0] = (uint16_t)c.p; t[1] = (uint16_t)a; t[2] = (uint16_t)b; }
    if (c.gl == 0) { fv.tbase[c.p] = (uint16_t)base; fv.tcnt[c.p] = (uint8_t)k; }
}

// ---------------------------------------------------------------------------------------------
// all stars of the staged point set; EMIT selects the consumer
// ---------------------------------------------------------------------------------------------
// (not a template and not inlined: one copy of the star builder in the kernel, whatever the number of call sites)
__device__ __noinline__ void run_stars(const bool EMIT, int n, const PointSet &ps, FrameView &fv, uint16_t *star_mem, unsigned char *warp_mem,
                                       uint16_t *defer, int tri_cap, int tslot) {
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctl *ctl = fv.ctl;
    int n_exact = 0;
    long long tc0 = clock64();
    if (EMIT) for (int i = tid; i < n; i += NT) { fv.tcnt[i] = 0; fv.tbase[i] = 0; }
    if (tid == 0) ctl->next_pos = 0;
    __syncthreads();
    // one 16-lane group per point, points fetched dynamically in cell order (spatially coherent)
    {
        GCtx c;
        c.gl = tid & (GL - 1);
        c.gmask = (lane & GL) ? 0xFFFF0000u : 0x0000FFFFu;
        c.n_exact = 0;
        for (;;) {
            int pos = 0;
            if (c.gl == 0) pos = atomicAdd(&ctl->next_pos, 1);
            pos = gshfl(c, pos, 0);
            if (pos >= n) break;
            int p = ps.cell_pts[pos];
            if (p == INF16) continue;                     // duplicate: in no triangle
            int r = g_build(c, p, ps);
            if (r == STAR_OK) {
                if (EMIT) g_consume_emit(c, fv, tri_cap); else g_consume_vote(c, fv);
            } else if (r != STAR_NONE) {
                // degree > 16 (or an inconsistency): the sequential-rule warp builder takes the point
                if (c.gl == 0) { int slot = atomicAdd(&ctl->n_defer, 1); defer[slot] = (uint16_t)p; }
            }
        }
        n_exact += c.n_exact;
    }
    __syncthreads();
    long long tc1 = clock64();
    int nd = ctl->n_defer;
    WarpStar ws;
    unsigned char *wm = warp_mem + warp * WARPSTAR_BYTES;
    ws.qx = (double *)wm; ws.qy = ws.qx + MAXDEG_W; ws.ql = ws.qy + MAXDEG_W; ws.vx = ws.ql + MAXDEG_W; ws.vy = ws.vx + MAXDEG_W;
    ws.r2 = ws.vy + MAXDEG_W; ws.id = (uint16_t *)(ws.r2 + MAXDEG_W);
    for (int k = warp; k < nd; k += NWARP) {
        int p = defer[k];
        int d;
        int r = build_star_warp(ws, d, p, ps, lane, n_exact);
        __syncwarp();
        if (lane == 0) {
            if (r == STAR_OK) {
                if (EMIT) consume_emit(ws, d, p, fv, tri_cap); else consume_vote(ws, d, p, fv);
            } else {
                if (r != STAR_NONE) atomicOr(&ctl->status, MVOSR_ST_OVERFLOW);
                if (EMIT) { fv.tcnt[p] = 0; fv.tbase[p] = 0; }
            }
        }
        __syncwarp();
    }
    if (n_exact) atomicAdd(&ctl->n_exact, n_exact);
    __syncthreads();
    if (tid == 0) {
        ctl->n_deferred_total += nd; ctl->n_defer = 0;
        long long tc2 = clock64();
        ctl->tphase[tslot] += tc1 - tc0; ctl->tphase[tslot + 1] += tc2 - tc1;
    }
    __syncthreads();
}

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
    int tid = threadIdx.x;
    unsigned long long prefix = 0;
    for (int pass = 0; pass < 8; ++pass) {
        int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += NT) ctl->hist[i] = 0;
        __syncthreads();
        unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int t = tid; t < T; t += NT) {
            if (!(flags[t] & 1)) continue;
            unsigned long long key = (unsigned long long)__double_as_longlong(h[t]);
            if ((key & mask) == prefix) atomicAdd(&ctl->hist[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, b = 0;
            for (; b < 256; ++b) { int c = (int)ctl->hist[b]; if (acc + c > k) break; acc += c; }
            ctl->sel_k = k - acc;
            ctl->sel_prefix = prefix | ((unsigned long long)b << shift);
        }
        __syncthreads();
        k = ctl->sel_k; prefix = ctl->sel_prefix;
        __syncthreads();
    }
    return prefix;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool FROM_CORR>
__global__ void __launch_bounds__(NT, 1) frame_kernel(FrameParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Ctl ctl;
    const SmemPlan pl = make_plan(P.cap);
    const int cap = P.cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FrameView fv;
    fv.px = (double *)(smem + pl.off_px); fv.py = (double *)(smem + pl.off_py);
    fv.X = (float *)(smem + pl.off_X); fv.Y = (float *)(smem + pl.off_Y); fv.Z = (float *)(smem + pl.off_Z);
    fv.pflag = smem + pl.off_pflag;
    fv.tri = (uint16_t *)(smem + pl.off_tri);
    fv.tbase = (uint16_t *)(smem + pl.off_tbase); fv.tcnt = smem + pl.off_tcnt;
    fv.ctl = &ctl; fv.pass_mask = P.cfg.graph_pass_mask;
    GridArrays ga;
    ga.cell_start = (uint16_t *)(smem + pl.off_cell_start); ga.cell_n = (uint16_t *)(smem + pl.off_cell_n);
    ga.cell_pts = (uint16_t *)(smem + pl.off_cell_pts); ga.scr = (uint32_t *)(smem + pl.off_scr);
    uint16_t *mult = (uint16_t *)(smem + pl.off_mult);
    uint16_t *defer = (uint16_t *)(smem + pl.off_defer);
    uint16_t *star_mem = (uint16_t *)(smem + pl.off_A);
    unsigned char *warp_mem = smem + pl.off_A;
    double *theight = (double *)(smem + pl.off_A);          // aliases the star storage (dead by then)
    uint8_t *tflags = smem + pl.off_flags;
    const int tri_cap = 2 * cap;
    const mvosr_config &cfg = P.cfg;

    long long tlast = 0;
#define TMARK(k) do { if (tid == 0) { long long tn_ = clock64(); ctl.tphase[k] += tn_ - tlast; tlast = tn_; } } while (0)
    for (;;) {
        __syncthreads();
        if (tid == 0) ctl.frame = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int f = ctl.frame;
        if (f >= P.n_frames) break;
        if (tid == 0) {
            ctl.n_roi = ctl.n_feat = ctl.n2 = ctl.status = ctl.bad = ctl.n_dup = ctl.n_kept = ctl.T = 0; ctl.n_dup1 = -1;
            ctl.n_defer = ctl.n_exact = ctl.n_deferred_total = 0;
            ctl.n_loose = ctl.n_tight = ctl.n_valid = 0; ctl.best_hyp = -1; ctl.best_ic = 0; ctl.hyps_used = 0;
            ctl.n_degenerate = 0; ctl.err = 0; ctl.height_level = CUDART_NAN;
            for (int k = 0; k < 16; ++k) ctl.tphase[k] = 0;
            tlast = clock64();
        }
        __syncthreads();
        const int base = P.offsets[f];
        const int n_in = P.counts ? P.counts[f] : (P.offsets[f + 1] - base);

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
                    double X, Y, Z, uu, vv;
                    feat = triangulate_point(P.cur_u[base + i], P.cur_v[base + i], P.ref_u[base + i], P.ref_v[base + i], pose,
                                             cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.triangulation_max_depth, X, Y, Z, uu, vv);
                    if (P.e_mask) feat = feat && P.e_mask[base + i] != 0;
                    fx3 = (float)X; fy3 = (float)Y; fz3 = (float)Z; fu = (float)uu; fv_ = (float)vv;
                } else {
                    feat = true;
                    fv_ = P.v[base + i];
                }
                ok = feat && (P.mode == MODE_DT_ONLY || fv_ > cfg.vanish);
                if (ok && !FROM_CORR) {
                    fu = P.u[base + i];
                    if (P.mode != MODE_DT_ONLY) { fx3 = P.x[base + i]; fy3 = P.y[base + i]; fz3 = P.z[base + i]; }
                }
            }
            unsigned bal = __ballot_sync(0xFFFFFFFFu, ok), balf = __ballot_sync(0xFFFFFFFFu, feat);
            if (lane == 0) { ctl.warp_cnt[warp] = __popc(bal); ctl.warp_cnt2[warp] = __popc(balf); }
            __syncthreads();
            int woff = 0, tot = 0, totf = 0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) { int c = ctl.warp_cnt[w]; if (w < warp) woff += c; tot += c; totf += ctl.warp_cnt2[w]; }
            if (ok) {
                int pos = total + woff + __popc(bal & ((1u << lane) - 1u));
                if (!(fabsf(fu) < 4096.f) || !(fabsf(fv_) < 4096.f)) ctl.bad = 1;
                if (fabsf(fu) < 7.62939453125e-06f) fu = 0.f;       // 2^-17: keeps every difference exact in float64
                if (fabsf(fv_) < 7.62939453125e-06f) fv_ = 0.f;
                if (pos < cap) {
                    fv.px[pos] = (double)fu; fv.py[pos] = (double)fv_; fv.X[pos] = fx3; fv.Y[pos] = fy3; fv.Z[pos] = fz3; fv.pflag[pos] = 0;
                }
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

        Grid g;
        PointSet ps;
        ps.px = fv.px; ps.py = fv.py; ps.cell_start = ga.cell_start; ps.cell_n = ga.cell_n; ps.cell_pts = ga.cell_pts;
        bool second = false;
        int n1 = n;
        if (!status) {
            // ---------------- Delaunay #1 -> graph vote (or triangles in DT-only mode) ----------------
            build_grid(n, fv.px, fv.py, fv.pflag, ga, cap, &ctl, g);
            TMARK(1);
            ps.g = g;
            if (P.mode == MODE_DT_ONLY) {
                run_stars(true, n, ps, fv, star_mem, warp_mem, defer, tri_cap, 2);
            } else {
                run_stars(false, n, ps, fv, star_mem, warp_mem, defer, tri_cap, 2);
            }
            status |= ctl.status;
            if (tid == 0) tlast = clock64();
        }
        if (!status && P.mode == MODE_DT_ONLY) {
            write_canonical(n, fv, ga.scr, ctl.warp_cnt, P.tri_out + 3 * (size_t)(2 * base), P.n_tri_out + f);
            if (ctl.T == 0) status |= MVOSR_ST_FEW_ROI;
        }
        if (!status && P.mode == MODE_FULL) {
            // ---------------- keep mask, survivor rule (rescale.py:131-137) ----------------
            int cnt = 0;
            for (int i = tid; i < n; i += NT) cnt += (fv.pflag[i] >> 1) & 1;
#pragma unroll
            for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (lane == 0) atomicAdd(&ctl.n_kept, cnt);
            __syncthreads();
            const int n_kept = ctl.n_kept;
            if (P.has_dbg && P.dbg.keep) for (int i = tid; i < n; i += NT) P.dbg.keep[base + i] = (fv.pflag[i] >> 1) & 1;
            second = n_kept > cfg.min_kept;
            if (P.has_dbg && P.dbg.tri1) {
                // parity probe: also materialise Delaunay #1 (re-runs the stars in emit mode)
                run_stars(true, n, ps, fv, star_mem, warp_mem, defer, tri_cap, 14);
                write_canonical(n, fv, ga.scr, ctl.warp_cnt, P.dbg.tri1 + 3 * (size_t)(2 * base), P.dbg.n_tri1 ? P.dbg.n_tri1 + f : nullptr);
                if (tid == 0) ctl.T = 0;
                __syncthreads();
            }
            if (second) {
                // order-preserving in-place compaction by the keep flag, chunk by chunk
                int run = 0;
                for (int c0 = 0; c0 < n; c0 += NT) {
                    int i = c0 + tid;
                    bool k = i < n && (fv.pflag[i] & 2);
                    double a0 = 0, a1 = 0; float a2 = 0, a3 = 0, a4 = 0;
                    if (k) { a0 = fv.px[i]; a1 = fv.py[i]; a2 = fv.X[i]; a3 = fv.Y[i]; a4 = fv.Z[i]; }
                    unsigned bal = __ballot_sync(0xFFFFFFFFu, k);
                    if (lane == 0) ctl.warp_cnt[warp] = __popc(bal);
                    __syncthreads();
                    int woff = 0, tot = 0;
#pragma unroll
                    for (int w = 0; w < NWARP; ++w) { int c = ctl.warp_cnt[w]; if (w < warp) woff += c; tot += c; }
                    if (k) {
                        int pos = run + woff + __popc(bal & ((1u << lane) - 1u));
                        fv.px[pos] = a0; fv.py[pos] = a1; fv.X[pos] = a2; fv.Y[pos] = a3; fv.Z[pos] = a4;
                    }
                    run += tot;
                    __syncthreads();
                }
                n = n_kept;
                for (int i = tid; i < n; i += NT) fv.pflag[i] = 0;
                if (tid == 0) { ctl.n_dup1 = ctl.n_dup; ctl.n_dup = 0; }
                __syncthreads();
                build_grid(n, fv.px, fv.py, fv.pflag, ga, cap, &ctl, g);
                ps.g = g;
            } else {
                for (int i = tid; i < n; i += NT) fv.pflag[i] &= 1;
                __syncthreads();
            }
            // ---------------- Delaunay #2 (or #1 again) with triangle emission ----------------
            TMARK(6);
            run_stars(true, n, ps, fv, star_mem, warp_mem, defer, tri_cap, 7);
            status |= ctl.status;
            if (tid == 0) tlast = clock64();
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
                double p0x = fv.X[i0], p0y = fv.Y[i0], p0z = fv.Z[i0];
                double e1x = (double)fv.X[i1] - p0x, e1y = (double)fv.Y[i1] - p0y, e1z = (double)fv.Z[i1] - p0z;
                double e2x = (double)fv.X[i2] - p0x, e2y = (double)fv.Y[i2] - p0y, e2z = (double)fv.Z[i2] - p0z;
                // n = P^-1 . 1 = (e1 x e2) / (p0 . (e1 x e2))
                double cx = e1y * e2z - e1z * e2y, cy = e1z * e2x - e1x * e2z, cz = e1x * e2y - e1y * e2x;
                double det = p0x * cx + p0y * cy + p0z * cz;
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
            TMARK(10);
            const int n_loose = ctl.n_loose;
            // height_level = 0.9 * median(height[loose])  (np.median: mean of the two middle values for even counts)
            double level = CUDART_NAN;
            if (n_loose > 0) {
                unsigned long long ka = block_select(theight, tflags, T, (n_loose - 1) / 2, &ctl);
                double a = __longlong_as_double((long long)ka), b = a;
                if ((n_loose & 1) == 0) {
                    unsigned long long kb = block_select(theight, tflags, T, n_loose / 2, &ctl);
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
            if (n_sel >= cfg.min_selected) {
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
                        sample3_positions(P.seed, (uint32_t)h, (uint32_t)(P.frame_index0 + f), (uint32_t)P.seq_id, (uint32_t)n_sel,
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
                        double p0x = fv.X[vtx[0]], p0y = fv.Y[vtx[0]], p0z = fv.Z[vtx[0]];
                        double e1x = (double)fv.X[vtx[1]] - p0x, e1y = (double)fv.Y[vtx[1]] - p0y, e1z = (double)fv.Z[vtx[1]] - p0z;
                        double e2x = (double)fv.X[vtx[2]] - p0x, e2y = (double)fv.Y[vtx[2]] - p0y, e2z = (double)fv.Z[vtx[2]] - p0z;
                        // null vector of [p 1] (3x4) in closed form: (n, -n.p0), n = e1 x e2 (estimate_road_norm.py:13-15)
                        nx = e1y * e2z - e1z * e2y; ny = e1z * e2x - e1x * e2z; nz = e1x * e2y - e1y * e2x;
                        dd = -(nx * p0x + ny * p0y + nz * p0z);
                        n4 = sqrt(nx * nx + ny * ny + nz * nz + dd * dd);
                        if (!(nx * nx + ny * ny + nz * nz > 0)) degenerate = true;
                        if (!degenerate) {
                            const double thr = cfg.ransac_threshold * n4;   // |m.[x 1]| < thr for the unit-4-norm model m
                            for (int q = lane; q < n; q += 32) {
                                int w = mult[q];
                                if (!w) continue;
                                double r = nx * (double)fv.X[q] + ny * (double)fv.Y[q] + nz * (double)fv.Z[q] + dd;
                                if (fabs(r) < thr) ic += w;
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
                    const int hi_ = min(H, h_done + NWARP);
                    for (int hh = h_done; hh < hi_ && !stop; ++hh) {
                        int w = hh - h_done, ic = ctl.round_ic[w];
                        used = hh + 1;
                        ndeg += ic < 0;
                        if (ic > best_ic) {
                            best_ic = ic; best = hh;
                            b_nx = ctl.round_model[w][0]; b_ny = ctl.round_model[w][1]; b_nz = ctl.round_model[w][2];
                            b_dd = ctl.round_model[w][3]; b_n4 = ctl.round_model[w][4];
                            if ((double)ic > goal && cfg.ransac_stop_at_goal) stop = true;
                        }
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
                            double r = b_nx * (double)fv.X[q] + b_ny * (double)fv.Y[q] + b_nz * (double)fv.Z[q] + b_dd;
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
                    for (int i = tid; i < n; i += NT) ga.scr[i] = (fv.pflag[i] & 1) ? 0u : fv.tcnt[i];
                    __syncthreads();
                    block_excl_scan(ga.scr, n, ctl.warp_cnt);
                    for (int p = tid; p < n; p += NT) {
                        if (fv.pflag[p] & 1) continue;
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
                P.raw_scale[f] = scale;
                if (P.stats) {
                    mvosr_frame_stats &s = P.stats[f];
                    s.model[0] = m_a; s.model[1] = m_b; s.model[2] = m_c; s.model[3] = m_d; s.height = height;
                }
            }
        } else if (P.mode == MODE_FULL) {
            if (tid == 0) {
                P.raw_scale[f] = CUDART_NAN;
                if (P.stats) { mvosr_frame_stats &s = P.stats[f]; s.model[0] = s.model[1] = s.model[2] = s.model[3] = CUDART_NAN; s.height = CUDART_NAN; }
            }
        }
        if (second) status |= MVOSR_ST_SECOND_DT;
        __syncthreads();
        if (tid == 0) {
            if (P.status) P.status[f] = (uint8_t)status;
            if (P.n_features) P.n_features[f] = n_feat;
            if (P.mode == MODE_DT_ONLY && (status & ~MVOSR_ST_SECOND_DT) && P.n_tri_out) P.n_tri_out[f] = 0;
            if (P.stats) {
                mvosr_frame_stats &s = P.stats[f];
                s.n_features = n_feat; s.n_roi = n1; s.n_dup = ctl.n_dup1 >= 0 ? ctl.n_dup1 : ctl.n_dup; s.n_kept = ctl.n_kept; s.n_tri = ctl.T;
                s.n_loose = ctl.n_loose; s.n_tight = ctl.n_tight; s.n_valid = ctl.n_valid;
                s.best_hyp = ctl.best_hyp; s.best_ic = ctl.best_ic; s.hyps_used = ctl.hyps_used; s.n_degenerate = ctl.n_degenerate;
                s.n_deferred = ctl.n_deferred_total; s.n_exact = ctl.n_exact; s.height_level = ctl.height_level;
            }
            if (P.phase_cycles) for (int k = 0; k < 16; ++k) P.phase_cycles[16 * (size_t)f + k] = ctl.tphase[k];
        }
    }
}

}  // namespace mvosr
