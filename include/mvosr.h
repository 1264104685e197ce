/* mvosr.h -- C ABI of the B200 scale-recovery library (libmvosr.so).
 *
 * Drop-in boundary for the per-frame scale-recovery path of TimingSpace/MVOScaleRecovery.
 * The reference is pure Python and has no FFI; each entry point below names the reference
 * call it replaces (file:line into the reference tree).  Signatures are plain C: pointers,
 * sizes, a POD config, an opaque handle and a CUDA stream passed as void*.  No torch types.
 *
 * Memory: unless a name ends in _host, every data pointer is a DEVICE pointer owned by the
 * caller (e.g. torch tensors); the library owns only its internal workspace.  Kernels are
 * enqueued on the given stream and the call returns without synchronising (the _host entry
 * points synchronise before returning because they copy results to host memory).
 *
 * Threading: a handle must not be used from two host threads concurrently (its workspace, scheduler counters and streams are
 * unguarded); different handles are independent.  Every call except mvosr_create makes the handle's device current for the
 * calling thread (cudaSetDevice) and leaves it so.
 * Errors: every function returns MVOSR_OK (0) or a negative MVOSR_E_* code; per-frame
 * conditions are reported in the status byte of each frame, never as a call failure.
 *
 * Data layout ("frame batch"): F frames packed CSR-style.  offsets[f]..offsets[f+1] delimit
 * the features of frame f inside structure-of-arrays float32 buffers (x[], y[], z[], u[], v[]
 * for triangulated features; cur_u[], cur_v[], ref_u[], ref_v[] for tracked correspondences).
 */
#ifndef MVOSR_H
#define MVOSR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVOSR_VERSION 200

/* ---- return codes ---- */
#define MVOSR_OK                 0
#define MVOSR_E_INVALID         -1   /* bad argument */
#define MVOSR_E_CUDA            -2   /* CUDA runtime error (see mvosr_last_cuda_error) */
#define MVOSR_E_NOMEM           -3
#define MVOSR_E_CAPACITY        -4   /* max_features exceeds what the fused kernel supports */
#define MVOSR_E_NO_DEVICE       -5

/* ---- per-frame status byte (bit flags) ---- */
#define MVOSR_ST_UPDATED      0x01   /* RANSAC ran (N_sel >= 12): raw_scale is valid (rescale.py:152) */
#define MVOSR_ST_SECOND_DT    0x02   /* graph check kept > 10 features -> second Delaunay (rescale.py:133) */
#define MVOSR_ST_FEW_ROI      0x04   /* < 3 ROI features or all collinear: reference raises QhullError */
#define MVOSR_ST_NO_MODEL     0x08   /* every hypothesis had 0 inliers: reference crashes in np.matrix(None) */
#define MVOSR_ST_BAD_INPUT    0x10   /* |u|,|v| >= 4096, or a non-finite pixel coordinate or 3-D coordinate in the ROI */
#define MVOSR_ST_OVERFLOW     0x20   /* frame exceeds the kernel capacity (ROI features per frame; a point with more than 254 Delaunay neighbours) */
#define MVOSR_ST_SKIPPED      0x40   /* frame not processed (not moving / too few features, main_offline.py:64,73) */
#define MVOSR_ST_SINGULAR     0x80   /* a triangle's vertex matrix is singular (plane through the origin): the reference raises
                                        LinAlgError in np.matrix(...).I (rescale.py:79); no RANSAC, the temporal state is held */

/* ---- configuration (POD). Defaults = the constants hard-coded in the reference. ---- */
typedef struct mvosr_config {
    double  absolute_reference;   /* camera height in metres; main.py:55 passes param.camera_h (1.75); harness uses 1.7 */
    double  fx, fy, cx, cy;       /* src/param.py:30-35 */
    float   vanish;               /* ROI row, features with v > vanish are kept (rescale.py:30,115) = 185 */
    int32_t min_features;         /* estimator is called only if n > this (param.py:37, main_offline.py:73) = 100 */
    int32_t min_kept;             /* second Delaunay only if kept > this (rescale.py:133) = 10 */
    int32_t min_selected;         /* RANSAC only if N_sel >= this (rescale.py:152) = 12 */
    double  sin_loose;            /* largest s with asin(s)*180/pi < -80 (rescale.py:85): loose <=> s <= sin_loose */
    double  sin_tight;            /* same for -85 (rescale.py:86) */
    double  height_level_factor;  /* 0.9 (rescale.py:91) */
    int32_t ransac_iterations;    /* 100 (rescale.py:155) */
    int32_t ransac_stop_at_goal;  /* 1: stop at first ic > goal (ransac.py:20); 0: evaluate all (sweep config C5) */
    double  ransac_threshold;     /* 0.005 (rescale.py:155) */
    double  ransac_goal_fraction; /* 0.8 (estimate_road_norm.py:68) */
    uint32_t graph_pass_mask;     /* 24-bit table: bit (idx*3+k) set <=> p_k(idx) > 0.6 (graph.py:131-145) for the
                                     canonical (ascending) vertex order; default from [[3,1],[2,2],[2,2],[0,4]] */
    double  slew_limit;           /* 0.3 (rescale.py:169-172) */
    int32_t window_size;          /* 5: both mains pass window_size=5 (main.py:55); class default is 6 */
    double  triangulation_max_depth; /* distanceThresh=100 (visual_odometry.py:133) */
    int32_t reserved[8];          /* reserved[0]: grid density of the Delaunay stage x100 (mean points per cell), 0 = built-in default */
} mvosr_config;

/* ---- per-frame counters (parity probes and logging; mirrors the reference's prints) ---- */
typedef struct mvosr_frame_stats {
    int32_t n_features;     /* features in the frame after the triangulation mask (pre-ROI) */
    int32_t n_roi;          /* after v > vanish */
    int32_t n_dup;          /* exact duplicate 2-D points dropped from Delaunay #1 (Qhull: .coplanar) */
    int32_t n_kept;         /* graph-check survivors ("feature left", rescale.py:132) */
    int32_t n_tri;          /* triangles fed to flat_selection (Delaunay #2, or #1 if not re-triangulated) */
    int32_t n_loose;        /* pitch < -80 ("triangle left", rescale.py:88) */
    int32_t n_tight;        /* pitch < -85 */
    int32_t n_valid;        /* tight & height > level ("triangle left final", rescale.py:97); N_sel = 3*n_valid */
    int32_t best_hyp;       /* index of the returned hypothesis, -1 if none */
    int32_t best_ic;        /* its inlier count over the vertex list with multiplicity (ransac.py:13-16) */
    int32_t hyps_used;      /* hypotheses the sequential loop would have evaluated (ransac.py:9,20) */
    int32_t n_degenerate;   /* evaluated hypotheses whose 3 positions do not span a plane (reference: LAPACK-arbitrary) */
    int32_t n_deferred;     /* stars finished by the warp-cooperative path (Delaunay #1 + #2) */
    int32_t n_exact;        /* predicate evaluations that fell through to exact arithmetic */
    double  height_level;   /* 0.9 * median(height[loose]) (rescale.py:91) */
    double  model[4];       /* returned plane (a,b,c,d), unit 4-norm, sign normalised so that b >= 0 */
    double  height;         /* h_bar / |n| (rescale.py:157-164) */
} mvosr_frame_stats;

/* ---- optional dense per-frame debug outputs (device pointers, any may be NULL) ----
 * Frame f writes at row offset offsets[f] (feature-indexed arrays) or 2*offsets[f] (triangle-indexed). */
typedef struct mvosr_debug_buffers {
    int32_t *tri1;          /* [2*M][3] canonical triangles of Delaunay #1, indices into the ROI-compacted frame */
    int32_t *n_tri1;        /* [F] */
    uint8_t *keep;          /* [M] graph-check keep flag per ROI feature (graph.py:33-36) */
    int32_t *tri2;          /* [2*M][3] canonical triangles fed to flat_selection, indices after keep-compaction */
    uint8_t *tri_flags;     /* [2*M] bit0 loose, bit1 tight, bit2 valid */
    double  *tri_height;    /* [2*M] 1/|n| per triangle */
    uint8_t *inlier;        /* [M] inlier flag of the returned model per kept feature (only selected vertices count) */
    int32_t *data_id;       /* [6*M] the vertex list handed to RANSAC (rescale.py:101), 3*n_valid entries */
} mvosr_debug_buffers;

/* ---- one frame's result as a single 16-byte record: what a rank contributes to the fleet's all-gather (fleet.py) ---- */
typedef struct mvosr_frame_record {
    double  raw_scale;      /* absolute_reference / height, NaN unless status has MVOSR_ST_UPDATED */
    int32_t n_features;     /* post-mask feature count (the n > min_features gate of main_offline.py:73) */
    uint8_t status;         /* MVOSR_ST_* */
    uint8_t pad[3];
} mvosr_frame_record;

typedef struct mvosr_handle mvosr_handle;

int  mvosr_version(void);
const char *mvosr_error_string(int code);
const char *mvosr_last_cuda_error(void);

/* Fill cfg with the reference's constants (see field comments). */
int  mvosr_default_config(mvosr_config *cfg);

/* Replaces ScaleEstimator.__init__ (src/rescale.py:23-35) for a batch context on CUDA device `device`. */
int  mvosr_create(const mvosr_config *cfg, int device, mvosr_handle **out);
int  mvosr_destroy(mvosr_handle *h);
int  mvosr_get_config(const mvosr_handle *h, mvosr_config *cfg);

/* Stage 1 -- replaces the triangulation inside cv2.recoverPose(..., distanceThresh=100) +
 * dehomogenisation (src/thirdparty/MonocularVO/visual_odometry.py:129-147) and the reprojection of
 * src/main.py:102-104.  poses: [F][12] float64 row-major [R|t], x_ref = R x_cur + t.
 * e_mask: optional per-correspondence uint8 (the findEssentialMat inlier mask, :134-136), may be NULL.
 * Outputs are written order-preserving and compacted at offsets[f]; n_out[f] = surviving features. */
int  mvosr_triangulate_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                              const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                              const uint8_t *e_mask, const double *poses,
                              float *x, float *y, float *z, float *u, float *v, int32_t *n_out,
                              void *stream);

/* Stages 2-5 -- replaces ScaleEstimator.feature_selection + the RANSAC/height/scale part of
 * scale_calculation_ransac (src/rescale.py:113-167) for F frames at once; the temporal state
 * (:168-178) is applied by mvosr_filter_sequences.  counts may be NULL (then offsets[f+1]-offsets[f]).
 * frame_index0 / seq_id / seed select the Philox hypothesis stream: hypothesis i of frame f uses
 * counter (i, frame_index0+f, seq_id, 0) and key (seed lo, seed hi).  max_features = upper bound on the
 * features of any frame (sizes the shared-memory staging).  raw_scale[f] = absolute_reference / height
 * when status has MVOSR_ST_UPDATED, NaN otherwise. */
int  mvosr_scale_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const int32_t *counts,
                        const float *x, const float *y, const float *z, const float *u, const float *v,
                        int32_t max_features, int32_t frame_index0, int32_t seq_id, uint64_t seed,
                        double *raw_scale, uint8_t *status, mvosr_frame_stats *stats,
                        const mvosr_debug_buffers *debug, void *stream);

/* The same on the float64 arrays of the reference's own hand-off: feature3d [M][3], feature2d [M][2] (array-of-structures, as numpy
 * holds what src/main.py:102-113 passes to scale_calculation), CSR by offsets.  The reference computes on these float64 values;
 * rounding them to float32 first moves a gate in about one frame in ten, which changes N_sel and with it the hypotheses drawn
 * (tests/test_f64_handoff.py).  Here the ROI cut, the depth-order votes, the planes, the gates and the RANSAC evaluate the float64
 * values; only the Delaunay triangulations run on the float32 roundings of the pixel coordinates (points that differ only below
 * float32 resolution count as duplicates). */
int  mvosr_scale_frames_f64(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const double *feature3d, const double *feature2d,
                        int32_t max_features, int32_t frame_index0, int32_t seq_id, uint64_t seed,
                        double *raw_scale, uint8_t *status, mvosr_frame_stats *stats,
                        const mvosr_debug_buffers *debug, void *stream);

/* One frame, HOST pointers in and out -- what rescale.ScaleEstimator.scale_calculation(feature3d, feature2d) does per call
 * (src/main.py:113, src/main_offline.py:75) up to the temporal state: the arrays are copied as they are, one launch of the float64
 * variant above, one copy back; synchronises.  record_out_host receives (raw_scale, n_features, status); stats_out_host is optional. */
int  mvosr_scale_frame_host_f64(mvosr_handle *h, int32_t n_features, const double *feature3d_host, const double *feature2d_host,
                        int32_t frame_index, int32_t seq_id, uint64_t seed,
                        mvosr_frame_record *record_out_host, mvosr_frame_stats *stats_out_host);

/* Stages 1-5 fused: tracked correspondences + relative poses in, raw scales out; triangulated
 * features never leave the SM.  n_features[f] receives the post-mask feature count (needed by the
 * n > min_features gate of main.py:110).  Same stream/seed conventions as mvosr_scale_frames. */
int  mvosr_scale_frames_from_correspondences(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                        const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                        const uint8_t *e_mask, const double *poses,
                        int32_t max_features, int32_t frame_index0, int32_t seq_id, uint64_t seed,
                        double *raw_scale, uint8_t *status, int32_t *n_features, mvosr_frame_stats *stats,
                        void *stream);

/* Stages 1-5 over a SHARD of a fleet (BASELINE configs[3]): one contiguous frame range of the concatenated sequences, which
 * may span sequence boundaries, in ONE launch.  Replaces the frame loop of src/main_offline.py:57-88 for that range.
 * frame_seq[f] / frame_index[f] (device int32 [F], either may be NULL = seq_id / frame_index0 + f) give the Philox stream of
 * every frame, so a frame draws the same hypotheses wherever the shard boundaries fall; order (device int32 [F], optional) is
 * the processing order of the persistent CTAs -- e.g. frames sorted by decreasing size, which shortens the tail of the last
 * wave; records[f] receives (raw_scale, n_features, status) of frame f (not of work item f). */
int  mvosr_scale_shard_from_correspondences(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                        const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                        const uint8_t *e_mask, const double *poses, int32_t max_features,
                        const int32_t *frame_seq, const int32_t *frame_index, int32_t frame_index0, int32_t seq_id,
                        const int32_t *order, uint64_t seed, mvosr_frame_record *records, void *stream);

/* Stage 6 over gathered records: as mvosr_filter_sequences, reading frame f of the global frame order from
 * records[slot ? slot[f] : f] -- slot maps a global frame to its place in the all-gathered buffer (rank-major, every rank's
 * block padded to the longest shard), so the gather output is consumed in place with no unpacking pass. */
int  mvosr_filter_records(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets,
                          const mvosr_frame_record *records, const int32_t *slot, const uint8_t *move_flags,
                          double *scale_out, double *filter10_out, void *stream);

/* Stage 6 -- replaces the driver gating of src/main_offline.py:57-88 and the temporal state of
 * src/rescale.py:168-178 (slew limiter + median of the last window_size states), then
 * filter(data, 10) of script/evaluate_scale.py:25-29.  One CTA per sequence; seq_offsets: [S+1]
 * frame ranges.  move_flags may be NULL (all moving).  n_features may be NULL (all above the gate).
 * scale_out[f] = what main_offline appends to `scales` (scales[1:]); filter10_out may be NULL. */
int  mvosr_filter_sequences(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets,
                            const double *raw_scale, const uint8_t *status, const uint8_t *move_flags,
                            const int32_t *n_features, double *scale_out, double *filter10_out, void *stream);

/* Canonical Delaunay triangles of F point sets (the replacement of scipy.spatial.Delaunay(...).simplices
 * at src/rescale.py:124-125, rows ascending, rows lexsorted).  tri: [2*M][3] written at 2*offsets[f]. */
int  mvosr_delaunay_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                           const float *u, const float *v, int32_t max_features,
                           int32_t *tri, int32_t *n_tri, uint8_t *status, void *stream);

/* Host-buffer convenience used for end-to-end timing and by callers without device memory: all pointers
 * are HOST pointers (pinned for full speed); copies, stages 1-6 and the copy back happen inside. */
int  mvosr_recover_scales_host(mvosr_handle *h, int32_t n_frames, const int32_t *offsets_host,
                        const float *cur_u_host, const float *cur_v_host, const float *ref_u_host, const float *ref_v_host,
                        const double *poses_host, const uint8_t *move_flags_host,
                        int32_t max_features, int32_t seq_id, uint64_t seed,
                        double *scale_out_host, double *raw_scale_out_host, uint8_t *status_out_host);

/* The same for S sequences in one call (BASELINE configs[3], the fleet): seq_offsets_host [S+1] are frame ranges inside the
 * CSR batch; sequence s uses Philox sequence id seq_id0 + s, frame counters and the temporal filter restart at every sequence. */
int  mvosr_recover_fleet_host(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets_host, const int32_t *offsets_host,
                        const float *cur_u_host, const float *cur_v_host, const float *ref_u_host, const float *ref_v_host,
                        const double *poses_host, const uint8_t *move_flags_host,
                        int32_t max_features, int32_t seq_id0, uint64_t seed,
                        double *scale_out_host, double *raw_scale_out_host, uint8_t *status_out_host);

/* ---- stand-alone primitives: the same arithmetic for callers that bring their own triangles / point lists ----
 * They back the API-surface variants of the reference that take explicit triangles or points (float64 arrays as numpy
 * hands them): ScaleEstimator.flat_selection (src/rescale.py:75-102), Reconstruct.triangle_model (src/reconstruct.py:70-90),
 * feature_selection_by_tri (src/scale_calculator.py:225-248), find_outliers (src/rescale.py:63-72,
 * src/scale_calculator.py:151-167), get_pitch_ransac (src/estimate_road_norm.py:66-70), the batch scripts
 * (src/calculate_height_pitch.py:62-204, src/triangle_batch.py:23-63) and get_path (src/main_offline.py:95-119). */

/* n = P^-1 . 1 per triangle (P rows = the three vertices; the loop body of src/rescale.py:77-84).
 * tri: [T][3] int32 indices into xyz ([N][3] float64 row-major).  Outputs, each may be NULL: normal [T][3] (not
 * normalised), height [T] = 1/|n|, mean_y [T] = mean Y of the vertices (src/scale_calculator.py:237).  Singular -> NaN. */
int  mvosr_triangle_planes(mvosr_handle *h, int32_t n_tri, const int32_t *tri, const double *xyz,
                           double *normal, double *height, double *mean_y, void *stream);

/* Depth-order votes per vertex (check_triangle + find_outliers, src/rescale.py:45-72): flagged[i] = number of triangles
 * that flag vertex i ([a|b, a|b|c, c] with a,b,c the (v_i-v_j)(d_i-d_j) > 0 tests of edges 01, 02, 12), incident[i]
 * (optional) = number of triangles vertex i belongs to.  v: pixel rows [N], d: depths [N], float64. */
int  mvosr_triangle_votes(mvosr_handle *h, int32_t n_tri, const int32_t *tri, const double *v, const double *d,
                          int32_t n_points, int32_t *flagged, int32_t *incident, void *stream);

/* Batched 3-point plane RANSAC over explicit point lists -- get_pitch_ransac / run_ransac
 * (src/estimate_road_norm.py:66-70, src/thirdparty/Ransac/ransac.py:3-23).  List s = rows offsets[s]..offsets[s+1] of
 * xyz ([M][3] float64).  Hypothesis i of list s draws its three distinct positions from the Philox stream with counter
 * (i, frame_index ? frame_index[s] : s, seq_id, 0); goal = goal_fraction * N (reference: 0.8).  Outputs: model [S][4] =
 * the returned plane, unit 4-norm, sign b >= 0 (NaN when no hypothesis had an inlier); ic, best_hyp (-1: none),
 * hyps_used [S] (optional). */
int  mvosr_ransac_planes(mvosr_handle *h, int32_t n_sets, const int32_t *offsets, const double *xyz,
                         int32_t iterations, double threshold, double goal_fraction, int32_t stop_at_goal,
                         uint64_t seed, const int32_t *frame_index, int32_t seq_id,
                         double *model, int32_t *ic, int32_t *best_hyp, int32_t *hyps_used, void *stream);

/* get_path / motion2pose (src/main_offline.py:95-119) for S sequences: translations scaled by the per-frame scale
 * (scales may be NULL = 1), then the running product of the relative motions.  motions: [F][12] row-major [R|t];
 * seq_offsets: [S+1] frame ranges; poses_out: [F+S][12], sequence s occupies rows seq_offsets[s]+s ...
 * seq_offsets[s+1]+s (its first row is the identity).  Computed as a parallel scan: equal to the reference's
 * left-to-right product up to rounding. */
int  mvosr_integrate_paths(mvosr_handle *h, int32_t n_sequences, const int32_t *seq_offsets, const double *motions,
                           const double *scales, double *poses_out, void *stream);

/* Pose from the essential matrix -- the selection step of cv2.recoverPose(E, px_cur, px_ref, K, distanceThresh)
 * (src/thirdparty/MonocularVO/visual_odometry.py:129-133; part of SURVEY N1): E is decomposed into its two rotations and
 * +-t, every correspondence is triangulated under the four candidates, and the candidate with the most points in front of
 * both cameras and nearer than triangulation_max_depth wins.  essential: [F][9] row-major, x_ref^T E x_cur = 0 in
 * normalised coordinates (what cv2.findEssentialMat(px_cur, px_ref, K) returns).  e_mask: optional, restricts the count
 * (the reference passes none).  Outputs: poses_out [F][12] = [R|t] with x_ref = R x_cur + t, |t| = 1 -- the input of
 * mvosr_triangulate_frames / mvosr_scale_frames_from_correspondences; n_good [F][4] (optional): counts of (R1,t), (R2,t),
 * (R1,-t), (R2,-t). */
int  mvosr_recover_pose_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                               const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                               const uint8_t *e_mask, const double *essential, double *poses_out, int32_t *n_good, void *stream);

/* Feature bucketing of the VO front-end -- bucket(features, bucket_size=30, density=2) of src/detector.py:65-95 (and
 * FeatureDetector.bucket, :18-47): features binned into bucket_size x bucket_size pixel cells, each cell keeps at
 * most `density` of them, survivors listed cell by cell (rows of cells top to bottom, cells left to right).  The reference
 * shuffles each cell with numpy's global RNG; here a cell keeps its `density` smallest (r_i, i), r_i = word 0 of
 * Philox4x32-10(counter = (i, frame_index ? frame_index[f] : f, seq_id, 3), key = seed), in that order.  u, v: [M] float32 pixel
 * coordinates (non-negative, below 1024 x 1023 cells), CSR by offsets; frames hold at most 4096 features.  Outputs: out_index
 * [M] = positions (within the frame) of the survivors, compacted at the frame's base offset; n_out [F]; status [F] (optional):
 * 0 ok, 1 coordinate out of range / not finite, 2 more than 4096 features (such frames keep nothing). */
int  mvosr_bucket_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets, const float *u, const float *v,
                         int32_t bucket_size, int32_t density, uint64_t seed, const int32_t *frame_index, int32_t seq_id,
                         int32_t *out_index, int32_t *n_out, uint8_t *status, void *stream);

/* recoverPose's per-correspondence mask under a given pose: mask_out [M] = 1 where the correspondence triangulates in front of
 * both cameras and nearer than triangulation_max_depth (and e_mask, when given, is set) -- `mask_bool & mask_e_bool` of
 * src/thirdparty/MonocularVO/visual_odometry.py:134-136; exactly the correspondences mvosr_triangulate_frames keeps, in order. */
int  mvosr_pose_mask_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                            const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                            const uint8_t *e_mask, const double *poses, uint8_t *mask_out, void *stream);

/* Essential matrix by five-point RANSAC -- replaces cv2.findEssentialMat(px_cur, px_ref, cameraMatrix=K, method=cv2.RANSAC,
 * prob=0.999, threshold=0.5) (src/thirdparty/MonocularVO/visual_odometry.py:100-102,129-130; SURVEY N1, first half).  Per
 * frame: `hypotheses` minimal samples of five correspondences drawn from the Philox stream (key = seed, counter = (hypothesis,
 * frame_index ? frame_index[f] : f, seq_id, 1|2), csrc/five_point.cuh), every real solution of the five-point problem scored
 * by the Sampson distance against threshold_px / ((fx + fy) / 2) (OpenCV's normalisation of its pixel threshold); the
 * candidate with the most inliers wins, ties to the lowest (hypothesis, candidate) pair.  `hypotheses` is the maximum (OpenCV's
 * maxIters, 1000 in the reference's call); confidence (OpenCV's prob, 0.999 in the reference's call) > 0 stops a frame after the
 * first round of 128 hypotheses at whose end  tried >= log(1 - confidence) / log(1 - w^5), w = best inlier ratio -- OpenCV's
 * adaptive count evaluated per round; confidence = 0 runs all hypotheses.  (OpenCV draws from its own RNG, which cannot be
 * reproduced; parity is defined on this stream.)
 * Outputs: essential [F][9] row-major with unit Frobenius norm, x_ref^T E x_cur = 0 in normalised coordinates -- the input of
 * mvosr_recover_pose_frames; e_mask_out [M] (optional) the winner's inlier mask -- the e_mask of the later stages; n_inliers
 * [F] (optional); best_hyp [F] (optional; -1 and a zero matrix for frames with fewer than five correspondences or no
 * solution -- e.g. a frame without motion, whose epipolar system is rank deficient; the reference gates such frames before its
 * estimator); hyps_used [F] (optional): hypotheses tried.  Non-finite correspondences are never inliers. */
int  mvosr_find_essential_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                                 const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                                 int32_t hypotheses, double threshold_px, double confidence, uint64_t seed, const int32_t *frame_index, int32_t seq_id,
                                 double *essential, uint8_t *e_mask_out, int32_t *n_inliers, int32_t *best_hyp, int32_t *hyps_used, void *stream);

/* Dense depth from the mesh -- Reconstruct.depth_generate (src/reconstruct.py:91-107): tri.find_simplex of every integer
 * pixel (u, v) of a width x height image + the depth of that triangle's plane along the pixel's ray,
 * depth = h / (n . ((u-cx)/fx, (v-cy)/fy, 1)).  tri: [T][3] int32 into uv ([N][2] float64 pixel coordinates); datas:
 * [T][4] float64 rows (unit normal, height) as Reconstruct.triangle_model returns them (src/reconstruct.py:70-90).
 * Outputs: depth [height][width] float64 (0 outside the triangulation), tri_id [height][width] int32 (-1 outside; on a
 * shared edge the smaller triangle index). */
int  mvosr_depth_from_mesh(mvosr_handle *h, int32_t width, int32_t height, double fx, double fy, double cx, double cy,
                           int32_t n_tri, const int32_t *tri, const double *uv, const double *datas,
                           double *depth, int32_t *tri_id, void *stream);

/* Profiling aid: when set (device pointer, [F][16] int64), the fused kernel stores per-frame SM cycles per phase:
 * 0 load+stage1+ROI, 1 grid#1, 2/3 stars#1 thread/warp path, 6 compaction+grid#2, 7/8 stars#2 thread/warp path,
 * 10 planes, 11 median, 12 valid list, 13 RANSAC. Pass NULL to disable. */
int  mvosr_set_phase_timing(mvosr_handle *h, int64_t *phase_cycles_device);

/* Number of kernels the library has launched on this handle since creation (bench.py's gpu_launches). */
int64_t mvosr_launch_count(const mvosr_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* MVOSR_H */
