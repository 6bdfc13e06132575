/*
 * svoxb.h -- C ABI of libsvoxb: the B200-native (sm_100a) replacement for the octree volume-rendering
 * hot path of HaiminLuo/svox_t.
 *
 * This header is the drop-in boundary. Every entry point below replaces one function that the reference
 * exposes through its pybind11 module `svox_t.csrc` (reference: svox_t/csrc/svox.cpp:119-144); the
 * reference file:line is cited per function. Conventions:
 *   - plain C: raw DEVICE pointers + sizes, no C++/torch types; all tensors contiguous, row-major;
 *   - outputs are caller-allocated (the reference allocates them inside C++ with torch::zeros/empty);
 *   - every call takes the CUDA stream to launch on (the reference always uses the legacy default stream);
 *   - every call returns 0 on success or a negative SVOXB_E* code; svoxb_last_error() gives the message
 *     (the reference raises TORCH_CHECK exceptions and only printf()s CUDA launch errors,
 *     include/common.cuh:108-111);
 *   - no host<->device synchronisation inside any call unless stated.
 * The library never falls back to a CPU path: without a CUDA device every compute call fails with
 * SVOXB_ECUDA.
 */
#ifndef SVOXB_H_
#define SVOXB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVOXB_ABI_VERSION 10

#if defined(__GNUC__)
#define SVOXB_API __attribute__((visibility("default")))
#else
#define SVOXB_API
#endif

enum {
    SVOXB_OK = 0,
    SVOXB_EINVAL = -1,      /* bad argument (null pointer, unsupported N / D / format ...) */
    SVOXB_ECUDA = -2,       /* CUDA runtime / launch error */
    SVOXB_EUNSUPPORTED = -3 /* feature of the reference deliberately not implemented */
};

/* DataFormat enum of the reference (include/data_spec.hpp:45-50). RGBA is the feature-level format of the hot path
 * (output = D-1 sigmoid features + opacity); SH/SG/ASG rows hold basis_dim coefficients per output channel
 * (output = (D-1)/basis_dim channels + opacity, rt_kernel.cu:1352-1358). */
enum { SVOXB_FORMAT_RGBA = 0, SVOXB_FORMAT_SH = 1, SVOXB_FORMAT_SG = 2, SVOXB_FORMAT_ASG = 3 };

/* Opaque acceleration structure built from (child, data) by svoxb_accel_create(): a dense top grid plus
 * dense bricks holding one tagged 32-bit word per cell (leaf depth | feature-row index). */
typedef struct svoxb_accel svoxb_accel;

/* Replaces TreeSpec / PackedTreeSpec (include/data_spec.hpp:67-111, include/data_spec_packed.cuh:57-100).
 * _weight_accum is an argument of svoxb_accumulate_weights; joint_features / skinning_weights / joint_index are arguments of the
 * motion-feature entry points instead of fields. */
typedef struct svoxb_tree {
    const float* features;      /* [M, D] float32; last channel = sigma                               */
    int64_t M;                  /* features.size(0); a data index >= M marks an empty leaf            */
    int32_t D;                  /* features.size(1)                                                   */
    int32_t N;                  /* branching factor per axis (child.size(1)); N == 2 is the fast path */
    const int32_t* child;       /* [n_nodes, N, N, N]  relative child offset, 0 = leaf                */
    const int32_t* data;        /* [n_nodes, N, N, N, 1] feature-row index per leaf slot              */
    const int32_t* parent_depth;/* [n_nodes, 2] (packed parent slot, depth); may be NULL              */
    int64_t n_nodes;            /* rows allocated in child/data (capacity)                            */
    int64_t n_internal;         /* rows in use (TreeSpec.n_internal = tree.filled)                    */
    const float* offset;        /* [3] device; world -> tree: q = offset + scaling * q                */
    const float* scaling;       /* [3] device (the reference's invradius)                             */
    const svoxb_accel* accel;   /* optional; NULL = walk child/data exactly like the reference        */
    const float* features_act;  /* optional [M, D]: features with the sigmoid already applied to channels  */
                                /* 0..D-2 (svoxb_activate_features); the march kernels then skip it       */
    const float* extra_data;    /* optional [extra_rows, extra_cols]: SG rows (lambda, mu[3]); ASG rows (a, b, x[3],  */
    int32_t extra_rows;         /* y[3], z[3]) -- the basis parameters of rt_kernel.cu:116-140. Unused for RGBA / SH.  */
    int32_t extra_cols;
    const float* transformation_matrices; /* optional [M,4,4]: per-row rotation of the view direction before the     */
                                /* basis is evaluated (rt_kernel.cu:283-291). No effect for RGBA.                    */
    int32_t features_act_stride;/* row stride of features_act in floats: 0 or D; for D % 4 != 0 the table holds the D-1   */
                                /* payload channels only, stride = D-1 rounded up to a multiple of 4 (aligned rows)   */
    const float* features_sigma;/* ... and the sigma channel travels as this compact [M] array (svoxb_activate_features)*/
    int32_t accel_marks_current;/* non-zero: the caller asserts that svoxb_accel_mark_hits(accel, features, ...) ran     */
                                /* after the last change of `features`; the march then skips rows marked sigma <= 0   */
} svoxb_tree;

/* Field-for-field the reference's RenderOptions (include/data_spec.hpp:129-145). */
typedef struct svoxb_render_options {
    float step_size;
    float background_brightness;
    int32_t format;
    int32_t basis_dim;
    int32_t ndc_width;          /* < 0 disables NDC (renderer.py:426). Like the reference, only the IMAGE entry   */
                                /* points convert their camera rays to NDC (rt_kernel.cu:1168-1191, 1204)          */
    int32_t ndc_height;
    float ndc_focal;
    int32_t min_comp;
    int32_t max_comp;
    float sigma_thresh;
    float stop_thresh;
} svoxb_render_options;

/* Replaces CameraSpec (include/data_spec.hpp:113-126). */
typedef struct svoxb_camera {
    const float* c2w;           /* device, row-major [3 or 4, 4] camera-to-world, OpenGL convention   */
    float fx, fy;
    int32_t width, height;
    int32_t row_begin, row_end; /* row_end > 0: render only image rows [row_begin, row_end) -- out / depth / grad_out   */
                                /* then hold (row_end - row_begin) x width pixels. 0, 0 = the whole image. One camera   */
                                /* frame split into row bands is how a single view shards over GPUs.        */
} svoxb_camera;

/* ---- housekeeping ------------------------------------------------------------------------------ */
SVOXB_API int svoxb_abi_version(void);
SVOXB_API const char* svoxb_last_error(void);          /* thread-local message of the last failing call         */
SVOXB_API int64_t svoxb_launch_count(void);            /* number of kernels this library has launched so far    */
SVOXB_API int svoxb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- acceleration structure (no reference counterpart; derived data, rebuilt when child/data change) */
/* max_depth <= 0: derive it from parent_depth (must then be non-NULL). Synchronises the stream (one small read-back
 * per stage). Memory comes from the device's default stream-ordered pool on `stream`; svoxb_accel_destroy releases it
 * in stream order on that same stream, after the work this library launched with the accelerator on other streams
 * (the four most recent distinct streams are remembered and waited for with an event each). */
SVOXB_API int svoxb_accel_create(const svoxb_tree* tree, int max_depth, void* stream, svoxb_accel** out);
/* With max_depth > 0 and depth <= 8 (two stages) NOTHING synchronises: the second stage is allocated for the most bricks
 * the top grid can point to, brick counts and flags stay on the device (svoxb_accel_describe fetches them), and
 * svoxb_accel_rebuild refills such an accelerator for another child/data of the same max_depth on the same stream
 * without any allocation (SVOXB_EUNSUPPORTED if it cannot: create a new one) -- the per-frame rebuild path. */
SVOXB_API int svoxb_accel_rebuild(svoxb_accel* accel, const svoxb_tree* tree, int max_depth, void* stream);
SVOXB_API void svoxb_accel_destroy(svoxb_accel* accel);
SVOXB_API int64_t svoxb_accel_bytes(const svoxb_accel* accel);
SVOXB_API int svoxb_accel_describe(const svoxb_accel* accel, int* n_stages, int* bits /*[4]*/, int64_t* bricks /*[4]*/);

/* Hit marks (derived data, refresh whenever features change): flags every leaf cell whose row has !(sigma > 0), the
 * negation of the march's hit predicate (rt_kernel.cu:279 at the default threshold; :382, :456). With
 * svoxb_tree.accel_marks_current set, the march kernels treat such rows as non-candidates and never fetch them
 * (about 20 % of the row traffic in the reference's headline configuration); results are unchanged. Mutates the
 * accelerator in place, in stream order on `stream` (~0.03 ms for 2M leaves). */
SVOXB_API int svoxb_accel_mark_hits(svoxb_accel* accel, const float* features, int64_t M, int32_t D, void* stream);

/* ---- per-row activation (no reference counterpart; derived data, rebuilt whenever features change) ------------ */
/* out[i, c] = sigmoid(features[i, c]) for c < D-1, out[i, D-1] = features[i, D-1] (sigma stays raw). A leaf row is
 * visited by ~77 rays in the reference's headline configuration, so applying the sigmoid once per row instead of once
 * per visit removes almost all transcendental work from the march. One streaming pass over the table. */
/* out_stride (floats): 0 or D writes the plain [M, D] layout (sigma_out ignored). For D % 4 != 0 pass
 * out_stride = D-1 rounded up to a multiple of 4 and sigma_out[M]: the table then holds the activated payload channels
 * in 16-byte aligned, zero-padded rows and sigma goes to the compact array (svoxb_tree.features_act_stride /
 * features_sigma). With them every width runs on the 128-bit row kernels: D = 33 costs what D = 32 costs plus one
 * 4-byte gather per sample, instead of 4x as much on the scalar-lane kernels. */
SVOXB_API int svoxb_activate_features(const float* features, int64_t M, int32_t D, float* out, int32_t out_stride,
                            float* sigma_out, void* stream);

/* sigma_out[i] = features[i, D-1]: the compact sigma array (svoxb_tree.features_sigma) on its own, for the marches that
 * read nothing else of a row -- svoxb_render_depth, svoxb_opacity_render_fwd/_bwd, svoxb_motion_render: 4 M bytes that
 * stay in L2 instead of one 32-byte sector of the [M, D] table per sample. Those entry points also honour the hit marks
 * (svoxb_tree.accel_marks_current): rows with sigma <= 0 are not fetched at all. Both are pure accelerations. */
SVOXB_API int svoxb_gather_sigma(const float* features, int64_t M, int32_t D, float* sigma_out, void* stream);

/* ---- per-step table pass (no reference counterpart) --------------------------------------------------------------- */
/* Everything a training step needs before its marches because `features` changed, in ONE streaming pass over the rows
 * (D % 4 == 0, 16-byte aligned tables; other shapes run the separate passes above, same results):
 *   - act[M, act_stride] (+ sigma_out): as svoxb_activate_features;
 *   - accel != NULL: the hit marks of svoxb_accel_mark_hits, refreshed through the accelerator's inverse map
 *     row -> leaf cell (falls back to the pass over the cells when a row is held by several leaves);
 *   - zero_table != NULL: zero-fill of that [M, D] float table -- the gradient table the backward of this step
 *     reduces into (the reference allocates zeros_like(features) per backward, rt_kernel.cu:1415).
 * 4 M D bytes read, 4 M D (+ 4 M D) written; C3: 0.11 ms against 0.19 ms for three separate passes. */
SVOXB_API int svoxb_prepare_step(svoxb_accel* accel, const float* features, int64_t M, int32_t D, float* act,
                       int32_t act_stride, float* sigma_out, float* zero_table, void* stream);

/* ---- octree descent ---------------------------------------------------------------------------- */
/* query_vertical, first kernel (svox_kernel.cu:66-81, 274-302): per point p (world coords) the leaf's packed
 * slot id node*N^3 + u*N^2 + v*N + w -> node_ids[q]; if the leaf holds a row (data idx < M): data_ids[q] = idx
 * and values[q,:] = features[idx,:]; rows of empty leaves are left untouched, as in the reference.
 * values / data_ids / slot_mask may be NULL. slot_mask[n_internal*N^3] (uint8, zeroed by the caller) gets 1 at
 * every visited leaf slot (empty leaves included, svox_kernel.cu:57-58). */
SVOXB_API int svoxb_query(const svoxb_tree* tree, const float* pts, int64_t Q,
                float* values, int64_t* node_ids, int64_t* data_ids, uint8_t* slot_mask, void* stream);

/* query_vertical, kernels 2+3 (svox_kernel.cu:239-269, 304-320) as a deterministic two-step compaction:
 * _scan counts the set slots (exclusive block offsets into scratch, total into *n_hit_dev); the caller reads
 * n_hit (the one device->host sync the reference also has, svox_kernel.cu:312), allocates leaf_node[n_hit,4]
 * and calls _emit, which writes [node, i, j, k] rows in increasing slot order. */
SVOXB_API size_t svoxb_leafset_scratch_bytes(int64_t n_slots);
SVOXB_API int svoxb_leafset_scan(const uint8_t* slot_mask, int64_t n_slots, void* scratch, int64_t* n_hit_dev, void* stream);
SVOXB_API int svoxb_leafset_emit(const uint8_t* slot_mask, int64_t n_slots, int32_t N, const void* scratch,
                       int64_t* leaf_node, void* stream);

/* construct_tree (svox_kernel.cu:110-121, 341-352): data[leaf(p_i)] = i. data_mut aliases tree->data. */
SVOXB_API int svoxb_construct_tree(const svoxb_tree* tree, int32_t* data_mut, const float* pts, int64_t P, void* stream);

/* query_vertical_backward (svox_kernel.cu:83-95, 380-403): grad_data[row(p_q), 0:K] += grad_out[q, 0:K] for every
 * point whose leaf holds a row; grad_data[M, K] is accumulated into (the caller zero-fills it, as the reference's
 * torch::zeros does). The reference's own kernel faults on a null data_id (svox_kernel.cu:61-62); this is the
 * computation its source states. */
SVOXB_API int svoxb_query_bwd(const svoxb_tree* tree, const float* pts, int64_t Q, const float* grad_out, int32_t K,
                    float* grad_data, void* stream);

/* assign_vertical (svox_kernel.cu:97-108, 326-339): features[row(p_q), 0:K] = values[q, 0:K], K <= D. features_mut
 * aliases tree->features. Where the reference lets racing threads interleave ("only one of them will be taken",
 * svox.py:172-173) the largest point index wins a shared leaf here, whole rows at a time (M ints of stream-ordered
 * scratch). */
SVOXB_API int svoxb_assign(const svoxb_tree* tree, float* features_mut, const float* pts, int64_t Q,
                 const float* values, int32_t K, void* stream);

/* calc_corners (svox_kernel.cu:213-237, 436-457): lower corner, in tree coordinates [0,1)^3, of each cell
 * indexer[q] = [node, i, j, k] (int64), walking parent_depth[n_nodes, 2] up to the root. out[Q, 3]. */
SVOXB_API int svoxb_calc_corners(const int32_t* parent_depth, int32_t N, int64_t n_nodes, const int64_t* indexer,
                       int64_t Q, float* out, void* stream);

/* ---- ray march --------------------------------------------------------------------------------- */
/* volume_render (rt_kernel.cu:654-671, 1362-1379) fused with render_depth (rt_kernel.cu:865-882, 1506-1523):
 * out[Q, Do] with Do = svoxb_out_data_dim(): RGBA: D-1 composited sigmoid features + opacity (1 - T);
 * SH/SG/ASG: (D-1)/basis_dim view-dependent channels + opacity. depth[Q] (nullable, RGBA only) = first-hit depth.
 * vdirs[Q,3] is read by the view-dependent formats only (may be NULL for RGBA). */
SVOXB_API int svoxb_out_data_dim(int32_t format, int32_t basis_dim, int32_t D);
SVOXB_API int svoxb_render_rays_fwd(const svoxb_tree* tree, const float* origins, const float* dirs, const float* vdirs,
                          int64_t Q, const svoxb_render_options* opt, float* out, float* depth, void* stream);

/* volume_render_backward (rt_kernel.cu:674-694, 1402-1426): grad_features[M, D] += dL/dfeatures.
 * One re-march instead of the reference's two: saved_out[Q, D] is the forward output computed with
 * sigma_thresh = 0 and stop_thresh < 0 (the backward's own hit predicate, rt_kernel.cu:382,456); with the
 * default options that is exactly what svoxb_render_rays_fwd returned. It MUST be that unmodified output for the same
 * tree, features, rays and options: the kernel takes T_end = 1 - saved_out[q, D-1] and <grad_out[q], saved_out[q]>
 * from it and cannot tell a stale or foreign buffer from the right one (the result would be plausible, wrong sigma
 * gradients). When in doubt, render again with those options and pass the fresh output. The caller zero-fills
 * grad_features (the reference allocates zeros_like(features), rt_kernel.cu:1415).
 * Widths and types (stated once for every march entry point): the RGBA format takes any D >= 2, like the reference's
 * loop over out_data_dim (rt_kernel.cu:302-306) -- D <= 128 runs on the tuned register / shared-memory kernels, wider
 * tables on the general kernels of svoxb_render_wide.cu (no accelerator, no NDC; same results, several times slower per
 * byte). The view-dependent formats (SH / SG / ASG) take D <= 128 with at most 31 output channels and basis_dim <= 25
 * (SVOXB_EINVAL beyond). These entry points are float32; the reference's float64 instantiation
 * (AT_DISPATCH_FLOATING_TYPES, rt_kernel.cu:1373) is the *_f64 family at the end of this header. */
SVOXB_API int svoxb_render_rays_bwd(const svoxb_tree* tree, const float* origins, const float* dirs, const float* vdirs,
                          int64_t Q, const svoxb_render_options* opt, const float* grad_out, const float* saved_out,
                          float* grad_features, void* stream);

/* The same two calls for a forward/backward PAIR over one ray batch, sharing a scheduling hint (no reference
 * counterpart). For batches of at least svoxb_ray_order_min_rays() rays (about 0.75 per resident lane) the forward also
 * writes ray_cost[Q] (int32: the march iterations of each ray; ray_cost[0] = -1 when the kernel that ran cannot count)
 * and the backward marches the rays longest first by that array (svoxb_order.cu): the rays of a warp are then alike
 * and busy at the same time -- 2^20 rays on the C3 tree 6.15 -> 4.94 ms, 128 k rays (one GPU's share of a batch split
 * over 8 GPUs) 1.10 -> 0.79 ms. Smaller batches neither write nor read ray_cost. Results are those of the plain calls
 * (only the lane a ray runs on, and with it the order of the floating-point reductions, changes). */
SVOXB_API int64_t svoxb_ray_order_max_rays(void);
SVOXB_API int64_t svoxb_ray_order_min_rays(void);
SVOXB_API int svoxb_render_rays_fwd_cost(const svoxb_tree* tree, const float* origins, const float* dirs, const float* vdirs,
                               int64_t Q, const svoxb_render_options* opt, float* out, float* depth, int32_t* ray_cost,
                               void* stream);
SVOXB_API int svoxb_render_rays_bwd_cost(const svoxb_tree* tree, const float* origins, const float* dirs, const float* vdirs,
                               int64_t Q, const svoxb_render_options* opt, const float* grad_out, const float* saved_out,
                               float* grad_features, const int32_t* ray_cost, void* stream);

/* volume_render_image / _backward (rt_kernel.cu:1152-1166, 1193-1238, 1381-1452): pinhole camera rays generated
 * in-kernel; out[H, W, D], depth[H, W] (nullable). */
SVOXB_API int svoxb_render_image_fwd(const svoxb_tree* tree, const svoxb_camera* cam, const svoxb_render_options* opt,
                           float* out, float* depth, void* stream);
SVOXB_API int svoxb_render_image_bwd(const svoxb_tree* tree, const svoxb_camera* cam, const svoxb_render_options* opt,
                           const float* grad_out, const float* saved_out, float* grad_features, void* stream);

/* render_depth alone (rt_kernel.cu:781-834, 865-882, 1506-1523): depth[Q] = delta_scale * t of the first
 * sample with sigma > sigma_thresh, 0 on miss. */
SVOXB_API int svoxb_render_depth(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                       const svoxb_render_options* opt, float* depth, void* stream);

/* opacity_render (rt_kernel.cu:499-560, 1574-1591): out[Q] = 1 - T, same thresholds / early stop as the forward. */
SVOXB_API int svoxb_opacity_render_fwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                             const svoxb_render_options* opt, float* out, void* stream);

/* Backward of opacity_render as the reference WROTE it (opacity_trace_ray_backward, rt_kernel.cu:562-651):
 * grad_features[idx, D-1] += delta_t * delta_scale * grad_out[q] * T_end for every sample with sigma > 0.
 * (The reference's own entry point launches the wrong kernel, rt_kernel.cu:1607; this is the intended semantics.) */
SVOXB_API int svoxb_opacity_render_bwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                             const svoxb_render_options* opt, const float* grad_out, float* grad_features, void* stream);
/* The same with the forward's own output: saved_out[Q] = svoxb_opacity_render_fwd's result for the SAME tree, rays and
 * default thresholds (sigma_thresh = 0, stop_thresh <= 0) gives T_end = 1 - saved_out, so the backward marches once
 * instead of twice. */
SVOXB_API int svoxb_opacity_render_bwd_saved(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                   const svoxb_render_options* opt, const float* grad_out, const float* saved_out,
                                   float* grad_features, void* stream);

/* motion_render (rt_kernel.cu:698-778, 1480-1504): at the first sample with sigma > sigma_thresh:
 * out[Q, J] = distance of the hit point to each row of extra_data[J, 3], depth[Q], hit_point[Q, 3], data_idx[Q]
 * (all zero when nothing is hit). The hit point reproduces the reference's arithmetic (see svoxb_render_x.cu). */
SVOXB_API int svoxb_motion_render(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                        const svoxb_render_options* opt, const float* extra_data, int32_t J, float* out, float* depth,
                        float* hit_point, int64_t* data_idx, void* stream);

/* The accumulate_weights side effect of volume_render / volume_render_image (rt_kernel.cu:266-267, 308-310; TreeSpec.
 * _weight_accum): weight_accum[n_nodes * N^3] (float, caller-zeroed, same shape as child) += T (1 - att) at every hit
 * leaf of every ray. cam != NULL: the camera's pixel rays (origins / dirs / Q ignored); else the explicit batch. Uses the
 * thresholds / early stop of the forward. Separate entry point: the render calls themselves never write it. */
SVOXB_API int svoxb_accumulate_weights(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                             const svoxb_camera* cam, const svoxb_render_options* opt, float* weight_accum, void* stream);

/* motion_feature_render (rt_kernel.cu:885-979, 1525-1543): out[Q, F] += weight * sigmoid(sum_j w_j * JF[joint_j][k]) at
 * every hit, where (w_j, joint_j) are the B skinning weights / joint indices of the hit ROW (skinning_weights[M,B],
 * joint_index[M,B]) and JF = joint_features[J,F]; F <= 127 for large batches (Q * 32 >= M: blend + sigmoid are tabulated once per row and the
 * feature render's kernels march that table), F <= 32 otherwise. No opacity channel; rays that miss the cube return 0. */
SVOXB_API int svoxb_motion_feature_render_fwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                    const svoxb_render_options* opt, const float* joint_features,
                                    const float* skinning_weights, const int32_t* joint_index, int32_t J, int32_t F,
                                    int32_t B, float* out, void* stream);

/* motion_feature_render_backward (rt_kernel.cu:981-1064, 1546-1572) with the arithmetic it was meant to have:
 * grad_joint_features[J,F] (zero-filled here) += w_j * weight * s (1 - s) * grad_out[q,k] over every sample with
 * sigma > 0. (The reference adds into an uninitialised array and indexes it by bone slot) */
SVOXB_API int svoxb_motion_feature_render_bwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                    const svoxb_render_options* opt, const float* joint_features,
                                    const float* skinning_weights, const int32_t* joint_index, int32_t J, int32_t F,
                                    int32_t B, const float* grad_out, float* grad_joint_features, void* stream);

/* grid_weight_render (rt_kernel.cu:1240-1344, 1454-1478): the camera's pixel rays (NDC when opt->ndc_width >= 0)
 * marched through a dense sigma grid[reso, reso, reso] spanning the tree cube; grid_weight[cell] = max compositing
 * weight T (1 - exp(-delta sigma)) any ray left in the cell, grid_hit[cell] = number of hits (sigma > sigma_thresh).
 * Both outputs are accumulated into (caller zero-fills, as the reference's zeros_like). Only step_size, sigma_thresh
 * and ndc_* of the options are read; cam->row_begin/row_end are ignored. */
SVOXB_API int svoxb_grid_weight_render(const float* grid, int32_t reso, const svoxb_camera* cam,
                             const svoxb_render_options* opt, const float* offset, const float* scaling,
                             float* grid_weight, float* grid_hit, void* stream);

/* ---- animated-frame rebuild -------------------------------------------------------------------- */
/* warp_vertices (svox_kernel.cu:123-154, 354-378): linear blend skinning. T[J,4,4], coords[P,3], w[P,B],
 * joint_index[P,B] -> coords_out[P,3], mats_out[P,4,4] (fully written, no zero-fill needed). */
SVOXB_API int svoxb_warp_vertices(const float* T, const float* coords, const float* w, const int32_t* joint_index,
                        int64_t P, int32_t B, float* coords_out, float* mats_out, void* stream);

/* p2v (p2v_kernel.cu:103-151, 240-261): Gaussian splat of point_features[:, F-1] into voxels[n,n,n,1]
 * (zero-filled by this call). corner[3], size[3] are device pointers like in the reference. */
SVOXB_API int svoxb_p2v(const float* points, const float* point_features, int64_t P, int32_t F,
              const float* corner, const float* size, int32_t n_voxels, float kernel_radius, float conv_radius,
              float* voxels, void* stream);

/* warp_vertices_backward (svox_kernel.cu:156-211, 404-436): grad_T[J,4,4] (zero-filled here; rows 0..2 receive),
 * grad_coords[P,3], grad_w[P,B] from the upstream gradients of coords_out[P,3] and mats_out[P,4,4]. */
SVOXB_API int svoxb_warp_vertices_bwd(const float* T, const float* coords, const float* w, const int32_t* joint_index,
                            const float* grad_coords_out, const float* grad_mats_out, int64_t P, int32_t B, int32_t J,
                            float* grad_T, float* grad_coords, float* grad_w, void* stream);

/* p2v_backward (p2v_kernel.cu:153-234, 263-285): grad_points[P,3], grad_features[P,F] (zero-filled here; the value
 * lands in channel 0 exactly as in the reference) from grad_voxels[n,n,n,1]. */
SVOXB_API int svoxb_p2v_bwd(const float* grad_voxels, const float* points, const float* point_features, int64_t P,
                  int32_t F, const float* corner, const float* size, int32_t n_voxels, float kernel_radius,
                  float conv_radius, float* grad_points, float* grad_features, void* stream);

/* One-shot octree build from points (replaces the reference's depth-1 rounds of
 * query_vertical + N3Tree.refine, svox_t/svox.py:488-560 + helpers.py:38-109, and the final construct_tree):
 * emits child/data/parent_depth in the reference tensor format for the octree whose depth-L leaves are the
 * occupied finest cells; nodes are numbered breadth-first, by Morton key within a level (deterministic).
 * data[leaf(p_i)] = the largest i among the points in that leaf. Two calls:
 *   _count: sorts the point keys and returns the node count in *n_nodes_host (synchronises the stream);
 *   _emit : fills the caller-allocated tensors (n_nodes rows). `work` comes from svoxb_build_work_bytes(P, L). */
SVOXB_API size_t svoxb_build_work_bytes(int64_t P, int32_t L);
SVOXB_API int svoxb_build_octree_count(const float* pts, int64_t P, int32_t L, const float* offset, const float* scaling,
                             void* work, int64_t* n_nodes_host, void* stream);
SVOXB_API int svoxb_build_octree_emit(int64_t P, int32_t L, const void* work, int64_t n_nodes,
                            int32_t* child, int32_t* data, int32_t* parent_depth, void* stream);

/* The same build without a sort and without any host read-back, for depths L <= svoxb_build_dense_max_depth() (10):
 * occupancy bitmaps + a pyramid of rank directories (svoxb_build_dense.cu; every kernel hand-written, no library
 * primitives). The caller sizes child / data / parent_depth for a node CAPACITY `cap_nodes`; rows beyond the nodes
 * actually needed are initialised as unreachable empty nodes. status_dev (device, optional) receives
 * [0] = nodes needed, [1] = 1 if that exceeds cap_nodes (the tree is then truncated: do not use it). Same tensors as
 * svoxb_build_octree_count/_emit, bit for bit, when the capacity suffices. `work`: 256-byte aligned device memory of
 * svoxb_build_dense_work_bytes(L) bytes. Nothing synchronises: a whole frame (warp -> splat -> rebuild ->
 * accelerator -> render) can be queued, or captured in a CUDA graph, in one go. */
SVOXB_API int32_t svoxb_build_dense_max_depth(void);
SVOXB_API size_t svoxb_build_dense_work_bytes(int32_t L);
SVOXB_API int svoxb_build_dense(const float* pts, int64_t P, int32_t L, const float* offset, const float* scaling,
                      void* work, int64_t cap_nodes, int32_t* child, int32_t* data, int32_t* parent_depth,
                      int64_t* status_dev, void* stream);

/* ---- float64 instantiation of the hot path ------------------------------------------------------------------- */
/* The reference dispatches AT_DISPATCH_FLOATING_TYPES on every entry point (rt_kernel.cu:1373, 1413, 1517;
 * svox_kernel.cu:290): features, offset, scaling, rays and outputs double, child / data int32, options unchanged
 * (float fields, promoted). These are that instantiation for the RGBA format: point query, ray-batch and camera
 * forward / backward, depth. Same argument meaning as the float32 calls above (saved_out = the forward output under the
 * backward's predicate, caller-zeroed grad_features, band cameras); no accelerator, no NDC, any D >= 2, any N.
 * Arithmetic is true fp64 throughout -- the reference's own double build calls expf() on doubles (rt_kernel.cu:280,
 * 304), so it is float-accurate only; results agree with it to ~1e-6 and with an fp64 evaluation to ~1e-12. */
typedef struct svoxb_tree_f64 {
    const double* features;     /* [M, D] float64; last channel = sigma                              */
    int64_t M;
    int32_t D;
    int32_t N;
    const int32_t* child;       /* [n_nodes, N, N, N]                                                */
    const int32_t* data;        /* [n_nodes, N, N, N, 1]                                             */
    int64_t n_nodes;
    int64_t n_internal;
    const double* offset;       /* [3] device                                                        */
    const double* scaling;      /* [3] device                                                        */
} svoxb_tree_f64;

typedef struct svoxb_camera_f64 {
    const double* c2w;          /* device, row-major [3 or 4, 4]                                     */
    double fx, fy;
    int32_t width, height;
    int32_t row_begin, row_end; /* as svoxb_camera                                                   */
} svoxb_camera_f64;

SVOXB_API int svoxb_query_f64(const svoxb_tree_f64* tree, const double* pts, int64_t Q, double* values,
                    int64_t* node_ids, int64_t* data_ids, uint8_t* slot_mask, void* stream);
SVOXB_API int svoxb_render_rays_fwd_f64(const svoxb_tree_f64* tree, const double* origins, const double* dirs, int64_t Q,
                              const svoxb_render_options* opt, double* out, double* depth, void* stream);
SVOXB_API int svoxb_render_rays_bwd_f64(const svoxb_tree_f64* tree, const double* origins, const double* dirs, int64_t Q,
                              const svoxb_render_options* opt, const double* grad_out, const double* saved_out,
                              double* grad_features, void* stream);
SVOXB_API int svoxb_render_image_fwd_f64(const svoxb_tree_f64* tree, const svoxb_camera_f64* cam,
                               const svoxb_render_options* opt, double* out, double* depth, void* stream);
SVOXB_API int svoxb_render_image_bwd_f64(const svoxb_tree_f64* tree, const svoxb_camera_f64* cam,
                               const svoxb_render_options* opt, const double* grad_out, const double* saved_out,
                               double* grad_features, void* stream);
SVOXB_API int svoxb_render_depth_f64(const svoxb_tree_f64* tree, const double* origins, const double* dirs, int64_t Q,
                           const svoxb_render_options* opt, double* depth, void* stream);

/* ---- multi-GPU exchange (no reference counterpart: the reference is single-GPU) --------------- */
/* The path's one exchange step: the leaf-gradient table grad[M, D] -- the reference's
 * zeros_like(features), rt_kernel.cu:1415 -- summed over the GPUs of one node, IN PLACE, by one kernel per rank over a
 * symmetric allocation (the same buffer mapped into every rank's address space, optionally with an NVSwitch multicast
 * mapping). The caller owns the allocation and the rendezvous (svox_t_b200/dist.py uses
 * torch.distributed._symmetric_memory); this struct carries raw addresses only. Every rank must call with the same
 * n_floats, blocks and epoch sequence, on a stream whose earlier work has produced the rank's table. */
typedef struct svoxb_peer_group {
    int32_t rank, world;        /* world <= 16 (one NVSwitch domain)                                            */
    void* const* buffers;       /* HOST array [world]: this process's mapping of every rank's buffer            */
    void* multicast;            /* multicast mapping of the buffer (multimem.ld_reduce / multimem.st), or NULL: */
                                /* the kernel then reads / writes the peers' mappings directly                  */
    int64_t table_offset;       /* byte offset of the float table inside the buffer (16-byte aligned)           */
    int64_t flags_offset;       /* byte offset of blocks * world uint32 flag words, zero at set-up              */
    int64_t status_offset;      /* byte offset of one uint32: 0, or 1 + the rank a barrier gave up waiting for  */
    int32_t blocks;             /* CTAs per rank (<= SM count: block b of every rank meets block b of the others) */
    uint32_t epoch;             /* first call 1, then += 2 per call (the call's two barriers use epoch, epoch+1) */
} svoxb_peer_group;

SVOXB_API int svoxb_exchange_max_blocks(void);
SVOXB_API int svoxb_exchange_sum(const svoxb_peer_group* group, int64_t n_floats, void* stream);
/* The same for the gradient table grad[M, D] that svoxb_render_*_bwd produced for `features` (this rank's replica of the
 * feature table): rows with !(features[r, D-1] > 0) are skipped. The backward's hit predicate is sigma > 0
 * (rt_kernel.cu:382, 456), so those rows hold zeros on every rank; skipping them saves their share of the NVLink
 * traffic (20 % in the reference's headline configuration). D % 4 != 0 or features == NULL: the dense form. */
SVOXB_API int svoxb_exchange_sum_rows(const svoxb_peer_group* group, int64_t M, int32_t D, const float* features, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVOXB_H_ */
