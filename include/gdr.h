/*
 * gdr.h -- C ABI of the B200-native Gaussian rasterizer (libgdr.so).
 *
 * This is the drop-in boundary for the reference's native rasterizer module
 * `diff_gaussian_rasterization._C` (reference tree, RAST/ =
 * third_party/diff-gaussian-rasterization/):
 *
 *   reference entry point (RAST/ext.cpp:15-19)                    replaced by
 *   ----------------------------------------------------------    ---------------------------------
 *   rasterize_gaussians  = RasterizeGaussiansCUDA                  gdr_forward_project + gdr_forward_render
 *       (RAST/rasterize_points.h:18-38, rasterize_points.cu:35-119;
 *        Rasterizer::forward RAST/cuda_rasterizer/rasterizer_impl.cu:197-339)
 *   rasterize_gaussians_backward = RasterizeGaussiansBackwardCUDA  gdr_backward
 *       (RAST/rasterize_points.h:40-65, rasterize_points.cu:121-208;
 *        Rasterizer::backward rasterizer_impl.cu:343-447)
 *   mark_visible = markVisible                                     gdr_mark_visible
 *       (RAST/rasterize_points.h:67-70, rasterize_points.cu:210-229)
 *   the three resizable byte buffers geomBuffer / binningBuffer /  gdr_*_bytes size queries; the caller
 *   imgBuffer (rasterize_points.cu:27-33,75-80;                    allocates and owns every buffer
 *   rasterizer_impl.h:22-73)
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer to contiguous FP32/INT32 data unless the
 *     parameter name ends in `_host`.  A null pointer means "not provided"
 *     exactly as an empty tensor does in the reference (colors_precomp == NULL
 *     -> colours from SH; cov3D_precomp == NULL -> covariance from
 *     scales/rotations).
 *   - viewmatrix / projmatrix are the reference's transposed (row-vector) 4x4
 *     matrices, 16 floats each, on the device; campos and bg are 3 device floats.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*).  No
 *     entry point synchronises the device or the stream.
 *   - Return value: GDR_OK (0) or a negative error code; gdr_last_error() returns
 *     a thread-local message.  Nothing throws across the ABI.
 *   - The library keeps no state between calls; all buffers (including the state
 *     saved between forward and backward) are caller-owned.
 *
 * The forward is split in two so that the host never has to drain the GPU to
 * learn the instance count R (the reference blocks on a cudaMemcpy of it,
 * rasterizer_impl.cu:282):
 *   1. gdr_forward_project  -- per-Gaussian projection AND binning in one kernel: every
 *      (Gaussian, tile) instance claims a slot of its tile's key segment (`sort_scratch`,
 *      `tile_capacity` slots per tile, chosen by the caller); enqueues an async copy of
 *      {R, flags, largest per-tile count, 0} into `counts_host` (pinned memory).
 *   2. gdr_forward_render   -- per-tile depth sort + record gather, blend; runs with a
 *      caller-chosen stream capacity.
 * Both capacities are predictions (the Python shim keeps the previous frame's values).  The
 * kernels stay in bounds whatever they are; the caller compares the counts with what it
 * passed and repeats step 1 (largest per-tile count > tile_capacity) or step 2
 * (R > capacity, with GDR_FLAG_RERUN) with larger buffers.  See INTEGRATION.md.
 */
#ifndef GDR_H_INCLUDED
#define GDR_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GDR_API __attribute__((visibility("default")))
#else
#define GDR_API
#endif

#define GDR_OK 0
#define GDR_ERR_INVALID_ARGUMENT (-1)
#define GDR_ERR_CUDA (-2)
#define GDR_ERR_UNSUPPORTED (-3)

#define GDR_ABI_VERSION 2

/* bit flags for gdr_backward(grad_mask): which input gradients the caller needs */
#define GDR_GRAD_MEANS2D 1
#define GDR_GRAD_MEANS3D 2
#define GDR_GRAD_COLOR 4   /* dL/dsh or dL/dcolors_precomp */
#define GDR_GRAD_OPACITY 8
#define GDR_GRAD_COV 16    /* dL/dscales + dL/drotations, or dL/dcov3D_precomp */
#define GDR_GRAD_ALL 31
#define GDR_GRAD_RAW_PARAMS 32 /* the forward ran with GDR_FLAG_RAW_PARAMS: return dL/dopacity, dL/dscales, dL/drotations
                                  w.r.t. the RAW parameters (through sigmoid / exp / normalise); `scales` and `rotations`
                                  passed to the backward are the raw ones */

#define GDR_GRAD_HWC_COLOR 64   /* the forward ran with GDR_FLAG_FUSED_EPILOGUE: dL_dout_color is [H][W][3] (the gradient of
                                  the clamped image); no gradient flows through a channel the clamp cut, as with
                                  torch.clamp */

#define GDR_GRAD_SCRATCH_CLEAN 128 /* gdr_backward / gdr_views_backward: backward_scratch holds zeros on entry (a buffer the
                                  caller keeps per stream, zero-filled once).  The call then skips its own fill, and the
                                  per-Gaussian kernel stores zeros back over every accumulator row after consuming it, so
                                  the buffer holds zeros again when the enqueued work has run.  A call that FAILS leaves
                                  the buffer in an unknown state: zero-fill it (or drop it) before the next use. */

/* flags for gdr_forward_project / gdr_forward_render (must be identical in both calls of a frame) */
#define GDR_FLAG_NO_TILE_CULL 1 /* bin every tile of the reference's 3-sigma rectangle, exactly like the reference.
                                   Default (0): drop (Gaussian, tile) pairs that provably cannot reach
                                   alpha = 1/255 in the tile -- outputs are unchanged, R shrinks. */
#define GDR_FLAG_RAW_PARAMS 2   /* activation-fused inputs: `opacities` are logits, `scales` log-scales,
                                   `rotations` un-normalised quaternions; sigmoid / exp / normalise (what
                                   Renderer.render_img applies first, lightning/renderer.py:95-101, 225-230) run inside
                                   the projection kernel with torch's exact roundings.  gdr_forward_project only. */
#define GDR_FLAG_RERUN 4        /* gdr_forward_render only: this render repeats an earlier one of the same projection
                                   (larger capacity); resets the stream cursor and the tile order lists first */
#define GDR_FLAG_FUSED_EPILOGUE 8 /* gdr_forward_render / gdr_views_forward_render: fuse the epilogue of Renderer.render_img
                                   (lightning/renderer.py:261-269) into the blend -- out_color is written as the image
                                   clamped to [0, 1] in [H][W][3] layout ([V][H][W][3] for a batch); out_depth / out_alpha
                                   are unchanged (their HW1 / HW forms are views).  The matching backward takes the
                                   gradient w.r.t. that image: pass GDR_GRAD_HWC_COLOR in grad_mask */

/* counts_host: 4 int32 per view in PINNED host memory that the device can address (cudaHostAlloc / cudaHostRegister,
 * unified addressing).  The projection kernel's last CTA stores words 0..2 and then word 3 = 1 with release / system
 * semantics: the caller zeroes word 3 before the call and polls it (or waits for the stream) -- no copy node or event
 * sits between the projection kernel and the kernels that follow it in the stream. */
#define GDR_COUNT_RENDERED 0  /* R: instances binned */
#define GDR_COUNT_FLAGS 1     /* GDR_COUNT_FLAG_* */
#define GDR_COUNT_MAX_TILE 2  /* largest per-tile instance count; > tile_capacity => keys were dropped: re-run */
#define GDR_COUNT_READY 3     /* set to 1 after the other words */
#define GDR_COUNT_FLAG_PREFILTERED 2 /* prefiltered != 0 but a Gaussian failed the near-plane test (the reference traps
                                        on the device here, auxiliary.h:154-158) */

GDR_API int gdr_abi_version(void);
GDR_API const char* gdr_last_error(void);

/* Size queries for the caller-owned opaque buffers (all 256-byte aligned device memory). */
GDR_API int gdr_geom_state_bytes(int P, int64_t* bytes);              /* per-Gaussian state (reference: GeometryState) */
GDR_API int gdr_image_state_bytes(int W, int H, int64_t* bytes);      /* per-pixel/per-tile state (reference: ImageState) */
GDR_API int gdr_splat_stream_bytes(int64_t capacity, int64_t* bytes); /* depth-sorted per-tile instance stream, saved for backward */
GDR_API int gdr_sort_scratch_bytes(int W, int H, int64_t tile_capacity, int64_t* bytes); /* per view; temporary: tiles x
                                                                         tile_capacity 8-byte keys (tile_capacity a positive
                                                                         multiple of 32), free after gdr_forward_render */
GDR_API int gdr_sort_scratch_exact_bytes(int64_t num_keys, int64_t* bytes); /* per view, the exact key layout (below):
                                                                         num_keys (a positive multiple of 32, >= R) keys */
GDR_API int gdr_backward_scratch_bytes(int P, int64_t* bytes);        /* temporary screen-space gradient accumulators */

/* Step 1 of the forward (replaces the first half of Rasterizer::forward, rasterizer_impl.cu:197-282). */
GDR_API int gdr_forward_project(int P, int sh_degree, int M, int W, int H,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* opacities, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos,
                        float tan_fovx, float tan_fovy, int prefiltered,
                        int32_t* radii /* out [P] */, void* geom_state, void* image_state,
                        void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets /* NULL: uniform layout */,
                        int32_t* counts_host /* pinned host memory [4] or NULL */, int flags, void* stream);

/* The two layouts of the key segments in sort_scratch (both steps of a frame must use the same):
 *   uniform (tile_offsets == NULL)  tile t owns keys [t * tile_capacity, (t + 1) * tile_capacity): one projection pass,
 *       nothing to scan -- the steady state, with tile_capacity predicted from the previous frame's
 *       GDR_COUNT_MAX_TILE.  A tile that receives more instances keeps counting (GDR_COUNT_MAX_TILE is exact) but
 *       drops the keys: project again.
 *   exact (tile_offsets != NULL)    tile t owns keys [tile_offsets[t], tile_offsets[t + 1]) and `tile_capacity` is the
 *       size of one view's key region (a multiple of 32, >= R; gdr_sort_scratch_exact_bytes).  tile_offsets comes from
 *       gdr_tile_offsets, which scans the per-tile counts a previous projection pass of the SAME inputs left in
 *       image_state: exactly R keys whatever the distribution -- the way out when the prediction failed, and the layout
 *       for scenes whose densest tile is so far above the average that uniform segments would waste memory.
 * gdr_tile_offsets: V = 1 for the single-view entry points; tile_offsets is device memory, [V][tiles + 1] uint32. */
GDR_API int gdr_tile_offsets(int V, int W, int H, const void* image_states, uint32_t* tile_offsets, void* stream);

/* Step 2 of the forward (replaces rasterizer_impl.cu:304-337). out_* are [3,H,W], [1,H,W], [1,H,W].
 * sort_scratch / tile_capacity are the ones step 1 filled. */
GDR_API int gdr_forward_render(int P, int W, int H, const float* bg,
                       const void* geom_state, void* image_state,
                       void* splat_stream, void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets,
                       int64_t capacity, float* out_color, float* out_depth, float* out_alpha, int flags, void* stream);

/* Backward (replaces Rasterizer::backward, rasterizer_impl.cu:343-447).  dL_dout_depth / dL_dout_alpha
 * may be NULL (treated as zero).  Output gradient pointers may be NULL when the matching bit of
 * grad_mask is clear; requested outputs are fully written (zeros for culled Gaussians), no pre-zeroing
 * needed.  dL_dmeans2D is [P,4] = (d/dx_ndc, d/dy_ndc, sum|.|, sum|.|) as in backward.cu:589-594. */
GDR_API int gdr_backward(int P, int sh_degree, int M, int W, int H, const float* bg,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* campos, float tan_fovx, float tan_fovy, const int32_t* radii,
                 const void* geom_state, const void* image_state, const void* splat_stream,
                 int64_t capacity, const float* out_alpha,
                 const float* dL_dout_color, const float* dL_dout_depth, const float* dL_dout_alpha,
                 void* backward_scratch, int grad_mask,
                 float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drotations, void* stream);

/* present[i] = view-space z > 0.2 (rasterizer_impl.cu:54-66). present is P bytes (bool). */
GDR_API int gdr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* ---- Batched multi-view entry points -------------------------------------------------
 * One set of Gaussians rendered from V cameras with ONE launch per stage (grid.y = view) instead of V
 * passes through the single-view entry points -- the shape of the reference's per-view loops
 * (lightning/network.py:827-838, 848-856, 964-972: `for j, c2w in enumerate(tar_c2ws)` around
 * Renderer.render_img, lightning/renderer.py:209-272).  Cameras live in DEVICE memory as an array of
 * gdr_camera blocks (what MiniCam, lightning/utils.py:22-48, plus Renderer.set_rasterizer,
 * renderer.py:106-126, hand to the rasterizer per view).  All V views share W, H, sh_degree and
 * scale_modifier.  Per-view buffers are the single-view buffers repeated V times back to back:
 *     radii [V][P];  geom_states V * gdr_geom_state_bytes(P);  image_states V * gdr_image_state_bytes(W,H);
 *     splat_streams gdr_splat_stream_bytes(V * capacity_per_view)  (view v starts at record v * capacity);
 *     sort_scratch  V * gdr_sort_scratch_bytes(W, H, tile_capacity);
 *     images [V][3|1][H][W];  counts_host [V][4] (pinned);  backward_scratch gdr_backward_scratch_bytes(V * P).
 * gdr_views_backward SUMS the gradients of the shared Gaussians over the views (what autograd does when
 * the reference renders the views one by one from the same tensors), deterministically in view order. */
typedef struct gdr_camera {
    float viewmatrix[16]; /* transposed world->view (MiniCam.world_view_transform) */
    float projmatrix[16]; /* transposed full projection (MiniCam.full_proj_transform) */
    float campos[3];      /* MiniCam.camera_center */
    float tan_fovx, tan_fovy;
    float bg[3];
    float reserved[8];
} gdr_camera; /* 48 floats */
#define GDR_CAMERA_FLOATS 48

GDR_API int gdr_views_forward_project(int V, int P, int sh_degree, int M, int W, int H,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* opacities, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp,
                        const gdr_camera* cameras /* device [V] */, int prefiltered,
                        int32_t* radii /* out [V][P] */, void* geom_states, void* image_states,
                        void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets /* [V][tiles+1] or NULL */,
                        int32_t* counts_host /* pinned [V][4] or NULL */, int flags, void* stream);
GDR_API int gdr_views_forward_render(int V, int P, int W, int H, const gdr_camera* cameras,
                        const void* geom_states, void* image_states, void* splat_streams, void* sort_scratch,
                        int64_t tile_capacity, const uint32_t* tile_offsets, int64_t capacity_per_view,
                        float* out_color /*[V,3,H,W]*/,
                        float* out_depth /*[V,1,H,W]*/, float* out_alpha /*[V,1,H,W]*/, int flags, void* stream);
GDR_API int gdr_views_backward(int V, int P, int sh_degree, int M, int W, int H,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* scales, float scale_modifier, const float* rotations,
                        const float* cov3D_precomp, const gdr_camera* cameras, const int32_t* radii,
                        const void* geom_states, const void* image_states, const void* splat_streams,
                        int64_t capacity_per_view, const float* out_alpha /*[V,1,H,W]*/,
                        const float* dL_dout_color /*[V,3,H,W]*/, const float* dL_dout_depth /* or NULL */,
                        const float* dL_dout_alpha /* or NULL */, void* backward_scratch, int grad_mask,
                        float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D,
                        float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drotations, void* stream);

/* ---- The densify select either side of the path, fused on the device -----------------
 * Reference: lightning/network.py:865-893 -- vjp of the image MSE through the n_views_sel-view render
 * w.r.t. ONE shared [P,4] screen-space tensor, then grad[mask][:, 2:4].norm(dim=-1) -> torch.topk(k_num) ->
 * boolean mask, followed by boolean-mask gathers of the selected / non-selected sets (:905-915, 955-959).
 *   gdr_mse_grad              dL/dcolor [V,3,H,W] of mean((clamp(color,0,1) - target)^2), target [V,H,W,3]
 *                             (renderer.py:261 clamp + network.py:855-862 loss); *loss (device, may be NULL).
 *   gdr_views_densify_scores  means2D-only backward blend of all V views, summed over the views:
 *                             grad_means2D [P,4] (may be NULL) and scores [P] = ||grad[:,2:4]||, or -1 where
 *                             candidate_mask[i] == 0 (candidate_mask NULL = every Gaussian is a candidate).
 *   gdr_topk_select           selected[i] = 1 for the k candidates with the largest scores (all candidates if
 *                             there are fewer than k; ties at the threshold go to the lowest indices), plus
 *                             selected_idx / rest_idx (ascending Gaussian indices of the selected / the
 *                             non-selected candidates) and counts[2] = their lengths -- all device memory,
 *                             any output may be NULL; exact radix select, no sort, no host round trip. */
GDR_API int gdr_mse_grad(int V, int W, int H, const float* color, const float* target, float* dL_dcolor, float* loss,
                        void* stream);
GDR_API int gdr_views_densify_scores(int V, int P, int W, int H, const gdr_camera* cameras, const void* image_states,
                        const void* splat_streams, int64_t capacity_per_view, const float* out_alpha,
                        const float* dL_dout_color, void* backward_scratch /* gdr_backward_scratch_bytes(V*P) */,
                        const uint8_t* candidate_mask, float* grad_means2D, float* scores, void* stream);
GDR_API int gdr_topk_select(int P, const float* scores, int k, uint8_t* selected, int32_t* selected_idx,
                        int32_t* rest_idx, int32_t* counts, void* stream);

/* ---- 2D Gaussian-surfel (2DGS) path: the diff_surfel_rasterization-shaped module -------------------------
 * Replaces the `diff_surfel_rasterization._C` extension that lightning/renderer_2dgs.py:7-10 of the reference
 * imports and calls at :224-233 (rasterize_gaussians -> (color, radii, allmap); allmap[7,H,W] = expected depth,
 * alpha, view-space normal x3, median depth, depth distortion, read at :241-257).  That extension's source is NOT in
 * the reference tree (not vendored, no submodule, no pinned version): PARITY UNPINNED.  The arithmetic is the
 * published 2DGS algorithm (ray-splat intersection through the splat->pixel homography, sqrt(2)/2 px low-pass, 1/255 and
 * 1e-4 cut-offs, depth-distortion accumulation); the checker is oracle/surfel_oracle.py.
 * Same conventions as the 3DGS entry points: raw device pointers, caller-owned state, no synchronisation, transposed
 * view / projection matrices.  geom_state / image_state / sort_scratch are the SAME opaque buffers (and sizes) as the
 * 3DGS path -- the binning kernels are shared.  `scales` is [P, scale_stride] with scale_stride >= 2: only the two
 * tangent scales are read (the reference's Gaussian heads carry 3).  `transmat_precomp` ([P,9] = rows Tu, Tv, Tw of the
 * splat->pixel homography) replaces scales + rotations when non-NULL (the module's cov3D_precomp argument). */
GDR_API int gdr_surfel_state_bytes(int P, int64_t* bytes);              /* 80-byte record per surfel, saved for backward */
GDR_API int gdr_surfel_stream_bytes(int64_t capacity, int64_t* bytes);  /* depth-sorted per-tile record stream, saved */
GDR_API int gdr_surfel_aux_bytes(int W, int H, int64_t* bytes);         /* per pixel: median contributor, M1, M2; saved */
GDR_API int gdr_surfel_backward_scratch_bytes(int P, int64_t* bytes);   /* temporary: 20 floats per surfel */

GDR_API int gdr_surfel_forward_project(int P, int sh_degree, int M, int W, int H,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* opacities, const float* scales, int scale_stride, float scale_modifier,
                        const float* rotations, const float* transmat_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos,
                        int32_t* radii /* out [P] */, void* geom_state, void* surfel_state, void* image_state,
                        void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets,
                        int32_t* counts_host /* pinned host memory [4] or NULL */, void* stream);

/* out_color [3,H,W], out_allmap [7,H,W].  flags: 0 or GDR_FLAG_RERUN. */
GDR_API int gdr_surfel_forward_render(int P, int W, int H, const float* bg,
                        const void* geom_state, const void* surfel_state, void* image_state,
                        void* surfel_stream, void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets,
                        int64_t capacity, float* out_color, float* out_allmap, void* surfel_aux, int flags, void* stream);

/* dL_dout_allmap may be NULL (zero).  Any output pointer may be NULL; requested outputs are fully written.
 * dL_dmeans2D is [P, means2D_cols] (3 or 4): columns 0:2 = the densification statistic of 2DGS (gradient w.r.t. the
 * centre's homogeneous pixel offsets scaled to NDC), column 2 = 0 (3 columns) or columns 2:4 = the sums of the
 * absolute per-pixel values (4 columns, the convention of the reference's 3DGS fork, backward.cu:592-594).
 * dL_dscales is [P, scale_stride] (columns >= 2 are zero); dL_dtransmat [P,9] only with transmat_precomp. */
GDR_API int gdr_surfel_backward(int P, int sh_degree, int M, int W, int H, const float* bg,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* scales, int scale_stride, float scale_modifier, const float* rotations,
                        const float* transmat_precomp, const float* viewmatrix, const float* projmatrix,
                        const float* campos, const int32_t* radii,
                        const void* geom_state, const void* surfel_state, const void* image_state,
                        const void* surfel_stream, int64_t capacity, const float* out_allmap, const void* surfel_aux,
                        const float* dL_dout_color, const float* dL_dout_allmap, void* backward_scratch,
                        int means2D_cols, float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                        float* dL_dmeans3D, float* dL_dtransmat, float* dL_dsh, float* dL_dscales,
                        float* dL_drotations, void* stream);

/* out[i] = mean squared distance of points[i] to its 3 nearest neighbours (exact, brute force over shared-memory
 * tiles) -- the quantity `simple_knn._C.distCUDA2` returns, imported by lightning/renderer_2dgs.py:11 and
 * lightning/point_decoder/layers/head.py:7 (simple_knn is not in the reference tree either). */
GDR_API int gdr_knn3_mean_dist2(int P, const float* points /*[P,3]*/, float* out /*[P]*/, void* stream);

/* Introspection for tests: copies of the per-Gaussian state in the reference's field layout
 * (geomState.means2D / depths / conic_opacity / rgb / tiles_touched / clamped, rasterizer_impl.h:33-48).
 * Any output pointer may be NULL. */
GDR_API int gdr_debug_unpack_geom(int P, const void* geom_state, float* means2D /*[P,2]*/, float* depths /*[P]*/,
                          float* conic_opacity /*[P,4]*/, float* rgb /*[P,3]*/, float* cov3D /*[P,6]*/,
                          uint32_t* tiles_touched /*[P]*/, uint8_t* clamped /*[P,3]*/, void* stream);
/* Sorted Gaussian ids per tile instance ([capacity] uint32) and tile ranges ([tiles,2] uint32), both in the
 * REFERENCE's layout (tiles in tile order, as its global sort leaves them, rasterizer_impl.cu:304-315) -- the stream
 * itself keeps the tiles in completion order.  `ranges` is required when `point_list` is requested. */
GDR_API int gdr_debug_unpack_bins(int W, int H, const void* image_state, const void* splat_stream, int64_t capacity,
                          uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib /*[H,W]*/, void* stream);

/* Opt-in per-stage device timing for benchmarks (not used on the product path).  While enabled, every
 * kernel launch is bracketed by cudaEvents on the launch stream.  gdr_profile_read synchronises on the
 * recorded events, adds the elapsed milliseconds and launch counts per stage into the two arrays
 * (GDR_NUM_STAGES entries each) and clears the log.  This event log is the one piece of process-wide state in the
 * library (guarded by a mutex: the forward and autograd's backward thread both append to it). */
#define GDR_STAGE_PROJECT 0   /* projection + binning */
#define GDR_STAGE_TILE_SORT 1
#define GDR_STAGE_BLEND_FWD 2
#define GDR_STAGE_BLEND_BWD 3
#define GDR_STAGE_GAUSS_BWD 4
#define GDR_NUM_STAGES 5
GDR_API int gdr_profile_enable(int on);
GDR_API int gdr_profile_read(double* stage_ms, int64_t* stage_launches);

#ifdef __cplusplus
}
#endif
#endif /* GDR_H_INCLUDED */
