// Host-side launchers of the sm_100a kernels (one translation unit per stage).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "state.cuh"

#include <utility>

namespace gdr {

// Launch `kernel` as a programmatic dependent of the previous kernel in stream `s`: its CTAs may become resident
// while that kernel's last CTAs are still running and wait in pdl_wait() (common.cuh), so launch latency and CTA
// ramp-up leave the critical path of the short kernels.  GDR_PDL=0 falls back to plain stream order.
bool pdl_enabled();
template <class... KArgs, class... Args>
cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                             Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Number of SMs of the current device (cached); grids of the per-Gaussian kernels are sized as a
// multiple of it and loop over virtual blocks, so no kernel ends with a nearly empty last wave.
int sm_count();

struct ProjectArgs {
    int P, sh_degree, M, W, H, gx, gy;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    int prefiltered;
    int raw_params;  // opacities are logits, scales are log-scales, rotations are un-normalised (GDR_FLAG_RAW_PARAMS)
    int cull;  // exact tile-level culling of (Gaussian, tile) pairs that cannot reach alpha = 1/255
    Views vw;          // cameras + per-view strides; every per-view pointer below is view 0's
    int32_t* radii;    // [V][P]
    GeomState geom;
    ImageState img;
    uint64_t* keys;      // [V][T][tile_cap] key segments (SortScratch, state.cuh)
    size_t keys_stride;  // keys between consecutive views
    uint32_t tile_cap;   // slots per tile segment (uniform layout)
    const uint32_t* tile_base;  // [V][T + 1] exact layout (see SortScratch in state.cuh), or nullptr
    int32_t* counts_host;  // device-accessible pinned host memory [V][4] or nullptr (see report_counts)
};

// project.cu: per-Gaussian projection + warp-aggregated tile counting
cudaError_t launch_project(const ProjectArgs& a, cudaStream_t s);
// zeroes the header + slot counters of every view; launch_project's kernel is its programmatic dependent
cudaError_t launch_zero_state(ImageState img, int W, int H, const Views& vw, cudaStream_t s);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                                cudaStream_t s);

// binning.cu: per-tile depth sort + record gather (the binning itself runs inside the projection kernels)
// Batched layout (V = vw.V views, blockIdx.y = view): every view has its own key segments (T x tile_cap keys) and
// `capacity` stream records, back to back; images are [V][C][H][W]; radii [V][P]; accum [V][P][12].
cudaError_t launch_tile_sort(int W, int H, GeomState geom, ImageState img, uint64_t* keys, int64_t tile_cap,
                             const uint32_t* tile_base, Splat* stream, int64_t capacity, const Views& vw,
                             cudaStream_t s);
// exclusive scan of a view's per-tile instance counts -> tile_offsets[T + 1] (the exact key layout)
cudaError_t launch_tile_offsets(int W, int H, ImageState img, uint32_t* tile_offsets, const Views& vw, cudaStream_t s);
// bytes of one view's key segments in either layout
size_t key_bytes_per_view(int W, int H, int64_t tile_cap, bool exact);

// The same per-tile sort with 80-byte Surfel records (surfel.cuh) gathered into the stream; single view.
cudaError_t launch_tile_sort_surfel(int W, int H, const void* surfel_records, ImageState img, uint64_t* keys,
                                    int64_t tile_cap, const uint32_t* tile_base, void* surfel_stream, int64_t capacity,
                                    cudaStream_t s);

// blend_fwd.cu
// hwc_clamp: write out_color as the clamped [H][W][3] image of Renderer.render_img (GDR_FLAG_FUSED_EPILOGUE)
cudaError_t launch_blend_forward(int W, int H, ImageState img, const Splat* stream, int64_t capacity,
                                 float* out_color, float* out_depth, float* out_alpha, const Views& vw, int hwc_clamp,
                                 cudaStream_t s);

// blend_bwd.cu: accum is [V][P][12] floats: (mean2D x,y,|x|,|y|), (conic a,b,c, opacity), (r,g,b, depth)
cudaError_t launch_blend_backward(int P, int W, int H, ImageState img, const Splat* stream, int64_t capacity,
                                  const float* out_alpha, const float* dL_dcolor, const float* dL_ddepth,
                                  const float* dL_dalpha, float* accum, int grad_mask, const Views& vw,
                                  cudaStream_t s);

struct GaussBackwardArgs {
    int P, sh_degree, M, W, H;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    Views vw;
    int view;          // which view of the batch this launch handles (its state is found through vw)
    int accumulate;    // 0: store the outputs; 1: add to them (views 1.. of a batch: gradients sum over views)
    const int32_t* radii;  // [V][P]
    GeomState geom;
    float* accum;          // [V][P][12]; read, and re-zeroed when `rezero` is set
    int grad_mask;
    int rezero;        // store zeros back over the accumulator rows after reading them (GDR_GRAD_SCRATCH_CLEAN)
    float* dL_dmeans2D;
    float* dL_dcolors;
    float* dL_dopacity;
    float* dL_dmeans3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscales;
    float* dL_drotations;
};
// gauss_bwd.cu: fused per-Gaussian chain rule (conic -> cov2D -> cov3D -> scale/rot, mean2D/depth -> mean3D, SH)
cudaError_t launch_gauss_backward(const GaussBackwardArgs& a, cudaStream_t s);

// densify.cu: the densify select either side of the raster path (lightning/network.py:865-893)
cudaError_t launch_mse_grad(int V, int W, int H, const float* color, const float* target, float* dL_dcolor,
                            float* loss, cudaStream_t s);
cudaError_t launch_densify_score(int V, int P, const float* accum, const uint8_t* candidate, float* grad, float* score,
                                 cudaStream_t s);
cudaError_t launch_topk_select(int P, const float* score, int k, uint8_t* selected, int32_t* selected_idx,
                               int32_t* rest_idx, int32_t* counts, cudaStream_t s);

// surfel.cu: the 2D Gaussian-surfel path behind the diff_surfel_rasterization-shaped module
// (lightning/renderer_2dgs.py:224-233).  Single view; binning is shared with the 3DGS path.
cudaError_t launch_surfel_project(int P, int sh_degree, int M, int W, int H, const float* means3D, const float* shs,
                                  const float* colors_precomp, const float* opacities, const float* scales,
                                  int scale_stride, float scale_modifier, const float* rotations,
                                  const float* transmat_precomp, const float* view, const float* proj,
                                  const float* campos, int32_t* radii, GeomState geom, void* surfel_state,
                                  ImageState img, uint64_t* keys, int64_t tile_cap, const uint32_t* tile_base,
                                  int32_t* counts_host, cudaStream_t s);
cudaError_t launch_surfel_blend_forward(int W, int H, ImageState img, const void* stream, int64_t capacity,
                                        const float* bg, float* out_color, float* out_allmap, float* aux,
                                        cudaStream_t s);
cudaError_t launch_surfel_blend_backward(int W, int H, ImageState img, const void* stream, int64_t capacity,
                                         const float* bg, const float* out_allmap, const float* aux,
                                         const float* dL_dcolor, const float* dL_dallmap, float* accum,
                                         cudaStream_t s);
cudaError_t launch_surfel_gauss_backward(int P, int sh_degree, int M, int W, int H, const float* means3D,
                                         const float* shs, const float* colors_precomp, const float* scales,
                                         int scale_stride, float scale_modifier, const float* rotations,
                                         const float* transmat_precomp, const float* view, const float* proj,
                                         const float* campos, const int32_t* radii, const void* surfel_state,
                                         const uint8_t* clamped, const float* accum, int means2D_cols,
                                         float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D,
                                         float* dL_dtransmat, float* dL_dsh, float* dL_dscales, float* dL_drotations,
                                         cudaStream_t s);

// knn.cu: mean squared distance to the 3 nearest neighbours (simple_knn's distCUDA2)
cudaError_t launch_knn3(int P, const float* points, float* out, cudaStream_t s);

// debug.cu
cudaError_t launch_unpack_geom(int P, GeomState geom, float* means2D, float* depths, float* conic_opacity, float* rgb,
                               float* cov3D, uint32_t* tiles_touched, uint8_t* clamped, cudaStream_t s);
cudaError_t launch_unpack_bins(int W, int H, ImageState img, const Splat* stream, int64_t capacity,
                               uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib, cudaStream_t s);

}  // namespace gdr
