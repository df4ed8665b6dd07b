// extern "C" entry points of libgdr.so (declared in include/gdr.h).
// Raw device pointers, sizes and a cudaStream_t in; an int status out.  No torch
// types, no exceptions, no device or stream synchronisation, no library state.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/gdr.h"
#include "kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s", where, cudaGetErrorString(e));
    return GDR_ERR_CUDA;
}

#define GDR_CUDA(call, where)                        \
    do {                                             \
        cudaError_t _e = (call);                     \
        if (_e != cudaSuccess) return cuda_fail(_e, where); \
    } while (0)

// ---- opt-in stage profiler (benchmarks only) ----
struct StageRecord {
    int stage;
    cudaEvent_t start, stop;
};
bool g_profile = false;
std::vector<StageRecord> g_records;

struct StageTimer {
    cudaStream_t s;
    int stage;
    cudaEvent_t start = nullptr, stop = nullptr;
    StageTimer(int stage_, cudaStream_t s_) : s(s_), stage(stage_) {
        if (g_profile) {
            cudaEventCreate(&start);
            cudaEventCreate(&stop);
            cudaEventRecord(start, s);
        }
    }
    ~StageTimer() {
        if (start) {
            cudaEventRecord(stop, s);
            g_records.push_back({stage, start, stop});
        }
    }
};

inline int tiles_of(int W, int H) { return ((W + gdr::TILE - 1) / gdr::TILE) * ((H + gdr::TILE - 1) / gdr::TILE); }

}  // namespace

extern "C" {

int gdr_abi_version(void) { return GDR_ABI_VERSION; }

const char* gdr_last_error(void) { return g_err; }

int gdr_geom_state_bytes(int P, int64_t* bytes) {
    if (P < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_geom_state_bytes: bad arguments");
    *bytes = (int64_t)gdr::GeomState::bytes((size_t)P);
    return GDR_OK;
}

int gdr_image_state_bytes(int W, int H, int64_t* bytes) {
    if (W <= 0 || H <= 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_image_state_bytes: bad arguments");
    *bytes = (int64_t)gdr::ImageState::bytes(W, H);
    return GDR_OK;
}

int gdr_splat_stream_bytes(int64_t capacity, int64_t* bytes) {
    if (capacity < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_splat_stream_bytes: bad arguments");
    *bytes = (int64_t)sizeof(gdr::Splat) * capacity + 256;
    return GDR_OK;
}

int gdr_sort_scratch_bytes(int64_t capacity, int64_t* bytes) {
    if (capacity < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_sort_scratch_bytes: bad arguments");
    *bytes = 2 * (int64_t)gdr::align_up(sizeof(uint64_t) * (size_t)capacity, 256) + 256;
    return GDR_OK;
}

int gdr_backward_scratch_bytes(int P, int64_t* bytes) {
    if (P < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_backward_scratch_bytes: bad arguments");
    *bytes = (int64_t)sizeof(float) * 12 * P + 256;
    return GDR_OK;
}

int gdr_forward_project(int P, int sh_degree, int M, int W, int H, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales,
                        float scale_modifier, const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                        float tan_fovy, int prefiltered, int32_t* radii, void* geom_state, void* image_state,
                        int32_t* num_rendered_host, int flags, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_project: bad sizes");
    if (!image_state) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_project: image_state is NULL");
    if (P > 0) {
        if (!means3D || !opacities || !radii || !geom_state || !viewmatrix || !projmatrix)
            return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_project: a required pointer is NULL");
        if (!shs && !colors_precomp)
            return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_project: provide SHs or precomputed colors");
        if (!colors_precomp && (!campos || M <= 0 || (sh_degree + 1) * (sh_degree + 1) > M || sh_degree < 0 ||
                                sh_degree > 3))
            return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_project: SH degree / coefficient count mismatch");
        if (!cov3D_precomp && (!scales || !rotations))
            return fail(GDR_ERR_INVALID_ARGUMENT,
                        "gdr_forward_project: provide scales+rotations or a precomputed 3D covariance");
        if (!cov3D_precomp && (((uintptr_t)rotations) & 15u))
            return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_project: rotations must be 16-byte aligned");
    }
    gdr::ImageState img = gdr::ImageState::carve(image_state, W, H);
    const int T = tiles_of(W, H);
    GDR_CUDA(cudaMemsetAsync(img.header, 0, sizeof(uint32_t) * gdr::IMG_HEADER_WORDS, s), "memset(header)");
    GDR_CUDA(cudaMemsetAsync(img.tile_counter, 0, sizeof(uint32_t) * (size_t)T, s), "memset(tile_counter)");
    if (P > 0) {
        gdr::ProjectArgs a;
        a.P = P; a.sh_degree = sh_degree; a.M = M; a.W = W; a.H = H;
        a.gx = (W + gdr::TILE - 1) / gdr::TILE;
        a.gy = (H + gdr::TILE - 1) / gdr::TILE;
        a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.opacities = opacities;
        a.scales = scales; a.scale_modifier = scale_modifier; a.rotations = rotations;
        a.cov3D_precomp = cov3D_precomp; a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.campos = campos;
        a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
        a.focal_y = H / (2.0f * tan_fovy);  // rasterizer_impl.cu:222-223
        a.focal_x = W / (2.0f * tan_fovx);
        a.prefiltered = prefiltered;
        a.cull = (flags & GDR_FLAG_NO_TILE_CULL) ? 0 : 1;
        a.radii = radii;
        a.geom = gdr::GeomState::carve(geom_state, (size_t)P);
        a.img = img;
        {
            StageTimer t(GDR_STAGE_PROJECT, s);
            GDR_CUDA(gdr::launch_project(a, s), "project");
        }
    }
    {
        StageTimer t(GDR_STAGE_TILE_SCAN, s);
        GDR_CUDA(gdr::launch_tile_scan(T, img, s), "tile_scan");
    }
    if (num_rendered_host)
        GDR_CUDA(cudaMemcpyAsync(num_rendered_host, img.header + gdr::HDR_NUM_RENDERED, sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, s),
                 "memcpy(num_rendered)");
    return GDR_OK;
}

int gdr_forward_render(int P, int W, int H, const float* bg, const int32_t* radii, const void* geom_state,
                       void* image_state, void* splat_stream, void* sort_scratch, int64_t capacity, float* out_color,
                       float* out_depth, float* out_alpha, int flags, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_render: bad sizes");
    if (!image_state || !bg || !out_color || !out_depth || !out_alpha)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_render: a required pointer is NULL");
    if (capacity > 0 && (!splat_stream || !sort_scratch))
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_render: stream/scratch is NULL with capacity > 0");
    if (P > 0 && (!geom_state || !radii)) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_forward_render: geom_state is NULL");
    gdr::ImageState img = gdr::ImageState::carve(image_state, W, H);
    gdr::Splat* strm = (gdr::Splat*)splat_stream;
    if (P > 0 && capacity > 0) {
        gdr::GeomState geom = gdr::GeomState::carve(const_cast<void*>(geom_state), (size_t)P);
        uint64_t* keys = (uint64_t*)sort_scratch;
        uint64_t* keys_alt = (uint64_t*)((char*)sort_scratch + gdr::align_up(sizeof(uint64_t) * (size_t)capacity, 256));
        // the emit cursors are zero here: tile_scan zeroes them and tile_sort re-zeroes them after use
        {
            StageTimer t(GDR_STAGE_EMIT, s);
            GDR_CUDA(gdr::launch_emit(P, W, H, radii, geom, img, keys, capacity, (flags & GDR_FLAG_NO_TILE_CULL) ? 0 : 1, s), "emit");
        }
        {
            StageTimer t(GDR_STAGE_TILE_SORT, s);
            GDR_CUDA(gdr::launch_tile_sort(W, H, geom, img, keys, keys_alt, strm, capacity, s), "tile_sort");
        }
    }
    {
        StageTimer t(GDR_STAGE_BLEND_FWD, s);
        GDR_CUDA(gdr::launch_blend_forward(W, H, bg, img, strm, (P > 0) ? capacity : 0, out_color, out_depth, out_alpha,
                                           s),
                 "blend_forward");
    }
    return GDR_OK;
}

int gdr_backward(int P, int sh_degree, int M, int W, int H, const float* bg, const float* means3D, const float* shs,
                 const float* colors_precomp, const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos,
                 float tan_fovx, float tan_fovy, const int32_t* radii, const void* geom_state,
                 const void* image_state, const void* splat_stream, int64_t capacity, const float* out_alpha,
                 const float* dL_dout_color, const float* dL_dout_depth, const float* dL_dout_alpha,
                 void* backward_scratch, int grad_mask, float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                 float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drotations,
                 void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_backward: bad sizes");
    if (P == 0) return GDR_OK;
    if (!means3D || !radii || !geom_state || !image_state || !out_alpha || !dL_dout_color || !backward_scratch ||
        !viewmatrix || !projmatrix || !bg)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_backward: a required pointer is NULL");
    if (capacity > 0 && !splat_stream) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_backward: splat_stream is NULL");
    if ((((uintptr_t)backward_scratch) & 15u) || (dL_dmeans2D && (((uintptr_t)dL_dmeans2D) & 15u)) ||
        (dL_drotations && (((uintptr_t)dL_drotations) & 15u)) || (rotations && (((uintptr_t)rotations) & 15u)))
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_backward: float4 buffers must be 16-byte aligned");
    gdr::ImageState img = gdr::ImageState::carve(const_cast<void*>(image_state), W, H);
    gdr::GeomState geom = gdr::GeomState::carve(const_cast<void*>(geom_state), (size_t)P);
    float* accum = (float*)backward_scratch;
    GDR_CUDA(cudaMemsetAsync(accum, 0, sizeof(float) * 12 * (size_t)P, s), "memset(accum)");
    {
        StageTimer t(GDR_STAGE_BLEND_BWD, s);
        GDR_CUDA(gdr::launch_blend_backward(W, H, bg, img, (const gdr::Splat*)splat_stream, capacity, out_alpha,
                                            dL_dout_color, dL_dout_depth, dL_dout_alpha, accum, grad_mask, s),
                 "blend_backward");
    }
    gdr::GaussBackwardArgs a;
    a.P = P; a.sh_degree = sh_degree; a.M = M; a.W = W; a.H = H;
    a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.scales = scales;
    a.scale_modifier = scale_modifier; a.rotations = rotations; a.cov3D_precomp = cov3D_precomp;
    a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.campos = campos;
    a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
    a.focal_y = H / (2.0f * tan_fovy);
    a.focal_x = W / (2.0f * tan_fovx);
    a.radii = radii; a.geom = geom; a.accum = accum; a.grad_mask = grad_mask;
    a.dL_dmeans2D = dL_dmeans2D; a.dL_dcolors = dL_dcolors; a.dL_dopacity = dL_dopacity;
    a.dL_dmeans3D = dL_dmeans3D; a.dL_dcov3D = dL_dcov3D; a.dL_dsh = dL_dsh; a.dL_dscales = dL_dscales;
    a.dL_drotations = dL_drotations;
    if (shs && !campos) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_backward: campos is NULL");
    {
        StageTimer t(GDR_STAGE_GAUSS_BWD, s);
        GDR_CUDA(gdr::launch_gauss_backward(a, s), "gauss_backward");
    }
    return GDR_OK;
}

int gdr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                     void* stream) {
    (void)projmatrix;  // the reference's test only uses the view-space depth (auxiliary.h:152)
    if (P < 0) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_mark_visible: bad P");
    if (P == 0) return GDR_OK;
    if (!means3D || !viewmatrix || !present) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_mark_visible: NULL pointer");
    GDR_CUDA(gdr::launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream), "mark_visible");
    return GDR_OK;
}

int gdr_debug_unpack_geom(int P, const void* geom_state, float* means2D, float* depths, float* conic_opacity,
                          float* rgb, float* cov3D, uint32_t* tiles_touched, uint8_t* clamped, void* stream) {
    if (P < 0 || (P > 0 && !geom_state)) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_debug_unpack_geom: bad arguments");
    if (P == 0) return GDR_OK;
    gdr::GeomState geom = gdr::GeomState::carve(const_cast<void*>(geom_state), (size_t)P);
    GDR_CUDA(gdr::launch_unpack_geom(P, geom, means2D, depths, conic_opacity, rgb, cov3D, tiles_touched, clamped,
                                     (cudaStream_t)stream),
             "unpack_geom");
    return GDR_OK;
}

int gdr_debug_unpack_bins(int W, int H, const void* image_state, const void* splat_stream, int64_t capacity,
                          uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib, void* stream) {
    if (W <= 0 || H <= 0 || !image_state) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_debug_unpack_bins: bad arguments");
    gdr::ImageState img = gdr::ImageState::carve(const_cast<void*>(image_state), W, H);
    GDR_CUDA(gdr::launch_unpack_bins(W, H, img, (const gdr::Splat*)splat_stream, capacity, point_list, ranges,
                                     n_contrib, (cudaStream_t)stream),
             "unpack_bins");
    return GDR_OK;
}

int gdr_profile_enable(int on) {
    g_profile = on != 0;
    return GDR_OK;
}

int gdr_profile_read(double* stage_ms, int64_t* stage_launches) {
    if (!stage_ms || !stage_launches) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_profile_read: NULL output");
    for (auto& r : g_records) {
        GDR_CUDA(cudaEventSynchronize(r.stop), "profile sync");
        float ms = 0.f;
        GDR_CUDA(cudaEventElapsedTime(&ms, r.start, r.stop), "profile elapsed");
        if (r.stage >= 0 && r.stage < GDR_NUM_STAGES) {
            stage_ms[r.stage] += ms;
            stage_launches[r.stage] += 1;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_records.clear();
    return GDR_OK;
}

}  // extern "C"
