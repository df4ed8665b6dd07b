// extern "C" entry points of libgdr.so (declared in include/gdr.h).
// Raw device pointers, sizes and a cudaStream_t in; an int status out.  No torch
// types, no exceptions, no device or stream synchronisation, no library state.
#include <stdio.h>
#include <string.h>

#include <mutex>
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
// The event log is the library's only process-wide state (include/gdr.h says so): the forward and autograd's backward
// thread both append to it, so it is guarded.
bool g_profile = false;
std::mutex g_records_mutex;
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
            std::lock_guard<std::mutex> lock(g_records_mutex);
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

int gdr_sort_scratch_bytes(int W, int H, int64_t tile_capacity, int64_t* bytes) {
    if (W <= 0 || H <= 0 || tile_capacity <= 0 || (tile_capacity & 31) || !bytes)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_sort_scratch_bytes: bad arguments (tile_capacity must be a positive multiple of 32)");
    *bytes = (int64_t)gdr::sort_scratch_bytes(W, H, tile_capacity);
    return GDR_OK;
}

int gdr_sort_scratch_exact_bytes(int64_t num_keys, int64_t* bytes) {
    if (num_keys <= 0 || (num_keys & 31) || !bytes)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_sort_scratch_exact_bytes: num_keys must be a positive multiple of 32");
    *bytes = (int64_t)gdr::key_bytes_per_view(1, 1, num_keys, true);
    return GDR_OK;
}

int gdr_backward_scratch_bytes(int P, int64_t* bytes) {
    if (P < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_backward_scratch_bytes: bad arguments");
    *bytes = (int64_t)sizeof(float) * 12 * P + 256;
    return GDR_OK;
}

}  // extern "C"

// ---- shared implementation of the single-view and the batched entry points -------------------------
namespace {

struct GaussianInputs {
    const float *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp;
    float scale_modifier;
};

gdr::Views single_view(int P, int W, int H, const float* viewmatrix, const float* projmatrix, const float* campos,
                       const float* bg, float tan_fovx, float tan_fovy) {
    gdr::Views vw;
    vw.V = 1;
    vw.geom_stride = vw.img_stride = vw.cam_stride = 0;
    vw.view = viewmatrix; vw.proj = projmatrix; vw.campos = campos; vw.bg = bg;
    vw.tanfov = nullptr; vw.tan_fovx = tan_fovx; vw.tan_fovy = tan_fovy;
    (void)P; (void)W; (void)H;
    return vw;
}

gdr::Views batched_views(int V, int P, int W, int H, const gdr_camera* cams) {
    gdr::Views vw;
    const float* c = reinterpret_cast<const float*>(cams);
    vw.V = V;
    vw.geom_stride = gdr::GeomState::bytes((size_t)P);
    vw.img_stride = gdr::ImageState::bytes(W, H);
    vw.cam_stride = GDR_CAMERA_FLOATS;
    vw.view = c; vw.proj = c + 16; vw.campos = c + 32; vw.tanfov = c + 35; vw.bg = c + 37;
    vw.tan_fovx = vw.tan_fovy = 0.f;
    return vw;
}

int project_impl(const char* who, const gdr::Views& vw, int P, int sh_degree, int M, int W, int H, const GaussianInputs& g,
                 int prefiltered, int32_t* radii, void* geom_state, void* image_state, void* sort_scratch,
                 int64_t tile_capacity, const uint32_t* tile_offsets, int32_t* counts_host, int flags, cudaStream_t s) {
    if (P < 0 || W <= 0 || H <= 0 || vw.V <= 0) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    if (P >= (1 << gdr::STREAM_REGION_SHIFT))
        return fail(GDR_ERR_UNSUPPORTED, "%s: at most 2^28 - 1 Gaussians per call", who);
    if (!image_state) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: image_state is NULL", who);
    if (tiles_of(W, H) >= (1 << 24)) return fail(GDR_ERR_UNSUPPORTED, "%s: at most 2^24 - 1 tiles per image", who);
    if (P > 0) {
        if (!g.means3D || !g.opacities || !radii || !geom_state || !vw.view || !vw.proj)
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: a required pointer is NULL", who);
        if (!sort_scratch || tile_capacity <= 0 || (tile_capacity & 31) || tile_capacity > 0x7fffffff)
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: sort_scratch is NULL or tile_capacity is not a positive multiple of 32", who);
        if (!g.shs && !g.colors_precomp)
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: provide SHs or precomputed colors", who);
        if (!g.colors_precomp && (!vw.campos || M <= 0 || (sh_degree + 1) * (sh_degree + 1) > M || sh_degree < 0 ||
                                  sh_degree > 3))
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: SH degree / coefficient count mismatch", who);
        if (!g.cov3D_precomp && (!g.scales || !g.rotations))
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: provide scales+rotations or a precomputed 3D covariance", who);
        if (!g.cov3D_precomp && (((uintptr_t)g.rotations) & 15u))
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: rotations must be 16-byte aligned", who);
    }
    gdr::ImageState img = gdr::ImageState::carve(image_state, W, H);
    const int T = tiles_of(W, H);
    // header and per-tile slot counters are adjacent: one zeroing launch for all views, which the projection kernel
    // follows as a programmatic dependent (project.cu)
    GDR_CUDA(gdr::launch_zero_state(img, W, H, vw, s), "zero(header, tile_count)");
    if (P > 0) {
        gdr::ProjectArgs a;
        a.P = P; a.sh_degree = sh_degree; a.M = M; a.W = W; a.H = H;
        a.gx = (W + gdr::TILE - 1) / gdr::TILE;
        a.gy = (H + gdr::TILE - 1) / gdr::TILE;
        a.means3D = g.means3D; a.shs = g.shs; a.colors_precomp = g.colors_precomp; a.opacities = g.opacities;
        a.scales = g.scales; a.scale_modifier = g.scale_modifier; a.rotations = g.rotations;
        a.cov3D_precomp = g.cov3D_precomp;
        a.prefiltered = prefiltered;
        a.raw_params = (flags & GDR_FLAG_RAW_PARAMS) ? 1 : 0;
        a.cull = (flags & GDR_FLAG_NO_TILE_CULL) ? 0 : 1;
        a.vw = vw;
        a.radii = radii;
        a.geom = gdr::GeomState::carve(geom_state, (size_t)P);
        a.img = img;
        a.keys = (uint64_t*)sort_scratch;
        a.keys_stride = gdr::key_bytes_per_view(W, H, tile_capacity, tile_offsets != nullptr) / sizeof(uint64_t);
        a.tile_cap = (uint32_t)tile_capacity;
        a.tile_base = tile_offsets;
        a.counts_host = counts_host;  // written by the kernel's last CTA: no copy node between it and tile_sort
        {
            StageTimer t(GDR_STAGE_PROJECT, s);
            GDR_CUDA(gdr::launch_project(a, s), "project");
        }
    } else if (counts_host) {
        for (int v = 0; v < vw.V; v++) {
            counts_host[4 * v] = counts_host[4 * v + 1] = counts_host[4 * v + 2] = 0;
            counts_host[4 * v + 3] = 1;
        }
    }
    return GDR_OK;
}

int render_impl(const char* who, const gdr::Views& vw, int P, int W, int H, const void* geom_state, void* image_state,
                void* splat_stream, void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets,
                int64_t capacity, float* out_color, float* out_depth, float* out_alpha, int flags, cudaStream_t s) {
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0 || vw.V <= 0) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    if (!image_state || !vw.bg || !out_color || !out_depth || !out_alpha)
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: a required pointer is NULL", who);
    if (capacity > 0 && !splat_stream) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: stream is NULL with capacity > 0", who);
    if (P > 0 && (!geom_state || !sort_scratch || tile_capacity <= 0 || (tile_capacity & 31)))
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: geom_state / sort_scratch is NULL or bad tile_capacity", who);
    if (capacity >= ((int64_t)1 << 32)) return fail(GDR_ERR_UNSUPPORTED, "%s: capacity must be below 2^32", who);
    gdr::ImageState img = gdr::ImageState::carve(image_state, W, H);
    gdr::Splat* strm = (gdr::Splat*)splat_stream;
    if (flags & GDR_FLAG_RERUN) {
        // a repeated render of the same projection (larger capacity): reset what tile_sort accumulates in the header
        // -- stream cursor, order-list fills, overflow flag -- but keep what the projection kernel left there
        const size_t off = sizeof(uint32_t) * gdr::HDR_CURSOR, len = sizeof(uint32_t) * (gdr::IMG_HEADER_WORDS - gdr::HDR_CURSOR);
        if (vw.V == 1)
            GDR_CUDA(cudaMemsetAsync((char*)img.header + off, 0, len, s), "memset(header tail)");
        else
            GDR_CUDA(cudaMemset2DAsync((char*)img.header + off, vw.img_stride, 0, len, (size_t)vw.V, s), "memset(header tail)");
    }
    {
        gdr::GeomState geom = gdr::GeomState::carve(const_cast<void*>(geom_state), (size_t)(P > 0 ? P : 0));
        StageTimer t(GDR_STAGE_TILE_SORT, s);
        // runs for P == 0 too: it files every (empty) tile in the order lists the blend kernel walks
        GDR_CUDA(gdr::launch_tile_sort(W, H, geom, img, (uint64_t*)sort_scratch, P > 0 ? tile_capacity : 32,
                                       P > 0 ? tile_offsets : nullptr, strm, capacity, vw, s),
                 "tile_sort");
    }
    {
        StageTimer t(GDR_STAGE_BLEND_FWD, s);
        GDR_CUDA(gdr::launch_blend_forward(W, H, img, strm, capacity, out_color, out_depth, out_alpha, vw,
                                           (flags & GDR_FLAG_FUSED_EPILOGUE) ? 1 : 0, s),
                 "blend_forward");
    }
    return GDR_OK;
}

struct GradOutputs {
    float *means2D, *colors, *opacity, *means3D, *cov3D, *sh, *scales, *rotations;
};

int backward_impl(const char* who, const gdr::Views& vw, int P, int sh_degree, int M, int W, int H,
                  const GaussianInputs& g, const int32_t* radii, const void* geom_state, const void* image_state,
                  const void* splat_stream, int64_t capacity, const float* out_alpha, const float* dL_dout_color,
                  const float* dL_dout_depth, const float* dL_dout_alpha, void* backward_scratch, int grad_mask,
                  const GradOutputs& o, cudaStream_t s) {
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0 || vw.V <= 0) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    if (P == 0) return GDR_OK;
    if (!g.means3D || !radii || !geom_state || !image_state || !out_alpha || !dL_dout_color || !backward_scratch ||
        !vw.view || !vw.proj || !vw.bg)
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: a required pointer is NULL", who);
    if (capacity > 0 && !splat_stream) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: splat_stream is NULL", who);
    if ((((uintptr_t)backward_scratch) & 15u) || (o.means2D && (((uintptr_t)o.means2D) & 15u)) ||
        (o.rotations && (((uintptr_t)o.rotations) & 15u)) || (g.rotations && (((uintptr_t)g.rotations) & 15u)))
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: float4 buffers must be 16-byte aligned", who);
    if (g.shs && !vw.campos) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: campos is NULL", who);
    gdr::ImageState img = gdr::ImageState::carve(const_cast<void*>(image_state), W, H);
    gdr::GeomState geom = gdr::GeomState::carve(const_cast<void*>(geom_state), (size_t)P);
    float* accum = (float*)backward_scratch;
    const bool clean = (grad_mask & GDR_GRAD_SCRATCH_CLEAN) != 0;  // the caller's buffer is zeroed and stays so
    if (!clean) GDR_CUDA(cudaMemsetAsync(accum, 0, sizeof(float) * 12 * (size_t)P * (size_t)vw.V, s), "memset(accum)");
    {
        StageTimer t(GDR_STAGE_BLEND_BWD, s);
        GDR_CUDA(gdr::launch_blend_backward(P, W, H, img, (const gdr::Splat*)splat_stream, capacity, out_alpha,
                                            dL_dout_color, dL_dout_depth, dL_dout_alpha, accum, grad_mask, vw, s),
                 "blend_backward");
    }
    gdr::GaussBackwardArgs a;
    a.P = P; a.sh_degree = sh_degree; a.M = M; a.W = W; a.H = H;
    a.means3D = g.means3D; a.shs = g.shs; a.colors_precomp = g.colors_precomp; a.scales = g.scales;
    a.scale_modifier = g.scale_modifier; a.rotations = g.rotations; a.cov3D_precomp = g.cov3D_precomp;
    a.vw = vw;
    a.radii = radii; a.geom = geom; a.accum = accum; a.grad_mask = grad_mask; a.rezero = clean ? 1 : 0;
    a.dL_dmeans2D = o.means2D; a.dL_dcolors = o.colors; a.dL_dopacity = o.opacity;
    a.dL_dmeans3D = o.means3D; a.dL_dcov3D = o.cov3D; a.dL_dsh = o.sh; a.dL_dscales = o.scales;
    a.dL_drotations = o.rotations;
    for (int v = 0; v < vw.V; v++) {  // gradients of shared Gaussians sum over the views, in view order
        a.view = v;
        a.accumulate = v > 0;
        StageTimer t(GDR_STAGE_GAUSS_BWD, s);
        GDR_CUDA(gdr::launch_gauss_backward(a, s), "gauss_backward");
    }
    return GDR_OK;
}

}  // namespace

extern "C" {

int gdr_forward_project(int P, int sh_degree, int M, int W, int H, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales,
                        float scale_modifier, const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                        float tan_fovy, int prefiltered, int32_t* radii, void* geom_state, void* image_state,
                        void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets, int32_t* counts_host,
                        int flags, void* stream) {
    const GaussianInputs g = {means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, scale_modifier};
    return project_impl("gdr_forward_project", single_view(P, W, H, viewmatrix, projmatrix, campos, nullptr, tan_fovx, tan_fovy),
                        P, sh_degree, M, W, H, g, prefiltered, radii, geom_state, image_state, sort_scratch, tile_capacity,
                        tile_offsets, counts_host, flags, (cudaStream_t)stream);
}

int gdr_forward_render(int P, int W, int H, const float* bg, const void* geom_state, void* image_state,
                       void* splat_stream, void* sort_scratch, int64_t tile_capacity, const uint32_t* tile_offsets,
                       int64_t capacity, float* out_color, float* out_depth, float* out_alpha, int flags, void* stream) {
    return render_impl("gdr_forward_render", single_view(P, W, H, nullptr, nullptr, nullptr, bg, 0.f, 0.f), P, W, H,
                       geom_state, image_state, splat_stream, sort_scratch, tile_capacity, tile_offsets, capacity,
                       out_color, out_depth, out_alpha, flags, (cudaStream_t)stream);
}

int gdr_tile_offsets(int V, int W, int H, const void* image_states, uint32_t* tile_offsets, void* stream) {
    if (V <= 0 || W <= 0 || H <= 0 || !image_states || !tile_offsets)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_tile_offsets: bad arguments");
    gdr::Views vw;
    memset(&vw, 0, sizeof(vw));
    vw.V = V;
    vw.img_stride = V > 1 ? gdr::ImageState::bytes(W, H) : 0;
    GDR_CUDA(gdr::launch_tile_offsets(W, H, gdr::ImageState::carve(const_cast<void*>(image_states), W, H), tile_offsets, vw,
                                      (cudaStream_t)stream),
             "tile_offsets");
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
    const GaussianInputs g = {means3D, shs, colors_precomp, nullptr, scales, rotations, cov3D_precomp, scale_modifier};
    const GradOutputs o = {dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations};
    return backward_impl("gdr_backward", single_view(P, W, H, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy), P,
                         sh_degree, M, W, H, g, radii, geom_state, image_state, splat_stream, capacity, out_alpha,
                         dL_dout_color, dL_dout_depth, dL_dout_alpha, backward_scratch, grad_mask, o,
                         (cudaStream_t)stream);
}

// ---- batched multi-view entry points: V cameras, one set of Gaussians, one launch per stage ----
int gdr_views_forward_project(int V, int P, int sh_degree, int M, int W, int H, const float* means3D, const float* shs,
                              const float* colors_precomp, const float* opacities, const float* scales,
                              float scale_modifier, const float* rotations, const float* cov3D_precomp,
                              const gdr_camera* cameras, int prefiltered, int32_t* radii, void* geom_states,
                              void* image_states, void* sort_scratch, int64_t tile_capacity,
                              const uint32_t* tile_offsets, int32_t* counts_host, int flags, void* stream) {
    if (V <= 0 || !cameras) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_views_forward_project: bad V or cameras is NULL");
    const GaussianInputs g = {means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, scale_modifier};
    return project_impl("gdr_views_forward_project", batched_views(V, P, W, H, cameras), P, sh_degree, M, W, H, g,
                        prefiltered, radii, geom_states, image_states, sort_scratch, tile_capacity, tile_offsets,
                        counts_host, flags, (cudaStream_t)stream);
}

int gdr_views_forward_render(int V, int P, int W, int H, const gdr_camera* cameras, const void* geom_states,
                             void* image_states, void* splat_streams, void* sort_scratch, int64_t tile_capacity,
                             const uint32_t* tile_offsets, int64_t capacity_per_view, float* out_color, float* out_depth,
                             float* out_alpha, int flags, void* stream) {
    if (V <= 0 || !cameras) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_views_forward_render: bad V or cameras is NULL");
    return render_impl("gdr_views_forward_render", batched_views(V, P, W, H, cameras), P, W, H, geom_states, image_states,
                       splat_streams, sort_scratch, tile_capacity, tile_offsets, capacity_per_view, out_color, out_depth,
                       out_alpha, flags, (cudaStream_t)stream);
}

int gdr_views_backward(int V, int P, int sh_degree, int M, int W, int H, const float* means3D, const float* shs,
                       const float* colors_precomp, const float* scales, float scale_modifier, const float* rotations,
                       const float* cov3D_precomp, const gdr_camera* cameras, const int32_t* radii,
                       const void* geom_states, const void* image_states, const void* splat_streams,
                       int64_t capacity_per_view, const float* out_alpha, const float* dL_dout_color,
                       const float* dL_dout_depth, const float* dL_dout_alpha, void* backward_scratch, int grad_mask,
                       float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D, float* dL_dcov3D,
                       float* dL_dsh, float* dL_dscales, float* dL_drotations, void* stream) {
    if (V <= 0 || !cameras) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_views_backward: bad V or cameras is NULL");
    const GaussianInputs g = {means3D, shs, colors_precomp, nullptr, scales, rotations, cov3D_precomp, scale_modifier};
    const GradOutputs o = {dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations};
    return backward_impl("gdr_views_backward", batched_views(V, P, W, H, cameras), P, sh_degree, M, W, H, g, radii,
                         geom_states, image_states, splat_streams, capacity_per_view, out_alpha, dL_dout_color,
                         dL_dout_depth, dL_dout_alpha, backward_scratch, grad_mask, o, (cudaStream_t)stream);
}

// ---- the densify select, fused on the device ----
int gdr_mse_grad(int V, int W, int H, const float* color, const float* target, float* dL_dcolor, float* loss,
                 void* stream) {
    if (V <= 0 || W <= 0 || H <= 0 || !color || !target || !dL_dcolor)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_mse_grad: bad arguments");
    GDR_CUDA(gdr::launch_mse_grad(V, W, H, color, target, dL_dcolor, loss, (cudaStream_t)stream), "mse_grad");
    return GDR_OK;
}

int gdr_views_densify_scores(int V, int P, int W, int H, const gdr_camera* cameras, const void* image_states,
                             const void* splat_streams, int64_t capacity_per_view, const float* out_alpha,
                             const float* dL_dout_color, void* backward_scratch, const uint8_t* candidate_mask,
                             float* grad_means2D, float* scores, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (V <= 0 || P < 0 || W <= 0 || H <= 0 || capacity_per_view < 0 || !cameras)
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_views_densify_scores: bad sizes or cameras is NULL");
    if (P == 0) return GDR_OK;
    if (!image_states || !out_alpha || !dL_dout_color || !backward_scratch || !scores ||
        (capacity_per_view > 0 && !splat_streams))
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_views_densify_scores: a required pointer is NULL");
    if ((((uintptr_t)backward_scratch) & 15u) || (grad_means2D && (((uintptr_t)grad_means2D) & 15u)))
        return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_views_densify_scores: float4 buffers must be 16-byte aligned");
    const gdr::Views vw = batched_views(V, P, W, H, cameras);
    gdr::ImageState img = gdr::ImageState::carve(const_cast<void*>(image_states), W, H);
    float* accum = (float*)backward_scratch;
    GDR_CUDA(cudaMemsetAsync(accum, 0, sizeof(float) * 12 * (size_t)P * (size_t)V, s), "memset(accum)");
    {
        StageTimer t(GDR_STAGE_BLEND_BWD, s);
        GDR_CUDA(gdr::launch_blend_backward(P, W, H, img, (const gdr::Splat*)splat_streams, capacity_per_view, out_alpha,
                                            dL_dout_color, nullptr, nullptr, accum, GDR_GRAD_MEANS2D, vw, s),
                 "blend_backward");
    }
    GDR_CUDA(gdr::launch_densify_score(V, P, accum, candidate_mask, grad_means2D, scores, s), "densify_score");
    return GDR_OK;
}

int gdr_topk_select(int P, const float* scores, int k, uint8_t* selected, int32_t* selected_idx, int32_t* rest_idx,
                    int32_t* counts, void* stream) {
    if (P < 0 || (P > 0 && !scores)) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_topk_select: bad arguments");
    GDR_CUDA(gdr::launch_topk_select(P, scores, k, selected, selected_idx, rest_idx, counts, (cudaStream_t)stream),
             "topk_select");
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

// ---- 2D Gaussian-surfel path (surfel.cu) -------------------------------------------------------------
int gdr_surfel_state_bytes(int P, int64_t* bytes) {
    if (P < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_surfel_state_bytes: bad arguments");
    *bytes = (int64_t)80 * P + 256;
    return GDR_OK;
}

int gdr_surfel_stream_bytes(int64_t capacity, int64_t* bytes) {
    if (capacity < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_surfel_stream_bytes: bad arguments");
    *bytes = (int64_t)80 * capacity + 256;
    return GDR_OK;
}

int gdr_surfel_aux_bytes(int W, int H, int64_t* bytes) {
    if (W <= 0 || H <= 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_surfel_aux_bytes: bad arguments");
    *bytes = (int64_t)12 * W * H + 256;
    return GDR_OK;
}

int gdr_surfel_backward_scratch_bytes(int P, int64_t* bytes) {
    if (P < 0 || !bytes) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_surfel_backward_scratch_bytes: bad arguments");
    *bytes = (int64_t)sizeof(float) * 20 * P + 256;
    return GDR_OK;
}

int gdr_surfel_forward_project(int P, int sh_degree, int M, int W, int H, const float* means3D, const float* shs,
                               const float* colors_precomp, const float* opacities, const float* scales,
                               int scale_stride, float scale_modifier, const float* rotations,
                               const float* transmat_precomp, const float* viewmatrix, const float* projmatrix,
                               const float* campos, int32_t* radii, void* geom_state, void* surfel_state,
                               void* image_state, void* sort_scratch, int64_t tile_capacity,
                               const uint32_t* tile_offsets, int32_t* counts_host, void* stream) {
    const char* who = "gdr_surfel_forward_project";
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    if (!image_state) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: image_state is NULL", who);
    if (P > 0) {
        if (!means3D || !opacities || !radii || !geom_state || !surfel_state || !viewmatrix || !projmatrix)
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: a required pointer is NULL", who);
        if (!sort_scratch || tile_capacity <= 0 || (tile_capacity & 31) || tile_capacity > 0x7fffffff)
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: sort_scratch is NULL or tile_capacity is not a positive multiple of 32", who);
        if (!shs && !colors_precomp) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: provide SHs or precomputed colors", who);
        if (!colors_precomp && (!campos || M <= 0 || sh_degree < 0 || sh_degree > 3 || (sh_degree + 1) * (sh_degree + 1) > M))
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: SH degree / coefficient count mismatch", who);
        if (!transmat_precomp && (!scales || !rotations || scale_stride < 2))
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: provide scales (>= 2 columns) + rotations or a precomputed homography", who);
        if ((rotations && (((uintptr_t)rotations) & 15u)) || (((uintptr_t)surfel_state) & 15u))
            return fail(GDR_ERR_INVALID_ARGUMENT, "%s: rotations / surfel_state must be 16-byte aligned", who);
    }
    gdr::ImageState img = gdr::ImageState::carve(image_state, W, H);
    const int T = tiles_of(W, H);
    GDR_CUDA(cudaMemsetAsync(img.header, 0, (size_t)((char*)(img.tile_count + (size_t)T * gdr::COUNT_STRIDE) - (char*)img.header), s),
             "memset(header, tile_count)");
    if (P > 0) {
        StageTimer t(GDR_STAGE_PROJECT, s);
        GDR_CUDA(gdr::launch_surfel_project(P, sh_degree, M, W, H, means3D, shs, colors_precomp, opacities, scales,
                                            scale_stride, scale_modifier, rotations, transmat_precomp, viewmatrix,
                                            projmatrix, campos, radii, gdr::GeomState::carve(geom_state, (size_t)P),
                                            surfel_state, img, (uint64_t*)sort_scratch, tile_capacity, tile_offsets,
                                            counts_host, s),
                 "surfel_project");
    } else if (counts_host) {
        counts_host[0] = counts_host[1] = counts_host[2] = 0;
        counts_host[3] = 1;
    }
    return GDR_OK;
}

int gdr_surfel_forward_render(int P, int W, int H, const float* bg, const void* geom_state, const void* surfel_state,
                              void* image_state, void* surfel_stream, void* sort_scratch, int64_t tile_capacity,
                              const uint32_t* tile_offsets, int64_t capacity, float* out_color, float* out_allmap,
                              void* surfel_aux, int flags, void* stream) {
    const char* who = "gdr_surfel_forward_render";
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    if (!image_state || !bg || !out_color || !out_allmap || !surfel_aux)
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: a required pointer is NULL", who);
    if (capacity > 0 && !surfel_stream) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: stream is NULL with capacity > 0", who);
    if (P > 0 && (!geom_state || !surfel_state || !sort_scratch || tile_capacity <= 0 || (tile_capacity & 31)))
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: state / sort_scratch is NULL or bad tile_capacity", who);
    if (capacity >= ((int64_t)1 << 32)) return fail(GDR_ERR_UNSUPPORTED, "%s: capacity must be below 2^32", who);
    gdr::ImageState img = gdr::ImageState::carve(image_state, W, H);
    if (flags & GDR_FLAG_RERUN)
        GDR_CUDA(cudaMemsetAsync(img.header + gdr::HDR_CURSOR, 0,
                                 sizeof(uint32_t) * (gdr::IMG_HEADER_WORDS - gdr::HDR_CURSOR), s),
                 "memset(header tail)");
    {
        StageTimer t(GDR_STAGE_TILE_SORT, s);
        GDR_CUDA(gdr::launch_tile_sort_surfel(W, H, surfel_state, img, (uint64_t*)sort_scratch,
                                              P > 0 ? tile_capacity : 32, P > 0 ? tile_offsets : nullptr, surfel_stream,
                                              capacity, s),
                 "tile_sort");
    }
    {
        StageTimer t(GDR_STAGE_BLEND_FWD, s);
        GDR_CUDA(gdr::launch_surfel_blend_forward(W, H, img, surfel_stream, capacity, bg, out_color, out_allmap,
                                                  (float*)surfel_aux, s),
                 "surfel_blend_forward");
    }
    return GDR_OK;
}

int gdr_surfel_backward(int P, int sh_degree, int M, int W, int H, const float* bg, const float* means3D,
                        const float* shs, const float* colors_precomp, const float* scales, int scale_stride,
                        float scale_modifier, const float* rotations, const float* transmat_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos, const int32_t* radii,
                        const void* geom_state, const void* surfel_state, const void* image_state,
                        const void* surfel_stream, int64_t capacity, const float* out_allmap, const void* surfel_aux,
                        const float* dL_dout_color, const float* dL_dout_allmap, void* backward_scratch,
                        int means2D_cols, float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                        float* dL_dmeans3D, float* dL_dtransmat, float* dL_dsh, float* dL_dscales,
                        float* dL_drotations, void* stream) {
    const char* who = "gdr_surfel_backward";
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    if (P == 0) return GDR_OK;
    if (!means3D || !radii || !geom_state || !surfel_state || !image_state || !out_allmap || !surfel_aux ||
        !dL_dout_color || !backward_scratch || !viewmatrix || !projmatrix || !bg)
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: a required pointer is NULL", who);
    if (capacity > 0 && !surfel_stream) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: surfel_stream is NULL", who);
    if (means2D_cols != 3 && means2D_cols != 4) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: means2D_cols must be 3 or 4", who);
    if (!transmat_precomp && (!scales || !rotations || scale_stride < 2))
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: scales / rotations missing", who);
    if ((rotations && (((uintptr_t)rotations) & 15u)) || (dL_drotations && (((uintptr_t)dL_drotations) & 15u)))
        return fail(GDR_ERR_INVALID_ARGUMENT, "%s: float4 buffers must be 16-byte aligned", who);
    if (!colors_precomp && (!shs || !campos)) return fail(GDR_ERR_INVALID_ARGUMENT, "%s: shs / campos is NULL", who);
    gdr::ImageState img = gdr::ImageState::carve(const_cast<void*>(image_state), W, H);
    gdr::GeomState geom = gdr::GeomState::carve(const_cast<void*>(geom_state), (size_t)P);
    float* accum = (float*)backward_scratch;
    GDR_CUDA(cudaMemsetAsync(accum, 0, sizeof(float) * 20 * (size_t)P, s), "memset(accum)");
    {
        StageTimer t(GDR_STAGE_BLEND_BWD, s);
        GDR_CUDA(gdr::launch_surfel_blend_backward(W, H, img, surfel_stream, capacity, bg, out_allmap,
                                                   (const float*)surfel_aux, dL_dout_color, dL_dout_allmap, accum, s),
                 "surfel_blend_backward");
    }
    {
        StageTimer t(GDR_STAGE_GAUSS_BWD, s);
        GDR_CUDA(gdr::launch_surfel_gauss_backward(P, sh_degree, M, W, H, means3D, shs, colors_precomp, scales,
                                                   scale_stride, scale_modifier, rotations, transmat_precomp, viewmatrix,
                                                   projmatrix, campos, radii, surfel_state, geom.clamped, accum,
                                                   means2D_cols, dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D,
                                                   dL_dtransmat, dL_dsh, dL_dscales, dL_drotations, s),
                 "surfel_gauss_backward");
    }
    return GDR_OK;
}

int gdr_knn3_mean_dist2(int P, const float* points, float* out, void* stream) {
    if (P < 0 || (P > 0 && (!points || !out))) return fail(GDR_ERR_INVALID_ARGUMENT, "gdr_knn3_mean_dist2: bad arguments");
    GDR_CUDA(gdr::launch_knn3(P, points, out, (cudaStream_t)stream), "knn3");
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
    std::vector<StageRecord> records;
    {
        std::lock_guard<std::mutex> lock(g_records_mutex);
        records.swap(g_records);
    }
    for (auto& r : records) {
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
    return GDR_OK;
}

}  // extern "C"
