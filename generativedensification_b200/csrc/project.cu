// Stage 1: per-Gaussian projection + tile binning (one launch).
//
// Replaces the reference's preprocessCUDA forward
// (RAST/cuda_rasterizer/forward.cu:155-256 with computeCov3D :118-152,
// computeCov2D :74-113, computeColorFromSH :20-71, in_frustum auxiliary.h:139-164,
// getRect auxiliary.h:46-56) AND its binning front end: the InclusiveSum over
// tiles_touched, the blocking read-back of the instance count and
// duplicateWithKeys (rasterizer_impl.cu:278-301, 70-111).  Differences in
// structure, not in values:
//   * inputs with a 12-byte stride (means, scales, SH rows) reach shared memory as whole-CTA
//     slices moved by the bulk-copy engine (cp.async.bulk + mbarrier; one thread issues every
//     slice), and the 48-byte records / 24-byte covariances leave the same way;
//   * the outputs of the blend are written as one packed 48-byte Splat record;
//   * the same launch bins the instances: every (Gaussian, tile) pair that
//     survives the exact culling test claims a slot of its tile's key segment with
//     one returning atomic and writes its sort key there (tile_iter.cuh) -- no
//     count pass, no prefix sum, no second pass over the Gaussians.
#include <stdlib.h>

#include "kernels.h"
#include "tile_iter.cuh"

namespace gdr {

namespace {

constexpr int PROJ_THREADS = 128;

__device__ constexpr float kSH0 = 0.28209479177387814f;
__device__ constexpr float kSH1 = 0.4886025119029199f;
__device__ constexpr float kSH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                      -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float kSH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                      0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                      -0.5900435899266435f};

// World covariance from scale + (un-normalised) quaternion: Sigma = (S R)^T (S R).
// The roundings are pinned with explicit intrinsics: which product of x*z +- r*y etc. gets fused
// into an FMA decides the last bit of the conic, and through it whether a pair sits above or
// below the alpha = 1/255 cut.  The sequence below is the one nvcc emits for the reference's
// computeCov3D (checked against the SASS of the reference build and bit-for-bit on the GPU).
__device__ __forceinline__ void cov3d_from_scale_rot(float3 scale, float mod, float4 q, float* cov3D) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    float R[3][3];  // R[c][k]: column c, row k of the reference's rotation matrix
    R[0][0] = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(yy, zz)));
    R[0][1] = __fmul_rn(2.f, __fmaf_rn(x, y, -rz));
    R[0][2] = __fmul_rn(2.f, __fmaf_rn(r, y, xz));
    R[1][0] = __fmul_rn(2.f, __fmaf_rn(x, y, rz));
    R[1][1] = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, zz)));
    R[1][2] = __fmul_rn(2.f, __fmaf_rn(y, z, -rx));
    R[2][0] = __fmul_rn(2.f, __fmaf_rn(-r, y, xz));
    R[2][1] = __fmul_rn(2.f, __fmaf_rn(y, z, rx));
    R[2][2] = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, yy)));
    const float s[3] = {__fmul_rn(mod, scale.x), __fmul_rn(mod, scale.y), __fmul_rn(mod, scale.z)};
    float M[3][3];  // M = S * R : M[c][k] = s_k * R[c][k]
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int k = 0; k < 3; k++) M[c][k] = __fmul_rn(s[k], R[c][k]);
    // Sigma[c][r] = sum_k M[r][k] * M[c][k]
    auto dot3 = [&](int rr, int cc) {
        return __fmaf_rn(M[rr][2], M[cc][2], __fmaf_rn(M[rr][0], M[cc][0], __fmul_rn(M[rr][1], M[cc][1])));
    };
    cov3D[0] = dot3(0, 0);
    cov3D[1] = dot3(1, 0);
    cov3D[2] = dot3(2, 0);
    cov3D[3] = dot3(1, 1);
    cov3D[4] = dot3(2, 1);
    cov3D[5] = dot3(2, 2);
}

// EWA screen-space covariance (a, b, c) with the 0.3 px low-pass.
__device__ __forceinline__ float3 cov2d_ewa(const float3 mean, float focal_x, float focal_y, float tan_fovx,
                                            float tan_fovy, const float* cov3D, const float* __restrict__ view) {
    float3 t = xform_point_4x3(mean, view);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;

    Mat3 J;
    J.m[0][0] = focal_x / t.z;
    J.m[0][1] = 0.0f;
    J.m[0][2] = -(focal_x * t.x) / (t.z * t.z);
    J.m[1][0] = 0.0f;
    J.m[1][1] = focal_y / t.z;
    J.m[1][2] = -(focal_y * t.y) / (t.z * t.z);
    J.m[2][0] = 0.0f;
    J.m[2][1] = 0.0f;
    J.m[2][2] = 0.0f;
    Mat3 Wm;
    Wm.m[0][0] = view[0]; Wm.m[0][1] = view[4]; Wm.m[0][2] = view[8];
    Wm.m[1][0] = view[1]; Wm.m[1][1] = view[5]; Wm.m[1][2] = view[9];
    Wm.m[2][0] = view[2]; Wm.m[2][1] = view[6]; Wm.m[2][2] = view[10];
    const Mat3 T = mat3_mul(Wm, J);
    Mat3 Vrk;
    Vrk.m[0][0] = cov3D[0]; Vrk.m[0][1] = cov3D[1]; Vrk.m[0][2] = cov3D[2];
    Vrk.m[1][0] = cov3D[1]; Vrk.m[1][1] = cov3D[3]; Vrk.m[1][2] = cov3D[4];
    Vrk.m[2][0] = cov3D[2]; Vrk.m[2][1] = cov3D[4]; Vrk.m[2][2] = cov3D[5];
    Mat3 cov = mat3_mul(mat3_mul(mat3_transpose(T), mat3_transpose(Vrk)), T);
    cov.m[0][0] += 0.3f;
    cov.m[1][1] += 0.3f;
    return make_float3(cov.m[0][0], cov.m[0][1], cov.m[1][1]);
}

// View-dependent colour from SH coefficients sh[k*3 + ch] (k < (deg+1)^2), +0.5, clamped at 0.
__device__ __forceinline__ float3 sh_to_rgb(int deg, float3 pos, float3 campos, const float* sh, unsigned& clamp_bits) {
    float3 dir = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x = dir.x / len;
    dir.y = dir.y / len;
    dir.z = dir.z / len;
    const float x = dir.x, y = dir.y, z = dir.z;
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float v = kSH0 * sh[ch];
        if (deg > 0) {
            v = v - kSH1 * y * sh[3 + ch] + kSH1 * z * sh[6 + ch] - kSH1 * x * sh[9 + ch];
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z;
                const float xy = x * y, yz = y * z, xz = x * z;
                v = v + kSH2[0] * xy * sh[12 + ch] + kSH2[1] * yz * sh[15 + ch] +
                    kSH2[2] * (2.0f * zz - xx - yy) * sh[18 + ch] + kSH2[3] * xz * sh[21 + ch] +
                    kSH2[4] * (xx - yy) * sh[24 + ch];
                if (deg > 2) {
                    v = v + kSH3[0] * y * (3.0f * xx - yy) * sh[27 + ch] + kSH3[1] * xy * z * sh[30 + ch] +
                        kSH3[2] * y * (4.0f * zz - xx - yy) * sh[33 + ch] +
                        kSH3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + ch] +
                        kSH3[4] * x * (4.0f * zz - xx - yy) * sh[39 + ch] + kSH3[5] * z * (xx - yy) * sh[42 + ch] +
                        kSH3[6] * x * (xx - 3.0f * yy) * sh[45 + ch];
                }
            }
        }
        v += 0.5f;
        if (v < 0) clamp_bits |= (1u << ch);
        res[ch] = fmaxf(v, 0.0f);
    }
    return make_float3(res[0], res[1], res[2]);
}

#ifndef GDR_PROJ_MINB
#define GDR_PROJ_MINB 1
#endif
__global__ void __launch_bounds__(PROJ_THREADS, GDR_PROJ_MINB) project_kernel(const ProjectArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int v = blockIdx.y;  // view of the batch
    const float* __restrict__ viewmatrix = a.vw.view + (size_t)v * a.vw.cam_stride;
    const float* __restrict__ projmatrix = a.vw.proj + (size_t)v * a.vw.cam_stride;
    const float tan_fovx = a.vw.tanx(v), tan_fovy = a.vw.tany(v);
    const float focal_y = a.H / (2.0f * tan_fovy);  // rasterizer_impl.cu:222-223
    const float focal_x = a.W / (2.0f * tan_fovx);
    const GeomState geom = a.geom.at(v, a.vw.geom_stride);
    const ImageState img = a.img.at(v, a.vw.img_stride);
    int32_t* __restrict__ radii = a.radii + (size_t)v * a.P;
    EmitTarget target;
    target.tile_count = img.tile_count;
    target.keys = a.keys + (size_t)v * a.keys_stride;
    target.tile_cap = a.tile_cap;
    target.tile_base = a.tile_base ? a.tile_base + (size_t)v * (a.gx * a.gy + 1) : nullptr;
    target.gx = a.gx;
    __shared__ uint32_t s_tot[2];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        s_tot[0] = s_tot[1] = 0u;
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    uint32_t kept = 0, max_fill = 0;
    const bool use_sh = a.colors_precomp == nullptr;
    const bool use_scale = a.cov3D_precomp == nullptr;
    // shared-memory slices of one virtual block (see the staging notes in common.cuh)
    float* s_mean = smem;                          // 3 * PROJ_THREADS
    float* s_scale = s_mean + 3 * PROJ_THREADS;    // 3 * PROJ_THREADS
    float* s_sh = s_scale + 3 * PROJ_THREADS;      // 3 * M * PROJ_THREADS
    float* s_splat = s_sh + 3 * (use_sh ? a.M : 0) * PROJ_THREADS;  // 12 * PROJ_THREADS: the block's Splat records
    float* s_cov = s_splat + 12 * PROJ_THREADS;    // 6 * PROJ_THREADS: the block's world covariances
    const int emit_offset_words = PROJ_THREADS * (6 + 12 + 6 + 3 * (use_sh ? a.M : 0));  // then one EmitRec per thread
    // Full, 16-byte-aligned blocks move with the bulk-copy engine (one thread issues every slice: all in flight
    // together, no per-thread copy loops); ragged or misaligned ones take the loops.
    const bool aligned = ((((uintptr_t)a.means3D) | ((uintptr_t)a.scales) | ((uintptr_t)a.shs)) & 15u) == 0;
    constexpr uint32_t ROW = sizeof(float) * PROJ_THREADS;  // bytes of one float per Gaussian
    const int n_vblocks = (a.P + PROJ_THREADS - 1) / PROJ_THREADS;
    for (int vb = blockIdx.x; vb < n_vblocks; vb += gridDim.x) {  // virtual blocks: balanced single wave
    if (vb != (int)blockIdx.x) {
        if (threadIdx.x == 0) bulk_wait_read();  // the previous block's records have left shared memory
        __syncthreads();                         // the staging buffers are reused
    }
    const int first = vb * PROJ_THREADS;
    const int n_items = min(PROJ_THREADS, a.P - first);
    const int idx = first + threadIdx.x;
    const bool in_range = threadIdx.x < n_items;
    const bool bulk = aligned && n_items == PROJ_THREADS;

    // ---- stage the 12-byte-stride inputs ----
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, ROW * (3 + (use_scale ? 3 : 0) + (use_sh ? 3 * a.M : 0)));
            bulk_g2s(s_mean, a.means3D + (size_t)first * 3, 3 * ROW, &bar);
            if (use_scale) bulk_g2s(s_scale, a.scales + (size_t)first * 3, 3 * ROW, &bar);
            if (use_sh) bulk_g2s(s_sh, a.shs + (size_t)first * 3 * a.M, 3 * a.M * ROW, &bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
    } else {
        stage_floats(s_mean, a.means3D + (size_t)first * 3, n_items * 3);
        if (use_scale) stage_floats(s_scale, a.scales + (size_t)first * 3, n_items * 3);
        if (use_sh) stage_floats(s_sh, a.shs + (size_t)first * 3 * a.M, n_items * 3 * a.M);
        __syncthreads();
    }

    int n_tiles = 0, rx0 = 0, ry0 = 0, rw = 0;
    Splat rec;
    rec.q0 = rec.q1 = rec.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_range) {
        int radius_out = 0;
        rec.q0 = make_float4(0.f, 0.f, 0.f, __int_as_float(idx));
        rec.q1 = make_float4(0.f, 0.f, 0.f, 0.f);
        rec.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned clamp_bits = 0;
        float cov3D[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

        const float3 p_orig = make_float3(s_mean[3 * threadIdx.x], s_mean[3 * threadIdx.x + 1],
                                          s_mean[3 * threadIdx.x + 2]);
        const float4 p_hom = xform_point_4x4(p_orig, projmatrix);
        const float p_w = 1.0f / (p_hom.w + 0.0000001f);
        const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);
        const float3 p_view = xform_point_4x3(p_orig, viewmatrix);

        // near-plane cull only (auxiliary.h:152: `if (p_view.z <= 0.2f) return false`, so a NaN depth passes, exactly as
        // in the reference -- such a Gaussian then ends with an empty tile rectangle there and here)
        if (!(p_view.z <= NEAR_Z)) {
            if (a.cov3D_precomp != nullptr) {
#pragma unroll
                for (int k = 0; k < 6; k++) cov3D[k] = __ldg(a.cov3D_precomp + (size_t)idx * 6 + k);
            } else {
                float3 sc = make_float3(s_scale[3 * threadIdx.x], s_scale[3 * threadIdx.x + 1],
                                        s_scale[3 * threadIdx.x + 2]);
                float4 q = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
                if (a.raw_params) {  // activation-fused inputs (renderer.py:225-230 done here)
                    sc = make_float3(act_scale(sc.x), act_scale(sc.y), act_scale(sc.z));
                    q = act_rotation(q);
                }
                cov3d_from_scale_rot(sc, a.scale_modifier, q, cov3D);
            }
            const float3 cov = cov2d_ewa(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix);
            const float det = (cov.x * cov.z - cov.y * cov.y);
            if (det != 0.0f) {
                const float det_inv = 1.f / det;
                const float3 conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);
                const float mid = 0.5f * (cov.x + cov.z);
                const float lambda1 = mid + sqrtf(max(0.1f, mid * mid - det));
                const float lambda2 = mid - sqrtf(max(0.1f, mid * mid - det));
                const float my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
                const float2 pix = make_float2(ndc_to_pix(p_proj.x, a.W), ndc_to_pix(p_proj.y, a.H));
                int x0, y0, x1, y1;
                tile_rect(pix.x, pix.y, (int)my_radius, a.gx, a.gy, x0, y0, x1, y1);
                if ((x1 - x0) * (y1 - y0) != 0) {
                    float3 rgb;
                    if (use_sh) {
                        const float* cp = a.vw.campos + (size_t)v * a.vw.cam_stride;
                        const float3 cam = make_float3(__ldg(cp), __ldg(cp + 1), __ldg(cp + 2));
                        rgb = sh_to_rgb(a.sh_degree, p_orig, cam, s_sh + (size_t)threadIdx.x * 3 * a.M, clamp_bits);
                    } else {
                        rgb = make_float3(__ldg(a.colors_precomp + (size_t)idx * 3),
                                          __ldg(a.colors_precomp + (size_t)idx * 3 + 1),
                                          __ldg(a.colors_precomp + (size_t)idx * 3 + 2));
                    }
                    float opacity = __ldg(a.opacities + idx);
                    if (a.raw_params) opacity = act_opacity(opacity);
                    // Conservative reject threshold: power < thr  =>  opacity*exp(power) < 1/255 for certain
                    // (1e-4 of slack in the exponent dwarfs every rounding error of the exact test).
                    // (opacity <= 0 can never reach 1/255: reject everything; NaN opacity: never reject.)
                    const float thr = opacity > 0.f ? (logf(ALPHA_MIN / opacity) - 1e-4f)
                                                    : (opacity <= 0.f ? INFINITY : -INFINITY);
                    radius_out = (int)my_radius;
                    rec.q0 = make_float4(pix.x, pix.y, thr, __int_as_float(idx));
                    rec.q1 = make_float4(conic.x, conic.y, conic.z, opacity);
                    rec.q2 = make_float4(rgb.x, rgb.y, rgb.z, p_view.z);
                    rx0 = x0;
                    ry0 = y0;
                    rw = x1 - x0;
                    n_tiles = (y1 - y0) * (x1 - x0);
                }
            }
        } else if (a.prefiltered) {
            pdl_wait();  // the header is being zeroed by the launch before this one (see launch_project)
            atomicOr(&img.header[HDR_PROJECT_FLAGS], HDR_FLAG_PREFILTERED);  // the reference traps here (auxiliary.h:154-158);
                                                                     // we flag, the host raises
        }
        radii[idx] = radius_out;
        geom.tiles_touched[idx] = (uint32_t)n_tiles;
        geom.clamped[idx] = (uint8_t)clamp_bits;
        // the 48-byte record and the 24-byte covariance: through shared memory for full blocks (they leave as two
        // bulk stores below), directly otherwise
        Splat* rec_dst = bulk ? reinterpret_cast<Splat*>(s_splat) + threadIdx.x : geom.splat + idx;
        *rec_dst = rec;
        if (use_scale) {
            float2* c = reinterpret_cast<float2*>(bulk ? s_cov + 6 * threadIdx.x : geom.cov3D + (size_t)idx * 6);
            c[0] = make_float2(cov3D[0], cov3D[1]);
            c[1] = make_float2(cov3D[2], cov3D[3]);
            c[2] = make_float2(cov3D[4], cov3D[5]);
        }
    }
    if (bulk) {
        fence_async_smem();  // this thread's rows -> visible to the bulk-copy engine
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_s2g(geom.splat + first, s_splat, 12 * ROW);
            if (use_scale) bulk_s2g(geom.cov3D + (size_t)first * 6, s_cov, 6 * ROW);
            bulk_commit();
        }
    }

    // ---- bin the instances: claim a slot of the tile's key segment per kept (Gaussian, tile) pair ----
    // With culling on, a pair is only binned if the splat can reach alpha >= 1/255 somewhere in the tile
    // (exact: see splat_misses_rect in common.cuh).
    // Everything above touched neither the header nor the slot counters: this launch is a programmatic dependent of the
    // kernel that zeroes them (zero_state_kernel), so the staging and the projection arithmetic overlap it; the
    // claims need the zeros.
    pdl_wait();
    EmitRec* s_rec = reinterpret_cast<EmitRec*>(smem + emit_offset_words) + (threadIdx.x & ~31u);
    const uint32_t depth_bits = __float_as_uint(rec.q2.w);
    if (a.cull)
        warp_emit_tiles<true>(s_rec, n_tiles, rx0, ry0, rw, rec.q0, rec.q1, depth_bits, (uint32_t)idx, target, kept,
                              max_fill);
    else
        warp_emit_tiles<false>(s_rec, n_tiles, rx0, ry0, rw, rec.q0, rec.q1, depth_bits, (uint32_t)idx, target, kept,
                               max_fill);
    }  // virtual blocks
    // instance total and largest claim of this CTA: one atomic each per CTA on the view's header
    kept = __reduce_add_sync(0xffffffffu, kept);
    max_fill = __reduce_max_sync(0xffffffffu, max_fill);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_tot[0], kept);
        atomicMax(&s_tot[1], max_fill);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_tot[0]) atomicAdd(&img.header[HDR_NUM_RENDERED], s_tot[0]);
        if (s_tot[1]) atomicMax(&img.header[HDR_MAX_TILE], s_tot[1]);
        bulk_wait_all();  // this CTA's record / covariance stores are performed before it reports and exits
        if (a.counts_host) report_counts(img.header, a.counts_host + 4 * v, gridDim.x);
    }
    pdl_trigger();  // tile_sort may start launching
}

// Zeroes the header and the per-tile slot counters of every view (they are adjacent in ImageState): a kernel rather
// than a memset node so that the projection kernel can be launched as its programmatic dependent -- it releases the
// dependent at once, and the projection only waits for it where it first needs the zeros.
__global__ void __launch_bounds__(256) zero_state_kernel(uint4* __restrict__ base, size_t n16, size_t view_stride_bytes) {
    pdl_trigger();
    uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<char*>(base) + (size_t)blockIdx.y * view_stride_bytes);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 v = xform_point_4x3(p, view);
    present[idx] = !(v.z <= NEAR_Z);  // in_frustum (auxiliary.h:152): NaN passes, as in the reference
}

}  // namespace

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("GDR_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

cudaError_t launch_project(const ProjectArgs& a, cudaStream_t s) {
    if (a.P <= 0) return cudaSuccess;
    const size_t smem = sizeof(float) * PROJ_THREADS * (6 + 12 + 6 + 3 * (size_t)(a.colors_precomp ? 0 : a.M)) +
                        sizeof(EmitRec) * PROJ_THREADS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int n_vblocks = (a.P + PROJ_THREADS - 1) / PROJ_THREADS;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, project_kernel, PROJ_THREADS, smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    const int V = max(1, a.vw.V);
    const int grid = min(n_vblocks, max(1, sm_count() * per_sm / V));  // one balanced wave over all views
    return launch_dependent(project_kernel, dim3(grid, V), dim3(PROJ_THREADS), smem, s, a);
}

cudaError_t launch_zero_state(ImageState img, int W, int H, const Views& vw, cudaStream_t s) {
    const size_t T = ImageState::tiles(W, H);
    const size_t bytes = (size_t)((char*)(img.tile_count + T * COUNT_STRIDE) - (char*)img.header);  // a multiple of 256
    const size_t n16 = bytes / sizeof(uint4);
    const int V = max(1, vw.V);
    const size_t blocks = (n16 + 255) / 256, cap = (size_t)sm_count() * 4;
    const int grid = (int)(blocks < cap ? blocks : cap);
    zero_state_kernel<<<dim3(grid, V), 256, 0, s>>>(reinterpret_cast<uint4*>(img.header), n16, vw.img_stride);
    return cudaGetLastError();
}

cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                                cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
    return cudaGetLastError();
}

}  // namespace gdr
