// Stage 5: per-Gaussian backward, one fused launch.
//
// Replaces the reference's computeCov2DCUDA + preprocessCUDA backward
// (RAST/cuda_rasterizer/backward.cu:144-274, 346-412) with their device helpers
// computeColorFromSH bwd (:20-139), computeCov3D bwd (:278-341) and dnormvdv
// (auxiliary.h:107-117), and the ten torch::zeros fills of
// RasterizeGaussiansBackwardCUDA (RAST/rasterize_points.cu:153-162): every
// requested output row is written here (zeros for culled Gaussians), so the
// caller allocates with empty() and nothing is memset.
//
// Inputs: the 12 screen-space sums per Gaussian produced by blend_bwd.cu.
#include "kernels.h"

namespace gdr {

namespace {

__device__ constexpr float kSH0 = 0.28209479177387814f;
__device__ constexpr float kSH1 = 0.4886025119029199f;
__device__ constexpr float kSH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                      -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float kSH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                      0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                      -0.5900435899266435f};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 operator*(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

constexpr int GB_THREADS = 128;
constexpr int GRAD_RAW_PARAMS = 32;  // == GDR_GRAD_RAW_PARAMS (include/gdr.h)

template <bool ACC>
__device__ __forceinline__ void gauss_backward_one(const GaussBackwardArgs& a, const int idx);

// ACC = false stores every requested output row; ACC = true adds to it (views 1.. of a batch launch in
// stream order after view 0, so the per-Gaussian gradients are summed over views deterministically).
template <bool ACC>
__global__ void __launch_bounds__(GB_THREADS) gauss_backward_kernel(const GaussBackwardArgs a) {
    for (int idx = blockIdx.x * GB_THREADS + threadIdx.x; idx < a.P; idx += gridDim.x * GB_THREADS)
        gauss_backward_one<ACC>(a, idx);
}

template <bool ACC>
__device__ __forceinline__ void put(float* p, float v) {
    if (ACC) *p += v; else *p = v;
}
template <bool ACC>
__device__ __forceinline__ void putv(V3* p, V3 v) {
    if (ACC) {
        const V3 o = *p;
        v = V3{o.x + v.x, o.y + v.y, o.z + v.z};
    }
    *p = v;
}
template <bool ACC>
__device__ __forceinline__ void put4(float* p, float4 v) {
    float4* q = reinterpret_cast<float4*>(p);
    if (ACC) {
        const float4 o = *q;
        v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
    }
    *q = v;
}

template <bool ACC>
__device__ __forceinline__ void gauss_backward_one(const GaussBackwardArgs& a, const int idx) {
    const int vi = a.view;
    const bool visible = a.radii[(size_t)vi * a.P + idx] > 0;
    const GeomState geom = a.geom.at(vi, a.vw.geom_stride);
    const int M = a.M;
    // Every per-Gaussian input is fetched here, in ONE round trip: the output stores below may alias the inputs as
    // far as the compiler knows, so it would issue these loads only after the stores (which wait for the
    // accumulators) -- two serialized round trips in a latency-bound kernel.  None of them is written by the
    // backward blend, so as a programmatic dependent of that kernel this launch fetches them while the blend's last
    // CTAs are still running and only then waits (pdl_wait) for the accumulators.
    const float3 mean = make_float3(__ldg(a.means3D + (size_t)idx * 3), __ldg(a.means3D + (size_t)idx * 3 + 1),
                                    __ldg(a.means3D + (size_t)idx * 3 + 2));
    float cov3D[6];
    {
        const float2* c2 = reinterpret_cast<const float2*>(
            a.cov3D_precomp ? a.cov3D_precomp + (size_t)idx * 6 : geom.cov3D + (size_t)idx * 6);
        const float2 c0 = c2[0], c1 = c2[1], c2v = c2[2];
        cov3D[0] = c0.x; cov3D[1] = c0.y; cov3D[2] = c1.x; cov3D[3] = c1.y; cov3D[4] = c2v.x; cov3D[5] = c2v.y;
    }
    float4 q_in = make_float4(1.f, 0.f, 0.f, 0.f);
    float3 sc_in = make_float3(0.f, 0.f, 0.f);
    const bool want_scale_rot = a.scales != nullptr && (a.dL_dscales || a.dL_drotations);
    if (want_scale_rot) {
        q_in = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
        sc_in = make_float3(__ldg(a.scales + (size_t)idx * 3), __ldg(a.scales + (size_t)idx * 3 + 1),
                            __ldg(a.scales + (size_t)idx * 3 + 2));
    }
    const unsigned clamp_in = a.shs != nullptr ? (unsigned)geom.clamped[idx] : 0u;
    V3 sh1 = {0, 0, 0}, sh2 = {0, 0, 0}, sh3 = {0, 0, 0};  // the degree-1 band (the repo's default degree)
    if (a.shs != nullptr && a.sh_degree > 0) {
        const float* sp = a.shs + ((size_t)idx * M + 1) * 3;
        sh1 = V3{__ldg(sp + 0), __ldg(sp + 1), __ldg(sp + 2)};
        sh2 = V3{__ldg(sp + 3), __ldg(sp + 4), __ldg(sp + 5)};
        sh3 = V3{__ldg(sp + 6), __ldg(sp + 7), __ldg(sp + 8)};
    }
    const float opacity_in = (a.dL_dopacity && (a.grad_mask & GRAD_RAW_PARAMS)) ? geom.splat[idx].q1.w : 0.f;
    pdl_wait();  // the backward blend (or the previous view's launch of this kernel) has completed
    const float4* acc = reinterpret_cast<const float4*>(a.accum + ((size_t)vi * a.P + idx) * 12);
    const float4 g_mean2D = acc[0];
    const float4 g_conic_op = acc[1];
    const float4 g_rgb_depth = acc[2];

    if (a.dL_dmeans2D) put4<ACC>(a.dL_dmeans2D + (size_t)idx * 4, g_mean2D);
    const bool raw = (a.grad_mask & GRAD_RAW_PARAMS) != 0;  // gradients w.r.t. logits / log-scales / raw quaternions
    if (a.dL_dopacity) {
        float g = g_conic_op.w;
        if (raw) {  // d sigmoid: o (1 - o); the activated opacity is in the saved record (0 for culled Gaussians)
            const float o = opacity_in;
            g = g * (1.f - o) * o;
        }
        put<ACC>(a.dL_dopacity + idx, g);
    }
    if (a.dL_dcolors) {
        put<ACC>(a.dL_dcolors + (size_t)idx * 3 + 0, g_rgb_depth.x);
        put<ACC>(a.dL_dcolors + (size_t)idx * 3 + 1, g_rgb_depth.y);
        put<ACC>(a.dL_dcolors + (size_t)idx * 3 + 2, g_rgb_depth.z);
    }
    const bool want_geo = a.dL_dmeans3D || a.dL_dcov3D || a.dL_dscales || a.dL_drotations;
    const bool want_sh = a.dL_dsh != nullptr && a.shs != nullptr;
    if (!want_geo && !want_sh) return;

    if (!visible) {
        if (ACC) return;  // adds nothing
        if (a.dL_dmeans3D)
            for (int k = 0; k < 3; k++) a.dL_dmeans3D[(size_t)idx * 3 + k] = 0.f;
        if (a.dL_dcov3D)
            for (int k = 0; k < 6; k++) a.dL_dcov3D[(size_t)idx * 6 + k] = 0.f;
        if (a.dL_dsh)
            for (int k = 0; k < 3 * M; k++) a.dL_dsh[(size_t)idx * 3 * M + k] = 0.f;
        if (a.dL_dscales)
            for (int k = 0; k < 3; k++) a.dL_dscales[(size_t)idx * 3 + k] = 0.f;
        if (a.dL_drotations) reinterpret_cast<float4*>(a.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }

    const float* view = a.vw.view + (size_t)vi * a.vw.cam_stride;
    const float* proj = a.vw.proj + (size_t)vi * a.vw.cam_stride;
    const float* campos = a.vw.campos + (size_t)vi * a.vw.cam_stride;
    const float tan_fovx = a.vw.tanx(vi), tan_fovy = a.vw.tany(vi);
    const float focal_y = a.H / (2.0f * tan_fovy), focal_x = a.W / (2.0f * tan_fovx);
    float dL_dcov[6];
    float3 dL_dmean;

    // ---- conic -> cov2D -> (cov3D, mean) : backward.cu:144-274 ----
    {
        const float3 dL_dconic = make_float3(g_conic_op.x, g_conic_op.y, g_conic_op.z);
        float3 t = xform_point_4x3(mean, view);
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = t.x / t.z, tytz = t.y / t.z;
        t.x = min(limx, max(-limx, txtz)) * t.z;
        t.y = min(limy, max(-limy, tytz)) * t.z;
        const float x_grad_mul = txtz < -limx || txtz > limx ? 0.f : 1.f;
        const float y_grad_mul = tytz < -limy || tytz > limy ? 0.f : 1.f;
        const float h_x = focal_x, h_y = focal_y;

        Mat3 J;
        J.m[0][0] = h_x / t.z; J.m[0][1] = 0.0f;      J.m[0][2] = -(h_x * t.x) / (t.z * t.z);
        J.m[1][0] = 0.0f;      J.m[1][1] = h_y / t.z; J.m[1][2] = -(h_y * t.y) / (t.z * t.z);
        J.m[2][0] = 0.0f;      J.m[2][1] = 0.0f;      J.m[2][2] = 0.0f;
        Mat3 Wm;
        Wm.m[0][0] = view[0]; Wm.m[0][1] = view[4]; Wm.m[0][2] = view[8];
        Wm.m[1][0] = view[1]; Wm.m[1][1] = view[5]; Wm.m[1][2] = view[9];
        Wm.m[2][0] = view[2]; Wm.m[2][1] = view[6]; Wm.m[2][2] = view[10];
        Mat3 Vrk;
        Vrk.m[0][0] = cov3D[0]; Vrk.m[0][1] = cov3D[1]; Vrk.m[0][2] = cov3D[2];
        Vrk.m[1][0] = cov3D[1]; Vrk.m[1][1] = cov3D[3]; Vrk.m[1][2] = cov3D[4];
        Vrk.m[2][0] = cov3D[2]; Vrk.m[2][1] = cov3D[4]; Vrk.m[2][2] = cov3D[5];
        const Mat3 T = mat3_mul(Wm, J);
        const Mat3 cov2D = mat3_mul(mat3_mul(mat3_transpose(T), mat3_transpose(Vrk)), T);
        const float ca = cov2D.m[0][0] + 0.3f;
        const float cb = cov2D.m[0][1];
        const float cc = cov2D.m[1][1] + 0.3f;
        const float denom = ca * cc - cb * cb;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (denom2inv != 0) {
            dL_da = denom2inv * (-cc * cc * dL_dconic.x + 2 * cb * cc * dL_dconic.y + (denom - ca * cc) * dL_dconic.z);
            dL_dc = denom2inv * (-ca * ca * dL_dconic.z + 2 * ca * cb * dL_dconic.y + (denom - ca * cc) * dL_dconic.x);
            dL_db = denom2inv * 2 * (cb * cc * dL_dconic.x - (denom + 2 * cb * cb) * dL_dconic.y + ca * cb * dL_dconic.z);
            dL_dcov[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
            dL_dcov[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
            dL_dcov[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
            dL_dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db +
                         2 * T.m[1][0] * T.m[1][1] * dL_dc;
            dL_dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db +
                         2 * T.m[1][0] * T.m[1][2] * dL_dc;
            dL_dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db +
                         2 * T.m[1][1] * T.m[1][2] * dL_dc;
        } else {
#pragma unroll
            for (int i = 0; i < 6; i++) dL_dcov[i] = 0;
        }
        float dL_dT0[3], dL_dT1[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const float tv0 = T.m[0][0] * Vrk.m[q][0] + T.m[0][1] * Vrk.m[q][1] + T.m[0][2] * Vrk.m[q][2];
            const float tv1 = T.m[1][0] * Vrk.m[q][0] + T.m[1][1] * Vrk.m[q][1] + T.m[1][2] * Vrk.m[q][2];
            dL_dT0[q] = 2 * tv0 * dL_da + tv1 * dL_db;
            dL_dT1[q] = 2 * tv1 * dL_dc + tv0 * dL_db;
        }
        const float dL_dJ00 = Wm.m[0][0] * dL_dT0[0] + Wm.m[0][1] * dL_dT0[1] + Wm.m[0][2] * dL_dT0[2];
        const float dL_dJ02 = Wm.m[2][0] * dL_dT0[0] + Wm.m[2][1] * dL_dT0[1] + Wm.m[2][2] * dL_dT0[2];
        const float dL_dJ11 = Wm.m[1][0] * dL_dT1[0] + Wm.m[1][1] * dL_dT1[1] + Wm.m[1][2] * dL_dT1[2];
        const float dL_dJ12 = Wm.m[2][0] * dL_dT1[0] + Wm.m[2][1] * dL_dT1[1] + Wm.m[2][2] * dL_dT1[2];
        const float tz = 1.f / t.z;
        const float tz2 = tz * tz;
        const float tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                             (2 * h_y * t.y) * tz3 * dL_dJ12;
        // mean = W^T-part of the view transform applied to dL/dt (this term initialises dL/dmean)
        dL_dmean.x = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        dL_dmean.y = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        dL_dmean.z = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;
    }

    // ---- mean2D and depth -> mean3D : backward.cu:362-393 ----
    {
        const float4 m_hom = xform_point_4x4(mean, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        dL_dmean.x += (proj[0] * m_w - proj[3] * mul1) * g_mean2D.x + (proj[1] * m_w - proj[3] * mul2) * g_mean2D.y;
        dL_dmean.y += (proj[4] * m_w - proj[7] * mul1) * g_mean2D.x + (proj[5] * m_w - proj[7] * mul2) * g_mean2D.y;
        dL_dmean.z += (proj[8] * m_w - proj[11] * mul1) * g_mean2D.x + (proj[9] * m_w - proj[11] * mul2) * g_mean2D.y;
        const float mul3 = view[2] * mean.x + view[6] * mean.y + view[10] * mean.z + view[14];
        const float g_depth = g_rgb_depth.w;
        dL_dmean.x += (view[2] - view[3] * mul3) * g_depth;
        dL_dmean.y += (view[6] - view[7] * mul3) * g_depth;
        dL_dmean.z += (view[10] - view[11] * mul3) * g_depth;
    }

    // ---- colour -> SH coefficients and view direction : backward.cu:20-139 ----
    if (a.shs != nullptr) {
        const V3 dir_orig = {mean.x - campos[0], mean.y - campos[1], mean.z - campos[2]};
        const float len = sqrtf(dot(dir_orig, dir_orig));
        const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
        const V3* sh = reinterpret_cast<const V3*>(a.shs) + (size_t)idx * M;
        const unsigned cl = clamp_in;
        V3 dRGB = {g_rgb_depth.x, g_rgb_depth.y, g_rgb_depth.z};
        dRGB.x *= (cl & 1u) ? 0.f : 1.f;
        dRGB.y *= (cl & 2u) ? 0.f : 1.f;
        dRGB.z *= (cl & 4u) ? 0.f : 1.f;
        V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
        V3* dsh = a.dL_dsh ? reinterpret_cast<V3*>(a.dL_dsh) + (size_t)idx * M : nullptr;
        const int deg = a.sh_degree;
        const int used = (deg + 1) * (deg + 1);
        if (dsh) {
            putv<ACC>(dsh + 0, kSH0 * dRGB);
            if (!ACC)
                for (int k = used; k < M; k++) dsh[k] = V3{0.f, 0.f, 0.f};
        }
        if (deg > 0) {
            if (dsh) {
                putv<ACC>(dsh + 1, (-kSH1 * y) * dRGB);
                putv<ACC>(dsh + 2, (kSH1 * z) * dRGB);
                putv<ACC>(dsh + 3, (-kSH1 * x) * dRGB);
            }
            dRGBdx = -kSH1 * sh3;
            dRGBdy = -kSH1 * sh1;
            dRGBdz = kSH1 * sh2;
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z;
                const float xy = x * y, yz = y * z, xz = x * z;
                if (dsh) {
                    putv<ACC>(dsh + 4, (kSH2[0] * xy) * dRGB);
                    putv<ACC>(dsh + 5, (kSH2[1] * yz) * dRGB);
                    putv<ACC>(dsh + 6, (kSH2[2] * (2.f * zz - xx - yy)) * dRGB);
                    putv<ACC>(dsh + 7, (kSH2[3] * xz) * dRGB);
                    putv<ACC>(dsh + 8, (kSH2[4] * (xx - yy)) * dRGB);
                }
                dRGBdx = dRGBdx + (kSH2[0] * y) * sh[4] + (kSH2[2] * 2.f * -x) * sh[6] + (kSH2[3] * z) * sh[7] +
                         (kSH2[4] * 2.f * x) * sh[8];
                dRGBdy = dRGBdy + (kSH2[0] * x) * sh[4] + (kSH2[1] * z) * sh[5] + (kSH2[2] * 2.f * -y) * sh[6] +
                         (kSH2[4] * 2.f * -y) * sh[8];
                dRGBdz = dRGBdz + (kSH2[1] * y) * sh[5] + (kSH2[2] * 2.f * 2.f * z) * sh[6] + (kSH2[3] * x) * sh[7];
                if (deg > 2) {
                    if (dsh) {
                        putv<ACC>(dsh + 9, (kSH3[0] * y * (3.f * xx - yy)) * dRGB);
                        putv<ACC>(dsh + 10, (kSH3[1] * xy * z) * dRGB);
                        putv<ACC>(dsh + 11, (kSH3[2] * y * (4.f * zz - xx - yy)) * dRGB);
                        putv<ACC>(dsh + 12, (kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dRGB);
                        putv<ACC>(dsh + 13, (kSH3[4] * x * (4.f * zz - xx - yy)) * dRGB);
                        putv<ACC>(dsh + 14, (kSH3[5] * z * (xx - yy)) * dRGB);
                        putv<ACC>(dsh + 15, (kSH3[6] * x * (xx - 3.f * yy)) * dRGB);
                    }
                    dRGBdx = dRGBdx + (kSH3[0] * 3.f * 2.f * xy) * sh[9] + (kSH3[1] * yz) * sh[10] +
                             (kSH3[2] * -2.f * xy) * sh[11] + (kSH3[3] * -3.f * 2.f * xz) * sh[12] +
                             (kSH3[4] * (-3.f * xx + 4.f * zz - yy)) * sh[13] + (kSH3[5] * 2.f * xz) * sh[14] +
                             (kSH3[6] * 3.f * (xx - yy)) * sh[15];
                    dRGBdy = dRGBdy + (kSH3[0] * 3.f * (xx - yy)) * sh[9] + (kSH3[1] * xz) * sh[10] +
                             (kSH3[2] * (-3.f * yy + 4.f * zz - xx)) * sh[11] + (kSH3[3] * -3.f * 2.f * yz) * sh[12] +
                             (kSH3[4] * -2.f * xy) * sh[13] + (kSH3[5] * -2.f * yz) * sh[14] +
                             (kSH3[6] * -3.f * 2.f * xy) * sh[15];
                    dRGBdz = dRGBdz + (kSH3[1] * xy) * sh[10] + (kSH3[2] * 4.f * 2.f * yz) * sh[11] +
                             (kSH3[3] * 3.f * (2.f * zz - xx - yy)) * sh[12] + (kSH3[4] * 4.f * 2.f * xz) * sh[13] +
                             (kSH3[5] * (xx - yy)) * sh[14];
                }
            }
        }
        const float3 dL_ddir = make_float3(dot(dRGBdx, dRGB), dot(dRGBdy, dRGB), dot(dRGBdz, dRGB));
        // through the normalisation of the view direction
        const float3 v = make_float3(dir_orig.x, dir_orig.y, dir_orig.z);
        const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
        const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        dL_dmean.x += ((+sum2 - v.x * v.x) * dL_ddir.x - v.y * v.x * dL_ddir.y - v.z * v.x * dL_ddir.z) * invsum32;
        dL_dmean.y += (-v.x * v.y * dL_ddir.x + (sum2 - v.y * v.y) * dL_ddir.y - v.z * v.y * dL_ddir.z) * invsum32;
        dL_dmean.z += (-v.x * v.z * dL_ddir.x - v.y * v.z * dL_ddir.y + (sum2 - v.z * v.z) * dL_ddir.z) * invsum32;
    } else if (a.dL_dsh && !ACC) {
        for (int k = 0; k < 3 * M; k++) a.dL_dsh[(size_t)idx * 3 * M + k] = 0.f;
    }

    if (a.dL_dmeans3D) {
        put<ACC>(a.dL_dmeans3D + (size_t)idx * 3 + 0, dL_dmean.x);
        put<ACC>(a.dL_dmeans3D + (size_t)idx * 3 + 1, dL_dmean.y);
        put<ACC>(a.dL_dmeans3D + (size_t)idx * 3 + 2, dL_dmean.z);
    }
    if (a.dL_dcov3D)
        for (int k = 0; k < 6; k++) put<ACC>(a.dL_dcov3D + (size_t)idx * 6 + k, dL_dcov[k]);

    // ---- cov3D -> scale, rotation : backward.cu:278-341 ----
    if (want_scale_rot) {
        float4 q = q_in;
        float3 sc = sc_in;
        float qn = 1.f;
        if (raw) {
            qn = fmaxf(quat_norm(q), 1e-12f);
            q = act_rotation(q);
            sc = make_float3(act_scale(sc.x), act_scale(sc.y), act_scale(sc.z));
        }
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        Mat3 R;
        R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
        R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
        R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
        const float3 s = make_float3(a.scale_modifier * sc.x, a.scale_modifier * sc.y, a.scale_modifier * sc.z);
        Mat3 S;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) S.m[c][rr] = 0.0f;
        S.m[0][0] = s.x; S.m[1][1] = s.y; S.m[2][2] = s.z;
        const Mat3 Mm = mat3_mul(S, R);
        Mat3 dL_dSigma;
        dL_dSigma.m[0][0] = dL_dcov[0];        dL_dSigma.m[0][1] = 0.5f * dL_dcov[1]; dL_dSigma.m[0][2] = 0.5f * dL_dcov[2];
        dL_dSigma.m[1][0] = 0.5f * dL_dcov[1]; dL_dSigma.m[1][1] = dL_dcov[3];        dL_dSigma.m[1][2] = 0.5f * dL_dcov[4];
        dL_dSigma.m[2][0] = 0.5f * dL_dcov[2]; dL_dSigma.m[2][1] = 0.5f * dL_dcov[4]; dL_dSigma.m[2][2] = dL_dcov[5];
        Mat3 M2;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) M2.m[c][rr] = 2.0f * Mm.m[c][rr];
        const Mat3 dL_dM = mat3_mul(M2, dL_dSigma);
        const Mat3 Rt = mat3_transpose(R);
        Mat3 dL_dMt = mat3_transpose(dL_dM);
        if (a.dL_dscales) {
            // d exp: multiply by the activated scale when the inputs are log-scales
            const float k0 = raw ? sc.x : 1.f, k1 = raw ? sc.y : 1.f, k2 = raw ? sc.z : 1.f;
            put<ACC>(a.dL_dscales + (size_t)idx * 3 + 0, k0 * (Rt.m[0][0] * dL_dMt.m[0][0] + Rt.m[0][1] * dL_dMt.m[0][1] + Rt.m[0][2] * dL_dMt.m[0][2]));
            put<ACC>(a.dL_dscales + (size_t)idx * 3 + 1, k1 * (Rt.m[1][0] * dL_dMt.m[1][0] + Rt.m[1][1] * dL_dMt.m[1][1] + Rt.m[1][2] * dL_dMt.m[1][2]));
            put<ACC>(a.dL_dscales + (size_t)idx * 3 + 2, k2 * (Rt.m[2][0] * dL_dMt.m[2][0] + Rt.m[2][1] * dL_dMt.m[2][1] + Rt.m[2][2] * dL_dMt.m[2][2]));
        }
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
            dL_dMt.m[0][rr] *= s.x;
            dL_dMt.m[1][rr] *= s.y;
            dL_dMt.m[2][rr] *= s.z;
        }
        if (a.dL_drotations) {
            float4 dq;
            dq.x = 2 * z * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * y * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) + 2 * x * (dL_dMt.m[1][2] - dL_dMt.m[2][1]);
            dq.y = 2 * y * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * z * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) + 2 * r * (dL_dMt.m[1][2] - dL_dMt.m[2][1]) - 4 * x * (dL_dMt.m[2][2] + dL_dMt.m[1][1]);
            dq.z = 2 * x * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * r * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) + 2 * z * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * y * (dL_dMt.m[2][2] + dL_dMt.m[0][0]);
            dq.w = 2 * r * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * x * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) + 2 * y * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * z * (dL_dMt.m[1][1] + dL_dMt.m[0][0]);
            if (raw) {  // through q / ||q||: (g - q^ (q^ . g)) / ||q||
                const float d = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w;
                const float inv = 1.f / qn;
                dq = make_float4((dq.x - q.x * d) * inv, (dq.y - q.y * d) * inv, (dq.z - q.z * d) * inv,
                                 (dq.w - q.w * d) * inv);
            }
            put4<ACC>(a.dL_drotations + (size_t)idx * 4, dq);
        }
    } else if (!ACC) {
        if (a.dL_dscales)
            for (int k = 0; k < 3; k++) a.dL_dscales[(size_t)idx * 3 + k] = 0.f;
        if (a.dL_drotations) reinterpret_cast<float4*>(a.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace

cudaError_t launch_gauss_backward(const GaussBackwardArgs& a, cudaStream_t s) {
    if (a.P <= 0) return cudaSuccess;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gauss_backward_kernel<false>, GB_THREADS, 0) !=
            cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    const int grid = min((a.P + GB_THREADS - 1) / GB_THREADS, sm_count() * per_sm);
    return a.accumulate ? launch_dependent(gauss_backward_kernel<true>, dim3(grid), dim3(GB_THREADS), 0, s, a)
                        : launch_dependent(gauss_backward_kernel<false>, dim3(grid), dim3(GB_THREADS), 0, s, a);
}

}  // namespace gdr
