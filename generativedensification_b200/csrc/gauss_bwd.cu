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
//
// The chain rule is written from the structure of the forward rather than as generic 3x3 products:
//   * the screen-space covariance only needs the two rows m0, m1 of M = J Wv (J has two non-zero entries per
//     row): (a, b, c) = (m0.S m0, m0.S m1, m1.S m1) with S the world covariance, and every gradient below is a
//     combination of u0 = S m0 and u1 = S m1 -- 45 multiply-adds where three general matrix products take 135;
//   * world covariance S = N N^T with N = R diag(s): dL/dN = 2 G N for the symmetric G = dL/dS; the scale gradient
//     is the column-wise dot of dL/dN with R and the rotation gradient follows from dL/dR = dL/dN diag(s);
//   * the SH colour gradient w.r.t. the view direction is sum_k (sh_k . dL/drgb) grad b_k (one scalar per
//     coefficient), and the direction's normalisation is (g - dir (dir . g)) / |d|.
// The reference's quirks are kept: denominators det^2 + 1e-7 and w + 1e-7, the clamp gating
// of the 1.3 tan(fov) limit applies to dL/dt.xy only, gradients w.r.t. the UN-normalised quaternion.
#include "kernels.h"

namespace gdr {

namespace {

constexpr float SH0 = 0.28209479177387814f;
constexpr float SH1 = 0.4886025119029199f;
constexpr float SH2_0 = 1.0925484305920792f, SH2_1 = -1.0925484305920792f, SH2_2 = 0.31539156525252005f,
                SH2_3 = -1.0925484305920792f, SH2_4 = 0.5462742152960396f;
constexpr float SH3_0 = -0.5900435899266435f, SH3_1 = 2.890611442640554f, SH3_2 = -0.4570457994644658f,
                SH3_3 = 0.3731763325901154f, SH3_4 = -0.4570457994644658f, SH3_5 = 1.445305721320277f,
                SH3_6 = -0.5900435899266435f;

constexpr int GB_THREADS = 128;
constexpr int GRAD_RAW_PARAMS = 32;  // == GDR_GRAD_RAW_PARAMS (include/gdr.h)

struct F3 {
    float x, y, z;
};
__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ F3 axpy(float s, F3 a, F3 b) { return F3{fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)}; }
__device__ __forceinline__ F3 scale3(float s, F3 a) { return F3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ void st3(float* p, F3 v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

// y = S x for the symmetric S stored as (xx, xy, xz, yy, yz, zz)
__device__ __forceinline__ F3 sym_mul(const float* S, F3 v) {
    return F3{fmaf(S[2], v.z, fmaf(S[1], v.y, S[0] * v.x)), fmaf(S[4], v.z, fmaf(S[3], v.y, S[1] * v.x)),
              fmaf(S[5], v.z, fmaf(S[4], v.y, S[2] * v.x))};
}

template <bool ACC>
__device__ __forceinline__ void put(float* p, float v) {
    if (ACC) v += *p;
    *p = v;
}
template <bool ACC>
__device__ __forceinline__ void put3(float* p, F3 v) {
    if (ACC) v = F3{v.x + p[0], v.y + p[1], v.z + p[2]};
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
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

// SH basis of degrees 2 and 3: b[k] and its gradient w.r.t. the (unit) direction, k = 4 .. 15.
__device__ __forceinline__ void sh_basis_high(int deg, float x, float y, float z, float* b, F3* gb) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = SH2_0 * xy;                    gb[4] = scale3(SH2_0, f3(y, x, 0.f));
    b[5] = SH2_1 * yz;                    gb[5] = scale3(SH2_1, f3(0.f, z, y));
    b[6] = SH2_2 * (2.f * zz - xx - yy);  gb[6] = scale3(SH2_2, f3(-2.f * x, -2.f * y, 4.f * z));
    b[7] = SH2_3 * xz;                    gb[7] = scale3(SH2_3, f3(z, 0.f, x));
    b[8] = SH2_4 * (xx - yy);             gb[8] = scale3(SH2_4, f3(2.f * x, -2.f * y, 0.f));
    if (deg < 3) return;
    b[9] = SH3_0 * y * (3.f * xx - yy);                gb[9] = scale3(SH3_0, f3(6.f * xy, 3.f * (xx - yy), 0.f));
    b[10] = SH3_1 * xy * z;                            gb[10] = scale3(SH3_1, f3(yz, xz, xy));
    b[11] = SH3_2 * y * (4.f * zz - xx - yy);          gb[11] = scale3(SH3_2, f3(-2.f * xy, 4.f * zz - xx - 3.f * yy, 8.f * yz));
    b[12] = SH3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
    gb[12] = scale3(SH3_3, f3(-6.f * xz, -6.f * yz, 6.f * zz - 3.f * xx - 3.f * yy));
    b[13] = SH3_4 * x * (4.f * zz - xx - yy);          gb[13] = scale3(SH3_4, f3(4.f * zz - 3.f * xx - yy, -2.f * xy, 8.f * xz));
    b[14] = SH3_5 * z * (xx - yy);                     gb[14] = scale3(SH3_5, f3(2.f * xz, -2.f * yz, xx - yy));
    b[15] = SH3_6 * x * (xx - 3.f * yy);               gb[15] = scale3(SH3_6, f3(3.f * (xx - yy), -6.f * xy, 0.f));
}

// Shared-memory slices of one virtual block of GB_THREADS Gaussians.  The inputs with a 12- / 24- / 48-byte row stride
// are staged with coalesced 128-bit accesses (stage_floats, common.cuh), and the outputs with such strides are written
// back into the SAME rows (a thread reads its row completely before it overwrites it) and leave coalesced too: with
// per-lane scalar accesses this kernel was bound by L1TEX wavefronts (~400 per warp), not by DRAM or issue.
struct Rows {
    float* mean;   // [T][3]  in: mean3D          out: dL/dmean3D
    float* scale;  // [T][3]  in: scale           out: dL/dscale
    float* cov;    // [T][6]  in: world covariance out: dL/dcov3D (precomputed-covariance inputs only)
    float* acc;    // [T][12] in: the screen-space sums
    float* sh;     // [T][3M] in: SH coefficients  out: dL/dsh
};

template <bool ACC>
__device__ __forceinline__ void gauss_backward_one(const GaussBackwardArgs& a, const Rows& rows, const int idx,
                                                   const bool visible, const float4 q_in, const unsigned clamp_in,
                                                   const float opacity_in) {
    const int vi = a.view;
    const int t = threadIdx.x;
    const int M = a.M;
    const bool raw = (a.grad_mask & GRAD_RAW_PARAMS) != 0;  // gradients w.r.t. logits / log-scales / raw quaternions
    const bool want_geo = a.dL_dmeans3D || a.dL_dcov3D || a.dL_dscales || a.dL_drotations;
    const bool want_sh = a.dL_dsh != nullptr && a.shs != nullptr;
    const bool want_scale_rot = a.scales != nullptr && (a.dL_dscales || a.dL_drotations);

    const float4* acc = reinterpret_cast<const float4*>(rows.acc + 12 * t);
    const float4 g_mean2D = acc[0];     // dL/d(mean2D) x, y, sum |x|, sum |y|
    const float4 g_conic_op = acc[1];   // dL/d(conic a, b, c), dL/d(opacity)
    const float4 g_rgb_depth = acc[2];  // dL/d(r, g, b), dL/d(depth)

    if (a.dL_dmeans2D) put4<ACC>(a.dL_dmeans2D + (size_t)idx * 4, g_mean2D);
    if (a.dL_dopacity) {
        float g = g_conic_op.w;
        if (raw) g = g * (1.f - opacity_in) * opacity_in;  // d sigmoid; the saved opacity is 0 for culled Gaussians
        put<ACC>(a.dL_dopacity + idx, g);
    }
    if (a.dL_dcolors) put3<ACC>(a.dL_dcolors + (size_t)idx * 3, f3(g_rgb_depth.x, g_rgb_depth.y, g_rgb_depth.z));
    if (!want_geo && !want_sh) return;

    const F3 mean = f3(rows.mean[3 * t], rows.mean[3 * t + 1], rows.mean[3 * t + 2]);
    float S[6];
    {
        const float2* c2 = reinterpret_cast<const float2*>(rows.cov + 6 * t);
        const float2 c0 = c2[0], c1 = c2[1], c2v = c2[2];
        S[0] = c0.x; S[1] = c0.y; S[2] = c1.x; S[3] = c1.y; S[4] = c2v.x; S[5] = c2v.y;
    }
    const F3 sc_in = want_scale_rot ? f3(rows.scale[3 * t], rows.scale[3 * t + 1], rows.scale[3 * t + 2]) : f3(0.f, 0.f, 0.f);
    float* const sh = rows.sh + (size_t)3 * M * t;  // this Gaussian's coefficients, then its dL/dsh row
    float* const o_mean = rows.mean + 3 * t;
    float* const o_scale = rows.scale + 3 * t;
    float* const o_cov = rows.cov + 6 * t;

    if (!visible) {  // the reference's kernels return early for radii <= 0 and leave the zero fill
        if (a.dL_dmeans3D) o_mean[0] = o_mean[1] = o_mean[2] = 0.f;
        if (a.dL_dcov3D)
            for (int k = 0; k < 6; k++) o_cov[k] = 0.f;
        if (a.dL_dsh)
            for (int k = 0; k < 3 * M; k++) sh[k] = 0.f;
        if (a.dL_dscales) o_scale[0] = o_scale[1] = o_scale[2] = 0.f;
        if (a.dL_drotations && !ACC) reinterpret_cast<float4*>(a.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }

    const float* vm = a.vw.view + (size_t)vi * a.vw.cam_stride;  // transposed storage: t = p * vm
    const float* pm = a.vw.proj + (size_t)vi * a.vw.cam_stride;
    const float tan_fovx = a.vw.tanx(vi), tan_fovy = a.vw.tany(vi);
    const float fy = a.H / (2.0f * tan_fovy), fx = a.W / (2.0f * tan_fovx);
    // rows of the world -> view rotation
    const F3 w0 = f3(__ldg(vm + 0), __ldg(vm + 4), __ldg(vm + 8));
    const F3 w1 = f3(__ldg(vm + 1), __ldg(vm + 5), __ldg(vm + 9));
    const F3 w2 = f3(__ldg(vm + 2), __ldg(vm + 6), __ldg(vm + 10));
    F3 g_mean;        // dL/d(mean3D)
    float G[6];       // dL/d(world covariance): (xx, xy, xz, yy, yz, zz), off-diagonals counted once for both entries

    // ---- conic -> screen covariance -> (world covariance, mean) ----
    {
        const float tz = dot(w2, mean) + __ldg(vm + 14);
        const float rz = 1.f / tz;
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float ux = (dot(w0, mean) + __ldg(vm + 12)) * rz, uy = (dot(w1, mean) + __ldg(vm + 13)) * rz;
        const bool clip_x = ux < -limx || ux > limx, clip_y = uy < -limy || uy > limy;
        const float tx = fminf(limx, fmaxf(-limx, ux)) * tz, ty = fminf(limy, fmaxf(-limy, uy)) * tz;
        // the two rows of M = J Wv:  m0 = j00 w0 + j02 w2,  m1 = j11 w1 + j12 w2
        const float j00 = fx * rz, j11 = fy * rz, j02 = -fx * tx * rz * rz, j12 = -fy * ty * rz * rz;
        const F3 m0 = axpy(j02, w2, scale3(j00, w0)), m1 = axpy(j12, w2, scale3(j11, w1));
        const F3 u0 = sym_mul(S, m0), u1 = sym_mul(S, m1);
        const float ca = dot(m0, u0) + 0.3f, cb = dot(m0, u1), cc = dot(m1, u1) + 0.3f;
        // conic = (cc, -cb, ca) / det: gradient w.r.t. (ca, cb, cc), with the reference's regularised denominator
        const float det = ca * cc - cb * cb;
        const float k = 1.0f / (det * det + 0.0000001f);
        const float gx = g_conic_op.x, gy = g_conic_op.y, gz = g_conic_op.z;
        float da = 0.f, db = 0.f, dc = 0.f;
        if (k != 0.f) {
            const float off = det - ca * cc;  // = -cb^2
            da = k * (-cc * cc * gx + 2.f * cb * cc * gy + off * gz);
            dc = k * (-ca * ca * gz + 2.f * ca * cb * gy + off * gx);
            db = k * 2.f * (cb * cc * gx - (det + 2.f * cb * cb) * gy + ca * cb * gz);
        }
        // (ca, cb, cc) are quadratic forms in m0, m1: G = da m0 m0^T + db sym(m0 m1^T) + dc m1 m1^T
        const F3 p0 = axpy(db, m1, scale3(da, m0));  // da m0 + db m1
        const F3 p1 = scale3(dc, m1);
        G[0] = fmaf(p0.x, m0.x, p1.x * m1.x);
        G[3] = fmaf(p0.y, m0.y, p1.y * m1.y);
        G[5] = fmaf(p0.z, m0.z, p1.z * m1.z);
        const F3 r0 = axpy(db, m1, scale3(2.f * da, m0));  // 2 da m0 + db m1
        const F3 r1 = axpy(db, m0, scale3(2.f * dc, m1));  // 2 dc m1 + db m0
        G[1] = fmaf(r0.x, m0.y, r1.x * m1.y);
        G[2] = fmaf(r0.x, m0.z, r1.x * m1.z);
        G[4] = fmaf(r0.y, m0.z, r1.y * m1.z);
        // ... and linear in S: dL/dm0 = 2 da u0 + db u1, dL/dm1 = 2 dc u1 + db u0
        const F3 gm0 = axpy(db, u1, scale3(2.f * da, u0)), gm1 = axpy(db, u0, scale3(2.f * dc, u1));
        const float gj00 = dot(w0, gm0), gj02 = dot(w2, gm0), gj11 = dot(w1, gm1), gj12 = dot(w2, gm1);
        const float rz2 = rz * rz, rz3 = rz2 * rz;
        const float gtx = clip_x ? 0.f : -fx * rz2 * gj02;
        const float gty = clip_y ? 0.f : -fy * rz2 * gj12;
        const float gtz = -fx * rz2 * gj00 - fy * rz2 * gj11 + 2.f * fx * tx * rz3 * gj02 + 2.f * fy * ty * rz3 * gj12;
        // t = Wv p + const: back through the rotation
        g_mean = f3(fmaf(w2.x, gtz, fmaf(w1.x, gty, w0.x * gtx)), fmaf(w2.y, gtz, fmaf(w1.y, gty, w0.y * gtx)),
                    fmaf(w2.z, gtz, fmaf(w1.z, gty, w0.z * gtx)));
    }

    // ---- mean2D (NDC) and depth -> mean3D ----
    {
        const float hx = fmaf(__ldg(pm + 8), mean.z, fmaf(__ldg(pm + 4), mean.y, __ldg(pm + 0) * mean.x)) + __ldg(pm + 12);
        const float hy = fmaf(__ldg(pm + 9), mean.z, fmaf(__ldg(pm + 5), mean.y, __ldg(pm + 1) * mean.x)) + __ldg(pm + 13);
        const float hw = fmaf(__ldg(pm + 11), mean.z, fmaf(__ldg(pm + 7), mean.y, __ldg(pm + 3) * mean.x)) + __ldg(pm + 15);
        const float rw = 1.0f / (hw + 0.0000001f);
        // ndc = h.xy rw:  d ndc_x / d p_i = rw (P_xi - P_wi ndc_x)
        const float s = (g_mean2D.x * hx + g_mean2D.y * hy) * rw;
        const float gd = g_rgb_depth.w;
        const float depth = dot(w2, mean) + __ldg(vm + 14);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float gi = rw * (g_mean2D.x * __ldg(pm + 4 * i) + g_mean2D.y * __ldg(pm + 4 * i + 1) - s * __ldg(pm + 4 * i + 3)) +
                             (__ldg(vm + 4 * i + 2) - __ldg(vm + 4 * i + 3) * depth) * gd;
            if (i == 0) g_mean.x += gi;
            if (i == 1) g_mean.y += gi;
            if (i == 2) g_mean.z += gi;
        }
    }

    // ---- colour -> SH coefficients and view direction ----
    if (a.shs != nullptr) {
        const float* cp = a.vw.campos + (size_t)vi * a.vw.cam_stride;
        const F3 d = f3(mean.x - __ldg(cp), mean.y - __ldg(cp + 1), mean.z - __ldg(cp + 2));
        const float rlen = 1.0f / sqrtf(dot(d, d));
        const float x = d.x * rlen, y = d.y * rlen, z = d.z * rlen;
        // the colour was clamped at 0 in the forward: no gradient through a clamped channel
        const F3 drgb = f3((clamp_in & 1u) ? 0.f : g_rgb_depth.x, (clamp_in & 2u) ? 0.f : g_rgb_depth.y,
                           (clamp_in & 4u) ? 0.f : g_rgb_depth.z);
        const bool dsh = a.dL_dsh != nullptr;  // the row of coefficients becomes the row of their gradients
        const int deg = a.sh_degree;
        const int used = (deg + 1) * (deg + 1);
        F3 gdir = f3(0.f, 0.f, 0.f);  // dL/d(unit direction) = sum_k (sh_k . drgb) grad b_k
        if (deg > 0) {
            gdir = f3(-SH1 * dot(f3(sh[9], sh[10], sh[11]), drgb), -SH1 * dot(f3(sh[3], sh[4], sh[5]), drgb),
                      SH1 * dot(f3(sh[6], sh[7], sh[8]), drgb));
            if (dsh) {
                st3(sh + 3, scale3(-SH1 * y, drgb));
                st3(sh + 6, scale3(SH1 * z, drgb));
                st3(sh + 9, scale3(-SH1 * x, drgb));
            }
            if (deg > 1) {
                float b[16];
                F3 gb[16];
                sh_basis_high(deg, x, y, z, b, gb);
                for (int k = 4; k < used; k++) {
                    gdir = axpy(dot(f3(sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]), drgb), gb[k], gdir);
                    if (dsh) st3(sh + 3 * k, scale3(b[k], drgb));
                }
            }
            // through dir = d / |d|
            const float along = x * gdir.x + y * gdir.y + z * gdir.z;
            g_mean.x += (gdir.x - x * along) * rlen;
            g_mean.y += (gdir.y - y * along) * rlen;
            g_mean.z += (gdir.z - z * along) * rlen;
        }
        if (dsh) {
            st3(sh, scale3(SH0, drgb));
            for (int k = used; k < M; k++) st3(sh + 3 * k, f3(0.f, 0.f, 0.f));
        }
    } else if (a.dL_dsh) {
        for (int k = 0; k < 3 * M; k++) sh[k] = 0.f;
    }

    if (a.dL_dmeans3D) st3(o_mean, g_mean);
    if (a.dL_dcov3D)
        for (int k = 0; k < 6; k++) o_cov[k] = G[k];

    // ---- world covariance -> scale, rotation:  S = N N^T,  N = R diag(s) ----
    if (want_scale_rot) {
        float4 q = q_in;
        F3 sc = sc_in;
        float qn = 1.f;
        if (raw) {
            qn = fmaxf(quat_norm(q), 1e-12f);
            q = act_rotation(q);
            sc = f3(act_scale(sc.x), act_scale(sc.y), act_scale(sc.z));
        }
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        // rows of R
        const F3 R0 = f3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y));
        const F3 R1 = f3(2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x));
        const F3 R2 = f3(2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
        const F3 s = scale3(a.scale_modifier, sc);
        // dL/dN = 2 Gs N with Gs the symmetric matrix whose off-diagonals are HALF the stored (counted-once) ones:
        // E = Gs R (row i of E = sum_k Gs_ik R_k), then dL/dN_ij = 2 E_ij s_j
        const float h1 = 0.5f * G[1], h2 = 0.5f * G[2], h4 = 0.5f * G[4];
        const F3 E0 = axpy(h2, R2, axpy(h1, R1, scale3(G[0], R0)));
        const F3 E1 = axpy(h4, R2, axpy(G[3], R1, scale3(h1, R0)));
        const F3 E2 = axpy(G[5], R2, axpy(h4, R1, scale3(h2, R0)));
        if (a.dL_dscales) {
            // dL/ds_j = sum_i dL/dN_ij R_ij (the scale modifier multiplies in as the reference does: through s only
            // via dL/dN); d exp: times the activated scale when the inputs are log-scales
            F3 gs = f3(2.f * s.x * (E0.x * R0.x + E1.x * R1.x + E2.x * R2.x),
                       2.f * s.y * (E0.y * R0.y + E1.y * R1.y + E2.y * R2.y),
                       2.f * s.z * (E0.z * R0.z + E1.z * R1.z + E2.z * R2.z));
            if (raw) gs = f3(gs.x * sc.x, gs.y * sc.y, gs.z * sc.z);
            st3(o_scale, gs);
        }
        if (a.dL_drotations) {
            // D = dL/dR = dL/dN diag(s):  D_ij = 2 E_ij s_j^2
            const F3 t = f3(2.f * s.x * s.x, 2.f * s.y * s.y, 2.f * s.z * s.z);
            const float D00 = E0.x * t.x, D01 = E0.y * t.y, D02 = E0.z * t.z;
            const float D10 = E1.x * t.x, D11 = E1.y * t.y, D12 = E1.z * t.z;
            const float D20 = E2.x * t.x, D21 = E2.y * t.y, D22 = E2.z * t.z;
            float4 dq;
            dq.x = 2.f * (z * (D10 - D01) + y * (D02 - D20) + x * (D21 - D12));
            dq.y = 2.f * (y * (D01 + D10) + z * (D02 + D20) + r * (D21 - D12)) - 4.f * x * (D11 + D22);
            dq.z = 2.f * (x * (D01 + D10) + r * (D02 - D20) + z * (D12 + D21)) - 4.f * y * (D00 + D22);
            dq.w = 2.f * (r * (D10 - D01) + x * (D02 + D20) + y * (D12 + D21)) - 4.f * z * (D00 + D11);
            if (raw) {  // through q / ||q||: (g - q^ (q^ . g)) / ||q||
                const float along = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w;
                const float inv = 1.f / qn;
                dq = make_float4((dq.x - q.x * along) * inv, (dq.y - q.y * along) * inv, (dq.z - q.z * along) * inv,
                                 (dq.w - q.w * along) * inv);
            }
            put4<ACC>(a.dL_drotations + (size_t)idx * 4, dq);
        }
    } else {
        if (a.dL_dscales) o_scale[0] = o_scale[1] = o_scale[2] = 0.f;
        if (a.dL_drotations && !ACC) reinterpret_cast<float4*>(a.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ACC = false stores every requested output row; ACC = true adds to it (views 1.. of a batch launch in
// stream order after view 0, so the per-Gaussian gradients are summed over views deterministically).
#ifndef GDR_GB_MINB
#define GDR_GB_MINB 1
#endif
template <bool ACC>
__global__ void __launch_bounds__(GB_THREADS, GDR_GB_MINB) gauss_backward_kernel(const GaussBackwardArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    Rows rows;
    rows.mean = smem;
    rows.scale = rows.mean + 3 * GB_THREADS;
    rows.cov = rows.scale + 3 * GB_THREADS;
    rows.acc = rows.cov + 6 * GB_THREADS;
    rows.sh = rows.acc + 12 * GB_THREADS;
    const int vi = a.view, M = a.M;
    const GeomState geom = a.geom.at(vi, a.vw.geom_stride);
    const bool raw = (a.grad_mask & GRAD_RAW_PARAMS) != 0;
    const bool want_geo = a.dL_dmeans3D || a.dL_dcov3D || a.dL_dscales || a.dL_drotations;
    const bool want_sh = a.dL_dsh != nullptr && a.shs != nullptr;
    const bool want_scale_rot = a.scales != nullptr && (a.dL_dscales || a.dL_drotations);
    const bool need_rows = want_geo || want_sh;
    const float* cov_src = a.cov3D_precomp ? a.cov3D_precomp : geom.cov3D;
    // Full, 16-byte-aligned blocks move with the bulk-copy engine: ONE thread issues every slice of the block
    // (cp.async.bulk + mbarrier), so all of them are in flight together and no thread spends instructions or
    // registers on the copies; the strided outputs leave the same way (bulk store, or bulk add for ACC).
    const bool aligned =
        ((((uintptr_t)a.means3D) | ((uintptr_t)cov_src) | ((uintptr_t)a.scales) | ((uintptr_t)a.shs) | ((uintptr_t)a.accum) |
          ((uintptr_t)a.dL_dmeans3D) | ((uintptr_t)a.dL_dscales) | ((uintptr_t)a.dL_dcov3D) | ((uintptr_t)a.dL_dsh)) & 15u) == 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    const int n_vblocks = (a.P + GB_THREADS - 1) / GB_THREADS;
    for (int vb = blockIdx.x; vb < n_vblocks; vb += gridDim.x) {  // virtual blocks: one balanced wave
        if (vb != (int)blockIdx.x) __syncthreads();               // the rows are reused
        const int first = vb * GB_THREADS;
        const int n_items = min(GB_THREADS, a.P - first);
        const int idx = first + (int)threadIdx.x;
        const bool in_range = (int)threadIdx.x < n_items;
        const bool bulk = aligned && n_items == GB_THREADS;
        // ---- every per-Gaussian input that the backward blend does not write, in ONE round trip: as a programmatic
        // dependent of that kernel this launch fetches them while the blend's last CTAs are still running and only
        // then waits (pdl_wait) for the accumulators ----
        bool visible = false;
        float4 q_in = make_float4(1.f, 0.f, 0.f, 0.f);
        unsigned clamp_in = 0u;
        float opacity_in = 0.f;
        constexpr uint32_t ROW = sizeof(float) * GB_THREADS;  // bytes of one float per Gaussian
        if (need_rows) {
            if (bulk) {
                if (threadIdx.x == 0) {
                    const uint32_t bytes = ROW * (3 + 6 + 12 + (want_scale_rot ? 3 : 0) + (a.shs != nullptr ? 3 * M : 0));
                    mbar_expect_tx(&bar, bytes);
                    bulk_g2s(rows.mean, a.means3D + (size_t)first * 3, 3 * ROW, &bar);
                    bulk_g2s(rows.cov, cov_src + (size_t)first * 6, 6 * ROW, &bar);
                    if (want_scale_rot) bulk_g2s(rows.scale, a.scales + (size_t)first * 3, 3 * ROW, &bar);
                    if (a.shs != nullptr) bulk_g2s(rows.sh, a.shs + (size_t)first * 3 * M, 3 * M * ROW, &bar);
                }
            } else {
                stage_floats(rows.mean, a.means3D + (size_t)first * 3, n_items * 3);
                stage_floats(rows.cov, cov_src + (size_t)first * 6, n_items * 6);
                if (want_scale_rot) stage_floats(rows.scale, a.scales + (size_t)first * 3, n_items * 3);
                if (a.shs != nullptr) stage_floats(rows.sh, a.shs + (size_t)first * 3 * M, n_items * 3 * M);
            }
            if (in_range) {
                visible = a.radii[(size_t)vi * a.P + idx] > 0;
                if (want_scale_rot) q_in = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
                if (a.shs != nullptr) clamp_in = (unsigned)geom.clamped[idx];
            }
        }
        if (in_range && a.dL_dopacity && raw) opacity_in = geom.splat[idx].q1.w;
        pdl_wait();  // the backward blend (or the previous view's launch of this kernel) has completed
        const float* acc_src = a.accum + ((size_t)vi * a.P + first) * 12;
        if (need_rows && bulk) {
            if (threadIdx.x == 0) bulk_g2s(rows.acc, acc_src, 12 * ROW, &bar);
            mbar_wait(&bar, parity);  // every slice has landed (and is visible to all threads that waited)
            parity ^= 1u;
        } else {
            stage_floats(rows.acc, acc_src, n_items * 12);
            __syncthreads();
        }
        if (in_range) gauss_backward_one<ACC>(a, rows, idx, visible, q_in, clamp_in, opacity_in);
        // GDR_GRAD_SCRATCH_CLEAN: the accumulator rows go back to zero once consumed, so the caller's buffer needs no
        // fill before its next backward (a 9.6 MB memset per view at 200k Gaussians, and a launch boundary).  Each
        // thread has read its own row by now; full blocks leave as one bulk store of the zeroed rows, others directly.
        float* acc_dst = a.accum + ((size_t)vi * a.P + first) * 12;
        const bool rezero_bulk = a.rezero && need_rows && bulk;
        if (a.rezero && in_range) {
            float4* z = reinterpret_cast<float4*>(rezero_bulk ? rows.acc + 12 * threadIdx.x : acc_dst + 12 * threadIdx.x);
            z[0] = z[1] = z[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- the outputs with a 12- / 24- / 48-byte row stride leave coalesced ----
        if (need_rows) {
            if (bulk) {
                fence_async_smem();  // this thread's rows -> visible to the bulk-copy engine
                __syncthreads();
                if (threadIdx.x == 0) {
                    auto out = [&](float* g, const float* srow, uint32_t bytes) {
                        if (ACC) bulk_s2g_add_f32(g, srow, bytes); else bulk_s2g(g, srow, bytes);
                    };
                    if (a.dL_dmeans3D) out(a.dL_dmeans3D + (size_t)first * 3, rows.mean, 3 * ROW);
                    if (a.dL_dscales) out(a.dL_dscales + (size_t)first * 3, rows.scale, 3 * ROW);
                    if (a.dL_dcov3D) out(a.dL_dcov3D + (size_t)first * 6, rows.cov, 6 * ROW);
                    if (a.dL_dsh) out(a.dL_dsh + (size_t)first * 3 * M, rows.sh, 3 * M * ROW);
                    if (rezero_bulk) bulk_s2g(acc_dst, rows.acc, 12 * ROW);
                    bulk_commit();
                    bulk_wait_read();  // the rows may be overwritten (next virtual block) once they have been read
                }
            } else {
                __syncthreads();
                if (a.dL_dmeans3D) unstage_floats<ACC>(a.dL_dmeans3D + (size_t)first * 3, rows.mean, n_items * 3);
                if (a.dL_dscales) unstage_floats<ACC>(a.dL_dscales + (size_t)first * 3, rows.scale, n_items * 3);
                if (a.dL_dcov3D) unstage_floats<ACC>(a.dL_dcov3D + (size_t)first * 6, rows.cov, n_items * 6);
                if (a.dL_dsh) unstage_floats<ACC>(a.dL_dsh + (size_t)first * 3 * M, rows.sh, n_items * 3 * M);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait_all();  // this CTA's bulk stores are performed before it exits
}

}  // namespace

cudaError_t launch_gauss_backward(const GaussBackwardArgs& a, cudaStream_t s) {
    if (a.P <= 0) return cudaSuccess;
    const size_t smem = sizeof(float) * GB_THREADS * (24 + 3 * (size_t)max(a.M, 0));
    static int per_sm_cached[2][17] = {};
    const int mi = min(max(a.M, 0), 16);
    int& per_sm = per_sm_cached[a.accumulate ? 1 : 0][mi];
    if (per_sm == 0) {
        int n = 1;
        cudaError_t e = a.accumulate ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gauss_backward_kernel<true>, GB_THREADS, smem)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gauss_backward_kernel<false>, GB_THREADS, smem);
        per_sm = (e == cudaSuccess && n >= 1) ? n : 1;
    }
    const int grid = min((a.P + GB_THREADS - 1) / GB_THREADS, sm_count() * per_sm);
    return a.accumulate ? launch_dependent(gauss_backward_kernel<true>, dim3(grid), dim3(GB_THREADS), smem, s, a)
                        : launch_dependent(gauss_backward_kernel<false>, dim3(grid), dim3(GB_THREADS), smem, s, a);
}

}  // namespace gdr
