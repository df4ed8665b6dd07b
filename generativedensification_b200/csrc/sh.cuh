// Spherical-harmonics colour (degree 0-3) and its backward, used by the surfel path.
// Same basis, +0.5 offset and clamp-at-zero as the reference's computeColorFromSH
// (RAST/cuda_rasterizer/forward.cu:20-71, backward.cu:20-139).  The 3DGS kernels keep their own
// copy in project.cu / gauss_bwd.cu whose rounding sequence is pinned against the reference build.
#pragma once
#include "common.cuh"

namespace gdr {
namespace sh {

__device__ constexpr float C0 = 0.28209479177387814f;
__device__ constexpr float C1 = 0.4886025119029199f;
__device__ constexpr float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                    -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                    0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                    -0.5900435899266435f};

// sh points at this Gaussian's coefficients, element (k, ch) at sh[k * 3 + ch].
__device__ __forceinline__ float3 eval(int deg, float3 pos, float3 campos, const float* __restrict__ sh,
                                       unsigned& clamp_bits) {
    float3 d = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
    const float inv = 1.0f / sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    const float x = d.x * inv, y = d.y * inv, z = d.z * inv;
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float v = C0 * sh[ch];
        if (deg > 0) {
            v = v - C1 * y * sh[3 + ch] + C1 * z * sh[6 + ch] - C1 * x * sh[9 + ch];
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                v = v + C2[0] * xy * sh[12 + ch] + C2[1] * yz * sh[15 + ch] + C2[2] * (2.0f * zz - xx - yy) * sh[18 + ch] +
                    C2[3] * xz * sh[21 + ch] + C2[4] * (xx - yy) * sh[24 + ch];
                if (deg > 2) {
                    v = v + C3[0] * y * (3.0f * xx - yy) * sh[27 + ch] + C3[1] * xy * z * sh[30 + ch] +
                        C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + ch] +
                        C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + ch] +
                        C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + ch] + C3[5] * z * (xx - yy) * sh[42 + ch] +
                        C3[6] * x * (xx - 3.0f * yy) * sh[45 + ch];
                }
            }
        }
        v += 0.5f;
        if (v < 0.f) clamp_bits |= 1u << ch;
        res[ch] = fmaxf(v, 0.f);
    }
    return make_float3(res[0], res[1], res[2]);
}

// dL/dsh (written to dsh[k * 3 + ch], k < (deg+1)^2) and the part of dL/dmean that flows through the
// view direction.  dL_drgb must already be zero in clamped channels.
__device__ __forceinline__ float3 backward(int deg, float3 pos, float3 campos, const float* __restrict__ shc,
                                           float3 dL_drgb, float* __restrict__ dsh) {
    const float3 d0 = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
    const float sum2 = d0.x * d0.x + d0.y * d0.y + d0.z * d0.z;
    const float inv = 1.0f / sqrtf(sum2);
    const float x = d0.x * inv, y = d0.y * inv, z = d0.z * inv;
    const float g[3] = {dL_drgb.x, dL_drgb.y, dL_drgb.z};
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;  // dL/d(dir)
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float* s = shc + ch;  // s[3 * k]
        float* o = dsh + ch;
        const float gc = g[ch];
        float rx = 0.f, ry = 0.f, rz = 0.f;  // d rgb[ch] / d(dir)
        o[0] = C0 * gc;
        if (deg > 0) {
            o[3] = -C1 * y * gc;
            o[6] = C1 * z * gc;
            o[9] = -C1 * x * gc;
            rx = -C1 * s[9];
            ry = -C1 * s[3];
            rz = C1 * s[6];
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                o[12] = C2[0] * xy * gc;
                o[15] = C2[1] * yz * gc;
                o[18] = C2[2] * (2.f * zz - xx - yy) * gc;
                o[21] = C2[3] * xz * gc;
                o[24] = C2[4] * (xx - yy) * gc;
                rx += C2[0] * y * s[12] + C2[2] * 2.f * -x * s[18] + C2[3] * z * s[21] + C2[4] * 2.f * x * s[24];
                ry += C2[0] * x * s[12] + C2[1] * z * s[15] + C2[2] * 2.f * -y * s[18] + C2[4] * 2.f * -y * s[24];
                rz += C2[1] * y * s[15] + C2[2] * 4.f * z * s[18] + C2[3] * x * s[21];
                if (deg > 2) {
                    o[27] = C3[0] * y * (3.f * xx - yy) * gc;
                    o[30] = C3[1] * xy * z * gc;
                    o[33] = C3[2] * y * (4.f * zz - xx - yy) * gc;
                    o[36] = C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * gc;
                    o[39] = C3[4] * x * (4.f * zz - xx - yy) * gc;
                    o[42] = C3[5] * z * (xx - yy) * gc;
                    o[45] = C3[6] * x * (xx - 3.f * yy) * gc;
                    rx += C3[0] * s[27] * 6.f * xy + C3[1] * s[30] * yz + C3[2] * s[33] * -2.f * xy +
                          C3[3] * s[36] * -6.f * xz + C3[4] * s[39] * (-3.f * xx + 4.f * zz - yy) +
                          C3[5] * s[42] * 2.f * xz + C3[6] * s[45] * 3.f * (xx - yy);
                    ry += C3[0] * s[27] * 3.f * (xx - yy) + C3[1] * s[30] * xz +
                          C3[2] * s[33] * (-3.f * yy + 4.f * zz - xx) + C3[3] * s[36] * -6.f * yz +
                          C3[4] * s[39] * -2.f * xy + C3[5] * s[42] * -2.f * yz + C3[6] * s[45] * -6.f * xy;
                    rz += C3[1] * s[30] * xy + C3[2] * s[33] * 8.f * yz + C3[3] * s[36] * 3.f * (2.f * zz - xx - yy) +
                          C3[4] * s[39] * 8.f * xz + C3[5] * s[42] * (xx - yy);
                }
            }
        }
        ddx = fmaf(rx, gc, ddx);
        ddy = fmaf(ry, gc, ddy);
        ddz = fmaf(rz, gc, ddz);
    }
    // through dir = d0 / |d0|
    const float inv3 = inv * inv * inv;
    return make_float3(((sum2 - d0.x * d0.x) * ddx - d0.y * d0.x * ddy - d0.z * d0.x * ddz) * inv3,
                       (-d0.x * d0.y * ddx + (sum2 - d0.y * d0.y) * ddy - d0.z * d0.y * ddz) * inv3,
                       (-d0.x * d0.z * ddx - d0.y * d0.z * ddy + (sum2 - d0.z * d0.z) * ddz) * inv3);
}

}  // namespace sh
}  // namespace gdr
