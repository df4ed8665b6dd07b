// Records and constants of the 2D Gaussian-surfel (2DGS) raster path.
//
// The reference calls a `diff_surfel_rasterization` extension from
// lightning/renderer_2dgs.py:224-233 whose source is not in the reference tree
// (PARITY UNPINNED).  The arithmetic here follows the published
// 2DGS algorithm (Huang et al. 2024: splat-to-screen homography, ray-splat
// intersection by two homogeneous planes, sqrt(2)/2 px object-space low-pass,
// depth-distortion accumulation) on top of this library's own binning.
#pragma once
#include "common.cuh"

namespace gdr {

// 80-byte record per surfel (and per sorted tile instance): five aligned 128-bit words, so a tile's
// list is staged with cp.async.bulk exactly like the 48-byte Splat stream.
//   r0 = {bbox centre x, bbox centre y, opacity, surfel index bits}
//   r1 = {Tu.x, Tu.y, Tu.z, red}      Tu, Tv, Tw: rows of the splat -> pixel homography acting on (u, v, 1)
//   r2 = {Tv.x, Tv.y, Tv.z, green}
//   r3 = {Tw.x, Tw.y, Tw.z, blue}
//   r4 = {view-space normal x, y, z (flipped towards the camera), w}: w = conservative screen-space reach in pixels in the
//        per-surfel state; in the sorted stream copies tile_sort replaces it by the 8 block bits of surfel_region_mask8
struct __align__(16) Surfel {
    float4 r0, r1, r2, r3, r4;
};
static_assert(sizeof(Surfel) == 80, "Surfel must be 80 bytes");

constexpr float SURFEL_FILTER_SIZE = 0.707106f;    // sqrt(2)/2 px low-pass (paper eq. 11)
constexpr float SURFEL_FILTER_INV_SQUARE = 2.0f;
constexpr float SURFEL_CUTOFF = 3.0f;              // bounding box at 3 sigma
constexpr float SURFEL_NEAR = 0.2f;
constexpr float SURFEL_FAR = 100.0f;
constexpr int SURFEL_ACC = 20;  // floats per surfel in the backward scratch:
// [0..2] dL/dTu  [3..5] dL/dTv  [6..8] dL/dTw  [9..10] dL/d(bbox centre)  [11] dL/dopacity  [12..14] dL/drgb
// [15..17] dL/dnormal  [18..19] sum over pixels of |dL/dTu.z|, |dL/dTv.z| (densification statistic)

// allmap channels (lightning/renderer_2dgs.py:241-257)
constexpr int AM_DEPTH = 0, AM_ALPHA = 1, AM_NORMAL = 2, AM_MIDDEPTH = 5, AM_DISTORTION = 6, AM_CHANNELS = 7;

// Which of a 16x16 tile's eight 8x4 pixel blocks (bit b: x half = b & 1, row quarter = b >> 1 -- the blocks the blend
// kernels' warps own) a surfel can reach with alpha >= 1/255.  tile_sort evaluates it once per (surfel, tile) instance
// and parks the bits in the stream copy's r4.w; a warp then walks only the records whose bit is set.
//
// alpha >= 1/255 needs min(rho3d, rho2d) <= tau = 2 ln(255 opacity):
//   * rho2d <= tau is the low-pass disc of radius sqrt(tau / 2) around the bounding-box centre (r0.xy);
//   * rho3d <= tau: with k = px Tw - Tu, l = py Tw - Tv and p = k x l the blend evaluates rho3d = (p.x^2 + p.y^2) / p.z^2,
//     and p is LINEAR in (px, py, 1):  p = px (Tv x Tw) + py (Tw x Tu) + Tu x Tv.  So the region is the conic
//     Q(px, py) = p.x^2 + p.y^2 - tau p.z^2 <= 0.  In coordinates relative to the projected splat centre
//     pc = (Tu.z, Tv.z) / Tw.z (where rho3d = 0: pc lies inside every level set) the third components of Tu, Tv vanish
//     and Q = Tw.z^2 |(x Tv'.y - y Tu'.y, y Tu'.x - x Tv'.x)|^2 - tau (x c1z + y c2z + D)^2 with small, well-scaled terms.
//     A block [X0, X1] x [Y0, Y1] meets the region iff pc lies in it or Q <= 0 somewhere on its boundary (the region is
//     convex and contains pc); on an edge Q is a 1-D quadratic whose minimum is at the clamped stationary point.
// Conservative by construction: a block is only dropped when the computed minimum of Q exceeds a bound of its own
// rounding error (the same polynomial evaluated on absolute values, times 64 ulp), and every case the argument does not
// cover -- a conic that is not an ellipse (the splat's plane passes near the eye), NaNs, a vanishing Tw.z -- keeps all
// blocks.  Skipping a record a block cannot reach changes no pixel: all its pairs fail the alpha test.
__device__ __forceinline__ unsigned surfel_region_mask8(float4 r0, float4 r1, float4 r2, float4 r3, float tx0, float ty0) {
    const float o = r0.z;
    if (!(o * 255.0f > 1.0f)) return o == o ? 0u : 0xffu;  // too faint to ever contribute (NaN: keep everything)
    const float tau = 2.0f * logf(o * 255.0f) + 1e-3f;
    // block bounds in absolute pixels
    unsigned disc = 0, ell = 0;
    const float r2sq = 0.5f * tau + 1e-3f;  // rho2d = 2 |d|^2 <= tau
    const float twz = r3.z;
    bool conic_ok = fabsf(twz) > 1e-20f;
    const float iw = conic_ok ? 1.0f / twz : 0.f;
    const float pcx = r1.z * iw, pcy = r2.z * iw;
    const float ux = r1.x - pcx * r3.x, uy = r1.y - pcx * r3.y;  // Tu' (third component 0)
    const float vx = r2.x - pcy * r3.x, vy = r2.y - pcy * r3.y;  // Tv'
    // Tu', Tv' are differences: give up where they cancelled most of their operands (their rounding error is then no
    // longer small against them; everywhere else it moves the boundary by far less than the 0.05 px the blocks are
    // grown by below)
    conic_ok = conic_ok && fabsf(ux) + fabsf(uy) > 1e-3f * (fabsf(r1.x) + fabsf(r1.y) + fabsf(pcx) * (fabsf(r3.x) + fabsf(r3.y))) &&
               fabsf(vx) + fabsf(vy) > 1e-3f * (fabsf(r2.x) + fabsf(r2.y) + fabsf(pcy) * (fabsf(r3.x) + fabsf(r3.y)));
    const float c1z = vx * r3.y - vy * r3.x, c2z = r3.x * uy - r3.y * ux, D = ux * vy - uy * vx;
    const float w2 = twz * twz;
    // Q = a x^2 + 2 b x y + c y^2 + 2 d x + 2 e y + f, and the same with every term's absolute value (error bound)
    const float a1 = w2 * (vx * vx + vy * vy), a2 = tau * c1z * c1z;
    const float b1 = -w2 * (vy * uy + ux * vx), b2 = tau * c1z * c2z;
    const float c1 = w2 * (ux * ux + uy * uy), c2 = tau * c2z * c2z;
    const float a = a1 - a2, b = b1 - b2, c = c1 - c2;
    const float d = -tau * c1z * D, e = -tau * c2z * D, f = -tau * D * D;
    const float aa = a1 + a2, ab = fabsf(b1) + fabsf(b2), ac = c1 + c2, ad = fabsf(d), ae = fabsf(e), af = fabsf(f);
    // an ellipse with margin: a, c clearly positive and the discriminant clearly positive
    conic_ok = conic_ok && a > 1e-3f * aa && c > 1e-3f * ac && (a * c - b * b) > 1e-3f * (aa * ac);
    const float ia = conic_ok ? 1.0f / a : 0.f, ic = conic_ok ? 1.0f / c : 0.f;
#pragma unroll
    for (int blk = 0; blk < 8; blk++) {
        const float bx0 = tx0 + (float)((blk & 1) * 8), by0 = ty0 + (float)((blk >> 1) * 4);
        const float bx1 = bx0 + 7.f, by1 = by0 + 3.f;
        // low-pass disc around the bounding-box centre
        const float ddx = r0.x - fminf(fmaxf(r0.x, bx0), bx1), ddy = r0.y - fminf(fmaxf(r0.y, by0), by1);
        if (!(ddx * ddx + ddy * ddy > r2sq)) disc |= 1u << blk;
        // the conic, in coordinates relative to pc
        const float X0 = bx0 - 0.05f - pcx, X1 = bx1 + 0.05f - pcx, Y0 = by0 - 0.05f - pcy, Y1 = by1 + 0.05f - pcy;
        bool hit = !conic_ok || (X0 <= 0.f && X1 >= 0.f && Y0 <= 0.f && Y1 >= 0.f);
        if (!hit) {
            const float Xm = fmaxf(fabsf(X0), fabsf(X1)), Ym = fmaxf(fabsf(Y0), fabsf(Y1));
            const float bound = 64.f * 1.1920929e-7f *
                                (aa * Xm * Xm + 2.f * ab * Xm * Ym + ac * Ym * Ym + 2.f * ad * Xm + 2.f * ae * Ym + af);
            auto on_vertical = [&](float X) {  // min over y in [Y0, Y1] of Q(X, y)
                const float lin = b * X + e;
                const float y = fminf(Y1, fmaxf(Y0, -lin * ic));
                return c * y * y + 2.f * lin * y + (a * X * X + 2.f * d * X + f);
            };
            auto on_horizontal = [&](float Y) {
                const float lin = b * Y + d;
                const float x = fminf(X1, fmaxf(X0, -lin * ia));
                return a * x * x + 2.f * lin * x + (c * Y * Y + 2.f * e * Y + f);
            };
            const float qmin = fminf(fminf(on_vertical(X0), on_vertical(X1)), fminf(on_horizontal(Y0), on_horizontal(Y1)));
            hit = !(qmin > bound);  // NaN: keep
        }
        if (hit) ell |= 1u << blk;
    }
    return disc | ell;
}

}  // namespace gdr
