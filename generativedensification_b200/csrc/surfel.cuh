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
//   r4 = {view-space normal x, y, z (flipped towards the camera), conservative screen-space reach in pixels}
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

}  // namespace gdr
