// Introspection kernels for the parity tests: unpack the opaque state into the
// reference's field layout (GeometryState / BinningState / ImageState,
// RAST/cuda_rasterizer/rasterizer_impl.h:33-66) so tests can compare stage by
// stage.  Not on the hot path.
#include "kernels.h"

namespace gdr {
namespace {

__global__ void unpack_geom_kernel(int P, GeomState g, float* means2D, float* depths, float* conic_opacity, float* rgb,
                                   float* cov3D, uint32_t* tiles_touched, uint8_t* clamped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const Splat s = g.splat[i];
    if (means2D) {
        means2D[2 * i] = s.q0.x;
        means2D[2 * i + 1] = s.q0.y;
    }
    if (depths) depths[i] = s.q2.w;
    if (conic_opacity) reinterpret_cast<float4*>(conic_opacity)[i] = s.q1;
    if (rgb) {
        rgb[3 * i] = s.q2.x;
        rgb[3 * i + 1] = s.q2.y;
        rgb[3 * i + 2] = s.q2.z;
    }
    if (cov3D)
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = g.cov3D[6 * (size_t)i + k];
    if (tiles_touched) tiles_touched[i] = g.tiles_touched[i];
    if (clamped) {
        const unsigned c = g.clamped[i];
        clamped[3 * i] = c & 1u;
        clamped[3 * i + 1] = (c >> 1) & 1u;
        clamped[3 * i + 2] = (c >> 2) & 1u;
    }
}

// The reference's `ranges` (rasterizer_impl.cu:116-138): tiles in tile order, [begin, end) into its globally sorted
// point_list; empty tiles stay (0, 0) from its memset (:311).  Our stream keeps the tiles in completion order, so the
// reference layout is an exclusive scan of the per-tile lengths (one CTA: introspection only).
__global__ void unpack_ranges_kernel(int T, const uint2* __restrict__ tile_range, uint32_t* __restrict__ ranges) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += blockDim.x) {
        const int t = base + threadIdx.x;
        const uint32_t n = t < T ? tile_range[t].y - tile_range[t].x : 0u;
        const int incl = warp_incl_scan((int)n);
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = (uint32_t)incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t ws = threadIdx.x < (blockDim.x >> 5) ? s_warp[threadIdx.x] : 0u;
            s_warp[threadIdx.x] = (uint32_t)warp_incl_scan((int)ws) - ws;
        }
        __syncthreads();
        const uint32_t begin = s_carry + s_warp[threadIdx.x >> 5] + (uint32_t)incl - n;
        if (t < T) {
            ranges[2 * t] = n ? begin : 0u;
            ranges[2 * t + 1] = n ? begin + n : 0u;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = begin + n;
        __syncthreads();
    }
}

// point_list in the reference's layout: tile t's sorted ids at ranges[t].
__global__ void unpack_list_kernel(int64_t capacity, const uint2* __restrict__ tile_range,
                                   const uint32_t* __restrict__ ranges, const Splat* __restrict__ stream,
                                   uint32_t* __restrict__ point_list) {
    const int t = blockIdx.x;
    const uint2 r = tile_range[t];
    const uint32_t dst = ranges[2 * t];
    for (uint32_t i = threadIdx.x; i < r.y - r.x; i += blockDim.x)
        if ((int64_t)dst + i < capacity)
            point_list[dst + i] = __float_as_uint(stream[r.x + i].q0.w) & STREAM_ID_MASK;
}

}  // namespace

cudaError_t launch_unpack_geom(int P, GeomState geom, float* means2D, float* depths, float* conic_opacity, float* rgb,
                               float* cov3D, uint32_t* tiles_touched, uint8_t* clamped, cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    unpack_geom_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, geom, means2D, depths, conic_opacity, rgb, cov3D,
                                                       tiles_touched, clamped);
    return cudaGetLastError();
}

cudaError_t launch_unpack_bins(int W, int H, ImageState img, const Splat* stream, int64_t capacity,
                               uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib, cudaStream_t s) {
    const int T = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    if (ranges) unpack_ranges_kernel<<<1, 1024, 0, s>>>(T, img.tile_range, ranges);
    if (point_list && capacity > 0) {
        if (!ranges) return cudaErrorInvalidValue;
        unpack_list_kernel<<<T, 128, 0, s>>>(capacity, img.tile_range, ranges, stream, point_list);
    }
    if (n_contrib) {
        cudaError_t e = cudaMemcpyAsync(n_contrib, img.n_contrib, sizeof(uint32_t) * (size_t)W * H,
                                        cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace gdr
