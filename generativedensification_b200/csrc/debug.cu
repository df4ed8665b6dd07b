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

__global__ void unpack_list_kernel(int64_t n, const Splat* stream, uint32_t* point_list) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) point_list[i] = __float_as_uint(stream[i].q0.w) & STREAM_ID_MASK;
}

__global__ void unpack_ranges_kernel(int T, int64_t capacity, const uint32_t* tile_offsets, uint32_t* ranges) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const uint32_t b = (uint32_t)min((int64_t)tile_offsets[t], capacity);
    const uint32_t e = (uint32_t)min((int64_t)tile_offsets[t + 1], capacity);
    // the reference leaves empty tiles at (0, 0) (memset at rasterizer_impl.cu:311)
    ranges[2 * t] = e > b ? b : 0;
    ranges[2 * t + 1] = e > b ? e : 0;
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
    if (point_list && capacity > 0)
        unpack_list_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, s>>>(capacity, stream, point_list);
    if (ranges) unpack_ranges_kernel<<<(T + 255) / 256, 256, 0, s>>>(T, capacity, img.tile_offsets, ranges);
    if (n_contrib) {
        cudaError_t e = cudaMemcpyAsync(n_contrib, img.n_contrib, sizeof(uint32_t) * (size_t)W * H,
                                        cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace gdr
