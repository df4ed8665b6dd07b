// Layout of the caller-owned opaque state buffers (see include/gdr.h).
//
// GeomState  (reference: GeometryState, rasterizer_impl.h:33-48) -- per Gaussian:
//     Splat    splat[P]          48 B   packed blend record (xy, depth, id | conic, opacity | rgb, reject threshold)
//     float    cov3D[6P]         24 B   world covariance (needed by the backward)
//     uint32   tiles_touched[P]   4 B
//     uint8    clamped[P]         1 B   bit c set <=> SH colour channel c was clamped at 0
//     uint64   tile_mask[P]       8 B   bit k set <=> tile k (row-major) of the Gaussian's tile rectangle is binned
//                                       (survived the exact culling test); valid for rectangles of <= 64 tiles
// ImageState (reference: ImageState, rasterizer_impl.h:50-56) -- per view:
//     uint32   header[64]               header[0] = R (instances), header[1] = overflow flag
//     uint32   tile_offsets[T + 1]      exclusive scan of per-tile instance counts (the reference's `ranges`)
//     uint32   tile_counter[T * SUBBINS] bin counters (count pass, then emit cursors); a tile's segment is the
//                                       concatenation of SUBBINS sub-segments chosen by (Gaussian index % SUBBINS),
//                                       so that same-address atomic traffic on the busiest tiles is split SUBBINS ways
//     uint32   sub_offsets[T * SUBBINS + 1]  exclusive scan of the sub-bin counts (tile_offsets[t] == sub_offsets[t * SUBBINS])
//     uint32   tile_order[T]            tiles in decreasing-work order (blend kernels: blockIdx -> tile)
//     uint32   n_contrib[H * W]
// SplatStream (reference: BinningState.point_list, but materialised):
//     Splat    stream[capacity]         per-tile, depth-sorted copies of the Gaussians' records,
//                                       contiguous per tile so a tile is staged with cp.async.bulk
// SortScratch: uint64 keys[capacity], uint64 keys_alt[capacity]
#pragma once
#include "common.cuh"

namespace gdr {

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct GeomState {
    Splat* splat;
    float* cov3D;
    uint32_t* tiles_touched;
    uint8_t* clamped;
    unsigned long long* tile_mask;
    static __host__ __device__ size_t bytes(size_t P) {
        size_t o = 0;
        o = align_up(o + sizeof(Splat) * P, 256);
        o = align_up(o + sizeof(float) * 6 * P, 256);
        o = align_up(o + sizeof(uint32_t) * P, 256);
        o = align_up(o + P, 256);
        o = align_up(o + sizeof(unsigned long long) * P, 256);
        return o + 256;
    }
    static __host__ __device__ GeomState carve(void* base, size_t P) {
        GeomState g;
        char* p = (char*)base;
        size_t o = 0;
        g.splat = (Splat*)(p + o);
        o = align_up(o + sizeof(Splat) * P, 256);
        g.cov3D = (float*)(p + o);
        o = align_up(o + sizeof(float) * 6 * P, 256);
        g.tiles_touched = (uint32_t*)(p + o);
        o = align_up(o + sizeof(uint32_t) * P, 256);
        g.clamped = (uint8_t*)(p + o);
        o = align_up(o + P, 256);
        g.tile_mask = (unsigned long long*)(p + o);
        return g;
    }
    // the same state of view v of a batch whose per-view states are `stride` bytes apart
    __host__ __device__ GeomState at(int v, size_t stride) const {
        GeomState g;
        const size_t off = (size_t)v * stride;
        g.splat = (Splat*)((char*)splat + off);
        g.cov3D = (float*)((char*)cov3D + off);
        g.tiles_touched = (uint32_t*)((char*)tiles_touched + off);
        g.clamped = clamped + off;
        g.tile_mask = (unsigned long long*)((char*)tile_mask + off);
        return g;
    }
};

constexpr int IMG_HEADER_WORDS = 64;
constexpr int SUBBINS = 8;  // sub-counters per tile (see ImageState)
constexpr int HDR_NUM_RENDERED = 0;
constexpr int HDR_OVERFLOW = 1;
constexpr int HDR_MAX_TILE = 2;

struct ImageState {
    uint32_t* header;
    uint32_t* tile_offsets;
    uint32_t* tile_counter;
    uint32_t* tile_order;
    uint32_t* n_contrib;
    uint32_t* sub_offsets;
    static __host__ __device__ size_t bytes(int W, int H) {
        const size_t T = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
        size_t o = 0;
        o = align_up(o + 4 * IMG_HEADER_WORDS, 256);
        o = align_up(o + 4 * (T + 1), 256);
        o = align_up(o + 4 * T * SUBBINS, 256);
        o = align_up(o + 4 * T, 256);
        o = align_up(o + 4 * (size_t)W * H, 256);
        o = align_up(o + 4 * (T * SUBBINS + 1), 256);
        return o + 256;
    }
    static __host__ __device__ ImageState carve(void* base, int W, int H) {
        const size_t T = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
        ImageState s;
        char* p = (char*)base;
        size_t o = 0;
        s.header = (uint32_t*)(p + o);
        o = align_up(o + 4 * IMG_HEADER_WORDS, 256);
        s.tile_offsets = (uint32_t*)(p + o);
        o = align_up(o + 4 * (T + 1), 256);
        s.tile_counter = (uint32_t*)(p + o);
        o = align_up(o + 4 * T * SUBBINS, 256);
        s.tile_order = (uint32_t*)(p + o);
        o = align_up(o + 4 * T, 256);
        s.n_contrib = (uint32_t*)(p + o);
        o = align_up(o + 4 * (size_t)W * H, 256);
        s.sub_offsets = (uint32_t*)(p + o);
        return s;
    }
    __host__ __device__ ImageState at(int v, size_t stride) const {
        ImageState s;
        const size_t off = (size_t)v * stride;
        s.header = (uint32_t*)((char*)header + off);
        s.tile_offsets = (uint32_t*)((char*)tile_offsets + off);
        s.tile_counter = (uint32_t*)((char*)tile_counter + off);
        s.tile_order = (uint32_t*)((char*)tile_order + off);
        s.n_contrib = (uint32_t*)((char*)n_contrib + off);
        s.sub_offsets = (uint32_t*)((char*)sub_offsets + off);
        return s;
    }
};

// How a kernel finds view v = blockIdx.y of a batch of V views that share one set of Gaussians.
// The single-view entry points use V = 1 with every stride 0 and the tangents passed by value; the
// batched entry points point view/proj/campos/bg/tanfov into an array of gdr_camera blocks
// (include/gdr.h) with cam_stride = 48 floats.
struct Views {
    int V;
    size_t geom_stride;  // bytes between consecutive views' GeomState
    size_t img_stride;   // bytes between consecutive views' ImageState
    size_t cam_stride;   // floats between consecutive views' camera parameters
    const float* view;   // [16] transposed world->view matrix of view 0
    const float* proj;   // [16] transposed full projection of view 0
    const float* campos; // [3]
    const float* bg;     // [3]
    const float* tanfov; // {tan_fovx, tan_fovy} of view 0 in device memory, or nullptr: use the two fields below
    float tan_fovx, tan_fovy;
    __device__ __forceinline__ float tanx(int v) const { return tanfov ? __ldg(tanfov + v * cam_stride) : tan_fovx; }
    __device__ __forceinline__ float tany(int v) const { return tanfov ? __ldg(tanfov + v * cam_stride + 1) : tan_fovy; }
};

}  // namespace gdr
