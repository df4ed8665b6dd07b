// Layout of the caller-owned opaque state buffers (see include/gdr.h).
//
// GeomState  (reference: GeometryState, rasterizer_impl.h:33-48) -- per Gaussian:
//     Splat    splat[P]          48 B   packed blend record (xy, reject threshold, id | conic, opacity | rgb, depth)
//     float    cov3D[6P]         24 B   world covariance (needed by the backward)
//     uint32   tiles_touched[P]   4 B   tiles of the reference's 3-sigma rectangle (before culling)
//     uint8    clamped[P]         1 B   bit c set <=> SH colour channel c was clamped at 0
// ImageState (reference: ImageState + the per-tile half of BinningState, rasterizer_impl.h:50-66) -- per view:
//     uint32   header[64]               see the HDR_* words below
//     uint32   tile_count[T * COUNT_STRIDE]  instances binned to each tile (word t * COUNT_STRIDE): the projection kernel
//                                       claims a tile's slots with returning atomics on this counter (no count pass, no
//                                       prefix sum).  One counter per 256 bytes: the L2 atomic units serialise per
//                                       line, and with 32 counters per 128-byte line ALL the frame's claims funnelled
//                                       through T / 32 lines (measured: the projection kernel's time was flat in P)
//     uint2    tile_range[T]            [begin, end) of the tile's depth-sorted records in the stream; tile_sort
//                                       allocates it from a global cursor, so tiles lie in completion order -- nothing
//                                       downstream needs them in tile order (the reference's `ranges` likewise only
//                                       holds [begin, end) per tile, rasterizer_impl.cu:116-138)
//     uint32   order[33][T]             tiles grouped by floor(log2(count)) + 1 (bucket 0 = empty tiles), filled by
//                                       tile_sort; the blend kernels walk the buckets from the heaviest down
//     uint32   n_contrib[H * W]
// SortScratch (temporary): uint64 keys -- key = depth bits << 32 | Gaussian index.  Two layouts:
//     uniform  keys[T][tile_capacity]   every tile's segment has the same capacity, predicted by the host from the
//                                       previous frame (the steady state: one projection pass, no prefix sum);
//     exact    keys[R], tile t's segment = [tile_offsets[t], tile_offsets[t + 1])  with tile_offsets the exclusive scan
//                                       of the per-tile counts a first projection pass measured (gdr_tile_offsets) --
//                                       taken when the prediction failed (first frame of a scene size, or a jump of the
//                                       densest tile) or when uniform segments would waste memory (a few very dense
//                                       tiles): exactly R keys, for any distribution
// SplatStream (reference: BinningState.point_list, but materialised):
//     Splat    stream[capacity]         per-tile, depth-sorted copies of the Gaussians' records,
//                                       contiguous per tile so a tile is staged with cp.async.bulk
#pragma once
#include "common.cuh"

namespace gdr {

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct GeomState {
    Splat* splat;
    float* cov3D;
    uint32_t* tiles_touched;
    uint8_t* clamped;
    static __host__ __device__ size_t bytes(size_t P) {
        size_t o = 0;
        o = align_up(o + sizeof(Splat) * P, 256);
        o = align_up(o + sizeof(float) * 6 * P, 256);
        o = align_up(o + sizeof(uint32_t) * P, 256);
        o = align_up(o + P, 256);
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
        return g;
    }
};

constexpr int IMG_HEADER_WORDS = 64;
constexpr int COUNT_STRIDE = 64;    // words between consecutive tiles' slot counters (256 B: see ImageState)
constexpr int ORDER_BUCKETS = 33;   // floor(log2(count)) + 1 for count > 0, bucket 0 = empty tiles
// header words.  The first four are what the host reads back (gdr_forward_project's counts_host); words from
// HDR_CURSOR on belong to tile_sort and are reset when a render is repeated with a larger stream capacity.
constexpr int HDR_NUM_RENDERED = 0;  // R: instances binned (after tile culling), summed by the projection kernel
constexpr int HDR_PROJECT_FLAGS = 1; // HDR_FLAG_PREFILTERED
constexpr int HDR_MAX_TILE = 2;      // largest per-tile instance count (may exceed the tile capacity: then re-run)
constexpr int HDR_TICKET = 3;        // CTAs of the projection kernel that have finished (the last one reports the counts)
constexpr int HDR_CURSOR = 4;        // tile_sort's stream allocation cursor
constexpr int HDR_SORT_FLAGS = 5;    // HDR_FLAG_STREAM_OVERFLOW
constexpr int HDR_BUCKET0 = 8;       // ORDER_BUCKETS fill counts of the order lists
constexpr uint32_t HDR_FLAG_STREAM_OVERFLOW = 1u;  // a tile's records did not fit the stream capacity (tile_sort)
constexpr uint32_t HDR_FLAG_PREFILTERED = 2u;      // prefiltered = true but a Gaussian failed the near-plane test
static_assert(HDR_BUCKET0 + ORDER_BUCKETS <= IMG_HEADER_WORDS, "header too small");

struct ImageState {
    uint32_t* header;
    uint32_t* tile_count;
    uint2* tile_range;
    uint32_t* order;
    uint32_t* n_contrib;
    static __host__ __device__ size_t tiles(int W, int H) {
        return (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    }
    static __host__ __device__ size_t bytes(int W, int H) {
        const size_t T = tiles(W, H);
        size_t o = 0;
        o = align_up(o + 4 * IMG_HEADER_WORDS, 256);
        o = align_up(o + 4 * T * COUNT_STRIDE, 256);
        o = align_up(o + 8 * T, 256);
        o = align_up(o + 4 * T * ORDER_BUCKETS, 256);
        o = align_up(o + 4 * (size_t)W * H, 256);
        return o + 256;
    }
    static __host__ __device__ ImageState carve(void* base, int W, int H) {
        const size_t T = tiles(W, H);
        ImageState s;
        char* p = (char*)base;
        size_t o = 0;
        s.header = (uint32_t*)(p + o);
        o = align_up(o + 4 * IMG_HEADER_WORDS, 256);
        s.tile_count = (uint32_t*)(p + o);
        o = align_up(o + 4 * T * COUNT_STRIDE, 256);
        s.tile_range = (uint2*)(p + o);
        o = align_up(o + 8 * T, 256);
        s.order = (uint32_t*)(p + o);
        o = align_up(o + 4 * T * ORDER_BUCKETS, 256);
        s.n_contrib = (uint32_t*)(p + o);
        return s;
    }
    __host__ __device__ ImageState at(int v, size_t stride) const {
        ImageState s;
        const size_t off = (size_t)v * stride;
        s.header = (uint32_t*)((char*)header + off);
        s.tile_count = (uint32_t*)((char*)tile_count + off);
        s.tile_range = (uint2*)((char*)tile_range + off);
        s.order = (uint32_t*)((char*)order + off);
        s.n_contrib = (uint32_t*)((char*)n_contrib + off);
        return s;
    }
};

// Bytes of one view's key segments: T tiles x tile_capacity keys (tile_capacity a multiple of 32).
__host__ __device__ inline size_t sort_scratch_bytes(int W, int H, int64_t tile_capacity) {
    return align_up(ImageState::tiles(W, H) * (size_t)tile_capacity * sizeof(uint64_t), 256);
}

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
