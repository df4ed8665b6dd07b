// Warp-cooperative binning of the (Gaussian, tile) pairs owned by a warp -- in ONE pass.
//
// Replaces the reference's tiles_touched prefix sum + duplicateWithKeys
// (RAST/cuda_rasterizer/rasterizer_impl.cu:278, 70-111) without a count pass: every kept pair claims the next
// slot of its tile's key segment with one returning atomic on the tile's counter and writes its key
// (depth bits << 32 | Gaussian index) there.  A tile's segment has a fixed capacity (ImageState / SortScratch in
// state.cuh); claims beyond it are counted but not stored, and the host re-runs the frame with a larger capacity.
//
// Each lane owns one Gaussian covering an n = w*h tile rectangle (n may be 0).  Instead of each lane looping over
// its own rectangle (divergent trip counts), the warp flattens all pairs into one index space and walks it 32 at a
// time, so lanes stay busy regardless of how uneven the rectangles are:
//   * the owners' parameters are parked once in a per-warp shared-memory table, compacted to the lanes with n > 0;
//   * per step, the owners whose first pair falls into the 32-pair window raise one bit each (REDUX.OR), and a lane
//     finds the owner of ITS pair with one popcount -- no binary search, no shuffles -- then reads the owner's
//     record with four LDS.128 (SHFL issues at one warp-instruction per clock per SM on this part,
//     profiles/r1_ubench.txt; the shuffle-based walk this replaces spent ~20 of them per step);
//   * the slot claims of up to EMIT_DEPTH windows are issued before the first key is stored, so the returning
//     atomics' round trips (ATOMG ~320 cycles unloaded) overlap each other and the culling tests.
#pragma once
#include "state.cuh"

namespace gdr {

// One owner (a lane with n > 0) of the warp's pair space.
struct __align__(16) EmitRec {
    float4 g0;  // centre x, y, reject threshold, conic a                  (culling test only)
    float4 g1;  // conic b, c, -b / c, -b / a                               (culling test only)
    uint4 r;    // first pair index (exclusive prefix), x0 | y0 << 16, w, ceil(2^32 / w)
    uint2 key;  // Gaussian index, depth bits
    uint2 pad;
};
static_assert(sizeof(EmitRec) == 64, "EmitRec must be 64 bytes");

#ifndef GDR_EMIT_DEPTH
#define GDR_EMIT_DEPTH 8
#endif
constexpr int EMIT_DEPTH = GDR_EMIT_DEPTH;  // 32-pair windows whose slot claims are in flight together

struct EmitTarget {
    uint32_t* tile_count;       // [T * COUNT_STRIDE] of this view
    uint64_t* keys;             // this view's key segments; T < 2^24 (checked at the API)
    uint32_t tile_cap;          // uniform layout: tile t's segment is keys[t * tile_cap .. + tile_cap)
    const uint32_t* tile_base;  // exact layout (non-null): tile t's segment is keys[tile_base[t] .. tile_base[t + 1])
    int gx;
};

// Where tile `tile`'s key segment starts and how many keys it holds (see SortScratch in state.cuh).
__device__ __forceinline__ void tile_segment(const uint32_t* __restrict__ tile_base, uint32_t tile_cap, int tile,
                                             size_t& first, uint32_t& cap) {
    if (tile_base) {
        first = tile_base[tile];
        cap = tile_base[tile + 1] - tile_base[tile];
    } else {
        first = (size_t)tile * tile_cap;
        cap = tile_cap;
    }
}

// All 32 lanes call.  `kept` / `max_fill` accumulate this lane's binned pairs and the largest slot + 1 it claimed.
// CULL: drop pairs whose tile the splat provably cannot reach with alpha >= 1/255 (splat_misses_rect, exact).
template <bool CULL>
__device__ __forceinline__ void warp_emit_tiles(EmitRec* __restrict__ s_rec, int n, int x0, int y0, int w, float4 q0,
                                                float4 q1, uint32_t depth_bits, uint32_t idx, const EmitTarget& t,
                                                uint32_t& kept, uint32_t& max_fill) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int incl = warp_incl_scan(n);
    const int total = __shfl_sync(full, incl, 31);
    if (total == 0) return;  // warp-uniform
    const int excl = incl - n;
    const unsigned owners = __ballot_sync(full, n > 0);
    if (n > 0) {
        EmitRec& r = s_rec[__popc(owners & lt)];
        if (CULL) {
            // (-b / c, -b / a): the same IEEE divisions splat_misses_rect performs, done once per splat
            r.g0 = make_float4(q0.x, q0.y, q0.z, q1.x);
            r.g1 = make_float4(q1.y, q1.z, __fdiv_rn(-q1.y, q1.z), __fdiv_rn(-q1.y, q1.x));
        }
        // local / w as a multiply-high by ceil(2^32 / w) (exact for local, w < 2^16); w <= 1: the row is `local`
        r.r = make_uint4((unsigned)excl, (unsigned)x0 | ((unsigned)y0 << 16), (unsigned)w,
                         w > 1 ? 0xffffffffu / (unsigned)w + 1u : 0u);
        r.key = make_uint2(idx, depth_bits);
    }
    __syncwarp();
    int o_start = 0;  // owners whose first pair lies before the current window
    // EMIT_DEPTH windows per round: all their slot claims are issued before the first key is stored, so a warp
    // waits for ONE atomic round trip per round instead of one per window (ncu on the one-window-at-a-time
    // version: 56 % of the kernel's stall samples sat on the claim's return value).
    for (int g0 = 0; g0 < total; g0 += 32 * EMIT_DEPTH) {
        uint32_t pos[EMIT_DEPTH], where[EMIT_DEPTH];  // claimed slot; tile | owner << 24, or ~0 for "no pair"
#pragma unroll
        for (int s = 0; s < EMIT_DEPTH; s++) {
            pos[s] = 0;
            where[s] = 0xffffffffu;
            const int base = g0 + 32 * s;
            if (base < total) {  // warp-uniform
                unsigned bit = 0;
                if (n > 0 && excl >= base && excl < base + 32) bit = 1u << (excl - base);
                const unsigned heads = __reduce_or_sync(full, bit);
                const int j = base + (int)lane;
                const bool valid = j < total;
                // the owner of pair j is the last owner whose first pair is <= j (lanes past the end: the last owner)
                const int oc = o_start + __popc(heads & (lt | (1u << lane))) - 1;
                o_start += __popc(heads);
                const EmitRec& rec = s_rec[oc];
                const uint4 rr = rec.r;
                const int local = valid ? j - (int)rr.x : 0;
                const int ow = (int)rr.z;
                const int row = ow > 1 ? (int)__umulhi((unsigned)local, rr.w) : local;
                const int tx = (int)(rr.y & 0xffffu) + (local - row * ow), ty = (int)(rr.y >> 16) + row;
                const int tile = ty * t.gx + tx;
                bool keep = valid;
                if (CULL) {
                    const float4 g0v = rec.g0, g1v = rec.g1;
                    const float tx0 = (float)(tx * TILE), ty0 = (float)(ty * TILE);
                    keep = valid && !splat_misses_rect_pre(g0v.x, g0v.y, g0v.w, g1v.x, g1v.y, g0v.z, g1v.z, g1v.w, tx0,
                                                           ty0, tx0 + (TILE - 1), ty0 + (TILE - 1));
                }
                if (keep) {
                    pos[s] = atomicAdd(&t.tile_count[(size_t)tile * COUNT_STRIDE], 1u);
                    where[s] = (unsigned)tile | ((unsigned)oc << 24);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < EMIT_DEPTH; s++) {
            if (where[s] != 0xffffffffu) {
                const uint2 key = s_rec[where[s] >> 24].key;
                size_t first;
                uint32_t cap;
                tile_segment(t.tile_base, t.tile_cap, (int)(where[s] & 0xffffffu), first, cap);
                if (pos[s] < cap) t.keys[first + pos[s]] = ((uint64_t)key.y << 32) | key.x;
                max_fill = max(max_fill, pos[s] + 1u);
                kept += 1u;
            }
        }
    }
    __syncwarp();  // the table is rewritten by the warp's next batch of Gaussians
}

// Called by one thread per CTA of a projection kernel after its header atomics: the last CTA of the view to arrive
// writes {R, flags, largest tile count} into the view's pinned host row and then sets word 3 (release, system scope)
// -- the host polls that word (see st_host_release in common.cuh).
__device__ __forceinline__ void report_counts(uint32_t* header, int32_t* counts_host, unsigned n_ctas) {
    __threadfence();  // this CTA's header atomics are ordered before its ticket
    if (atomicAdd(&header[HDR_TICKET], 1u) != n_ctas - 1) return;
    __threadfence();
    st_host_relaxed(counts_host + 0, (int32_t)ld_device_acquire(&header[HDR_NUM_RENDERED]));
    st_host_relaxed(counts_host + 1, (int32_t)ld_device_acquire(&header[HDR_PROJECT_FLAGS]));
    st_host_relaxed(counts_host + 2, (int32_t)ld_device_acquire(&header[HDR_MAX_TILE]));
    st_host_release(counts_host + 3, 1);
}

// The blend kernels' blockIdx -> tile map: tiles in decreasing-work order.  tile_sort files every tile under
// bucket floor(log2(count)) + 1 (0 = empty); CTA i takes the i-th tile counting from the heaviest bucket down.
// All 32 lanes call; every lane gets the tile.
__device__ __forceinline__ int tile_from_order(const uint32_t* __restrict__ header, const uint32_t* __restrict__ order,
                                               int T, int i) {
    const unsigned full = 0xffffffffu;
    const int lane = (int)lane_id();
    const uint32_t cnt = header[HDR_BUCKET0 + 32 - lane];  // lane l owns bucket 32 - l: heaviest first
    const int incl = warp_incl_scan((int)cnt);
    const unsigned after = __ballot_sync(full, incl > i);
    int b = 0, pos = i - __shfl_sync(full, incl, 31);  // bucket 0 (empty tiles) follows all the others
    if (after) {
        const int f = __ffs(after) - 1;
        b = 32 - f;
        pos = i - (__shfl_sync(full, incl, f) - (int)__shfl_sync(full, cnt, f));
    }
    return (int)order[(size_t)b * T + pos];
}

}  // namespace gdr
