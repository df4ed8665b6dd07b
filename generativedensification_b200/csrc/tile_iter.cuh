// Warp-cooperative iteration over the (Gaussian, tile) pairs owned by a warp.
//
// Each lane owns one Gaussian covering an n = w*h tile rectangle (n may be 0).
// Instead of each lane looping over its own rectangle (divergent trip counts,
// one atomic per pair), the warp flattens all pairs into one index space and
// walks it 32 at a time, so lanes stay busy regardless of how uneven the
// rectangles are.  Callers then aggregate the bin-counter atomics per distinct
// tile with __match_any_sync (one atomic per distinct tile per step): Gaussians
// that are neighbours in memory are usually neighbours on screen (the model's
// coarse Gaussians sit on a voxel grid in index order), so this removes most of
// the same-address contention of the counters.
#pragma once
#include "common.cuh"

namespace gdr {

// f(tile_id, owner_lane, local_index, valid, active_mask, tile_x, tile_y); invoked by all 32 lanes each step.
// local / w is a multiply-high by ceil(2^32 / w), computed once per lane (exact for local, w < 2^16): the
// per-step integer divisions were a fifth of the walk.
template <class F>
__device__ __forceinline__ void warp_foreach_tile(int n, int x0, int y0, int w, int gx, F&& f) {
    const unsigned full = 0xffffffffu;
    const int lane = (int)lane_id();
    const int incl = warp_incl_scan(n);
    const int total = __shfl_sync(full, incl, 31);
    const unsigned inv_w = w > 1 ? 0xffffffffu / (unsigned)w + 1u : 0u;  // w <= 1: rows are `local` itself
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        int lo = 0, hi = 31;  // smallest lane whose inclusive prefix exceeds j
#pragma unroll
        for (int step = 0; step < 5; step++) {
            const int mid = (lo + hi) >> 1;
            const int v = __shfl_sync(full, incl, mid);
            if (v > j) hi = mid; else lo = mid + 1;
        }
        const int owner = lo;
        const int o_excl = __shfl_sync(full, incl - n, owner);
        const int o_x0 = __shfl_sync(full, x0, owner);
        const int o_y0 = __shfl_sync(full, y0, owner);
        const int o_w = max(1, __shfl_sync(full, w, owner));
        const unsigned o_inv = __shfl_sync(full, inv_w, owner);
        const bool valid = j < total;
        const int local = valid ? j - o_excl : 0;
        const int row = o_w > 1 ? (int)__umulhi((unsigned)local, o_inv) : local;
        const int tx = o_x0 + (local - row * o_w), ty = o_y0 + row;
        const int tile = ty * gx + tx;
        const unsigned active = __ballot_sync(full, valid);
        f(tile, owner, local, valid, active, tx, ty);
    }
}

}  // namespace gdr
