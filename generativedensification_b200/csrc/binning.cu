// Stage 2: per-tile depth sort + record gather.
//
// Replaces the reference's global 64-bit DeviceRadixSort + identifyTileRanges
// (RAST/cuda_rasterizer/rasterizer_impl.cu:304-309, 116-138).  The order contract is
// unchanged: inside a tile, instances are ordered by (view-depth bits ascending, Gaussian
// index ascending) -- exactly what the reference's stable LSD sort on
// (tile << 32 | depth bits) of idx-major emitted pairs yields.
//
// B200-first structure (the binning itself happens inside the projection kernel, tile_iter.cuh):
//   tile_sort   one CTA per tile.  Reads the tile's instance count (the projection kernel's
//               slot counter), allocates the tile's range of the record stream from a global
//               cursor (tiles lie in completion order; nothing downstream needs tile order),
//               files the tile in the heaviest-first order lists of the blend kernels, sorts the
//               tile's key segment in shared memory (monotone depth-bucket sort; a bitonic
//               network only for short lists and exact depth ties) and writes the tile's
//               depth-sorted 48-byte Splat records contiguously, ready for bulk-copy staging in
//               the blend kernels.
// Compared with a global 64-bit radix sort of all R pairs (6+ passes over 12 B/pair) plus a
// prefix sum and a duplication pass, each instance is written once as an 8-byte key and read once.
#include "kernels.h"
#include "surfel.cuh"
#include "tile_iter.cuh"

namespace gdr {

namespace {

#ifndef GDR_SORT_THREADS
#define GDR_SORT_THREADS 256
#endif
constexpr int SORT_THREADS = GDR_SORT_THREADS;
// Keys per shared-memory sort (8 bytes each) come in two sizes, chosen per launch from the frame's tile capacity:
// 2048 (16 KB: 8 CTAs per SM) when the densest tile is expected below ~2000 instances -- lists up to 1024 sort in
// shared memory, the few longer ones through global memory -- and 4096 (32 KB: 5 CTAs per SM) for dense frames, whose
// long-list path then gets 4096 depth buckets instead of 2048 (measured: 29 vs 33 us at the 200k / 800^2 benchmark
// frame with the small buffer, 745 vs 672 us at 2M / 1600^2).
constexpr int SORT_CHUNK_SMALL = 2048, SORT_CHUNK_LARGE = 4096;
constexpr int BUCKET_BITS = 10;
constexpr int BUCKETS = 1 << BUCKET_BITS;   // depth buckets of the per-tile bucket sort
constexpr int BUCKET_MIN = 65;              // shorter lists: the bitonic network is already cheap
constexpr int BUCKET_MAX_FILL = 48;         // largest bucket the quadratic in-bucket ranking accepts
static_assert(BUCKETS % SORT_THREADS == 0, "bucket scan assigns BUCKETS / SORT_THREADS buckets per thread");

// One compare-exchange of the bitonic network: pair index i of sub-stage (k, j).
__device__ __forceinline__ void bitonic_cmpxchg(uint64_t* s, int i, int j, int k) {
    const int a = 2 * i - (i & (j - 1));
    const int b = a + j;
    const uint64_t va = s[a], vb = s[b];
    const bool up = (a & k) == 0;
    if ((va > vb) == up) {
        s[a] = vb;
        s[b] = va;
    }
}

// Bitonic sort of s[0 .. n_pad) ascending; n_pad is a power of two; all threads call.
// Sub-stages with partner distance j <= 32 only exchange elements inside aligned 64-element blocks, so a
// warp that owns a block runs all of them back to back with warp-level synchronisation; only the
// sub-stages with j >= 64 need the CTA barrier (6 instead of 45 barriers for a 512-entry tile list).
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, int n_pad) {  // n_pad <= the caller's key buffer
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int N_WARPS = SORT_THREADS / 32;
    const int n_blocks = max(n_pad >> 6, 1);  // 64-element blocks (a single partial block when n_pad < 64)
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j >= 64; j >>= 1) {
            for (int i = threadIdx.x; i < (n_pad >> 1); i += SORT_THREADS) bitonic_cmpxchg(s, i, j, k);
            __syncthreads();
        }
        for (int b = warp; b < n_blocks; b += N_WARPS) {
            for (int j = min(k >> 1, 32); j > 0; j >>= 1) {
                const int i = b * 32 + lane;
                if (i < (n_pad >> 1)) bitonic_cmpxchg(s, i, j, k);
                __syncwarp();
            }
        }
        if (k >= 64) __syncthreads();  // the next stage (or the caller) reads across blocks
    }
    __syncthreads();
}

__device__ __forceinline__ int lower_bound_u64(const uint64_t* a, int n, uint64_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// RQ = 128-bit words per record: 3 for the 48-byte Splat, 5 for the 80-byte Surfel (surfel.cuh).
template <int RQ, int SORT_CHUNK>
__global__ void __launch_bounds__(SORT_THREADS)
tile_sort_kernel(const float4* __restrict__ records0, ImageState img0, uint64_t* keys0, float4* __restrict__ stream0,
                 int64_t capacity, uint32_t tile_cap, const uint32_t* __restrict__ tile_base0, size_t keys_stride,
                 size_t geom_stride, size_t img_stride, int gx, int T) {
    constexpr int BUCKET_MAX = SORT_CHUNK / 2;  // the grouped copy lives in the second half of the key buffer
    constexpr int BIG_BITS = SORT_CHUNK >= 4096 ? 12 : 11;  // long lists: counters + scan alias the key buffer
    static_assert((2 << BIG_BITS) * 4 <= SORT_CHUNK * 8, "the long-list counters live in the key buffer");
    __shared__ uint64_t s_keys[SORT_CHUNK];
    __shared__ uint32_t s_cnt[BUCKETS];        // bucket histogram, then fill cursors
    __shared__ uint32_t s_start[BUCKETS + 1];  // exclusive scan of the histogram
    __shared__ uint32_t s_wsum[SORT_THREADS / 32];
    __shared__ uint32_t s_misc[4];             // min depth bits, max depth bits, largest bucket
    __shared__ uint32_t s_alloc[2];            // the tile's stream range: first record, records that fit
    pdl_wait();  // launched as a programmatic dependent of the projection kernel
    const int v = blockIdx.y;
    const float4* __restrict__ records =
        reinterpret_cast<const float4*>(reinterpret_cast<const char*>(records0) + (size_t)v * geom_stride);
    const ImageState img = img0.at(v, img_stride);
    float4* __restrict__ stream = stream0 + (size_t)v * capacity * RQ;
    const int tile = blockIdx.x;
    size_t seg_first;
    uint32_t seg_cap;
    tile_segment(tile_base0 ? tile_base0 + (size_t)v * (T + 1) : nullptr, tile_cap, tile, seg_first, seg_cap);
    const uint32_t n_binned = min(img.tile_count[(size_t)tile * COUNT_STRIDE], seg_cap);  // claims beyond the segment were not stored
    if (n_binned == 0) {  // an empty tile: no stream range, last in the blend kernels' order
        if (threadIdx.x == 0) {
            img.tile_range[tile] = make_uint2(0u, 0u);
            img.order[atomicAdd(&img.header[HDR_BUCKET0], 1u)] = (uint32_t)tile;
        }
        return;
    }
    // The tile's range of the stream comes from a global cursor.  Thread 0 issues the claim here and only looks at the
    // answer when the sorted keys are ready (allocate() below): the atomic's round trip overlaps the sort.
    uint32_t claimed = 0;
    if (threadIdx.x == 0) claimed = atomicAdd(&img.header[HDR_CURSOR], n_binned);
    auto allocate = [&]() {  // thread 0 publishes (base, records that fit) in s_misc[3], s_misc[2]; callers synchronise
        if (threadIdx.x == 0) {
            uint32_t base = claimed, n_fit = 0;
            if ((int64_t)base < capacity) n_fit = (uint32_t)min((int64_t)n_binned, capacity - (int64_t)base);
            else base = 0;
            if (n_fit < n_binned) atomicOr(&img.header[HDR_SORT_FLAGS], HDR_FLAG_STREAM_OVERFLOW);  // the host re-runs
            img.tile_range[tile] = make_uint2(base, base + n_fit);
            s_alloc[0] = base;
            s_alloc[1] = n_fit;
        }
    };
    const int n = (int)n_binned;
    uint64_t* seg = keys0 + (size_t)v * keys_stride + seg_first;
    const uint64_t* sorted;  // where the sorted keys end up (shared memory or `seg`)
    uint64_t* alt = nullptr;
    if (n > BUCKET_MAX) {
        // second key buffer of the long-list paths: the tile's own (not yet written) range of the record stream --
        // RQ * 16 bytes per record, of which 8 are used; the gather below overwrites it after the sorted keys are back
        // in `seg`.  These paths need the range first.
        allocate();
        __syncthreads();
        if ((int)s_alloc[1] < n) {  // the stream is too small for this tile (the host re-runs the render): nothing to sort into
            if (threadIdx.x == 0) {
                img.tile_range[tile] = make_uint2(0u, 0u);
                img.order[atomicAdd(&img.header[HDR_BUCKET0], 1u)] = (uint32_t)tile;
            }
            return;
        }
        alt = reinterpret_cast<uint64_t*>(stream + (size_t)s_alloc[0] * RQ);
    }

    // ---- long lists (dense scenes: 2M Gaussians at 1600^2 give ~5000 instances per covered tile) ----
    // The same monotone depth-bucket sort, with the grouped copy and the result in global memory (both L2-resident
    // for one tile) and 4096 buckets whose counters live in the otherwise unused key buffer.  Four linear passes
    // plus the in-bucket ranking replace a 4096-entry bitonic network per chunk and log2(n / 4096) merge passes.
    bool sorted_big = false;
    if (n > BUCKET_MAX) {
        constexpr int NBIG = 1 << BIG_BITS;
        uint32_t* big_cnt = reinterpret_cast<uint32_t*>(s_keys);  // [NBIG] histogram, then fill cursors
        uint32_t* big_start = big_cnt + NBIG;                     // [NBIG] exclusive scan
        uint64_t* grouped = alt;
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
            const uint32_t d = (uint32_t)(seg[i] >> 32);
            lo = min(lo, d);
            hi = max(hi, d);
        }
        for (int i = threadIdx.x; i < NBIG; i += SORT_THREADS) big_cnt[i] = 0;
        if (threadIdx.x == 0) {
            s_misc[0] = 0xffffffffu;
            s_misc[1] = 0u;
            s_misc[2] = 0u;
        }
        __syncthreads();
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&s_misc[0], lo);
            atomicMax(&s_misc[1], hi);
        }
        __syncthreads();
        const uint32_t dmin = s_misc[0], range = s_misc[1] - s_misc[0];
        const int sh = max(0, (32 - __clz(range)) - BIG_BITS);  // (d - dmin) >> sh < NBIG
        for (int i = threadIdx.x; i < n; i += SORT_THREADS)
            atomicAdd(&big_cnt[((uint32_t)(seg[i] >> 32) - dmin) >> sh], 1u);
        __syncthreads();
        {   // exclusive scan of the bucket counts (NBIG / SORT_THREADS consecutive buckets per thread)
            constexpr int PER = NBIG / SORT_THREADS;
            uint32_t sum = 0, mx = 0;
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const uint32_t c = big_cnt[threadIdx.x * PER + q];
                sum += c;
                mx = max(mx, c);
            }
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            const int incl = warp_incl_scan((int)sum);
            if (lane == 31) s_wsum[wid] = (uint32_t)incl;
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0) atomicMax(&s_misc[2], mx);
            __syncthreads();
            uint32_t run = (uint32_t)incl - sum;
            for (int w = 0; w < wid; w++) run += s_wsum[w];
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const uint32_t c = big_cnt[threadIdx.x * PER + q];
                big_start[threadIdx.x * PER + q] = run;
                big_cnt[threadIdx.x * PER + q] = 0;  // becomes the fill cursor
                run += c;
            }
        }
        __syncthreads();
        if (s_misc[2] <= 2 * BUCKET_MAX_FILL) {  // exact depth ties pile up in one bucket: take the network below
            for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                const uint64_t k = seg[i];
                const uint32_t bkt = ((uint32_t)(k >> 32) - dmin) >> sh;
                grouped[big_start[bkt] + atomicAdd(&big_cnt[bkt], 1u)] = k;
            }
            __threadfence_block();
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                const uint64_t k = grouped[i];
                const uint32_t bkt = ((uint32_t)(k >> 32) - dmin) >> sh;
                const uint32_t b0 = big_start[bkt], b1 = b0 + big_cnt[bkt];
                uint32_t rank = b0;
                for (uint32_t j = b0; j < b1; j++) rank += grouped[j] < k;
                seg[rank] = k;  // keys are unique: ranks are a permutation
            }
            __threadfence_block();
            __syncthreads();
            sorted_big = true;
        } else {
            __syncthreads();  // the counters alias the key buffer the fallback is about to fill
        }
    }
    if (sorted_big) {
        sorted = seg;
    } else if (n <= SORT_CHUNK) {
        bool sorted_by_buckets = false;
        if (n >= BUCKET_MIN && n <= BUCKET_MAX) {
            // Bucket sort: O(n) shared-memory operations instead of the bitonic network's O(n log^2 n).
            // The depth bits (positive floats: monotone as integers) are mapped monotonically onto BUCKETS
            // buckets spanning the tile's [min, max] depth; the keys are grouped by bucket with shared-memory
            // atomics and each key then finds its exact rank among the (few) keys of its own bucket.  Tiles
            // whose depths pile up in one bucket (exact depth ties) fall back to the bitonic network.
            uint32_t lo = 0xffffffffu, hi = 0u;
            for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                const uint64_t k = seg[i];
                s_keys[i] = k;
                const uint32_t d = (uint32_t)(k >> 32);
                lo = min(lo, d);
                hi = max(hi, d);
            }
            for (int i = threadIdx.x; i < BUCKETS; i += SORT_THREADS) s_cnt[i] = 0;
            if (threadIdx.x == 0) {
                s_misc[0] = 0xffffffffu;
                s_misc[1] = 0u;
                s_misc[2] = 0u;
            }
            __syncthreads();
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if ((threadIdx.x & 31) == 0) {
                atomicMin(&s_misc[0], lo);
                atomicMax(&s_misc[1], hi);
            }
            __syncthreads();
            const uint32_t dmin = s_misc[0], range = s_misc[1] - s_misc[0];
            const int sh = max(0, (32 - __clz(range)) - BUCKET_BITS);  // (d - dmin) >> sh < BUCKETS
            for (int i = threadIdx.x; i < n; i += SORT_THREADS)
                atomicAdd(&s_cnt[((uint32_t)(s_keys[i] >> 32) - dmin) >> sh], 1u);
            __syncthreads();
            {   // exclusive scan of the bucket counts (BUCKETS / SORT_THREADS consecutive buckets per thread)
                constexpr int PER = BUCKETS / SORT_THREADS;
                uint32_t c[PER], sum = 0, mx = 0;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    c[q] = s_cnt[threadIdx.x * PER + q];
                    sum += c[q];
                    mx = max(mx, c[q]);
                }
                const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
                const int incl = warp_incl_scan((int)sum);
                if (lane == 31) s_wsum[wid] = (uint32_t)incl;
                mx = __reduce_max_sync(0xffffffffu, mx);
                if (lane == 0) atomicMax(&s_misc[2], mx);
                __syncthreads();
                uint32_t run = (uint32_t)incl - sum;
                for (int w = 0; w < wid; w++) run += s_wsum[w];
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    s_start[threadIdx.x * PER + q] = run;
                    s_cnt[threadIdx.x * PER + q] = 0;  // becomes the fill cursor
                    run += c[q];
                }
                if (threadIdx.x == SORT_THREADS - 1) s_start[BUCKETS] = run;
            }
            __syncthreads();
            if (s_misc[2] <= BUCKET_MAX_FILL) {
                uint64_t* grouped = s_keys + BUCKET_MAX;  // second half of the key buffer
                for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                    const uint64_t k = s_keys[i];
                    const uint32_t bkt = ((uint32_t)(k >> 32) - dmin) >> sh;
                    grouped[s_start[bkt] + atomicAdd(&s_cnt[bkt], 1u)] = k;
                }
                __syncthreads();
                for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                    const uint64_t k = grouped[i];
                    const uint32_t bkt = ((uint32_t)(k >> 32) - dmin) >> sh;
                    const uint32_t b0 = s_start[bkt], b1 = s_start[bkt + 1];
                    uint32_t rank = b0;
                    for (uint32_t j = b0; j < b1; j++) rank += grouped[j] < k;
                    s_keys[rank] = k;  // keys are unique: ranks are a permutation
                }
                __syncthreads();
                sorted_by_buckets = true;
            }
        }
        if (!sorted_by_buckets) {
            int n_pad = 2;
            while (n_pad < n) n_pad <<= 1;
            for (int i = threadIdx.x; i < n_pad; i += SORT_THREADS) s_keys[i] = i < n ? seg[i] : ~0ull;
            __syncthreads();
            bitonic_sort_smem(s_keys, n_pad);
        }
        sorted = s_keys;
    } else {
        // chunk sort in shared memory, written back in place
        for (int c0 = 0; c0 < n; c0 += SORT_CHUNK) {
            const int cn = min(SORT_CHUNK, n - c0);
            for (int i = threadIdx.x; i < SORT_CHUNK; i += SORT_THREADS) s_keys[i] = i < cn ? seg[c0 + i] : ~0ull;
            __syncthreads();
            bitonic_sort_smem(s_keys, SORT_CHUNK);
            for (int i = threadIdx.x; i < cn; i += SORT_THREADS) seg[c0 + i] = s_keys[i];
            __syncthreads();
        }
        // pairwise merges through global memory (keys are unique, so ranks are unambiguous)
        uint64_t* src = seg;
        uint64_t* dst = alt;
        for (int width = SORT_CHUNK; width < n; width <<= 1) {
            __threadfence_block();
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                const int pair0 = (i / (2 * width)) * (2 * width);
                const int mid = min(n, pair0 + width), end = min(n, pair0 + 2 * width);
                const uint64_t k = src[i];
                int pos;
                if (i < mid)
                    pos = i + lower_bound_u64(src + mid, end - mid, k);
                else
                    pos = (i - mid) + pair0 + lower_bound_u64(src + pair0, mid - pair0, k);
                dst[pos] = k;
            }
            uint64_t* t = src;
            src = dst;
            dst = t;
        }
        __threadfence_block();
        __syncthreads();
        if (src != seg) {  // the gather writes over `alt`: the sorted keys must be in the key segment
            for (int i = threadIdx.x; i < n; i += SORT_THREADS) seg[i] = src[i];
            __threadfence_block();
            __syncthreads();
        }
        sorted = seg;
    }

    // gather the Gaussians' records into the tile's contiguous, depth-ordered stream
    if (alt == nullptr) {
        allocate();
        __syncthreads();
    }
    const int n_fit = (int)s_alloc[1];
    if (threadIdx.x == 0) {  // the tile's place in the blend kernels' heaviest-first order
        const int bucket = n_fit ? 32 - __clz(n_fit) : 0;
        img.order[(size_t)bucket * T + atomicAdd(&img.header[HDR_BUCKET0 + bucket], 1u)] = (uint32_t)tile;
    }
    float4* out = stream + (size_t)s_alloc[0] * RQ;
    for (int i = threadIdx.x; i < n_fit; i += SORT_THREADS) {
        const uint32_t id = (uint32_t)(sorted[i] & 0xffffffffu);
        const float4* src4 = records + (size_t)id * RQ;
        float4 w[RQ];
#pragma unroll
        for (int q = 0; q < RQ; q++) w[q] = __ldg(src4 + q);
        if constexpr (RQ == 3) {
            // Which of the tile's four 8x8 regions the splat can reach with alpha >= 1/255 (the exact rectangle
            // bound the blend kernels would otherwise evaluate once per warp, in the forward AND the backward):
            // bit 28 + r of the id word.  Gaussian indices stay below 2^28 (checked at the API).
            const float tx0 = (float)((tile % gx) * TILE), ty0 = (float)((tile / gx) * TILE);
            const unsigned m = region_mask4(w[0].x - tx0, w[0].y - ty0, w[1].x, w[1].y, w[1].z, w[0].z);
            w[0].w = __uint_as_float(__float_as_uint(w[0].w) | (m << STREAM_REGION_SHIFT));
        } else {
            // the surfel path's eight 8x4 blocks (surfel.cuh): the bits replace the reach in the stream copy's r4.w
            const float tx0 = (float)((tile % gx) * TILE), ty0 = (float)((tile / gx) * TILE);
            w[4].w = __uint_as_float(surfel_region_mask8(w[0], w[1], w[2], w[3], tx0, ty0));
        }
        float4* dst4 = out + (size_t)i * RQ;
#pragma unroll
        for (int q = 0; q < RQ; q++) dst4[q] = w[q];
    }
    pdl_trigger();  // the forward blend may start launching (empty tiles left above: exiting counts as a trigger)
}

// One CTA per view: exclusive scan of the per-tile instance counts (the exact key layout of state.cuh).  Only runs
// after a projection whose uniform segments overflowed, or for scenes whose uniform segments would waste memory.
__global__ void tile_offsets_kernel(int T, ImageState img0, size_t img_stride, uint32_t* __restrict__ offsets0) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const ImageState img = img0.at(blockIdx.x, img_stride);
    uint32_t* __restrict__ offsets = offsets0 + (size_t)blockIdx.x * (T + 1);
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += blockDim.x) {
        const int t = base + threadIdx.x;
        const uint32_t n = t < T ? img.tile_count[(size_t)t * COUNT_STRIDE] : 0u;
        const int incl = warp_incl_scan((int)n);
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = (uint32_t)incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t ws = threadIdx.x < (blockDim.x >> 5) ? s_warp[threadIdx.x] : 0u;
            s_warp[threadIdx.x] = (uint32_t)warp_incl_scan((int)ws) - ws;
        }
        __syncthreads();
        const uint32_t begin = s_carry + s_warp[threadIdx.x >> 5] + (uint32_t)incl - n;
        if (t < T) offsets[t] = begin;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = begin + n;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[T] = s_carry;
}

}  // namespace

size_t key_bytes_per_view(int W, int H, int64_t tile_cap, bool exact) {
    // exact layout: tile_cap is the number of keys of one view's region (>= the largest R of the batch)
    return exact ? align_up((size_t)tile_cap * sizeof(uint64_t), 256) : sort_scratch_bytes(W, H, tile_cap);
}

// The uniform layout's tile capacity is the host's forecast of the densest tile (2x the previous frame's + 256): up to
// 4096 slots the frame is a sparse one.  The exact layout (a frame whose forecast failed, or a degenerate distribution)
// always takes the large buffer.
#ifndef GDR_SORT_SMALL_CAP
#define GDR_SORT_SMALL_CAP 4096
#endif
static bool small_sort_buffer(int64_t tile_cap, const uint32_t* tile_base) {
    return tile_base == nullptr && tile_cap <= GDR_SORT_SMALL_CAP;
}

cudaError_t launch_tile_offsets(int W, int H, ImageState img, uint32_t* tile_offsets, const Views& vw, cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    tile_offsets_kernel<<<max(1, vw.V), 1024, 0, s>>>(gx * gy, img, vw.img_stride, tile_offsets);
    return cudaGetLastError();
}

cudaError_t launch_tile_sort(int W, int H, GeomState geom, ImageState img, uint64_t* keys, int64_t tile_cap,
                             const uint32_t* tile_base, Splat* stream, int64_t capacity, const Views& vw,
                             cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    auto launch = [&](auto kernel) {
        return launch_dependent(kernel, dim3(gx * gy, max(1, vw.V)), dim3(SORT_THREADS), 0, s,
                                reinterpret_cast<const float4*>(geom.splat), img, keys, reinterpret_cast<float4*>(stream),
                                capacity, (uint32_t)tile_cap, tile_base,
                                key_bytes_per_view(W, H, tile_cap, tile_base != nullptr) / sizeof(uint64_t),
                                vw.geom_stride, vw.img_stride, gx, gx * gy);
    };
    return small_sort_buffer(tile_cap, tile_base) ? launch(tile_sort_kernel<3, SORT_CHUNK_SMALL>)
                                                  : launch(tile_sort_kernel<3, SORT_CHUNK_LARGE>);
}

cudaError_t launch_tile_sort_surfel(int W, int H, const void* surfel_records, ImageState img, uint64_t* keys,
                                    int64_t tile_cap, const uint32_t* tile_base, void* surfel_stream, int64_t capacity,
                                    cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    auto launch = [&](auto kernel) {
        return launch_dependent(kernel, dim3(gx * gy, 1), dim3(SORT_THREADS), 0, s,
                                reinterpret_cast<const float4*>(surfel_records), img, keys,
                                reinterpret_cast<float4*>(surfel_stream), capacity, (uint32_t)tile_cap, tile_base,
                                (size_t)0, (size_t)0, (size_t)0, gx, gx * gy);
    };
    return small_sort_buffer(tile_cap, tile_base) ? launch(tile_sort_kernel<5, SORT_CHUNK_SMALL>)
                                                  : launch(tile_sort_kernel<5, SORT_CHUNK_LARGE>);
}

}  // namespace gdr
