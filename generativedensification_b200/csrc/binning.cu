// Stage 2: tile binning and per-tile depth sort.
//
// Replaces the reference's InclusiveSum + duplicateWithKeys + global 64-bit
// DeviceRadixSort + identifyTileRanges (RAST/cuda_rasterizer/rasterizer_impl.cu:
// 278, 70-111, 304-309, 116-138).  The order contract is unchanged: inside a
// tile, instances are ordered by (view-depth bits ascending, Gaussian index
// ascending) -- exactly what the reference's stable LSD sort on
// (tile << 32 | depth bits) of idx-major emitted pairs yields.
//
// B200-first structure:
//   tile_scan   a cluster of 8 CTAs scans the per-tile bin counters filled by the
//               project kernel (partial sums exchanged through distributed shared
//               memory) -> tile_offsets (the reference's `ranges`), R.
//   emit        one thread per Gaussian replays the kept-tile bitmask the projection
//               kernel recorded (rectangles of <= 64 tiles: no culling test, no
//               cooperation, the slot claims of one thread overlap in flight);
//               larger rectangles take a warp-cooperative walk over (Gaussian,
//               tile) pairs with one aggregated atomic per distinct tile.  Writes
//               the key (depth bits << 32 | idx).  Slot order within a segment is
//               arbitrary -- the keys are unique, so the sort below makes the
//               result deterministic.
//   tile_sort   one CTA per tile sorts its segment in shared memory (bitonic on
//               u64; segments longer than the smem chunk are chunk-sorted and
//               merged through global memory) and writes the tile's depth-sorted
//               48-byte Splat records contiguously, ready for bulk-copy staging
//               in the blend kernels.
// Compared with a global 64-bit radix sort of all R pairs (6+ passes over
// 12 B/pair), each instance is written once as an 8-byte key and read once.
#include <cooperative_groups.h>

#include "kernels.h"
#include "tile_iter.cuh"

namespace gdr {

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_CLUSTER = 8;  // CTAs (SMs) that share one view's scan through distributed shared memory

static_assert(SUBBINS % 4 == 0, "tile_scan_kernel moves a tile's sub-bin counters as 128-bit words");
constexpr int SUBQ = SUBBINS / 4;  // 128-bit words per tile

struct ScanShared {
    uint32_t bucket_count[33];  // tiles of this CTA per floor(log2(count)) + 1 bucket; bucket 0 = empty tiles
    uint32_t bucket_above[33];  // tiles of this CTA in heavier buckets
    uint32_t total;             // instances in this CTA's tile range
    uint32_t max_tile;          // largest tile of this CTA's range
};

// One CLUSTER of SCAN_CLUSTER CTAs per view (a single CTA was bound by one SM's load / store path and by the
// chain of block-wide barriers: 15 us for 2500 tiles x 8 sub-bins).  CTA r owns the contiguous tile range
// [r * per_cta, (r + 1) * per_cta): it sums its tiles' sub-bin counters (aligned 128-bit words), publishes its
// total / largest tile / log2-bucket histogram in its shared memory, and after one cluster barrier every CTA reads
// its peers' values through DSMEM to get (a) its base offset and (b) where its tiles of each bucket start in the
// heaviest-first tile order the blend kernels use.  A second cluster barrier keeps every CTA's shared memory alive
// until all peers have read it.
__global__ void __cluster_dims__(SCAN_CLUSTER, 1, 1) __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(int T, ImageState img0, size_t img_stride) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    pdl_wait();  // launched as a programmatic dependent of the projection kernel
    const ImageState img = img0.at(blockIdx.y, img_stride);
    __shared__ ScanShared sh;
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t bucket_base[33];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int per_cta = (T + SCAN_CLUSTER - 1) / SCAN_CLUSTER;
    const int t0 = min(T, rank * per_cta), t1 = min(T, t0 + per_cta);
    if (tid < 33) sh.bucket_count[tid] = 0;
    if (tid == 0) {
        sh.total = 0;
        sh.max_tile = 0;
        s_carry = 0;
    }
    __syncthreads();
    uint4* counters = reinterpret_cast<uint4*>(img.tile_counter);
    // ---- pass 1: totals, largest tile, bucket histogram of this CTA's range ----
    uint32_t lmax = 0, lsum = 0;
    for (int i = t0 + tid; i < t1; i += SCAN_THREADS) {
        uint32_t c = 0;
#pragma unroll
        for (int q = 0; q < SUBQ; q++) {
            const uint4 a = counters[SUBQ * i + q];
            c += a.x + a.y + a.z + a.w;
        }
        lmax = max(lmax, c);
        lsum += c;
        atomicAdd(&sh.bucket_count[c ? 32 - __clz(c) : 0], 1u);
    }
    lmax = __reduce_max_sync(0xffffffffu, lmax);
    lsum = __reduce_add_sync(0xffffffffu, lsum);
    if (lane == 0) {
        atomicMax(&sh.max_tile, lmax);
        atomicAdd(&sh.total, lsum);
    }
    __syncthreads();
    if (wid == 0) {  // lane l owns bucket 32 - l: an inclusive scan from the heaviest bucket down
        const uint32_t cnt = sh.bucket_count[32 - lane];
        const uint32_t incl = (uint32_t)warp_incl_scan((int)cnt);
        sh.bucket_above[32 - lane] = incl - cnt;
        if (lane == 31) sh.bucket_above[0] = incl;
    }
    cluster.sync();  // every CTA's ScanShared is complete and visible cluster-wide
    // ---- exchange through distributed shared memory ----
    if (tid < 33) {
        // heaviest bucket first; inside a bucket the CTAs' tiles follow each other in rank order
        uint32_t before = 0;
#pragma unroll
        for (int r = 0; r < SCAN_CLUSTER; r++) {
            const ScanShared* peer = cluster.map_shared_rank(&sh, r);
            before += peer->bucket_above[tid];
            if (r < rank) before += peer->bucket_count[tid];
        }
        bucket_base[tid] = before;
    }
    if (tid == 64) {
        uint32_t base = 0, mx = 0, total = 0;
#pragma unroll
        for (int r = 0; r < SCAN_CLUSTER; r++) {
            const ScanShared* peer = cluster.map_shared_rank(&sh, r);
            const uint32_t t = peer->total;
            if (r < rank) base += t;
            total += t;
            mx = max(mx, peer->max_tile);
        }
        s_carry = base;
        if (rank == 0) {
            img.header[HDR_MAX_TILE] = mx;
            img.header[HDR_NUM_RENDERED] = total;
            img.tile_offsets[T] = total;
            img.sub_offsets[T * SUBBINS] = total;
        }
    }
    __syncthreads();
    // ---- pass 2: chunked block scan of this CTA's range with a running carry ----
    for (int base = t0; base < t1; base += SCAN_THREADS) {
        const int i = base + tid;
        uint4 a[SUBQ];
        uint32_t c = 0;
#pragma unroll
        for (int q = 0; q < SUBQ; q++) {
            a[q] = i < t1 ? counters[SUBQ * i + q] : make_uint4(0, 0, 0, 0);
            c += a[q].x + a[q].y + a[q].z + a[q].w;
        }
        const int incl = warp_incl_scan((int)c);
        if (lane == 31) warp_sums[wid] = (uint32_t)incl;
        __syncthreads();
        if (wid == 0) {  // exclusive scan of the warp totals
            const uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
            const uint32_t ex = (uint32_t)warp_incl_scan((int)ws) - ws;
            if (lane < SCAN_THREADS / 32) warp_sums[lane] = ex;
        }
        __syncthreads();
        const uint32_t run = s_carry + warp_sums[wid] + (uint32_t)incl - c;
        if (i < t1) {
            img.tile_offsets[i] = run;
            uint4* so = reinterpret_cast<uint4*>(img.sub_offsets);
            uint32_t r = run;
#pragma unroll
            for (int q = 0; q < SUBQ; q++) {
                uint4 o;
                o.x = r;
                o.y = o.x + a[q].x;
                o.z = o.y + a[q].y;
                o.w = o.z + a[q].z;
                r = o.w + a[q].w;
                so[SUBQ * i + q] = o;
                counters[SUBQ * i + q] = o;  // the counters become the emit cursors: they start at the sub-bin's offset
            }
            img.tile_order[atomicAdd(&bucket_base[c ? 32 - __clz(c) : 0], 1u)] = (uint32_t)i;
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = run + c;  // total up to and including this chunk
        __syncthreads();
    }
    cluster.sync();  // no CTA leaves while a peer may still read its shared memory
}

constexpr int EMIT_THREADS = 128;

// One instance: write the key (depth bits << 32 | Gaussian index) into the claimed slot.  The cursors start at
// their sub-bin's offset (tile_scan), so a claim IS the position in the key array; the count pass and this pass
// replay the same kept-tile decisions, so a claim can only leave its segment by exceeding `capacity`.
__device__ __forceinline__ bool emit_one(uint64_t key, uint64_t* __restrict__ keys, int64_t capacity, uint32_t pos) {
    if ((int64_t)pos < capacity) {
        keys[pos] = key;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(int P, int gx, int gy, const int32_t* __restrict__ radii0, GeomState geom0, ImageState img0,
            uint64_t* __restrict__ keys0, int64_t capacity, int cull, size_t geom_stride, size_t img_stride) {
    const int v = blockIdx.y;
    const int32_t* __restrict__ radii = radii0 + (size_t)v * P;
    const GeomState geom = geom0.at(v, geom_stride);
    const Splat* __restrict__ splat = geom.splat;
    const ImageState img = img0.at(v, img_stride);
    uint32_t* __restrict__ cursor = img.tile_counter;
    uint32_t* __restrict__ header = img.header;
    uint64_t* __restrict__ keys = keys0 + (size_t)v * capacity;
    const int n_vblocks = (P + EMIT_THREADS - 1) / EMIT_THREADS;
    bool overflow = false;
    for (int vb = blockIdx.x; vb < n_vblocks; vb += gridDim.x) {  // virtual blocks: balanced single wave
        const int idx = vb * EMIT_THREADS + threadIdx.x;
        int n = 0, x0 = 0, y0 = 0, w = 0;
        uint32_t depth_bits = 0;
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
        unsigned long long m = 0ull;
        if (idx < P) {
            // four independent loads (one round trip), then the rectangle
            const int r = radii[idx];
            q0 = __ldg(&splat[idx].q0);
            depth_bits = __float_as_uint(__ldg(&splat[idx].q2.w));
            m = geom.tile_mask[idx];
            if (r > 0) {
                int x1, y1;
                tile_rect(q0.x, q0.y, r, gx, gy, x0, y0, x1, y1);
                w = x1 - x0;
                n = w * (y1 - y0);
            }
        }
        const uint64_t key = ((uint64_t)depth_bits << 32) | (uint32_t)idx;
        if (n > 0 && n <= 64) {
            // The common case: replay the kept-tile bitmask the projection kernel recorded for this Gaussian,
            // one slot claim per kept tile.  Four claims per round with their segment bounds: the atomics and
            // the loads are all issued before the first result is needed, so their round trips overlap.
            const uint32_t inv_w = (65536u + (uint32_t)w - 1u) / (uint32_t)w;  // local / w for local < 64, w <= 64
            while (m) {
                int bins[4];  // sub-bin ids: tile * SUBBINS + idx % SUBBINS
                uint32_t base[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    bins[u] = -1;
                    if (m) {
                        const int local = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        const int row = (int)(((uint32_t)local * inv_w) >> 16);
                        bins[u] = ((y0 + row) * gx + x0 + (local - row * w)) * SUBBINS + (idx & (SUBBINS - 1));
                        base[u] = atomicAdd(&cursor[bins[u]], 1u);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (bins[u] >= 0 && !emit_one(key, keys, capacity, base[u])) overflow = true;
                }
            }
            n = 0;  // done; takes no part in the cooperative walk below
        } else if (n > 64) {
            q1 = __ldg(&splat[idx].q1);
        }
        // Rectangles of more than 64 tiles (large splats): warp-cooperative walk over the (Gaussian, tile) pairs
        // with the culling test repeated exactly as the projection kernel counted it.
        if (__any_sync(0xffffffffu, n > 0)) {
            const int lane = (int)lane_id();
            warp_foreach_tile(n, x0, y0, w, gx, [&](int tile, int owner, int, bool valid, unsigned, int tx, int ty) {
                const uint32_t o_depth = __shfl_sync(0xffffffffu, depth_bits, owner);
                const int o_idx = __shfl_sync(0xffffffffu, idx, owner);
                bool keep = valid;
                if (cull) {  // the identical test project_kernel used when it counted this tile
                    const float cx = __shfl_sync(0xffffffffu, q0.x, owner), cy = __shfl_sync(0xffffffffu, q0.y, owner);
                    const float thr = __shfl_sync(0xffffffffu, q0.z, owner);
                    const float A = __shfl_sync(0xffffffffu, q1.x, owner), B = __shfl_sync(0xffffffffu, q1.y, owner);
                    const float C = __shfl_sync(0xffffffffu, q1.z, owner);
                    const float tx0 = (float)(tx * TILE), ty0 = (float)(ty * TILE);
                    keep = valid && !splat_misses_rect(cx, cy, A, B, C, thr, tx0, ty0, tx0 + (TILE - 1), ty0 + (TILE - 1));
                }
                const unsigned active = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int bin = tile * SUBBINS + (o_idx & (SUBBINS - 1));
                    const unsigned peers = __match_any_sync(active, bin);
                    const int leader = __ffs(peers) - 1;
                    uint32_t base = 0;
                    if (lane == leader) base = atomicAdd(&cursor[bin], (unsigned)__popc(peers));
                    base = __shfl_sync(peers, base, leader);
                    if (!emit_one(((uint64_t)o_depth << 32) | (uint32_t)o_idx, keys, capacity,
                                  base + __popc(peers & lanemask_lt())))
                        overflow = true;
                }
            });
        }
    }  // virtual blocks
    if (overflow) atomicOr(&header[HDR_OVERFLOW], 1u);
    pdl_trigger();  // tile_sort may start launching
}

constexpr int SORT_THREADS = 256;
constexpr int SORT_CHUNK = 4096;  // keys per shared-memory sort (32 KB)
constexpr int BUCKET_BITS = 10;
constexpr int BUCKETS = 1 << BUCKET_BITS;   // depth buckets of the per-tile bucket sort
constexpr int BUCKET_MIN = 65;              // shorter lists: the bitonic network is already cheap
constexpr int BUCKET_MAX = SORT_CHUNK / 2;  // the grouped copy lives in the second half of the key buffer
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
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, int n_pad) {
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
template <int RQ>
__global__ void __launch_bounds__(SORT_THREADS)
tile_sort_kernel(const float4* __restrict__ records0, ImageState img0, uint64_t* keys0, uint64_t* keys_alt0,
                 float4* __restrict__ stream0, int64_t capacity, size_t geom_stride, size_t img_stride, int gx) {
    __shared__ uint64_t s_keys[SORT_CHUNK];
    __shared__ uint32_t s_cnt[BUCKETS];        // bucket histogram, then fill cursors
    __shared__ uint32_t s_start[BUCKETS + 1];  // exclusive scan of the histogram
    __shared__ uint32_t s_wsum[SORT_THREADS / 32];
    __shared__ uint32_t s_misc[4];             // min depth bits, max depth bits, largest bucket
    pdl_wait();  // launched as a programmatic dependent of emit
    const int v = blockIdx.y;
    const float4* __restrict__ records =
        reinterpret_cast<const float4*>(reinterpret_cast<const char*>(records0) + (size_t)v * geom_stride);
    const ImageState img = img0.at(v, img_stride);
    const uint32_t* __restrict__ tile_offsets = img.tile_offsets;
    uint32_t* __restrict__ cursor = img.tile_counter;
    uint64_t* keys = keys0 + (size_t)v * capacity;
    uint64_t* keys_alt = keys_alt0 + (size_t)v * capacity;
    float4* __restrict__ stream = stream0 + (size_t)v * capacity * RQ;
    const int tile = blockIdx.x;
    if (threadIdx.x < SUBBINS)  // cursors back at their offsets for a (speculative) re-run
        cursor[tile * SUBBINS + threadIdx.x] = img.sub_offsets[tile * SUBBINS + threadIdx.x];
    const int64_t b = min((int64_t)tile_offsets[tile], capacity);
    const int64_t e = min((int64_t)tile_offsets[tile + 1], capacity);
    const int n = (int)(e - b);
    if (n == 0) return;
    uint64_t* seg = keys + b;
    const uint64_t* sorted;  // where the sorted keys end up (shared or global)

    // ---- long lists (dense scenes: 2M Gaussians at 1600^2 give ~5000 instances per covered tile) ----
    // The same monotone depth-bucket sort, with the grouped copy and the result in global memory (both L2-resident
    // for one tile) and 4096 buckets whose counters live in the otherwise unused key buffer.  Four linear passes
    // plus the in-bucket ranking replace a 4096-entry bitonic network per chunk and log2(n / 4096) merge passes.
    bool sorted_big = false;
    if (n > BUCKET_MAX) {
        constexpr int NBIG = 4096;
        uint32_t* big_cnt = reinterpret_cast<uint32_t*>(s_keys);  // [NBIG] histogram, then fill cursors
        uint32_t* big_start = big_cnt + NBIG;                     // [NBIG] exclusive scan
        uint64_t* grouped = keys_alt + b;
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
        const int sh = max(0, (32 - __clz(range)) - 12);  // (d - dmin) >> sh < NBIG
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
        uint64_t* dst = keys_alt + b;
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
        sorted = src;
    }

    // gather the Gaussians' records into the tile's contiguous, depth-ordered stream
    float4* out = stream + (size_t)b * RQ;
    for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
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
            unsigned m = 0;
            const float inv_c = __fdiv_rn(-w[1].y, w[1].z), inv_a = __fdiv_rn(-w[1].y, w[1].x);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const float rx = tx0 + (float)((r & 1) * 8), ry = ty0 + (float)((r >> 1) * 8);
                if (!splat_misses_rect_pre(w[0].x - rx, w[0].y - ry, w[1].x, w[1].y, w[1].z, w[0].z, inv_c, inv_a, 0.f,
                                           0.f, 7.f, 7.f))
                    m |= 1u << r;
            }
            w[0].w = __uint_as_float(__float_as_uint(w[0].w) | (m << STREAM_REGION_SHIFT));
        }
        float4* dst4 = out + (size_t)i * RQ;
#pragma unroll
        for (int q = 0; q < RQ; q++) dst4[q] = w[q];
    }
    pdl_trigger();  // the forward blend may start launching (empty tiles left above: exiting counts as a trigger)
}

}  // namespace

cudaError_t launch_tile_scan(int T, ImageState img, const Views& vw, cudaStream_t s) {
    return launch_dependent(tile_scan_kernel, dim3(SCAN_CLUSTER, max(1, vw.V)), dim3(SCAN_THREADS), 0, s, T, img,
                            vw.img_stride);
}

cudaError_t launch_emit(int P, int W, int H, const int32_t* radii, GeomState geom, ImageState img, uint64_t* keys,
                        int64_t capacity, int cull, const Views& vw, cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emit_kernel, EMIT_THREADS, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const int V = max(1, vw.V);
    const int grid = min((P + EMIT_THREADS - 1) / EMIT_THREADS, max(1, sm_count() * per_sm / V));
    emit_kernel<<<dim3(grid, V), EMIT_THREADS, 0, s>>>(P, gx, gy, radii, geom, img, keys, capacity, cull, vw.geom_stride,
                                                       vw.img_stride);
    return cudaGetLastError();
}

cudaError_t launch_tile_sort(int W, int H, GeomState geom, ImageState img, uint64_t* keys, uint64_t* keys_alt,
                             Splat* stream, int64_t capacity, const Views& vw, cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    return launch_dependent(tile_sort_kernel<3>, dim3(gx * gy, max(1, vw.V)), dim3(SORT_THREADS), 0, s,
                            reinterpret_cast<const float4*>(geom.splat), img, keys, keys_alt,
                            reinterpret_cast<float4*>(stream), capacity, vw.geom_stride, vw.img_stride, gx);
}

cudaError_t launch_tile_sort_surfel(int W, int H, const void* surfel_records, ImageState img, uint64_t* keys,
                                    uint64_t* keys_alt, void* surfel_stream, int64_t capacity, cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    tile_sort_kernel<5><<<dim3(gx * gy, 1), SORT_THREADS, 0, s>>>(reinterpret_cast<const float4*>(surfel_records), img,
                                                                  keys, keys_alt, reinterpret_cast<float4*>(surfel_stream),
                                                                  capacity, 0, 0, gx);
    return cudaGetLastError();
}

}  // namespace gdr
