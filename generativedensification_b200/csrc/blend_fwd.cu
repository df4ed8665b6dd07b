// Stage 3: forward alpha blend (front to back), one CTA per 16x16 tile.
//
// Replaces the reference's renderCUDA forward (RAST/cuda_rasterizer/forward.cu:
// 261-381).  Per pixel the arithmetic and its order are the reference's: the
// same exponent expression, full-precision expf, alpha = min(0.99, o*G), skip
// below 1/255, stop (without blending) when T*(1-alpha) < 1e-4, and
// out_alpha = sum(alpha*T).  What differs is how the work is organised:
//   * the tile's depth-sorted Splat records are contiguous (binning.cu), so each
//     256-record chunk is staged with ONE cp.async.bulk (TMA engine) into a
//     double-buffered shared-memory ring with mbarrier completion, while the
//     previous chunk is blended; the reference gathers 28 B per record through
//     per-thread loads and re-reads colour and depth from global memory for
//     every contributing (pixel, Gaussian) pair;
//   * the kernel is instruction-issue bound (ncu: ~90 % issue-slot utilisation), so
//     the lever is fewer (pixel, splat) evaluations.  A warp owns an 8x4 pixel
//     block; after a chunk lands, each thread classifies ONE record against the
//     eight 8x4 blocks of the tile with the exact rectangle bound of common.cuh
//     (an 8-bit mask), and every warp then walks only the records whose bit is set
//     for its block (ballot + find-first-set).  Records that cannot reach
//     alpha = 1/255 anywhere in a warp's block cost that warp nothing;
//   * for the pairs that are evaluated, a per-record conservative threshold on the
//     exponent skips the expf when alpha cannot reach 1/255 (exact: the slack is far
//     larger than any rounding error; everything near the cut takes the reference's
//     exact test).
#include "kernels.h"

namespace gdr {

namespace {

constexpr int BLEND_THREADS = 256;
constexpr int CHUNK = 256;

// Bit w of the result is set iff the record may contribute to the 8x4 pixel block of warp w.
// (lx, ly) = splat centre relative to the tile's first pixel.
__device__ __forceinline__ unsigned subblock_mask(float lx, float ly, float4 con_o, float thr) {
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const float x0 = (float)((w & 1) * 8), y0 = (float)((w >> 1) * 4);
        if (!splat_misses_rect(lx, ly, con_o.x, con_o.y, con_o.z, thr, x0, y0, x0 + 7.f, y0 + 3.f)) m |= 1u << w;
    }
    return m;
}

__global__ void __launch_bounds__(BLEND_THREADS)
blend_forward_kernel(int W, int H, int gx, ImageState img0, const Splat* __restrict__ stream0, int64_t capacity,
                     float* __restrict__ out_color0, float* __restrict__ out_depth0, float* __restrict__ out_alpha0,
                     const Views vw) {
    __shared__ __align__(128) Splat buf[2][CHUNK];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ uint8_t s_mask[CHUNK];

    const int v = blockIdx.y;  // view of the batch
    const ImageState img = img0.at(v, vw.img_stride);
    const uint32_t* __restrict__ tile_offsets = img.tile_offsets;
    const uint32_t* __restrict__ tile_order = img.tile_order;
    uint32_t* __restrict__ n_contrib = img.n_contrib;
    const Splat* __restrict__ stream = stream0 + (size_t)v * capacity;
    const float* __restrict__ bg = vw.bg + (size_t)v * vw.cam_stride;
    float* __restrict__ out_color = out_color0 + (size_t)v * 3 * H * W;
    float* __restrict__ out_depth = out_depth0 + (size_t)v * H * W;
    float* __restrict__ out_alpha = out_alpha0 + (size_t)v * H * W;

    const int tile = (int)tile_order[blockIdx.x];  // heaviest tiles first
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int64_t rb = min((int64_t)tile_offsets[tile], capacity);
    const int64_t re = min((int64_t)tile_offsets[tile + 1], capacity);
    const int n = (int)(re - rb);
    const int n_chunks = (n + CHUNK - 1) / CHUNK;
    const Splat* src = stream + rb;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tile_x * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float2 pixf = make_float2((float)px, (float)py);
    const float tile_fx = (float)(tile_x * TILE), tile_fy = (float)(tile_y * TILE);

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && n_chunks > 0) {
        const uint32_t bytes = (uint32_t)(min(CHUNK, n) * sizeof(Splat));
        mbar_expect_tx(&full[0], bytes);
        bulk_g2s(&buf[0][0], src, bytes, &full[0]);
    }

    bool done = !inside;
    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, weight = 0.f, D = 0.f;
    uint32_t last_contributor = 0;

    int c = 0;
    for (; c < n_chunks; c++) {
        // prefetch the next chunk into the other buffer (its previous contents were
        // released by the barrier at the end of the previous iteration)
        if (threadIdx.x == 0 && c + 1 < n_chunks) {
            const int cnt1 = min(CHUNK, n - (c + 1) * CHUNK);
            const uint32_t bytes = (uint32_t)(cnt1 * sizeof(Splat));
            mbar_expect_tx(&full[(c + 1) & 1], bytes);
            bulk_g2s(&buf[(c + 1) & 1][0], src + (size_t)(c + 1) * CHUNK, bytes, &full[(c + 1) & 1]);
        }
        mbar_wait(&full[c & 1], (c >> 1) & 1);
        const int cnt = min(CHUNK, n - c * CHUNK);
        const Splat* sp = &buf[c & 1][0];

        // classify: one record per thread against the eight 8x4 blocks of the tile
        {
            unsigned m = 0;
            if ((int)threadIdx.x < cnt) {
                const float4 q0 = sp[threadIdx.x].q0;
                m = subblock_mask(q0.x - tile_fx, q0.y - tile_fy, sp[threadIdx.x].q1, q0.z);
            }
            s_mask[threadIdx.x] = (uint8_t)m;
        }
        __syncthreads();

        if (!__all_sync(0xffffffffu, done)) {
            for (int k = 0; k * 32 < cnt; k++) {
                unsigned word = __ballot_sync(0xffffffffu, (s_mask[k * 32 + lane] >> warp) & 1u);
                while (word) {
                    const int j = k * 32 + __ffs(word) - 1;
                    word &= word - 1;
                    if (done) continue;
                    const float4 q0 = sp[j].q0;
                    const float4 con_o = sp[j].q1;
                    const float2 d = make_float2(q0.x - pixf.x, q0.y - pixf.y);
                    const float power = pair_power(con_o, d.x, d.y);
                    if (power > 0.0f) continue;
                    if (power < q0.z) continue;  // certainly alpha < 1/255
                    const float alpha = min(0.99f, con_o.w * expf(power));
                    if (alpha < ALPHA_MIN) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < T_MIN) {
                        done = true;
                        continue;
                    }
                    const float4 q2 = sp[j].q2;
                    C0 += q2.x * alpha * T;
                    C1 += q2.y * alpha * T;
                    C2 += q2.z * alpha * T;
                    weight += alpha * T;
                    D += q2.w * alpha * T;
                    T = test_T;
                    last_contributor = (uint32_t)(c * CHUNK + j + 1);
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        // everyone is finished with buf[c & 1] and s_mask; also the tile-wide early exit
        if (__syncthreads_and(done)) break;
    }
    // never leave with a bulk copy still in flight into our shared memory
    if (threadIdx.x == 0 && c < n_chunks && c + 1 < n_chunks) mbar_wait(&full[(c + 1) & 1], ((c + 1) >> 1) & 1);

    if (inside) {
        const size_t HW = (size_t)H * W;
        const size_t pid = (size_t)py * W + px;
        n_contrib[pid] = last_contributor;
        out_color[pid] = C0 + T * __ldg(bg);
        out_color[HW + pid] = C1 + T * __ldg(bg + 1);
        out_color[2 * HW + pid] = C2 + T * __ldg(bg + 2);
        out_alpha[pid] = weight;
        out_depth[pid] = D;
    }
}

}  // namespace

cudaError_t launch_blend_forward(int W, int H, ImageState img, const Splat* stream, int64_t capacity,
                                 float* out_color, float* out_depth, float* out_alpha, const Views& vw,
                                 cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    blend_forward_kernel<<<dim3(gx * gy, max(1, vw.V)), BLEND_THREADS, 0, s>>>(W, H, gx, img, stream, capacity, out_color,
                                                                               out_depth, out_alpha, vw);
    return cudaGetLastError();
}

}  // namespace gdr
