// Stage 3: forward alpha blend (front to back), one CTA per 16x16 tile.
//
// Replaces the reference's renderCUDA forward (RAST/cuda_rasterizer/forward.cu:
// 261-381).  Per pixel the arithmetic and its order are the reference's: the
// same exponent expression, full-precision expf, alpha = min(0.99, o*G), skip
// below 1/255, stop (without blending) when T*(1-alpha) < 1e-4, and
// out_alpha = sum(alpha*T).  What differs is how the work is organised:
//   * the tile's depth-sorted Splat records are contiguous (binning.cu), so each warp
//     stages them in 32-record chunks with ONE cp.async.bulk (TMA engine) per chunk into
//     its own 3-stage shared-memory ring with mbarrier completion -- the four warps of a
//     CTA never wait for each other; the reference gathers 28 B per record through
//     per-thread loads and re-reads colour and depth from global memory for every
//     contributing (pixel, Gaussian) pair;
//   * the kernel is instruction-issue bound (ncu: 82 % issue-slot utilisation), so
//     the levers are fewer (pixel, splat) evaluations and fewer issue slots per
//     evaluation.  A warp owns an 8x8 pixel region, each lane TWO pixels (rows y and
//     y + 4) whose values travel as packed FP32 pairs: the exponent, the exp range
//     reduction and the blend are FFMA2 / FMUL2 / FADD2 (sm_100 packed FP32, the same
//     IEEE roundings as the scalar code, one issue slot for both pixels), and the
//     list walk, the shared-memory loads and the votes are paid once per two pixels.
//     tile_sort classified every record against the four 8x8 regions of the tile
//     with the exact rectangle bound of common.cuh (a 4-bit mask in the id word of
//     the stream copy), and every warp walks only the records whose bit is set for
//     its region (ballot + find-first-set).  Records that cannot reach alpha = 1/255 anywhere
//     in a warp's region cost that warp nothing;
//   * every pair of a visited record takes the reference's exact test (full-precision expf as a packed
//     bit-exact replica, common.cuh: expf2); a pixel that does not blend a record carries alpha = 0
//     through the updates, which makes them the exact identity -- no divergent branch inside a visit;
//   * optional fused epilogue of Renderer.render_img (clamped HWC image, clamp mask parked in n_contrib).
#include "kernels.h"
#include "tile_iter.cuh"

namespace gdr {

namespace {

constexpr int BLEND_THREADS = 128;  // 4 warps, each owns an 8x8 pixel region: 2 pixels (rows y and y+4) per lane
constexpr int WARPS = BLEND_THREADS / 32;
constexpr int WCHUNK = 32;  // records per staged chunk: one record per lane to classify
constexpr int STAGES = 3;   // per-warp ring depth (two chunks in flight behind the one being blended)

__global__ void __launch_bounds__(BLEND_THREADS)
blend_forward_kernel(int W, int H, int gx, int T, ImageState img0, const Splat* __restrict__ stream0, int64_t capacity,
                     float* __restrict__ out_color0, float* __restrict__ out_depth0, float* __restrict__ out_alpha0,
                     const Views vw, const int hwc_clamp) {
    __shared__ __align__(128) Splat buf[WARPS][STAGES][WCHUNK];
    __shared__ __align__(8) uint64_t full[WARPS][STAGES];

    const int v = blockIdx.y;  // view of the batch
    const ImageState img = img0.at(v, vw.img_stride);
    uint32_t* __restrict__ n_contrib = img.n_contrib;
    const Splat* __restrict__ stream = stream0 + (size_t)v * capacity;
    const float* __restrict__ bg = vw.bg + (size_t)v * vw.cam_stride;
    float* __restrict__ out_color = out_color0 + (size_t)v * 3 * H * W;
    float* __restrict__ out_depth = out_depth0 + (size_t)v * H * W;
    float* __restrict__ out_alpha = out_alpha0 + (size_t)v * H * W;

    // Launched as a programmatic dependent of tile_sort, which wrote everything read from here on (the tile order
    // lists, the tile ranges, the stream).
    pdl_wait();
    const int tile = tile_from_order(img.header, img.order, T, (int)blockIdx.x);  // heaviest tiles first
    const int tile_x = tile % gx, tile_y = tile / gx;
    const uint2 range = img.tile_range[tile];
    const int n = (int)(range.y - range.x);
    const int n_chunks = (n + WCHUNK - 1) / WCHUNK;
    const Splat* src = stream + range.x;

    // From here on the four warps of the CTA never synchronise with each other: each streams the tile's
    // record list through its own ring and leaves as soon as its own 64 pixels are finished.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rx = tile_x * TILE + (warp & 1) * 8, ry = tile_y * TILE + (warp >> 1) * 8;  // region origin
    const int px = rx + (lane & 7);
    const int pya = ry + (lane >> 3), pyb = pya + 4;
    const bool inside_a = px < W && pya < H, inside_b = px < W && pyb < H;
    const float pxf = (float)px;
    const f32x2 pyf2 = pk((float)pya, (float)pyb);
    uint64_t* my_full = full[warp];
    Splat(*my_buf)[WCHUNK] = buf[warp];

    auto issue = [&](int c) {  // lane 0 only
        const int cnt = min(WCHUNK, n - c * WCHUNK);
        const uint32_t bytes = (uint32_t)(cnt * sizeof(Splat));
        mbar_expect_tx(&my_full[c % STAGES], bytes);
        bulk_g2s(&my_buf[c % STAGES][0], src + (size_t)c * WCHUNK, bytes, &my_full[c % STAGES]);
    };
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < STAGES; st++) mbar_init(&my_full[st], 1);
        mbar_fence_init();
        for (int c = 0; c < min(STAGES - 1, n_chunks); c++) issue(c);
    }
    __syncwarp();

    // A pixel's alpha floor: 1/255 while it is live, +inf once it is finished (T would drop below 1e-4) or if it
    // lies outside the image -- "finished" then needs no flag of its own in the per-record tests.
    float amin_a = inside_a ? ALPHA_MIN : INFINITY, amin_b = inside_b ? ALPHA_MIN : INFINITY;
    f32x2 T2 = pk2(1.0f);
    // (the colour / depth sums stay scalar: ptxas will not accumulate an FFMA2 in place when its product
    // operand dies, and the register copies that follow cost more issue slots than the packing saves)
    float C0a = 0.f, C0b = 0.f, C1a = 0.f, C1b = 0.f, C2a = 0.f, C2b = 0.f, Da = 0.f, Db = 0.f;
    f32x2 weight = pk2(0.f);
    uint32_t last_a = 0, last_b = 0;

    int c = 0;
    for (; c < n_chunks; c++) {
        // every lane has finished chunk c - 1, whose ring slot is the one chunk c + STAGES - 1 lands in
        __syncwarp();
        if (lane == 0 && c + STAGES - 1 < n_chunks) issue(c + STAGES - 1);
        mbar_wait(&my_full[c % STAGES], (c / STAGES) & 1);
        const int cnt = min(WCHUNK, n - c * WCHUNK);
        const Splat* sp = &my_buf[c % STAGES][0];

        // lane l reads record l's region bit (tile_sort evaluated the exact rectangle bound for the four regions)
        const bool hit = lane < cnt && ((__float_as_uint(sp[lane].q0.w) >> (STREAM_REGION_SHIFT + warp)) & 1u);
        unsigned word = __ballot_sync(0xffffffffu, hit);
        while (word) {
            const int j = __ffs(word) - 1;
            word &= word - 1;
            const float4 q0 = sp[j].q0;
            const float4 con_o = sp[j].q1;
            const float dx = q0.x - pxf;
            const f32x2 dy2 = sub2(pk2(q0.y), pyf2);
            const f32x2 power2 = pair_power2(con_o, pk2(dx), dy2);
            float pa, pb;
            upk(power2, pa, pb);
            const f32x2 og2 = mul2(pk2(con_o.w), expf2(power2));
            float aa, ab;
            upk(og2, aa, ab);
            aa = min(0.99f, aa);
            ab = min(0.99f, ab);
            // a pixel blends this record unless power > 0, alpha < 1/255, or the pixel is finished
            bool ma = !(pa > 0.0f) && !(aa < amin_a);
            bool mb = !(pb > 0.0f) && !(ab < amin_b);
            if (__any_sync(0xffffffffu, ma || mb)) {
                float Ta, Tb, tta, ttb;
                upk(T2, Ta, Tb);
                upk(mul2(T2, sub2(pk2(1.f), pk(aa, ab))), tta, ttb);  // test_T = T * (1 - alpha)
                if (ma && tta < T_MIN) {
                    amin_a = INFINITY;
                    ma = false;
                }
                if (mb && ttb < T_MIN) {
                    amin_b = INFINITY;
                    mb = false;
                }
                // a pixel that does not blend this record gets alpha = 0: every update below is then
                // the exact identity
                const f32x2 al2 = pk(ma ? aa : 0.f, mb ? ab : 0.f);
                const float4 q2 = sp[j].q2;
                float p0a, p0b, p1a, p1b, p2a, p2b, pda, pdb;
                upk(mul2(pk2(q2.x), al2), p0a, p0b);
                upk(mul2(pk2(q2.y), al2), p1a, p1b);
                upk(mul2(pk2(q2.z), al2), p2a, p2b);
                upk(mul2(pk2(q2.w), al2), pda, pdb);
                C0a = fmaf(p0a, Ta, C0a);
                C0b = fmaf(p0b, Tb, C0b);
                C1a = fmaf(p1a, Ta, C1a);
                C1b = fmaf(p1b, Tb, C1b);
                C2a = fmaf(p2a, Ta, C2a);
                C2b = fmaf(p2b, Tb, C2b);
                Da = fmaf(pda, Ta, Da);
                Db = fmaf(pdb, Tb, Db);
                fma2_acc(weight, al2, T2);
                T2 = pk(ma ? tta : Ta, mb ? ttb : Tb);
                const uint32_t pos1 = (uint32_t)(c * WCHUNK + j + 1);
                last_a = ma ? pos1 : last_a;
                last_b = mb ? pos1 : last_b;
            }
        }
        if (__all_sync(0xffffffffu, amin_a == INFINITY && amin_b == INFINITY)) {
            c++;
            break;
        }
    }
    // never leave with a bulk copy still in flight into our shared memory: chunks c .. c + STAGES - 2 were issued
    if (lane == 0)
        for (int k = c; k < min(n_chunks, c + STAGES - 1); k++) mbar_wait(&my_full[k % STAGES], (k / STAGES) & 1);

    const size_t HW = (size_t)H * W;
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    float Ta, Tb, wa, wb;
    upk(T2, Ta, Tb);
    upk(weight, wa, wb);
    // hwc_clamp: the epilogue of Renderer.render_img (lightning/renderer.py:261-265) fused in -- the colour leaves
    // clamped to [0, 1] in HWC layout, and the channels the clamp cut are remembered for the backward
    auto write_pixel = [&](int py, float c0, float c1, float c2, float T, float w, float d, uint32_t last) {
        const size_t pid = (size_t)py * W + px;
        c0 += T * bg0;
        c1 += T * bg1;
        c2 += T * bg2;
        if (hwc_clamp) {
            const unsigned cut = (!(c0 >= 0.f && c0 <= 1.f) ? 1u : 0u) | (!(c1 >= 0.f && c1 <= 1.f) ? 2u : 0u) |
                                 (!(c2 >= 0.f && c2 <= 1.f) ? 4u : 0u);
            float* o = out_color + 3 * pid;  // NaN stays NaN, as with torch.clamp
            o[0] = c0 < 0.f ? 0.f : (c0 > 1.f ? 1.f : c0);
            o[1] = c1 < 0.f ? 0.f : (c1 > 1.f ? 1.f : c1);
            o[2] = c2 < 0.f ? 0.f : (c2 > 1.f ? 1.f : c2);
            last |= cut << NCONTRIB_CLAMP_SHIFT;
        } else {
            out_color[pid] = c0;
            out_color[HW + pid] = c1;
            out_color[2 * HW + pid] = c2;
        }
        n_contrib[pid] = last;
        out_alpha[pid] = w;
        out_depth[pid] = d;
    };
    if (inside_a) write_pixel(pya, C0a, C1a, C2a, Ta, wa, Da, last_a);
    if (inside_b) write_pixel(pyb, C0b, C1b, C2b, Tb, wb, Db, last_b);
    pdl_trigger();  // a backward blend enqueued right behind this launch may start launching (it waits for our stores)
}

}  // namespace

cudaError_t launch_blend_forward(int W, int H, ImageState img, const Splat* stream, int64_t capacity,
                                 float* out_color, float* out_depth, float* out_alpha, const Views& vw, int hwc_clamp,
                                 cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    return launch_dependent(blend_forward_kernel, dim3(gx * gy, max(1, vw.V)), dim3(BLEND_THREADS), 0, s, W, H, gx, gx * gy, img,
                            stream, capacity, out_color, out_depth, out_alpha, vw, hwc_clamp);
}

}  // namespace gdr
