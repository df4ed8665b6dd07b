// 2D Gaussian-surfel (2DGS) raster path: projection, forward blend, backward blend, per-surfel backward.
//
// Serves the `diff_surfel_rasterization`-shaped module that lightning/renderer_2dgs.py:7-10,224-233
// of the reference imports.  PARITY UNPINNED: that extension's source is not in the reference tree
// (SURVEY.md 8c / 8f-3), so the arithmetic follows the published 2DGS algorithm and is checked against
// oracle/surfel_oracle.py (dense torch restatement, autograd backward).  Binning (slot claims inside the
// projection kernel, tile_iter.cuh; per-tile sort + record gather, binning.cu) is shared with the 3DGS path; the stream records are the
// 80-byte Surfel of surfel.cuh, staged per tile with cp.async.bulk like the 48-byte Splat stream.
#include <cstdlib>

#include "kernels.h"
#include "sh.cuh"
#include "surfel.cuh"
#include "tile_iter.cuh"

namespace gdr {

namespace {

constexpr int SP_THREADS = 128;

struct Rot3 {  // R[row][col] of the normalised quaternion (r, x, y, z)
    float m[3][3];
};

__device__ __forceinline__ float4 quat_normalised(float4 q) {
    const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
}

__device__ __forceinline__ Rot3 quat_to_rot(float4 qn) {
    const float r = qn.x, x = qn.y, y = qn.z, z = qn.w;
    Rot3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z);       R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z);       R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y);       R.m[2][1] = 2.f * (y * z + r * x);       R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    return R;
}

// Q[i][c]: world homogeneous coordinate i -> (pixel x * w, pixel y * w, w); pixel = ((ndc + 1) * S - 1) / 2.
__device__ __forceinline__ void world_to_pixel_hom(const float* __restrict__ proj, int W, int H, float Q[4][3]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float p0 = __ldg(proj + 4 * i), p1 = __ldg(proj + 4 * i + 1), p3 = __ldg(proj + 4 * i + 3);
        Q[i][0] = 0.5f * W * p0 + 0.5f * (W - 1) * p3;
        Q[i][1] = 0.5f * H * p1 + 0.5f * (H - 1) * p3;
        Q[i][2] = p3;
    }
}

struct SurfelProjectArgs {
    int P, sh_degree, M, W, H, gx, gy;
    const float *means3D, *shs, *colors_precomp, *opacities, *scales;
    int scale_stride;
    float scale_modifier;
    const float *rotations, *transmat_precomp;
    const float *view, *proj, *campos;
    int32_t* radii;
    GeomState geom;
    Surfel* surfel;
    ImageState img;
    uint64_t* keys;     // [T][tile_cap] key segments (SortScratch, state.cuh)
    uint32_t tile_cap;
    const uint32_t* tile_base;  // exact key layout [T + 1], or nullptr (state.cuh)
    int32_t* counts_host;  // device-accessible pinned host memory [4] or nullptr
};

// Bounding box of the ellipse rho3d <= cutoff^2 of the homography T (rows Tu, Tv, Tw).
__device__ __forceinline__ bool surfel_aabb_at(const float Tu[3], const float Tv[3], const float Tw[3], float cutoff2,
                                               float2& centre, float2& extent, float& dist_out) {
    const float t[3] = {cutoff2, cutoff2, -1.0f};
    const float dist = t[0] * Tw[0] * Tw[0] + t[1] * Tw[1] * Tw[1] + t[2] * Tw[2] * Tw[2];
    dist_out = dist;
    if (dist == 0.0f) return false;
    const float inv = 1.0f / dist;
    const float f[3] = {t[0] * inv, t[1] * inv, t[2] * inv};
    centre.x = f[0] * Tu[0] * Tw[0] + f[1] * Tu[1] * Tw[1] + f[2] * Tu[2] * Tw[2];
    centre.y = f[0] * Tv[0] * Tw[0] + f[1] * Tv[1] * Tw[1] + f[2] * Tv[2] * Tw[2];
    const float tx = f[0] * Tu[0] * Tu[0] + f[1] * Tu[1] * Tu[1] + f[2] * Tu[2] * Tu[2];
    const float ty = f[0] * Tv[0] * Tv[0] + f[1] * Tv[1] * Tv[1] + f[2] * Tv[2] * Tv[2];
    extent.x = sqrtf(fmaxf(1e-4f, centre.x * centre.x - tx));
    extent.y = sqrtf(fmaxf(1e-4f, centre.y * centre.y - ty));
    return true;
}

// Conservative reach of a surfel around its bounding-box centre `c`: outside the square |x - c.x|, |y - c.y| <= reach no
// pixel can get alpha >= 1/255 from it.  alpha >= 1/255 needs rho = min(rho3d, rho2d) <= 2 ln(255 o): either the
// low-pass disc of radius sqrt(ln(255 o)) around c, or the projected ellipse rho3d <= 2 ln(255 o), whose bounding box
// follows from the same formula as the 3-sigma box.  +inf when the ellipse is unbounded on screen (the surfel's plane
// passes close to the eye); -1 when the opacity is too low to ever contribute.  The blend kernels use it to let a warp
// skip the records that cannot touch its 8x4 pixel block; the slack dwarfs every rounding error of the exact tests.
__device__ __forceinline__ float surfel_reach(const float Tu[3], const float Tv[3], const float Tw[3], float2 c,
                                              float opacity) {
    if (!(opacity * 255.0f > 1.0f)) return opacity == opacity ? -1.0f : INFINITY;  // NaN opacity: never skip
    const float rho_max = 2.0f * logf(opacity * 255.0f) + 1e-3f;
    const float r2 = sqrtf(0.5f * rho_max);
    float2 ce, ex;
    float dist;
    if (!surfel_aabb_at(Tu, Tv, Tw, rho_max, ce, ex, dist) || !(dist < 0.f)) return INFINITY;
    const float h3 = fmaxf(fabsf(ce.x - c.x) + ex.x, fabsf(ce.y - c.y) + ex.y);
    const float h = fmaxf(r2, h3);
    return h * 1.001f + 0.05f;
}

// Bounding box of the 3-sigma ellipse of the homography T (rows Tu, Tv, Tw).
__device__ __forceinline__ bool surfel_aabb(const float Tu[3], const float Tv[3], const float Tw[3], float2& centre,
                                            float2& extent) {
    float dist;
    return surfel_aabb_at(Tu, Tv, Tw, SURFEL_CUTOFF * SURFEL_CUTOFF, centre, extent, dist);
}

__global__ void __launch_bounds__(SP_THREADS) surfel_project_kernel(const SurfelProjectArgs a) {
    __shared__ EmitRec s_emit[SP_THREADS];
    __shared__ uint32_t s_tot[2];
    float Q[4][3];
    world_to_pixel_hom(a.proj, a.W, a.H, Q);
    EmitTarget target;
    target.tile_count = a.img.tile_count;
    target.keys = a.keys;
    target.tile_cap = a.tile_cap;
    target.tile_base = a.tile_base;
    target.gx = a.gx;
    if (threadIdx.x == 0) s_tot[0] = s_tot[1] = 0u;
    __syncthreads();
    uint32_t kept = 0, max_fill = 0;
    const int n_vblocks = (a.P + SP_THREADS - 1) / SP_THREADS;
    for (int vb = blockIdx.x; vb < n_vblocks; vb += gridDim.x) {
        const int idx = vb * SP_THREADS + threadIdx.x;
        int n_tiles = 0, rx0 = 0, ry0 = 0, rw = 0;
        uint32_t depth_bits = 0;
        if (idx < a.P) {
            int radius_out = 0;
            unsigned clamp_bits = 0;
            Surfel rec;
            rec.r0 = make_float4(0.f, 0.f, 0.f, __int_as_float(idx));
            rec.r1 = rec.r2 = rec.r3 = rec.r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float3 p = make_float3(__ldg(a.means3D + 3 * (size_t)idx), __ldg(a.means3D + 3 * (size_t)idx + 1),
                                         __ldg(a.means3D + 3 * (size_t)idx + 2));
            const float3 pv = xform_point_4x3(p, a.view);
            if (!(pv.z <= SURFEL_NEAR)) {  // the near-plane test of in_frustum: a NaN depth passes, as in the 3DGS path
                float Tu[3], Tv[3], Tw[3];
                float3 normal;
                if (a.transmat_precomp) {
                    const float* t = a.transmat_precomp + 9 * (size_t)idx;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        Tu[k] = __ldg(t + k);
                        Tv[k] = __ldg(t + 3 + k);
                        Tw[k] = __ldg(t + 6 + k);
                    }
                    normal = make_float3(0.f, 0.f, 1.f);
                } else {
                    const float4 qn = quat_normalised(__ldg(reinterpret_cast<const float4*>(a.rotations) + idx));
                    const Rot3 R = quat_to_rot(qn);
                    const float su = a.scale_modifier * __ldg(a.scales + (size_t)idx * a.scale_stride);
                    const float sv = a.scale_modifier * __ldg(a.scales + (size_t)idx * a.scale_stride + 1);
                    const float pw[3] = {p.x, p.y, p.z};
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float tu = 0.f, tv = 0.f, tc = Q[3][c];
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            tu = fmaf(R.m[i][0] * su, Q[i][c], tu);
                            tv = fmaf(R.m[i][1] * sv, Q[i][c], tv);
                            tc = fmaf(pw[i], Q[i][c], tc);
                        }
                        float* row = c == 0 ? Tu : (c == 1 ? Tv : Tw);
                        row[0] = tu;
                        row[1] = tv;
                        row[2] = tc;
                    }
                    // view-space normal = third column of R through the rotation part of the view matrix
                    normal = make_float3(
                        __ldg(a.view + 0) * R.m[0][2] + __ldg(a.view + 4) * R.m[1][2] + __ldg(a.view + 8) * R.m[2][2],
                        __ldg(a.view + 1) * R.m[0][2] + __ldg(a.view + 5) * R.m[1][2] + __ldg(a.view + 9) * R.m[2][2],
                        __ldg(a.view + 2) * R.m[0][2] + __ldg(a.view + 6) * R.m[1][2] + __ldg(a.view + 10) * R.m[2][2]);
                }
                const float cosv = -(pv.x * normal.x + pv.y * normal.y + pv.z * normal.z);
                float2 centre, extent;
                if (cosv != 0.f && surfel_aabb(Tu, Tv, Tw, centre, extent)) {
                    const float mult = cosv > 0.f ? 1.f : -1.f;  // both faces are visible: flip towards the camera
                    const float radius = ceilf(fmaxf(fmaxf(extent.x, extent.y), SURFEL_CUTOFF * SURFEL_FILTER_SIZE));
                    int x0, y0, x1, y1;
                    tile_rect(centre.x, centre.y, (int)radius, a.gx, a.gy, x0, y0, x1, y1);
                    if ((x1 - x0) * (y1 - y0) != 0) {
                        float3 rgb;
                        if (a.colors_precomp) {
                            rgb = make_float3(__ldg(a.colors_precomp + 3 * (size_t)idx),
                                              __ldg(a.colors_precomp + 3 * (size_t)idx + 1),
                                              __ldg(a.colors_precomp + 3 * (size_t)idx + 2));
                        } else {
                            const float3 cam = make_float3(__ldg(a.campos), __ldg(a.campos + 1), __ldg(a.campos + 2));
                            rgb = sh::eval(a.sh_degree, p, cam, a.shs + (size_t)idx * 3 * a.M, clamp_bits);
                        }
                        radius_out = (int)radius;
                        rec.r0 = make_float4(centre.x, centre.y, __ldg(a.opacities + idx), __int_as_float(idx));
                        rec.r1 = make_float4(Tu[0], Tu[1], Tu[2], rgb.x);
                        rec.r2 = make_float4(Tv[0], Tv[1], Tv[2], rgb.y);
                        rec.r3 = make_float4(Tw[0], Tw[1], Tw[2], rgb.z);
                        rec.r4 = make_float4(mult * normal.x, mult * normal.y, mult * normal.z,
                                             surfel_reach(Tu, Tv, Tw, centre, __ldg(a.opacities + idx)));
                        rx0 = x0;
                        ry0 = y0;
                        rw = x1 - x0;
                        n_tiles = (y1 - y0) * (x1 - x0);
                    }
                }
            }
            a.radii[idx] = radius_out;
            a.surfel[idx] = rec;
            // debug introspection reads the centre and the depth from the 3DGS-shaped state
            Splat b;
            b.q0 = make_float4(rec.r0.x, rec.r0.y, 0.f, __int_as_float(idx));
            b.q1 = make_float4(0.f, 0.f, 0.f, 0.f);
            b.q2 = make_float4(0.f, 0.f, 0.f, pv.z);
            a.geom.splat[idx] = b;
            a.geom.tiles_touched[idx] = (uint32_t)n_tiles;
            a.geom.clamped[idx] = (uint8_t)clamp_bits;
            depth_bits = __float_as_uint(pv.z);
        }
        // bin the instances (no tile-level culling on this path): one slot claim + key per (surfel, tile) pair
        const float4 unused = make_float4(0.f, 0.f, 0.f, 0.f);
        warp_emit_tiles<false>(s_emit + (threadIdx.x & ~31u), n_tiles, rx0, ry0, rw, unused, unused, depth_bits,
                               (uint32_t)idx, target, kept, max_fill);
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    max_fill = __reduce_max_sync(0xffffffffu, max_fill);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_tot[0], kept);
        atomicMax(&s_tot[1], max_fill);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_tot[0]) atomicAdd(&a.img.header[HDR_NUM_RENDERED], s_tot[0]);
        if (s_tot[1]) atomicMax(&a.img.header[HDR_MAX_TILE], s_tot[1]);
        if (a.counts_host) report_counts(a.img.header, a.counts_host, gridDim.x);
    }
    pdl_trigger();  // tile_sort may start launching
}

// ---------------------------------------------------------------------------------------------------
// Blend kernels: one CTA per 16x16 tile, one pixel per thread (a warp owns an 8x4 block), the tile's
// depth-sorted records staged in 128-record chunks with cp.async.bulk into a double-buffered ring.
constexpr int SB_THREADS = 256;
constexpr int SCHUNK = 128;

struct SurfelSmem {
    Surfel buf[2][SCHUNK];
    uint64_t full[2];
    uint32_t warp_max[SB_THREADS / 32];
};

struct PairEval {  // geometry of one (pixel, surfel) pair
    float sx, sy, pz, depth, G, alpha;
    float3 k, l;
    bool use3d;
};

// Returns false when the pair is skipped (same tests, same order as the forward).
__device__ __forceinline__ bool surfel_pair(const Surfel& s, float px, float py, PairEval& e) {
    const float3 Tu = make_float3(s.r1.x, s.r1.y, s.r1.z), Tv = make_float3(s.r2.x, s.r2.y, s.r2.z);
    const float3 Tw = make_float3(s.r3.x, s.r3.y, s.r3.z);
    e.k = make_float3(px * Tw.x - Tu.x, px * Tw.y - Tu.y, px * Tw.z - Tu.z);
    e.l = make_float3(py * Tw.x - Tv.x, py * Tw.y - Tv.y, py * Tw.z - Tv.z);
    const float3 p = make_float3(e.k.y * e.l.z - e.k.z * e.l.y, e.k.z * e.l.x - e.k.x * e.l.z, e.k.x * e.l.y - e.k.y * e.l.x);
    if (p.z == 0.0f) return false;
    e.pz = p.z;
    e.sx = p.x / p.z;
    e.sy = p.y / p.z;
    const float rho3d = e.sx * e.sx + e.sy * e.sy;
    const float dx = s.r0.x - px, dy = s.r0.y - py;
    const float rho2d = SURFEL_FILTER_INV_SQUARE * (dx * dx + dy * dy);
    e.use3d = rho3d <= rho2d;
    const float rho = fminf(rho3d, rho2d);
    e.depth = e.use3d ? (e.sx * Tw.x + e.sy * Tw.y) + Tw.z : Tw.z;
    if (e.depth < SURFEL_NEAR) return false;
    const float power = -0.5f * rho;
    if (power > 0.0f) return false;
    e.G = expf(power);
    e.alpha = fminf(0.99f, s.r0.z * e.G);
    return !(e.alpha < ALPHA_MIN);
}

// Bit i of words[k] set iff record 32 k + i of the chunk can reach the warp's 8x4 pixel block: the block's bit of the
// mask tile_sort left in the stream copy's r4.w (surfel_region_mask8, surfel.cuh).  all_blocks: ignore the masks
// (GDR_SURFEL_MASKS=0; the tests compare the two modes bit for bit).
__device__ __forceinline__ void classify_chunk(const Surfel* sp, int cnt, int lane, int warp, bool all_blocks,
                                               unsigned (&words)[SCHUNK / 32]) {
#pragma unroll
    for (int k = 0; k < SCHUNK / 32; k++) {
        const int j = k * 32 + lane;
        const bool hit = j < cnt && (all_blocks || ((__float_as_uint(sp[j].r4.w) >> warp) & 1u));
        words[k] = __ballot_sync(0xffffffffu, hit);
    }
}

__global__ void __launch_bounds__(SB_THREADS)
surfel_blend_forward_kernel(int W, int H, int gx, int n_tiles_total, ImageState img, const Surfel* __restrict__ stream, int64_t capacity,
                            const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_allmap,
                            float* __restrict__ aux, const int all_blocks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SurfelSmem& sm = *reinterpret_cast<SurfelSmem*>(smem_raw);
    pdl_wait();  // launched as a programmatic dependent of tile_sort
    const int tile = tile_from_order(img.header, img.order, n_tiles_total, (int)blockIdx.x);
    const int tile_x = tile % gx, tile_y = tile / gx;
    const uint2 range = img.tile_range[tile];
    const int64_t rb = (int64_t)range.x;
    const int n = (int)(range.y - range.x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tile_x * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    if (threadIdx.x == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int n_chunks = (n + SCHUNK - 1) / SCHUNK;
    const Surfel* src = stream + rb;
    auto issue = [&](int it) {
        const int cnt = min(SCHUNK, n - it * SCHUNK);
        const uint32_t bytes = (uint32_t)(cnt * sizeof(Surfel));
        mbar_expect_tx(&sm.full[it & 1], bytes);
        bulk_g2s(&sm.buf[it & 1][0], src + (size_t)it * SCHUNK, bytes, &sm.full[it & 1]);
    };
    if (threadIdx.x == 0 && n_chunks > 0) issue(0);

    bool done = !inside;
    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f;
    float D = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f, median_depth = 0.f;
    uint32_t contributor = 0, last_contributor = 0, median_contributor = 0;

    for (int it = 0; it < n_chunks; it++) {
        if (threadIdx.x == 0 && it + 1 < n_chunks) issue(it + 1);
        mbar_wait(&sm.full[it & 1], (it >> 1) & 1);
        const int cnt = min(SCHUNK, n - it * SCHUNK);
        const Surfel* sp = &sm.buf[it & 1][0];
        // Each warp walks only the records that can reach its 8x4 block.  (The walk has no divergent break /
        // continue: with them the lanes do not reconverge until the loop ends -- ncu showed 2 active threads per
        // instruction and a 15x slower kernel.)
        unsigned words[SCHUNK / 32];
        classify_chunk(sp, cnt, lane, warp, all_blocks != 0, words);
#pragma unroll
        for (int k = 0; k < SCHUNK / 32; k++) {
            unsigned word = words[k];
            while (word) {
                if (__all_sync(0xffffffffu, done)) break;  // warp-uniform
                const int j = k * 32 + __ffs(word) - 1;
                word &= word - 1;
                PairEval e;
                const Surfel& s = sp[j];
                const bool hit = !done && surfel_pair(s, pxf, pyf, e);
                if (hit) {
                    contributor = (uint32_t)(it * SCHUNK + j + 1);
                    const float test_T = T * (1.f - e.alpha);
                    if (test_T < T_MIN) {
                        done = true;
                    } else {
                        const float w = e.alpha * T;
                        // depth distortion (paper appendix): sum_i sum_{k<i} w_i w_k (m_i - m_k)^2 in one pass
                        const float A = 1.f - T;
                        const float m = SURFEL_FAR / (SURFEL_FAR - SURFEL_NEAR) * (1.f - SURFEL_NEAR / e.depth);
                        distortion += (m * m * A + M2 - 2.f * m * M1) * w;
                        D += e.depth * w;
                        M1 += m * w;
                        M2 += m * m * w;
                        if (T > 0.5f) {
                            median_depth = e.depth;
                            median_contributor = contributor;
                        }
                        N0 = fmaf(s.r4.x, w, N0);
                        N1 = fmaf(s.r4.y, w, N1);
                        N2 = fmaf(s.r4.z, w, N2);
                        C0 = fmaf(s.r1.w, w, C0);
                        C1 = fmaf(s.r2.w, w, C1);
                        C2 = fmaf(s.r3.w, w, C2);
                        T = test_T;
                        last_contributor = contributor;
                    }
                }
            }
        }
        if (__syncthreads_and(done)) {  // also: everyone is finished with buf[it & 1]
            if (it + 1 < n_chunks) mbar_wait(&sm.full[(it + 1) & 1], ((it + 1) >> 1) & 1);  // drain the copy in flight
            break;
        }
    }
    if (inside) {
        const size_t HW = (size_t)H * W, pid = (size_t)py * W + px;
        img.n_contrib[pid] = last_contributor;
        out_color[pid] = C0 + T * __ldg(bg);
        out_color[HW + pid] = C1 + T * __ldg(bg + 1);
        out_color[2 * HW + pid] = C2 + T * __ldg(bg + 2);
        out_allmap[AM_DEPTH * HW + pid] = D;
        out_allmap[AM_ALPHA * HW + pid] = 1.f - T;
        out_allmap[(AM_NORMAL + 0) * HW + pid] = N0;
        out_allmap[(AM_NORMAL + 1) * HW + pid] = N1;
        out_allmap[(AM_NORMAL + 2) * HW + pid] = N2;
        out_allmap[AM_MIDDEPTH * HW + pid] = median_depth;
        out_allmap[AM_DISTORTION * HW + pid] = distortion;
        aux[pid] = __uint_as_float(median_contributor);
        aux[HW + pid] = M1;
        aux[2 * HW + pid] = M2;
    }
}

// Sum of v[i] over the warp's lanes, delivered to lane i (transposing butterfly: 31 shuffles for 32 values).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(SB_THREADS)
surfel_blend_backward_kernel(int W, int H, int gx, int n_tiles_total, ImageState img, const Surfel* __restrict__ stream, int64_t capacity,
                             const float* __restrict__ bg, const float* __restrict__ out_allmap,
                             const float* __restrict__ aux, const float* __restrict__ dL_dcolor,
                             const float* __restrict__ dL_dallmap, float* __restrict__ accum, const int all_blocks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SurfelSmem& sm = *reinterpret_cast<SurfelSmem*>(smem_raw);
    const int tile = tile_from_order(img.header, img.order, n_tiles_total, (int)blockIdx.x);
    const int tile_x = tile % gx, tile_y = tile / gx;
    const uint2 range = img.tile_range[tile];
    const int64_t rb = (int64_t)range.x;
    const int n_all = (int)(range.y - range.x);
    if (n_all == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tile_x * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)H * W, pid = (size_t)py * W + px;

    const uint32_t last_contributor = inside ? img.n_contrib[pid] : 0u;
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) sm.warp_max[warp] = warp_last;
    if (threadIdx.x == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t tile_max = 0;
#pragma unroll
    for (int i = 0; i < SB_THREADS / 32; i++) tile_max = max(tile_max, sm.warp_max[i]);
    const int n = min(n_all, (int)tile_max);
    if (n == 0) return;
    const int n_chunks = (n + SCHUNK - 1) / SCHUNK;
    const Surfel* src = stream + rb;
    auto issue = [&](int it) {  // iteration `it` handles chunk n_chunks - 1 - it
        const int ch = n_chunks - 1 - it;
        const int cnt = min(SCHUNK, n - ch * SCHUNK);
        const uint32_t bytes = (uint32_t)(cnt * sizeof(Surfel));
        mbar_expect_tx(&sm.full[it & 1], bytes);
        bulk_g2s(&sm.buf[it & 1][0], src + (size_t)ch * SCHUNK, bytes, &sm.full[it & 1]);
    };
    if (threadIdx.x == 0) issue(0);

    float dC0 = 0.f, dC1 = 0.f, dC2 = 0.f, dD = 0.f, dA = 0.f, dN0 = 0.f, dN1 = 0.f, dN2 = 0.f, dMed = 0.f, dDist = 0.f;
    float A_tot = 0.f, M1_tot = 0.f, M2_tot = 0.f;
    uint32_t median_contributor = 0;
    if (inside) {
        dC0 = dL_dcolor[pid];
        dC1 = dL_dcolor[HW + pid];
        dC2 = dL_dcolor[2 * HW + pid];
        if (dL_dallmap) {
            dD = dL_dallmap[AM_DEPTH * HW + pid];
            dA = dL_dallmap[AM_ALPHA * HW + pid];
            dN0 = dL_dallmap[(AM_NORMAL + 0) * HW + pid];
            dN1 = dL_dallmap[(AM_NORMAL + 1) * HW + pid];
            dN2 = dL_dallmap[(AM_NORMAL + 2) * HW + pid];
            dMed = dL_dallmap[AM_MIDDEPTH * HW + pid];
            dDist = dL_dallmap[AM_DISTORTION * HW + pid];
        }
        A_tot = out_allmap[AM_ALPHA * HW + pid];
        median_contributor = __float_as_uint(aux[pid]);
        M1_tot = aux[HW + pid];
        M2_tot = aux[2 * HW + pid];
    }
    const float T_final = 1.f - A_tot;
    float T = T_final;
    const float bg_dot = __ldg(bg) * dC0 + __ldg(bg + 1) * dC1 + __ldg(bg + 2) * dC2;
    const float neg_Tf_bg = -T_final * bg_dot;
    float beta = 0.f;  // sum over the pairs behind the current one of g_k w_k, divided by the transmittance behind it
    constexpr float dm_scale = SURFEL_FAR / (SURFEL_FAR - SURFEL_NEAR);

    for (int it = 0; it < n_chunks; it++) {
        if (threadIdx.x == 0 && it + 1 < n_chunks) issue(it + 1);
        mbar_wait(&sm.full[it & 1], (it >> 1) & 1);
        const int ch = n_chunks - 1 - it;
        const int cnt = min(SCHUNK, n - ch * SCHUNK);
        const Surfel* sp = &sm.buf[it & 1][0];
        int j_hi = cnt - 1;
        if ((uint32_t)(ch * SCHUNK + cnt) > warp_last) j_hi = (int)warp_last - ch * SCHUNK - 1;
        unsigned words[SCHUNK / 32];
        classify_chunk(sp, j_hi + 1, lane, warp, all_blocks != 0, words);  // j_hi < 0: nothing to do
#pragma unroll
        for (int k = SCHUNK / 32 - 1; k >= 0; k--) {
        unsigned word = words[k];
        while (word) {
            const int j = k * 32 + 31 - __clz(word);  // back to front
            word &= ~(1u << (j & 31));
            const uint32_t pos = (uint32_t)(ch * SCHUNK + j);
            const Surfel& s = sp[j];
            PairEval e;
            const bool contrib = (pos < last_contributor) && surfel_pair(s, pxf, pyf, e);
            if (!__any_sync(0xffffffffu, contrib)) continue;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = 0.f;
            if (contrib) {
                const float inv_1ma = 1.f / (1.f - e.alpha);
                T = T * inv_1ma;  // transmittance in front of this pair
                const float w = e.alpha * T;
                const float m = dm_scale * (1.f - SURFEL_NEAR / e.depth);
                // g = dL/dw with every w independent
                float g = s.r1.w * dC0 + s.r2.w * dC1 + s.r3.w * dC2 + s.r4.x * dN0 + s.r4.y * dN1 + s.r4.z * dN2 +
                          e.depth * dD + dA + dDist * (m * m * A_tot - 2.f * m * M1_tot + M2_tot);
                const float dL_dalpha = T * (g - beta) + inv_1ma * neg_Tf_bg;
                beta = fmaf(e.alpha, g - beta, beta);
                // colour, normal
                v[12] = w * dC0;
                v[13] = w * dC1;
                v[14] = w * dC2;
                v[15] = w * dN0;
                v[16] = w * dN1;
                v[17] = w * dN2;
                // depth of the intersection
                const float dm_dd = dm_scale * SURFEL_NEAR / (e.depth * e.depth);
                float dL_dz = w * dD + dDist * 2.f * w * (m * A_tot - M1_tot) * dm_dd;
                if (pos + 1 == median_contributor) dL_dz += dMed;
                v[11] = e.G * dL_dalpha;                // dL/dopacity (the 0.99 clamp is not masked, as in the 3DGS path)
                const float dL_dG = s.r0.z * dL_dalpha;
                const float3 Tw = make_float3(s.r3.x, s.r3.y, s.r3.z);
                if (e.use3d) {
                    const float gsx = dL_dG * -e.G * e.sx + dL_dz * Tw.x;
                    const float gsy = dL_dG * -e.G * e.sy + dL_dz * Tw.y;
                    const float ipz = 1.f / e.pz;
                    const float3 dp = make_float3(gsx * ipz, gsy * ipz, -(gsx * e.sx + gsy * e.sy) * ipz);
                    // p = k x l
                    const float3 dk = make_float3(e.l.y * dp.z - e.l.z * dp.y, e.l.z * dp.x - e.l.x * dp.z,
                                                  e.l.x * dp.y - e.l.y * dp.x);
                    const float3 dl = make_float3(dp.y * e.k.z - dp.z * e.k.y, dp.z * e.k.x - dp.x * e.k.z,
                                                  dp.x * e.k.y - dp.y * e.k.x);
                    v[0] = -dk.x; v[1] = -dk.y; v[2] = -dk.z;
                    v[3] = -dl.x; v[4] = -dl.y; v[5] = -dl.z;
                    v[6] = pxf * dk.x + pyf * dl.x + dL_dz * e.sx;
                    v[7] = pxf * dk.y + pyf * dl.y + dL_dz * e.sy;
                    v[8] = pxf * dk.z + pyf * dl.z + dL_dz;
                    v[18] = fabsf(dk.z);
                    v[19] = fabsf(dl.z);
                } else {
                    const float c = dL_dG * -e.G * SURFEL_FILTER_INV_SQUARE;
                    v[9] = c * (s.r0.x - pxf);
                    v[10] = c * (s.r0.y - pyf);
                    v[8] = dL_dz;
                }
            }
            const float total = warp_transpose_reduce(v, lane);
            if (lane < SURFEL_ACC && total != 0.f)
                atomicAdd(accum + (size_t)__float_as_int(s.r0.w) * SURFEL_ACC + lane, total);
        }
        }
        __syncthreads();  // everyone is finished with buf[it & 1]
    }
}

// ---------------------------------------------------------------------------------------------------
struct SurfelGaussBackwardArgs {
    int P, sh_degree, M, W, H;
    const float *means3D, *shs, *colors_precomp, *scales;
    int scale_stride;
    float scale_modifier;
    const float *rotations, *transmat_precomp;
    const float *view, *proj, *campos;
    const int32_t* radii;
    const Surfel* surfel;
    const uint8_t* clamped;
    const float* accum;
    int means2D_cols;  // 3 or 4 columns of dL_dmeans2D
    float *dL_dmeans2D, *dL_dcolors, *dL_dopacity, *dL_dmeans3D, *dL_dtransmat, *dL_dsh, *dL_dscales, *dL_drotations;
};

__global__ void __launch_bounds__(SP_THREADS) surfel_gauss_backward_kernel(const SurfelGaussBackwardArgs a) {
    const int idx = blockIdx.x * SP_THREADS + threadIdx.x;
    if (idx >= a.P) return;
    const bool visible = a.radii[idx] > 0;
    float acc[SURFEL_ACC];
#pragma unroll
    for (int i = 0; i < SURFEL_ACC; i++) acc[i] = visible ? a.accum[(size_t)idx * SURFEL_ACC + i] : 0.f;
    const Surfel s = a.surfel[idx];
    const float Tu[3] = {s.r1.x, s.r1.y, s.r1.z}, Tv[3] = {s.r2.x, s.r2.y, s.r2.z}, Tw[3] = {s.r3.x, s.r3.y, s.r3.z};
    float dTu[3] = {acc[0], acc[1], acc[2]}, dTv[3] = {acc[3], acc[4], acc[5]}, dTw[3] = {acc[6], acc[7], acc[8]};

    // screen-space statistic for densification: the gradient w.r.t. the homogeneous pixel offsets of the
    // centre, scaled to NDC; columns 2:4 carry the sums of absolute per-pixel values.
    if (a.dL_dmeans2D) {
        float* o = a.dL_dmeans2D + (size_t)idx * a.means2D_cols;
        const float depth = Tw[2];
        o[0] = acc[2] * depth * 0.5f * a.W;
        o[1] = acc[5] * depth * 0.5f * a.H;
        if (a.means2D_cols == 3) {
            o[2] = 0.f;
        } else {
            o[2] = acc[18] * fabsf(depth) * 0.5f * a.W;
            o[3] = acc[19] * fabsf(depth) * 0.5f * a.H;
        }
    }
    if (a.dL_dopacity) a.dL_dopacity[idx] = acc[11];

    // colour
    const float3 p = make_float3(a.means3D[3 * (size_t)idx], a.means3D[3 * (size_t)idx + 1], a.means3D[3 * (size_t)idx + 2]);
    float3 dmean = make_float3(0.f, 0.f, 0.f);
    if (a.colors_precomp) {
        if (a.dL_dcolors) {
            a.dL_dcolors[3 * (size_t)idx] = acc[12];
            a.dL_dcolors[3 * (size_t)idx + 1] = acc[13];
            a.dL_dcolors[3 * (size_t)idx + 2] = acc[14];
        }
    } else if (a.dL_dsh || a.dL_dmeans3D) {
        const unsigned cl = a.clamped[idx];
        const float3 drgb = make_float3((cl & 1u) ? 0.f : acc[12], (cl & 2u) ? 0.f : acc[13], (cl & 4u) ? 0.f : acc[14]);
        float dsh[48];
        const int deg = a.sh_degree;
        const float3 cam = make_float3(__ldg(a.campos), __ldg(a.campos + 1), __ldg(a.campos + 2));
        dmean = sh::backward(deg, p, cam, a.shs + (size_t)idx * 3 * a.M, drgb, dsh);
        if (a.dL_dsh) {
            float* o = a.dL_dsh + (size_t)idx * 3 * a.M;
            const int used = 3 * (deg + 1) * (deg + 1);
            for (int i = 0; i < 3 * a.M; i++) o[i] = (visible && i < used) ? dsh[i] : 0.f;
        }
        if (!visible) dmean = make_float3(0.f, 0.f, 0.f);
    }

    // bounding-box centre (the low-pass branch measures its distance to the pixel) -> homography
    {
        const float t[3] = {SURFEL_CUTOFF * SURFEL_CUTOFF, SURFEL_CUTOFF * SURFEL_CUTOFF, -1.0f};
        const float d = t[0] * Tw[0] * Tw[0] + t[1] * Tw[1] * Tw[1] + t[2] * Tw[2] * Tw[2];
        if (visible && d != 0.f && (acc[9] != 0.f || acc[10] != 0.f)) {
            const float inv = 1.f / d;
            float dL_dd = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float f = t[k] * inv;
                dTu[k] += acc[9] * f * Tw[k];
                dTv[k] += acc[10] * f * Tw[k];
                dTw[k] += acc[9] * f * Tu[k] + acc[10] * f * Tv[k];
                const float dL_df = acc[9] * Tu[k] * Tw[k] + acc[10] * Tv[k] * Tw[k];
                dL_dd -= dL_df * f * inv;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) dTw[k] += dL_dd * 2.f * t[k] * Tw[k];
        }
    }
    if (a.transmat_precomp) {
        if (a.dL_dtransmat) {
            float* o = a.dL_dtransmat + 9 * (size_t)idx;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                o[k] = dTu[k];
                o[3 + k] = dTv[k];
                o[6 + k] = dTw[k];
            }
        }
        if (a.dL_dmeans3D) {
            a.dL_dmeans3D[3 * (size_t)idx] = dmean.x;
            a.dL_dmeans3D[3 * (size_t)idx + 1] = dmean.y;
            a.dL_dmeans3D[3 * (size_t)idx + 2] = dmean.z;
        }
        return;
    }

    // homography -> splat-to-world matrix M (columns: tangent u * su, tangent v * sv, centre)
    float Q[4][3];
    world_to_pixel_hom(a.proj, a.W, a.H, Q);
    float dL0[3], dL1[3], dP[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        dL0[i] = dTu[0] * Q[i][0] + dTv[0] * Q[i][1] + dTw[0] * Q[i][2];
        dL1[i] = dTu[1] * Q[i][0] + dTv[1] * Q[i][1] + dTw[1] * Q[i][2];
        dP[i] = dTu[2] * Q[i][0] + dTv[2] * Q[i][1] + dTw[2] * Q[i][2];
    }
    if (a.dL_dmeans3D) {
        a.dL_dmeans3D[3 * (size_t)idx] = dP[0] + dmean.x;
        a.dL_dmeans3D[3 * (size_t)idx + 1] = dP[1] + dmean.y;
        a.dL_dmeans3D[3 * (size_t)idx + 2] = dP[2] + dmean.z;
    }
    if (!a.dL_dscales && !a.dL_drotations) return;
    const float4 qn = quat_normalised(__ldg(reinterpret_cast<const float4*>(a.rotations) + idx));
    const Rot3 R = quat_to_rot(qn);
    const float su = a.scale_modifier * a.scales[(size_t)idx * a.scale_stride];
    const float sv = a.scale_modifier * a.scales[(size_t)idx * a.scale_stride + 1];
    if (a.dL_dscales) {
        float* o = a.dL_dscales + (size_t)idx * a.scale_stride;
        o[0] = a.scale_modifier * (R.m[0][0] * dL0[0] + R.m[1][0] * dL0[1] + R.m[2][0] * dL0[2]);
        o[1] = a.scale_modifier * (R.m[0][1] * dL1[0] + R.m[1][1] * dL1[1] + R.m[2][1] * dL1[2]);
        for (int k = 2; k < a.scale_stride; k++) o[k] = 0.f;
    }
    if (a.dL_drotations) {
        // normal (view space, flipped towards the camera) -> third column of R
        const float3 pv = xform_point_4x3(p, a.view);
        const float3 n0 = make_float3(
            __ldg(a.view + 0) * R.m[0][2] + __ldg(a.view + 4) * R.m[1][2] + __ldg(a.view + 8) * R.m[2][2],
            __ldg(a.view + 1) * R.m[0][2] + __ldg(a.view + 5) * R.m[1][2] + __ldg(a.view + 9) * R.m[2][2],
            __ldg(a.view + 2) * R.m[0][2] + __ldg(a.view + 6) * R.m[1][2] + __ldg(a.view + 10) * R.m[2][2]);
        const float mult = -(pv.x * n0.x + pv.y * n0.y + pv.z * n0.z) > 0.f ? 1.f : -1.f;
        float dR[3][3];  // dL/dR[row][col]
#pragma unroll
        for (int i = 0; i < 3; i++) {
            dR[i][0] = su * dL0[i];
            dR[i][1] = sv * dL1[i];
            dR[i][2] = mult * (acc[15] * __ldg(a.view + 4 * i) + acc[16] * __ldg(a.view + 4 * i + 1) +
                               acc[17] * __ldg(a.view + 4 * i + 2));
        }
        const float r = qn.x, x = qn.y, y = qn.z, z = qn.w;
        float4 dq;  // w.r.t. the normalised quaternion (callers pass normalised rotations)
        dq.x = 2.f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
        dq.y = 2.f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.f * x * dR[1][1] - r * dR[1][2] + z * dR[2][0] +
                      r * dR[2][1] - 2.f * x * dR[2][2]);
        dq.z = 2.f * (-2.f * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] - r * dR[2][0] +
                      z * dR[2][1] - 2.f * y * dR[2][2]);
        dq.w = 2.f * (-2.f * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2.f * z * dR[1][1] + y * dR[1][2] +
                      x * dR[2][0] + y * dR[2][1]);
        reinterpret_cast<float4*>(a.dL_drotations)[idx] = visible ? dq : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace

// ---- host launchers --------------------------------------------------------------------------------
cudaError_t launch_surfel_project(int P, int sh_degree, int M, int W, int H, const float* means3D, const float* shs,
                                  const float* colors_precomp, const float* opacities, const float* scales,
                                  int scale_stride, float scale_modifier, const float* rotations,
                                  const float* transmat_precomp, const float* view, const float* proj,
                                  const float* campos, int32_t* radii, GeomState geom, void* surfel_state,
                                  ImageState img, uint64_t* keys, int64_t tile_cap, const uint32_t* tile_base,
                                  int32_t* counts_host, cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    SurfelProjectArgs a;
    a.P = P; a.sh_degree = sh_degree; a.M = M; a.W = W; a.H = H;
    a.gx = (W + TILE - 1) / TILE; a.gy = (H + TILE - 1) / TILE;
    a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.opacities = opacities; a.scales = scales;
    a.scale_stride = scale_stride; a.scale_modifier = scale_modifier; a.rotations = rotations;
    a.transmat_precomp = transmat_precomp; a.view = view; a.proj = proj; a.campos = campos;
    a.radii = radii; a.geom = geom; a.surfel = (Surfel*)surfel_state; a.img = img;
    a.keys = keys; a.tile_cap = (uint32_t)tile_cap; a.tile_base = tile_base; a.counts_host = counts_host;
    const int n_vblocks = (P + SP_THREADS - 1) / SP_THREADS;
    const int grid = min(n_vblocks, sm_count() * 8);
    surfel_project_kernel<<<grid, SP_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

// GDR_SURFEL_MASKS=0: the blend kernels ignore the block masks and walk every record of the tile (read per launch,
// so a test can compare both modes in one process)
static int surfel_masks_off() {
    const char* e = getenv("GDR_SURFEL_MASKS");
    return (e && e[0] == '0') ? 1 : 0;
}

cudaError_t launch_surfel_blend_forward(int W, int H, ImageState img, const void* stream, int64_t capacity,
                                        const float* bg, float* out_color, float* out_allmap, float* aux,
                                        cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    cudaError_t e = cudaFuncSetAttribute(surfel_blend_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(SurfelSmem));
    if (e != cudaSuccess) return e;
    surfel_blend_forward_kernel<<<gx * gy, SB_THREADS, sizeof(SurfelSmem), s>>>(
        W, H, gx, gx * gy, img, (const Surfel*)stream, capacity, bg, out_color, out_allmap, aux, surfel_masks_off());
    return cudaGetLastError();
}

cudaError_t launch_surfel_blend_backward(int W, int H, ImageState img, const void* stream, int64_t capacity,
                                         const float* bg, const float* out_allmap, const float* aux,
                                         const float* dL_dcolor, const float* dL_dallmap, float* accum,
                                         cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    cudaError_t e = cudaFuncSetAttribute(surfel_blend_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(SurfelSmem));
    if (e != cudaSuccess) return e;
    surfel_blend_backward_kernel<<<gx * gy, SB_THREADS, sizeof(SurfelSmem), s>>>(
        W, H, gx, gx * gy, img, (const Surfel*)stream, capacity, bg, out_allmap, aux, dL_dcolor, dL_dallmap, accum,
        surfel_masks_off());
    return cudaGetLastError();
}

cudaError_t launch_surfel_gauss_backward(int P, int sh_degree, int M, int W, int H, const float* means3D,
                                         const float* shs, const float* colors_precomp, const float* scales,
                                         int scale_stride, float scale_modifier, const float* rotations,
                                         const float* transmat_precomp, const float* view, const float* proj,
                                         const float* campos, const int32_t* radii, const void* surfel_state,
                                         const uint8_t* clamped, const float* accum, int means2D_cols,
                                         float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D,
                                         float* dL_dtransmat, float* dL_dsh, float* dL_dscales, float* dL_drotations,
                                         cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    SurfelGaussBackwardArgs a;
    a.P = P; a.sh_degree = sh_degree; a.M = M; a.W = W; a.H = H;
    a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.scales = scales;
    a.scale_stride = scale_stride; a.scale_modifier = scale_modifier; a.rotations = rotations;
    a.transmat_precomp = transmat_precomp; a.view = view; a.proj = proj; a.campos = campos;
    a.radii = radii; a.surfel = (const Surfel*)surfel_state; a.clamped = clamped; a.accum = accum;
    a.means2D_cols = means2D_cols;
    a.dL_dmeans2D = dL_dmeans2D; a.dL_dcolors = dL_dcolors; a.dL_dopacity = dL_dopacity; a.dL_dmeans3D = dL_dmeans3D;
    a.dL_dtransmat = dL_dtransmat; a.dL_dsh = dL_dsh; a.dL_dscales = dL_dscales; a.dL_drotations = dL_drotations;
    surfel_gauss_backward_kernel<<<(P + SP_THREADS - 1) / SP_THREADS, SP_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace gdr
