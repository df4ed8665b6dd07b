/*
 * gs_oracle.c -- CPU restatement of the reference Gaussian-splatting rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * generativedensification_b200/csrc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product never
 * routes through it.
 *
 * PARITY PIN: the reference has no tests or golden vectors of its own
 * (SURVEY.md section 4).  This restatement is pinned against outputs of the
 * UNMODIFIED reference extension (oracle/_ref, built by oracle/build_ref.py)
 * run on a B200 and committed as tests/golden/*.npz (generator:
 * tests/golden/make_golden.py).
 *
 * Every stage cites the reference file:line it follows.  RAST/ abbreviates
 * third_party/diff-gaussian-rasterization/ in the reference tree.
 *
 * Arithmetic notes.  The reference is FP32 CUDA compiled without fast-math but
 * with nvcc's default FMA contraction.  Sums of three products a*b + c*d + e*f
 * are contracted by nvcc into fma(e,f, fma(a,b, round(c*d))) (probed with
 * nvcc -ptx / cuobjdump -sass); FM3() below restates that so the
 * discrete decisions (radius, tile rectangle, depth order) agree with the GPU
 * wherever possible.  Where exact agreement cannot be guaranteed (CUDA's expf
 * is not glibc's), the blend reports an `ambiguous` mask of pixels/Gaussians
 * that sit within a rounding error of one of the rasterizer's discontinuities
 * (alpha < 1/255, T < 1e-4, ceil() of the radius, tile-rectangle edges, near
 * plane), so parity tests can separate rounding flips from real errors.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -mfma -shared -fPIC (oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16 /* RAST/cuda_rasterizer/config.h:16-17 (BLOCK_X = BLOCK_Y = 16) */

/* RAST/cuda_rasterizer/auxiliary.h:22-39 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* a0*b0 + a1*b1 + a2*b2 as nvcc contracts it (left product fused into the rounded
 * middle product, then the last product fused). */
static inline float FM3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

static inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

int gso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void gso_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* Stage 1: per-Gaussian projection.                                           */
/* RAST/cuda_rasterizer/forward.cu:155-256 (preprocessCUDA), :118-152           */
/* (computeCov3D), :74-113 (computeCov2D), :20-71 (computeColorFromSH);         */
/* RAST/cuda_rasterizer/auxiliary.h:41-56 (ndc2Pix, getRect), :139-164          */
/* (in_frustum).                                                               */
/* ------------------------------------------------------------------------- */

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* auxiliary.h:46-56; the C int cast truncates toward zero. */
static void tile_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
    *x0 = imin(gx, imax(0, (int)((px - radius) / TILE)));
    *y0 = imin(gy, imax(0, (int)((py - radius) / TILE)));
    *x1 = imin(gx, imax(0, (int)((px + radius + TILE - 1) / TILE)));
    *y1 = imin(gy, imax(0, (int)((py + radius + TILE - 1) / TILE)));
}

static inline int near_int(float v, float eps) {
    float r = v - floorf(v);
    return (r < eps) || (r > 1.0f - eps);
}

/*
 * Outputs (all caller-allocated, P entries unless noted):
 *   radii[P] int32, means2D[2P], depths[P], cov3D[6P], conic_opacity[4P], rgb[3P],
 *   clamped[3P] uint8, tiles_touched[P] uint32, ambiguous[P] uint8.
 * Culled Gaussians get radii = tiles_touched = 0 and zeros elsewhere (the
 * reference leaves those fields uninitialised, forward.cu:184-185).
 */
void gso_project(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                 const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                 const float* colors_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* campos, int W, int H, float tan_fovx, float tan_fovy, int32_t* radii,
                 float* means2D, float* depths, float* cov3Ds, float* conic_opacity, float* rgb,
                 uint8_t* clamped, uint32_t* tiles_touched, uint8_t* ambiguous) {
    /* rasterizer_impl.cu:222-223 */
    const float focal_y = H / (2.0f * tan_fovy);
    const float focal_x = W / (2.0f * tan_fovx);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float* V = viewmatrix;
    const float* F = projmatrix;

#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        means2D[2 * i] = means2D[2 * i + 1] = 0.f;
        depths[i] = 0.f;
        for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.f;
        for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.f; clamped[3 * i + k] = 0; }
        if (!cov3D_precomp) for (int k = 0; k < 6; k++) cov3Ds[6 * i + k] = 0.f;
        ambiguous[i] = 0;

        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];

        /* in_frustum, auxiliary.h:139-164: only the near plane (z <= 0.2) culls. */
        float hom[4], view[3];
        for (int k = 0; k < 4; k++) hom[k] = FM3(F[k], x, F[4 + k], y, F[8 + k], z) + F[12 + k];
        for (int k = 0; k < 3; k++) view[k] = FM3(V[k], x, V[4 + k], y, V[8 + k], z) + V[12 + k];
        if (fabsf(view[2] - 0.2f) < 1e-6f) ambiguous[i] = 3;
        if (view[2] <= 0.2f) continue;
        const float p_w = 1.0f / (hom[3] + 0.0000001f);
        const float ndc_x = hom[0] * p_w, ndc_y = hom[1] * p_w;

        /* computeCov3D, forward.cu:118-152 (quaternion used as given, not normalised). */
        float c3[6];
        if (cov3D_precomp) {
            for (int k = 0; k < 6; k++) c3[k] = cov3D_precomp[6 * i + k];
        } else {
            const float sx = scale_modifier * scales[3 * i], sy = scale_modifier * scales[3 * i + 1],
                        sz = scale_modifier * scales[3 * i + 2];
            const float qr = rotations[4 * i], qx = rotations[4 * i + 1], qy = rotations[4 * i + 2],
                        qz = rotations[4 * i + 3];
            /* Rm[k][c]: row k, column c of the matrix the reference calls R (GLM column c, row k).
             * Which product of x*z +- r*y etc. is fused follows the SASS nvcc emits for the
             * reference's computeCov3D ("numerical contract"). */
            float Rm[3][3];
            const float xz = qx * qz, rx = qr * qx, rz = qr * qz, yy = qy * qy, zz = qz * qz;
            Rm[0][0] = 1.f - 2.f * (yy + zz);
            Rm[1][0] = 2.f * fmaf(qx, qy, -rz);
            Rm[2][0] = 2.f * fmaf(qr, qy, xz);
            Rm[0][1] = 2.f * fmaf(qx, qy, rz);
            Rm[1][1] = 1.f - 2.f * fmaf(qx, qx, zz);
            Rm[2][1] = 2.f * fmaf(qy, qz, -rx);
            Rm[0][2] = 2.f * fmaf(-qr, qy, xz);
            Rm[1][2] = 2.f * fmaf(qy, qz, rx);
            Rm[2][2] = 1.f - 2.f * fmaf(qx, qx, yy);
            const float s[3] = {sx, sy, sz};
            float Mm[3][3]; /* M = S * R : row k scaled by s_k */
            for (int k = 0; k < 3; k++)
                for (int c = 0; c < 3; c++) Mm[k][c] = s[k] * Rm[k][c];
            /* Sigma = M^T M : Sigma(r,c) = sum_k M(k,r) M(k,c) */
            const int idx[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
            for (int e = 0; e < 6; e++) {
                const int c = idx[e][0], r = idx[e][1];
                c3[e] = FM3(Mm[0][r], Mm[0][c], Mm[1][r], Mm[1][c], Mm[2][r], Mm[2][c]);
            }
            for (int k = 0; k < 6; k++) cov3Ds[6 * i + k] = c3[k];
        }

        /* computeCov2D, forward.cu:74-113 (EWA projection + 0.3 low-pass). */
        float tx = view[0], ty = view[1];
        const float tz = view[2];
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float J00 = focal_x / tz, J02 = -(focal_x * tx) / (tz * tz);
        const float J11 = focal_y / tz, J12 = -(focal_y * ty) / (tz * tz);
        /* T = W * J (GLM), T0[r] = first column, T1[r] = second column */
        float T0[3], T1[3];
        for (int r = 0; r < 3; r++) {
            const float w0 = V[4 * r + 0], w1 = V[4 * r + 1], w2 = V[4 * r + 2];
            T0[r] = fmaf(w2, J02, w0 * J00);
            T1[r] = fmaf(w2, J12, w1 * J11);
        }
        const float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        float U0[3], U1[3]; /* U = T^T * Vrk ; U0[c] = U(c, row 0), U1[c] = U(c, row 1) */
        for (int c = 0; c < 3; c++) {
            U0[c] = FM3(T0[0], Vrk[c][0], T0[1], Vrk[c][1], T0[2], Vrk[c][2]);
            U1[c] = FM3(T1[0], Vrk[c][0], T1[1], Vrk[c][1], T1[2], Vrk[c][2]);
        }
        float cov_a = FM3(U0[0], T0[0], U0[1], T0[1], U0[2], T0[2]);
        const float cov_b = FM3(U1[0], T0[0], U1[1], T0[1], U1[2], T0[2]);
        float cov_c = FM3(U1[0], T1[0], U1[1], T1[1], U1[2], T1[2]);
        cov_a += 0.3f;
        cov_c += 0.3f;

        /* forward.cu:219-237: conic, radius, rectangle */
        const float det = fmaf(cov_a, cov_c, -(cov_b * cov_b));
        if (det == 0.0f) continue;
        const float det_inv = 1.f / det;
        const float con_x = cov_c * det_inv, con_y = -cov_b * det_inv, con_z = cov_a * det_inv;
        const float mid = 0.5f * (cov_a + cov_c);
        const float disc = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        const float lambda1 = mid + disc, lambda2 = mid - disc;
        const float r_unrounded = 3.f * sqrtf(fmaxf(lambda1, lambda2));
        const float my_radius = ceilf(r_unrounded);
        /* ndc2Pix is evaluated in double (auxiliary.h:41-44) */
        const float pix_x = (float)((((double)ndc_x + 1.0) * W - 1.0) * 0.5);
        const float pix_y = (float)((((double)ndc_y + 1.0) * H - 1.0) * 0.5);
        int x0, y0, x1, y1;
        tile_rect(pix_x, pix_y, (int)my_radius, gx, gy, &x0, &y0, &x1, &y1);

        /* discrete-decision ambiguity: the pixel position is reproduced bit for bit (see the golden
         * tests), the eigenvalue only to a few ulp -- flag the Gaussian if ceil() could go either way
         * (bit 1), and its tiles if that would also move the tile rectangle (bit 0). */
        {
            int amb = 0;
            if (near_int(r_unrounded, 4e-6f * fmaxf(1.f, r_unrounded))) {
                int a0, b0, a1, b1, c0, d0, c1, d1;
                tile_rect(pix_x, pix_y, (int)my_radius - 1, gx, gy, &a0, &b0, &a1, &b1);
                tile_rect(pix_x, pix_y, (int)my_radius + 1, gx, gy, &c0, &d0, &c1, &d1);
                if (a0 != x0 || b0 != y0 || a1 != x1 || b1 != y1) amb |= 1;
                if (c0 != x0 || d0 != y0 || c1 != x1 || d1 != y1) amb |= 1;
                amb |= 2;
            }
            ambiguous[i] |= (uint8_t)amb;
        }
        if ((x1 - x0) * (y1 - y0) == 0) continue;

        /* computeColorFromSH, forward.cu:20-71 */
        if (!colors_precomp) {
            float dx = x - campos[0], dy = y - campos[1], dz = z - campos[2];
            const float len = sqrtf(FM3(dx, dx, dy, dy, dz, dz));
            dx = dx / len; dy = dy / len; dz = dz / len;
            const float* sh = shs + (size_t)i * M * 3;
            for (int ch = 0; ch < 3; ch++) {
                float res = SH_C0 * sh[ch];
                if (D > 0) {
                    res = res - SH_C1 * dy * sh[3 + ch] + SH_C1 * dz * sh[6 + ch] - SH_C1 * dx * sh[9 + ch];
                    if (D > 1) {
                        const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
                        const float xy = dx * dy, yz = dy * dz, xz = dx * dz;
                        res = res + SH_C2[0] * xy * sh[12 + ch] + SH_C2[1] * yz * sh[15 + ch] +
                              SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + ch] + SH_C2[3] * xz * sh[21 + ch] +
                              SH_C2[4] * (xx - yy) * sh[24 + ch];
                        if (D > 2) {
                            res = res + SH_C3[0] * dy * (3.0f * xx - yy) * sh[27 + ch] +
                                  SH_C3[1] * xy * dz * sh[30 + ch] +
                                  SH_C3[2] * dy * (4.0f * zz - xx - yy) * sh[33 + ch] +
                                  SH_C3[3] * dz * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + ch] +
                                  SH_C3[4] * dx * (4.0f * zz - xx - yy) * sh[39 + ch] +
                                  SH_C3[5] * dz * (xx - yy) * sh[42 + ch] +
                                  SH_C3[6] * dx * (xx - 3.0f * yy) * sh[45 + ch];
                        }
                    }
                }
                res += 0.5f;
                clamped[3 * i + ch] = (res < 0);
                rgb[3 * i + ch] = fmaxf(res, 0.0f);
            }
        } else {
            for (int ch = 0; ch < 3; ch++) rgb[3 * i + ch] = colors_precomp[3 * i + ch];
        }

        depths[i] = view[2];
        radii[i] = (int32_t)my_radius;
        means2D[2 * i] = pix_x;
        means2D[2 * i + 1] = pix_y;
        conic_opacity[4 * i] = con_x;
        conic_opacity[4 * i + 1] = con_y;
        conic_opacity[4 * i + 2] = con_z;
        conic_opacity[4 * i + 3] = opacities[i];
        tiles_touched[i] = (uint32_t)((y1 - y0) * (x1 - x0));
    }
}

/* RAST/cuda_rasterizer/rasterizer_impl.cu:54-66 (checkFrustum / markVisible) */
void gso_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present) {
    const float* V = viewmatrix;
    for (int i = 0; i < P; i++) {
        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        const float vz = FM3(V[2], x, V[6], y, V[10], z) + V[14];
        present[i] = vz > 0.2f;
    }
}

/* ------------------------------------------------------------------------- */
/* Stage 2: binning.  rasterizer_impl.cu:70-111 (duplicateWithKeys), :304-309  */
/* (stable radix sort on (tile << 32 | depth bits)), :116-138 (ranges).         */
/* Final order contract: (tile, depth bits ascending, Gaussian index           */
/* ascending).                                                                 */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint32_t depth_bits;
    uint32_t idx;
} inst_t;

static int inst_cmp(const void* a, const void* b) {
    const inst_t* p = (const inst_t*)a;
    const inst_t* q = (const inst_t*)b;
    if (p->depth_bits != q->depth_bits) return p->depth_bits < q->depth_bits ? -1 : 1;
    if (p->idx != q->idx) return p->idx < q->idx ? -1 : 1;
    return 0;
}

/* point_list has R = sum(tiles_touched) entries, ranges has 2 * tiles entries.
 * Returns the number of instances written (== R) or -1 on mismatch. */
int64_t gso_bin(int P, int W, int H, const float* means2D, const float* depths, const int32_t* radii,
                int64_t R, uint32_t* point_list, uint32_t* ranges) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int T = gx * gy;
    int64_t* count = (int64_t*)calloc((size_t)T + 1, sizeof(int64_t));
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        tile_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        for (int ty = y0; ty < y1; ty++)
            for (int tx = x0; tx < x1; tx++) count[ty * gx + tx + 1]++;
    }
    for (int t = 0; t < T; t++) count[t + 1] += count[t];
    if (count[T] != R) {
        free(count);
        return -1;
    }
    inst_t* inst = (inst_t*)malloc(sizeof(inst_t) * (size_t)(R > 0 ? R : 1));
    int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)T);
    memcpy(cursor, count, sizeof(int64_t) * (size_t)T);
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        tile_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        const uint32_t db = f2u(depths[i]);
        for (int ty = y0; ty < y1; ty++)
            for (int tx = x0; tx < x1; tx++) {
                inst_t* e = &inst[cursor[ty * gx + tx]++];
                e->depth_bits = db;
                e->idx = (uint32_t)i;
            }
    }
#pragma omp parallel for schedule(dynamic, 8)
    for (int t = 0; t < T; t++) {
        const int64_t b = count[t], e = count[t + 1];
        if (e > b) {
            qsort(inst + b, (size_t)(e - b), sizeof(inst_t), inst_cmp);
            ranges[2 * t] = (uint32_t)b;
            ranges[2 * t + 1] = (uint32_t)e;
        } else {
            ranges[2 * t] = ranges[2 * t + 1] = 0; /* memset at rasterizer_impl.cu:311 */
        }
    }
    for (int64_t k = 0; k < R; k++) point_list[k] = inst[k].idx;
    free(inst);
    free(cursor);
    free(count);
    return R;
}

/* ------------------------------------------------------------------------- */
/* Stage 3: forward blend.  RAST/cuda_rasterizer/forward.cu:261-381.            */
/* ------------------------------------------------------------------------- */

/* power exactly as ptxas contracts the reference expression (forward.cu:338):
 * -0.5f*(A*dx*dx + C*dy*dy) - B*dx*dy  ->  fma(fma(dx, A*dx, (C*dy)*dy), -0.5, -((B*dx)*dy)) */
static inline float pair_power(float A, float B, float C, float dx, float dy) {
    const float q = fmaf(dx, dx * A, dy * (dy * C));
    return fmaf(q, -0.5f, -(dy * (dx * B)));
}

#define ALPHA_MIN (1.0f / 255.0f)

void gso_blend_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                       const float* colors, const float* depths, const float* conic_opacity, const float* bg,
                       const uint8_t* gauss_ambiguous, float* out_color, float* out_depth, float* out_alpha,
                       uint32_t* n_contrib, uint8_t* pix_ambiguous) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < gx * gy; t++) {
        const int tx = t % gx, ty = t / gx;
        const uint32_t rb = ranges[2 * t], re = ranges[2 * t + 1];
        int tile_amb = 0;
        if (gauss_ambiguous)
            for (uint32_t k = rb; k < re; k++)
                if (gauss_ambiguous[point_list[k]] & 1) tile_amb = 1;
        for (int py = ty * TILE; py < imin(H, ty * TILE + TILE); py++)
            for (int px = tx * TILE; px < imin(W, tx * TILE + TILE); px++) {
                const float pxf = (float)px, pyf = (float)py;
                float T = 1.0f, C[3] = {0, 0, 0}, weight = 0, Dacc = 0;
                uint32_t contributor = 0, last = 0;
                int amb = tile_amb;
                for (uint32_t k = rb; k < re; k++) {
                    const uint32_t g = point_list[k];
                    contributor++;
                    const float dx = means2D[2 * g] - pxf, dy = means2D[2 * g + 1] - pyf;
                    const float* co = conic_opacity + 4 * g;
                    const float power = pair_power(co[0], co[1], co[2], dx, dy);
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, co[3] * expf(power));
                    if (fabsf(alpha - ALPHA_MIN) < 4e-5f * ALPHA_MIN) amb = 1;
                    if (alpha < ALPHA_MIN) continue;
                    const float test_T = T * (1 - alpha);
                    if (fabsf(test_T - 0.0001f) < 3e-4f * 0.0001f) amb = 1;
                    if (test_T < 0.0001f) break; /* done = true: nothing later contributes */
                    for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(colors[3 * g + ch] * alpha, T, C[ch]);
                    weight = fmaf(alpha, T, weight);
                    Dacc = fmaf(depths[g] * alpha, T, Dacc);
                    T = test_T;
                    last = contributor;
                }
                const size_t pid = (size_t)py * W + px;
                n_contrib[pid] = last;
                for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pid] = fmaf(T, bg[ch], C[ch]);
                out_alpha[pid] = weight; /* forward.cu:378: sum of alpha*T, not 1 - T */
                out_depth[pid] = Dacc;
                if (pix_ambiguous) pix_ambiguous[pid] = (uint8_t)amb;
            }
    }
}

/* ------------------------------------------------------------------------- */
/* Stage 4: backward blend.  RAST/cuda_rasterizer/backward.cu:415-605.          */
/* Per-pair terms in FP32 as the reference; the cross-pixel sums that the      */
/* reference does with float atomics (order-nondeterministic) are accumulated  */
/* in double here.                                                             */
/* Outputs (double, zero-initialised by the caller):                           */
/*   g_mean2D[4P] (x, y, |x|, |y|), g_conic[4P] (reference layout x,y,_,w),     */
/*   g_opacity[P], g_color[3P], g_depth[P].                                    */
/* ------------------------------------------------------------------------- */
static inline void atomic_add_d(double* p, double v) {
#pragma omp atomic
    *p += v;
}

void gso_blend_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                        const float* means2D, const float* conic_opacity, const float* colors,
                        const float* depths, const float* alphas, const uint32_t* n_contrib,
                        const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dpix_alpha,
                        double* g_mean2D, double* g_conic, double* g_opacity, double* g_color,
                        double* g_depth) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < gx * gy; t++) {
        const int tx = t % gx, ty = t / gx;
        const uint32_t rb = ranges[2 * t], re = ranges[2 * t + 1];
        for (int py = ty * TILE; py < imin(H, ty * TILE + TILE); py++)
            for (int px = tx * TILE; px < imin(W, tx * TILE + TILE); px++) {
                const size_t pid = (size_t)py * W + px;
                const float pxf = (float)px, pyf = (float)py;
                const float T_final = 1 - alphas[pid];
                float T = T_final;
                const uint32_t last_contributor = n_contrib[pid];
                float accum_rec[3] = {0, 0, 0}, accum_depth_rec = 0, accum_alpha_rec = 0;
                float last_alpha = 0, last_color[3] = {0, 0, 0}, last_depth = 0;
                const float dpix[3] = {dL_dpix[pid], dL_dpix[HW + pid], dL_dpix[2 * HW + pid]};
                const float dpix_depth = dL_dpix_depth[pid], dpix_alpha = dL_dpix_alpha[pid];
                float bg_dot_dpixel = 0;
                for (int ch = 0; ch < 3; ch++) bg_dot_dpixel += bg[ch] * dpix[ch];
                /* back to front; entry k (0-based) has position k - rb + 1 */
                for (uint32_t k = rb + last_contributor; k-- > rb;) {
                    const uint32_t g = point_list[k];
                    const float dx = means2D[2 * g] - pxf, dy = means2D[2 * g + 1] - pyf;
                    const float* co = conic_opacity + 4 * g;
                    const float power = pair_power(co[0], co[1], co[2], dx, dy);
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < ALPHA_MIN) continue;
                    T = T / (1.f - alpha);
                    const float w = alpha * T;
                    float dL_dopa = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = colors[3 * g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dopa += (c - accum_rec[ch]) * dpix[ch];
                        atomic_add_d(&g_color[3 * g + ch], (double)(w * dpix[ch]));
                    }
                    const float c_d = depths[g];
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dopa += (c_d - accum_depth_rec) * dpix_depth;
                    atomic_add_d(&g_depth[g], (double)(w * dpix_depth));
                    accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dopa += (1 - accum_alpha_rec) * dpix_alpha;
                    dL_dopa *= T;
                    last_alpha = alpha;
                    dL_dopa += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

                    const float dL_dG = co[3] * dL_dopa;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    const float mx = dL_dG * dG_ddelx * ddelx_dx, my = dL_dG * dG_ddely * ddely_dy;
                    atomic_add_d(&g_mean2D[4 * g + 0], (double)mx);
                    atomic_add_d(&g_mean2D[4 * g + 1], (double)my);
                    atomic_add_d(&g_mean2D[4 * g + 2], (double)fabsf(mx));
                    atomic_add_d(&g_mean2D[4 * g + 3], (double)fabsf(my));
                    atomic_add_d(&g_conic[4 * g + 0], (double)(-0.5f * gdx * dx * dL_dG));
                    atomic_add_d(&g_conic[4 * g + 1], (double)(-0.5f * gdx * dy * dL_dG));
                    atomic_add_d(&g_conic[4 * g + 3], (double)(-0.5f * gdy * dy * dL_dG));
                    atomic_add_d(&g_opacity[g], (double)(G * dL_dopa));
                }
            }
    }
}

/* ------------------------------------------------------------------------- */
/* Stage 5: per-Gaussian backward.                                             */
/* backward.cu:144-274 (computeCov2DCUDA), :346-412 (preprocessCUDA bwd),       */
/* :278-341 (computeCov3D bwd), :20-139 (computeColorFromSH bwd).               */
/* Inputs are the float sums produced by stage 4 (cast to float first, as the  */
/* reference's kernels read float buffers).                                    */
/* ------------------------------------------------------------------------- */
void gso_preprocess_backward(int P, int D, int M, const float* means3D, const int32_t* radii, const float* shs,
                             const uint8_t* clamped, const float* scales, const float* rotations,
                             float scale_modifier, const float* cov3Ds, const float* viewmatrix,
                             const float* projmatrix, float focal_x, float focal_y, float tan_fovx,
                             float tan_fovy, const float* campos, const float* dL_dmean2D /*4P*/,
                             const float* dL_dconic /*4P*/, const float* dL_dcolor /*3P*/,
                             const float* dL_ddepth /*P*/, float* dL_dmeans3D /*3P*/, float* dL_dcov3D /*6P*/,
                             float* dL_dsh /*3MP*/, float* dL_dscale /*3P*/, float* dL_drot /*4P*/) {
    const float* V = viewmatrix;
    const float* F = projmatrix;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue; /* outputs stay zero (caller zero-fills) */
        const float mx = means3D[3 * i], my = means3D[3 * i + 1], mz = means3D[3 * i + 2];
        const float* c3 = cov3Ds + 6 * i;

        /* ---- computeCov2DCUDA, backward.cu:144-274 ---- */
        const float dcon_x = dL_dconic[4 * i], dcon_y = dL_dconic[4 * i + 1], dcon_z = dL_dconic[4 * i + 3];
        float t[3];
        for (int k = 0; k < 3; k++) t[k] = V[k] * mx + V[4 + k] * my + V[8 + k] * mz + V[12 + k];
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = t[0] / t[2], tytz = t[1] / t[2];
        t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
        t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float J00 = focal_x / t[2], J02 = -(focal_x * t[0]) / (t[2] * t[2]);
        const float J11 = focal_y / t[2], J12 = -(focal_y * t[1]) / (t[2] * t[2]);
        /* Tm[a][r] = T[a][r] in the reference's GLM indexing (column a, row r); third column is zero */
        float Tm[2][3];
        for (int r = 0; r < 3; r++) {
            Tm[0][r] = V[4 * r + 0] * J00 + V[4 * r + 2] * J02;
            Tm[1][r] = V[4 * r + 1] * J11 + V[4 * r + 2] * J12;
        }
        const float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        float a = 0, b = 0, c = 0;
        for (int p = 0; p < 3; p++)
            for (int q = 0; q < 3; q++) {
                a += Tm[0][p] * Vrk[p][q] * Tm[0][q];
                b += Tm[0][p] * Vrk[p][q] * Tm[1][q];
                c += Tm[1][p] * Vrk[p][q] * Tm[1][q];
            }
        a += 0.3f;
        c += 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcon_x + 2 * b * c * dcon_y + (denom - a * c) * dcon_z);
            dL_dc = denom2inv * (-a * a * dcon_z + 2 * a * b * dcon_y + (denom - a * c) * dcon_x);
            dL_db = denom2inv * 2 * (b * c * dcon_x - (denom + 2 * b * b) * dcon_y + a * b * dcon_z);
            dL_dcov3D[6 * i + 0] = Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc;
            dL_dcov3D[6 * i + 3] = Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc;
            dL_dcov3D[6 * i + 5] = Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc;
            dL_dcov3D[6 * i + 1] = 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
            dL_dcov3D[6 * i + 2] = 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
            dL_dcov3D[6 * i + 4] = 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
        } else {
            for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0;
        }
        /* dL/dT (upper 2x3) */
        float dL_dT[2][3];
        for (int q = 0; q < 3; q++) {
            const float tv0 = Tm[0][0] * Vrk[q][0] + Tm[0][1] * Vrk[q][1] + Tm[0][2] * Vrk[q][2];
            const float tv1 = Tm[1][0] * Vrk[q][0] + Tm[1][1] * Vrk[q][1] + Tm[1][2] * Vrk[q][2];
            dL_dT[0][q] = 2 * tv0 * dL_da + tv1 * dL_db;
            dL_dT[1][q] = 2 * tv1 * dL_dc + tv0 * dL_db;
        }
        /* T = W * J ; W[k][r] = V[4r + k] */
        const float dL_dJ00 = V[0] * dL_dT[0][0] + V[4] * dL_dT[0][1] + V[8] * dL_dT[0][2];
        const float dL_dJ02 = V[2] * dL_dT[0][0] + V[6] * dL_dT[0][1] + V[10] * dL_dT[0][2];
        const float dL_dJ11 = V[1] * dL_dT[1][0] + V[5] * dL_dT[1][1] + V[9] * dL_dT[1][2];
        const float dL_dJ12 = V[2] * dL_dT[1][0] + V[6] * dL_dT[1][1] + V[10] * dL_dT[1][2];
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -focal_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -focal_y * tz2 * dL_dJ12;
        const float dL_dtz = -focal_x * tz2 * dL_dJ00 - focal_y * tz2 * dL_dJ11 +
                             (2 * focal_x * t[0]) * tz3 * dL_dJ02 + (2 * focal_y * t[1]) * tz3 * dL_dJ12;
        /* transformVec4x3Transpose, auxiliary.h:97-105; this term OVERWRITES dL_dmeans (backward.cu:273) */
        float dmean[3];
        dmean[0] = V[0] * dL_dtx + V[1] * dL_dty + V[2] * dL_dtz;
        dmean[1] = V[4] * dL_dtx + V[5] * dL_dty + V[6] * dL_dtz;
        dmean[2] = V[8] * dL_dtx + V[9] * dL_dty + V[10] * dL_dtz;

        /* ---- preprocessCUDA bwd, backward.cu:346-412 ---- */
        const float hx = F[0] * mx + F[4] * my + F[8] * mz + F[12];
        const float hy = F[1] * mx + F[5] * my + F[9] * mz + F[13];
        const float hw = F[3] * mx + F[7] * my + F[11] * mz + F[15];
        const float m_w = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        const float g2x = dL_dmean2D[4 * i], g2y = dL_dmean2D[4 * i + 1];
        dmean[0] += (F[0] * m_w - F[3] * mul1) * g2x + (F[1] * m_w - F[3] * mul2) * g2y;
        dmean[1] += (F[4] * m_w - F[7] * mul1) * g2x + (F[5] * m_w - F[7] * mul2) * g2y;
        dmean[2] += (F[8] * m_w - F[11] * mul1) * g2x + (F[9] * m_w - F[11] * mul2) * g2y;
        const float mul3 = V[2] * mx + V[6] * my + V[10] * mz + V[14];
        const float gd = dL_ddepth[i];
        dmean[0] += (V[2] - V[3] * mul3) * gd;
        dmean[1] += (V[6] - V[7] * mul3) * gd;
        dmean[2] += (V[10] - V[11] * mul3) * gd;

        /* ---- computeColorFromSH bwd, backward.cu:20-139 ---- */
        if (shs) {
            const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox / len, y = oy / len, z = oz / len;
            const float* sh = shs + (size_t)i * M * 3;
            float* dsh = dL_dsh + (size_t)i * M * 3;
            float dRGB[3];
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[3 * i + ch] * (clamped[3 * i + ch] ? 0.f : 1.f);
            float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
#define SHV(k, ch) sh[3 * (k) + (ch)]
            for (int ch = 0; ch < 3; ch++) {
                dsh[ch] = SH_C0 * dRGB[ch];
                if (D > 0) {
                    dsh[3 + ch] = (-SH_C1 * y) * dRGB[ch];
                    dsh[6 + ch] = (SH_C1 * z) * dRGB[ch];
                    dsh[9 + ch] = (-SH_C1 * x) * dRGB[ch];
                    dRGBdx[ch] = -SH_C1 * SHV(3, ch);
                    dRGBdy[ch] = -SH_C1 * SHV(1, ch);
                    dRGBdz[ch] = SH_C1 * SHV(2, ch);
                    if (D > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        dsh[12 + ch] = (SH_C2[0] * xy) * dRGB[ch];
                        dsh[15 + ch] = (SH_C2[1] * yz) * dRGB[ch];
                        dsh[18 + ch] = (SH_C2[2] * (2.f * zz - xx - yy)) * dRGB[ch];
                        dsh[21 + ch] = (SH_C2[3] * xz) * dRGB[ch];
                        dsh[24 + ch] = (SH_C2[4] * (xx - yy)) * dRGB[ch];
                        dRGBdx[ch] += SH_C2[0] * y * SHV(4, ch) + SH_C2[2] * 2.f * -x * SHV(6, ch) +
                                      SH_C2[3] * z * SHV(7, ch) + SH_C2[4] * 2.f * x * SHV(8, ch);
                        dRGBdy[ch] += SH_C2[0] * x * SHV(4, ch) + SH_C2[1] * z * SHV(5, ch) +
                                      SH_C2[2] * 2.f * -y * SHV(6, ch) + SH_C2[4] * 2.f * -y * SHV(8, ch);
                        dRGBdz[ch] += SH_C2[1] * y * SHV(5, ch) + SH_C2[2] * 2.f * 2.f * z * SHV(6, ch) +
                                      SH_C2[3] * x * SHV(7, ch);
                        if (D > 2) {
                            dsh[27 + ch] = (SH_C3[0] * y * (3.f * xx - yy)) * dRGB[ch];
                            dsh[30 + ch] = (SH_C3[1] * xy * z) * dRGB[ch];
                            dsh[33 + ch] = (SH_C3[2] * y * (4.f * zz - xx - yy)) * dRGB[ch];
                            dsh[36 + ch] = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dRGB[ch];
                            dsh[39 + ch] = (SH_C3[4] * x * (4.f * zz - xx - yy)) * dRGB[ch];
                            dsh[42 + ch] = (SH_C3[5] * z * (xx - yy)) * dRGB[ch];
                            dsh[45 + ch] = (SH_C3[6] * x * (xx - 3.f * yy)) * dRGB[ch];
                            dRGBdx[ch] += SH_C3[0] * SHV(9, ch) * 3.f * 2.f * xy + SH_C3[1] * SHV(10, ch) * yz +
                                          SH_C3[2] * SHV(11, ch) * -2.f * xy + SH_C3[3] * SHV(12, ch) * -3.f * 2.f * xz +
                                          SH_C3[4] * SHV(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                                          SH_C3[5] * SHV(14, ch) * 2.f * xz + SH_C3[6] * SHV(15, ch) * 3.f * (xx - yy);
                            dRGBdy[ch] += SH_C3[0] * SHV(9, ch) * 3.f * (xx - yy) + SH_C3[1] * SHV(10, ch) * xz +
                                          SH_C3[2] * SHV(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                                          SH_C3[3] * SHV(12, ch) * -3.f * 2.f * yz + SH_C3[4] * SHV(13, ch) * -2.f * xy +
                                          SH_C3[5] * SHV(14, ch) * -2.f * yz + SH_C3[6] * SHV(15, ch) * -3.f * 2.f * xy;
                            dRGBdz[ch] += SH_C3[1] * SHV(10, ch) * xy + SH_C3[2] * SHV(11, ch) * 4.f * 2.f * yz +
                                          SH_C3[3] * SHV(12, ch) * 3.f * (2.f * zz - xx - yy) +
                                          SH_C3[4] * SHV(13, ch) * 4.f * 2.f * xz + SH_C3[5] * SHV(14, ch) * (xx - yy);
                        }
                    }
                }
            }
#undef SHV
            const float ddx = dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2];
            const float ddy = dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2];
            const float ddz = dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2];
            /* dnormvdv, auxiliary.h:107-117 */
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
            dmean[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
            dmean[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
        }
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * i + k] = dmean[k];

        /* ---- computeCov3D bwd, backward.cu:278-341 ---- */
        if (scales) {
            const float qr = rotations[4 * i], qx = rotations[4 * i + 1], qy = rotations[4 * i + 2],
                        qz = rotations[4 * i + 3];
            /* Rg[c][r]: the reference's GLM matrix R, column c, row r */
            float Rg[3][3];
            Rg[0][0] = 1.f - 2.f * (qy * qy + qz * qz); Rg[0][1] = 2.f * (qx * qy - qr * qz); Rg[0][2] = 2.f * (qx * qz + qr * qy);
            Rg[1][0] = 2.f * (qx * qy + qr * qz); Rg[1][1] = 1.f - 2.f * (qx * qx + qz * qz); Rg[1][2] = 2.f * (qy * qz - qr * qx);
            Rg[2][0] = 2.f * (qx * qz - qr * qy); Rg[2][1] = 2.f * (qy * qz + qr * qx); Rg[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
            const float s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1],
                                scale_modifier * scales[3 * i + 2]};
            float Mg[3][3]; /* M = S * R : Mg[c][r] = s[r] * Rg[c][r] */
            for (int cc = 0; cc < 3; cc++)
                for (int r = 0; r < 3; r++) Mg[cc][r] = s[r] * Rg[cc][r];
            const float* d = dL_dcov3D + 6 * i;
            /* dL_dSigma (symmetric), GLM column-major == row-major here */
            const float dS[3][3] = {{d[0], 0.5f * d[1], 0.5f * d[2]}, {0.5f * d[1], d[3], 0.5f * d[4]}, {0.5f * d[2], 0.5f * d[4], d[5]}};
            /* dL_dM = 2 * M * dL_dSigma : (c,r) = 2 * sum_k M[k][r] * dS[c][k] */
            float dM[3][3];
            for (int cc = 0; cc < 3; cc++)
                for (int r = 0; r < 3; r++)
                    dM[cc][r] = 2.0f * Mg[0][r] * dS[cc][0] + 2.0f * Mg[1][r] * dS[cc][1] + 2.0f * Mg[2][r] * dS[cc][2];
            /* Rt[c][r] = Rg[r][c]; dMt[c][r] = dM[r][c]; dL_dscale_k = dot(Rt[k], dMt[k]) */
            float dMt[3][3];
            for (int cc = 0; cc < 3; cc++)
                for (int r = 0; r < 3; r++) dMt[cc][r] = dM[r][cc];
            for (int k = 0; k < 3; k++)
                dL_dscale[3 * i + k] = Rg[0][k] * dMt[k][0] + Rg[1][k] * dMt[k][1] + Rg[2][k] * dMt[k][2];
            for (int k = 0; k < 3; k++)
                for (int r = 0; r < 3; r++) dMt[k][r] *= s[k];
            float* dq = dL_drot + 4 * i;
            dq[0] = 2 * qz * (dMt[0][1] - dMt[1][0]) + 2 * qy * (dMt[2][0] - dMt[0][2]) + 2 * qx * (dMt[1][2] - dMt[2][1]);
            dq[1] = 2 * qy * (dMt[1][0] + dMt[0][1]) + 2 * qz * (dMt[2][0] + dMt[0][2]) + 2 * qr * (dMt[1][2] - dMt[2][1]) - 4 * qx * (dMt[2][2] + dMt[1][1]);
            dq[2] = 2 * qx * (dMt[1][0] + dMt[0][1]) + 2 * qr * (dMt[2][0] - dMt[0][2]) + 2 * qz * (dMt[1][2] + dMt[2][1]) - 4 * qy * (dMt[2][2] + dMt[0][0]);
            dq[3] = 2 * qr * (dMt[0][1] - dMt[1][0]) + 2 * qx * (dMt[2][0] + dMt[0][2]) + 2 * qy * (dMt[1][2] + dMt[2][1]) - 4 * qz * (dMt[1][1] + dMt[0][0]);
        }
    }
}
