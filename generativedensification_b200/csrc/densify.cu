// The densify select either side of the raster path, fused on the device.
//
// Reference (lightning/network.py:865-893): vjp of the image MSE through the 4-view render w.r.t. one shared
// [P,4] screen-space tensor, then  grad[mask][:, 2:4].norm(dim=-1) -> torch.topk(k_num) -> boolean mask, then
// boolean-mask gathers of the selected / non-selected sets (:905-915, 955-959).  Here:
//   mse_grad_kernel     dL/dcolour of mean((clamp(render,0,1) - target)^2) for all views in one pass
//                       (renderer.py:261 clamp, network.py:855-862 loss), plus the loss itself;
//   score_kernel        sums the means2D-only backward's per-view accumulators over the views and emits
//                       the [P,4] gradient and the selection score ||(sum|d/dx|, sum|d/dy|)||;
//   topk_select_kernel  exact k-th-largest by 4-pass MSB radix select on the float bits, then the
//                       boolean mask and the two compacted index lists (ascending), all on the device:
//                       no sort, no host round trip.  Ties at the threshold go to the lowest indices
//                       (torch.topk leaves the choice among equal values unspecified).
#include "kernels.h"

namespace gdr {

namespace {

constexpr int MSE_THREADS = 256;

// color [V,3,H,W] (unclamped render), target [V,H,W,3]; dL_dcolor [V,3,H,W]; loss_sum += sum of squared errors.
__global__ void __launch_bounds__(MSE_THREADS)
mse_grad_kernel(size_t n_pix /* V*H*W */, size_t HW, const float* __restrict__ color, const float* __restrict__ target,
                float inv_count, float* __restrict__ dL_dcolor, float* __restrict__ loss_sum) {
    float local = 0.f;
    for (size_t i = (size_t)blockIdx.x * MSE_THREADS + threadIdx.x; i < n_pix; i += (size_t)gridDim.x * MSE_THREADS) {
        const size_t v = i / HW, p = i - v * HW;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const size_t ci = (v * 3 + c) * HW + p;
            const float x = color[ci];
            const float y = fminf(fmaxf(x, 0.f), 1.f);
            const float d = y - __ldg(target + i * 3 + c);
            local = fmaf(d, d, local);
            // torch.clamp passes the gradient where 0 <= x <= 1
            dL_dcolor[ci] = (x >= 0.f && x <= 1.f) ? 2.f * d * inv_count : 0.f;
        }
    }
    if (loss_sum) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        __shared__ float wsum[MSE_THREADS / 32];
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < MSE_THREADS / 32; w++) t += wsum[w];
            atomicAdd(loss_sum, t * inv_count);
        }
    }
}

// accum [V][P][12] (only [0..3] written by the means2D-only backward) -> grad [P,4] (sum over views), score [P].
__global__ void score_kernel(int V, int P, const float* __restrict__ accum, const uint8_t* __restrict__ candidate,
                             float* __restrict__ grad, float* __restrict__ score) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int v = 0; v < V; v++) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(accum + ((size_t)v * P + idx) * 12));
        g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
    }
    if (grad) reinterpret_cast<float4*>(grad)[idx] = g;
    // torch.norm(grad[:, 2:4], dim=-1); negative = not a candidate (outside `mask`)
    score[idx] = (candidate == nullptr || candidate[idx]) ? sqrtf(g.z * g.z + g.w * g.w) : -1.f;
}

constexpr int SEL_THREADS = 1024;

// Order-preserving key of a candidate's score: 0 = not a candidate, NaN sorts above everything (as in torch.topk).
__device__ __forceinline__ uint32_t score_key(float s) {
    if (s != s) return 0xffffffffu;
    if (s < 0.f) return 0u;
    return __float_as_uint(s) + 1u;
}

// One CTA.  counts[0] = number selected, counts[1] = number of candidates not selected.
__global__ void __launch_bounds__(SEL_THREADS)
topk_select_kernel(int P, const float* __restrict__ score, int k, uint8_t* __restrict__ selected,
                   int32_t* __restrict__ selected_idx, int32_t* __restrict__ rest_idx, int32_t* __restrict__ counts) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_k;
    __shared__ uint32_t wsum[3][SEL_THREADS / 32];
    __shared__ uint32_t run[3];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // ---- radix select: the key T of the k-th largest candidate and how many keys equal to T to take ----
    if (tid == 0) {
        s_prefix = 0;
        s_k = (uint32_t)max(k, 0);
    }
    uint32_t threshold = 1u, need_ties = 0xffffffffu;  // defaults: every candidate is selected
    bool all = false;
    for (int pass = 0; pass < 4 && !all; pass++) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, kk = s_k;
        const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < P; i += SEL_THREADS) {
            const uint32_t key = score_key(score[i]);
            if (key != 0u && (key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = 0;
            int b = 255;
            for (; b >= 0; b--) {
                if (acc + hist[b] >= kk) break;
                acc += hist[b];
            }
            if (b < 0) {  // fewer candidates than k (only possible in pass 0): take them all
                s_k = 0xffffffffu;
            } else {
                s_prefix = prefix | ((uint32_t)b << shift);
                s_k = kk - acc;
            }
        }
        __syncthreads();
        if (s_k == 0xffffffffu) all = true;
    }
    if (!all && k > 0) {
        threshold = s_prefix;
        need_ties = s_k;
    } else if (k <= 0) {
        threshold = 0xffffffffu;
        need_ties = 0;
    }
    if (tid < 3) run[tid] = 0;
    __syncthreads();

    // ---- one ordered sweep: mask + the two compacted index lists ----
    for (int base = 0; base < P; base += SEL_THREADS) {
        const int i = base + tid;
        const uint32_t key = i < P ? score_key(score[i]) : 0u;
        const bool cand = key != 0u;
        const bool tie = cand && key == threshold && need_ties != 0xffffffffu;
        const bool above = cand && (need_ties == 0xffffffffu ? true : key > threshold);
        // block-exclusive ranks of (tie) in index order
        const unsigned bt = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) wsum[0][wid] = __popc(bt);
        __syncthreads();
        uint32_t tie_rank = run[0] + __popc(bt & ((1u << lane) - 1u));
        for (int w = 0; w < wid; w++) tie_rank += wsum[0][w];
        const bool sel = above || (tie && tie_rank < need_ties);
        const bool rest = cand && !sel;
        const unsigned bs = __ballot_sync(0xffffffffu, sel), br = __ballot_sync(0xffffffffu, rest);
        if (lane == 0) {
            wsum[1][wid] = __popc(bs);
            wsum[2][wid] = __popc(br);
        }
        __syncthreads();
        uint32_t ps = run[1] + __popc(bs & ((1u << lane) - 1u)), pr = run[2] + __popc(br & ((1u << lane) - 1u));
        for (int w = 0; w < wid; w++) {
            ps += wsum[1][w];
            pr += wsum[2][w];
        }
        if (i < P) {
            if (selected) selected[i] = sel ? 1 : 0;
            if (sel && selected_idx) selected_idx[ps] = i;
            if (rest && rest_idx) rest_idx[pr] = i;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t t0 = 0, t1 = 0, t2 = 0;
            for (int w = 0; w < SEL_THREADS / 32; w++) {
                t0 += wsum[0][w];
                t1 += wsum[1][w];
                t2 += wsum[2][w];
            }
            run[0] += t0;
            run[1] += t1;
            run[2] += t2;
        }
        __syncthreads();
    }
    if (tid == 0 && counts) {
        counts[0] = (int32_t)run[1];
        counts[1] = (int32_t)run[2];
    }
}

}  // namespace

cudaError_t launch_mse_grad(int V, int W, int H, const float* color, const float* target, float* dL_dcolor,
                            float* loss, cudaStream_t s) {
    const size_t HW = (size_t)W * H, n = HW * (size_t)V;
    if (n == 0) return cudaSuccess;
    if (loss) {
        cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
        if (e != cudaSuccess) return e;
    }
    const int grid = (int)min((n + MSE_THREADS - 1) / MSE_THREADS, (size_t)sm_count() * 8);
    mse_grad_kernel<<<grid, MSE_THREADS, 0, s>>>(n, HW, color, target, 1.0f / (float)(n * 3), dL_dcolor, loss);
    return cudaGetLastError();
}

cudaError_t launch_densify_score(int V, int P, const float* accum, const uint8_t* candidate, float* grad, float* score,
                                 cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    score_kernel<<<(P + 255) / 256, 256, 0, s>>>(V, P, accum, candidate, grad, score);
    return cudaGetLastError();
}

cudaError_t launch_topk_select(int P, const float* score, int k, uint8_t* selected, int32_t* selected_idx,
                               int32_t* rest_idx, int32_t* counts, cudaStream_t s) {
    topk_select_kernel<<<1, SEL_THREADS, 0, s>>>(P, score, k, selected, selected_idx, rest_idx, counts);
    return cudaGetLastError();
}

}  // namespace gdr
