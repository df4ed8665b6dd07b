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
#include <cooperative_groups.h>

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
constexpr int SEL_CLUSTER = 8;  // CTAs (SMs) of the one cluster that selects; partial results travel through DSMEM

// Order-preserving key of a candidate's score: 0 = not a candidate, NaN sorts above everything (as in torch.topk).
__device__ __forceinline__ uint32_t score_key(float s) {
    if (s != s) return 0xffffffffu;
    if (s < 0.f) return 0u;
    return __float_as_uint(s) + 1u;
}

struct SelShared {
    uint32_t hist[256];   // this CTA's histogram of the current digit
    uint32_t counts[3];   // this CTA's range: candidates above the threshold, ties at it, candidates in total
};

constexpr int SEL_VEC = 4;                         // consecutive elements per thread and step (one 128-bit load)
constexpr int SEL_STEP = SEL_THREADS * SEL_VEC;    // elements a CTA covers per step

// keys of elements i .. i + 3 (0 beyond `end`); one LDG.128 when the row is whole and aligned
__device__ __forceinline__ void load_keys(const float* __restrict__ score, bool aligned, int i, int end, uint32_t key[SEL_VEC]) {
    if (aligned && i + SEL_VEC <= end) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(score + i));
        key[0] = score_key(v.x);
        key[1] = score_key(v.y);
        key[2] = score_key(v.z);
        key[3] = score_key(v.w);
    } else {
#pragma unroll
        for (int j = 0; j < SEL_VEC; j++) key[j] = i + j < end ? score_key(__ldg(score + i + j)) : 0u;
    }
}

// ONE thread-block cluster of SEL_CLUSTER CTAs (a single CTA sweeping P five times was fine at 262 144 candidates and
// took ~1 ms at 2 M).  CTA r owns the contiguous index range [r * chunk, (r + 1) * chunk); a thread handles four
// consecutive elements per step (one 128-bit load, and a quarter of the barriers of the ordered sweep).
//   radix select   4 passes over the keys' bytes, most significant first: every CTA histograms its range, the cluster
//                  barrier publishes the histograms, every CTA sums its peers' through distributed shared memory
//                  (cluster.map_shared_rank) and picks the digit -- redundantly, from identical data;
//   count          one sweep: per range the candidates above the threshold, the ties at it, the candidates in total --
//                  exchanged the same way, which gives every CTA the number of ties / selected / rest before its range;
//   write          one ordered sweep: the mask and the two ascending index lists at their final positions.
// counts[0] = number selected, counts[1] = number of candidates not selected.
__global__ void __cluster_dims__(SEL_CLUSTER, 1, 1) __launch_bounds__(SEL_THREADS)
topk_select_kernel(int P, const float* __restrict__ score, int k, uint8_t* __restrict__ selected,
                   int32_t* __restrict__ selected_idx, int32_t* __restrict__ rest_idx, int32_t* __restrict__ counts) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ SelShared sh;
    __shared__ uint32_t total_hist[256];
    __shared__ uint32_t s_prefix, s_k;
    __shared__ uint32_t wsum[3][SEL_THREADS / 32];
    __shared__ uint32_t run[3];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int chunk = ((P + SEL_CLUSTER - 1) / SEL_CLUSTER + SEL_STEP - 1) / SEL_STEP * SEL_STEP;
    const int i0 = min(P, rank * chunk), i1 = min(P, i0 + chunk);
    const bool aligned = (((uintptr_t)score) & 15u) == 0;

    // ---- radix select: the key T of the k-th largest candidate and how many keys equal to T to take ----
    if (tid == 0) {
        s_prefix = 0;
        s_k = (uint32_t)max(k, 0);
    }
    uint32_t threshold = 1u, need_ties = 0xffffffffu;  // defaults: every candidate is selected
    bool all = false;
    for (int pass = 0; pass < 4 && !all; pass++) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) sh.hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, kk = s_k;
        const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int base = i0; base < i1; base += SEL_STEP) {
            uint32_t key[SEL_VEC];
            load_keys(score, aligned, base + SEL_VEC * tid, i1, key);
#pragma unroll
            for (int j = 0; j < SEL_VEC; j++)
                if (key[j] != 0u && (key[j] & hi_mask) == prefix) atomicAdd(&sh.hist[(key[j] >> shift) & 255u], 1u);
        }
        cluster.sync();  // every CTA's histogram is complete and visible cluster-wide
        if (tid < 256) {
            uint32_t t = 0;
#pragma unroll
            for (int r = 0; r < SEL_CLUSTER; r++) t += cluster.map_shared_rank(&sh, r)->hist[tid];
            total_hist[tid] = t;
        }
        cluster.sync();  // peers have read this CTA's histogram: it may be cleared for the next pass
        if (tid == 0) {  // the same decision in every CTA (identical totals)
            uint32_t acc = 0;
            int b = 255;
            for (; b >= 0; b--) {
                if (acc + total_hist[b] >= kk) break;
                acc += total_hist[b];
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
    const bool take_all = need_ties == 0xffffffffu;

    // ---- count: this range's candidates above the threshold / at it / in total ----
    if (tid < 3) sh.counts[tid] = 0;
    __syncthreads();
    {
        uint32_t n_above = 0, n_tie = 0, n_cand = 0;
        for (int base = i0; base < i1; base += SEL_STEP) {
            uint32_t key[SEL_VEC];
            load_keys(score, aligned, base + SEL_VEC * tid, i1, key);
#pragma unroll
            for (int j = 0; j < SEL_VEC; j++) {
                const bool cand = key[j] != 0u;
                n_cand += cand;
                n_above += cand && (take_all || key[j] > threshold);
                n_tie += cand && !take_all && key[j] == threshold;
            }
        }
        n_above = __reduce_add_sync(0xffffffffu, n_above);
        n_tie = __reduce_add_sync(0xffffffffu, n_tie);
        n_cand = __reduce_add_sync(0xffffffffu, n_cand);
        if (lane == 0) {
            atomicAdd(&sh.counts[0], n_above);
            atomicAdd(&sh.counts[1], n_tie);
            atomicAdd(&sh.counts[2], n_cand);
        }
    }
    cluster.sync();
    if (tid == 0) {  // what lies before this range, in index order
        uint32_t ties_before = 0, sel_before = 0, rest_before = 0, sel_total = 0, rest_total = 0;
        for (int r = 0; r < SEL_CLUSTER; r++) {
            const SelShared* peer = cluster.map_shared_rank(&sh, r);
            const uint32_t above = peer->counts[0], tie = peer->counts[1], cand = peer->counts[2];
            // ties are taken in index order until need_ties of them are selected
            const uint32_t tie_taken = take_all ? 0u : min(tie, need_ties - min(need_ties, ties_before));
            const uint32_t sel = above + tie_taken;
            if (r < rank) {
                sel_before += sel;
                rest_before += cand - sel;
            }
            if (r == rank) run[0] = ties_before;  // ties before this CTA's first element
            ties_before += tie;
            sel_total += sel;
            rest_total += cand - sel;
        }
        run[1] = sel_before;
        run[2] = rest_before;
        if (rank == 0 && counts) {
            counts[0] = (int32_t)sel_total;
            counts[1] = (int32_t)rest_total;
        }
    }
    __syncthreads();

    // ---- one ordered sweep over this range: mask + the two compacted index lists ----
    // block-exclusive prefix of a per-thread count, in thread (= index) order; `slot` selects the scratch row
    auto block_excl = [&](uint32_t mine, int slot) {
        const uint32_t incl = (uint32_t)warp_incl_scan((int)mine);
        if (lane == 31) wsum[slot][wid] = incl;
        __syncthreads();
        uint32_t before = incl - mine;
        for (int w = 0; w < wid; w++) before += wsum[slot][w];
        return before;
    };
    for (int base = i0; base < i1; base += SEL_STEP) {
        const int i = base + SEL_VEC * tid;
        uint32_t key[SEL_VEC];
        load_keys(score, aligned, i, i1, key);
        bool cand[SEL_VEC], tie[SEL_VEC], sel[SEL_VEC];
        uint32_t n_tie = 0;
#pragma unroll
        for (int j = 0; j < SEL_VEC; j++) {
            cand[j] = key[j] != 0u;
            tie[j] = cand[j] && key[j] == threshold && !take_all;
            n_tie += tie[j];
        }
        uint32_t tie_rank = run[0] + block_excl(n_tie, 0);
        uint32_t n_sel = 0, n_rest = 0;
#pragma unroll
        for (int j = 0; j < SEL_VEC; j++) {
            const bool above = cand[j] && (take_all ? true : key[j] > threshold);
            sel[j] = above || (tie[j] && tie_rank < need_ties);
            tie_rank += tie[j];
            n_sel += sel[j];
            n_rest += cand[j] && !sel[j];
        }
        uint32_t ps = run[1] + block_excl(n_sel, 1), pr = run[2] + block_excl(n_rest, 2);
#pragma unroll
        for (int j = 0; j < SEL_VEC; j++) {
            if (i + j < i1) {
                if (selected) selected[i + j] = sel[j] ? 1 : 0;
                if (sel[j]) {
                    if (selected_idx) selected_idx[ps] = i + j;
                    ps++;
                } else if (cand[j]) {
                    if (rest_idx) rest_idx[pr] = i + j;
                    pr++;
                }
            }
        }
        __syncthreads();  // every thread has read run[] and wsum[]
        if (tid == SEL_THREADS - 1) {  // the last thread's inclusive totals close the step
            run[0] = tie_rank;
            run[1] = ps;
            run[2] = pr;
        }
        __syncthreads();
    }
    cluster.sync();  // no CTA leaves while a peer may still read its shared memory
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
    topk_select_kernel<<<SEL_CLUSTER, SEL_THREADS, 0, s>>>(P, score, k, selected, selected_idx, rest_idx, counts);
    return cudaGetLastError();
}

}  // namespace gdr
