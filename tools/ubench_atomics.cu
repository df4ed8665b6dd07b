// How fast does the L2 serve atomics that hit one 128-byte line?  (The per-view header of the rasterizer takes one
// cursor claim + one order-list claim per tile and three counter updates per projection CTA -- all on one or two lines.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_atomics.cu -o /tmp/ubench_atomics && /tmp/ubench_atomics
// Each CTA's thread 0 issues K atomics (returning or not) to address base[(cta % n_lines) * stride_words + k * kstep].
#include <cstdio>
#include <cuda_runtime.h>

template <bool RETURNING>
__global__ void __launch_bounds__(256) k_atomics(unsigned* base, int n_lines, int stride_words, int K, int kstep, unsigned* sink) {
    __shared__ unsigned s;
    if (threadIdx.x == 0) {
        unsigned* p = base + (size_t)(blockIdx.x % n_lines) * stride_words;
        unsigned acc = 0;
        for (int k = 0; k < K; k++) {
            if (RETURNING) acc += atomicAdd(p + k * kstep, 1u);
            else atomicAdd(p + k * kstep, 1u);
        }
        s = acc;
    }
    __syncthreads();  // the CTA lives until its atomics have returned (as tile_sort's does)
    if (RETURNING && threadIdx.x == 1 && s == 0xffffffffu) *sink = s;
}

__global__ void __launch_bounds__(256) k_empty(unsigned* sink) {
    if (threadIdx.x == 1 && sink == nullptr) *sink = 0;
}

template <typename F>
float time_us(F f, int reps = 20) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int i = 0; i < reps; i++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best * 1e3f;
}

int main() {
    unsigned *buf, *sink;
    cudaMalloc(&buf, 64 << 20);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, 64 << 20);
    printf("# atomics from thread 0 of each CTA (256 threads), best of 20 launches, us per launch\n");
    for (int ctas : {1563, 2500, 10000}) {
        printf("empty kernel, %5d CTAs: %6.1f us\n", ctas, time_us([&] { k_empty<<<ctas, 256>>>(sink); }));
        for (int K : {1, 2, 3}) {
            for (int lines : {1, 8, 64, 100000}) {
                for (int kstep : {1, 64}) {  // the CTA's K atomics on the same line, or on K lines 256 bytes apart
                    if (K == 1 && kstep != 1) continue;
                    float r = time_us([&] { k_atomics<true><<<ctas, 256>>>(buf, lines, 64 * 4, K, kstep, sink); });
                    float n = time_us([&] { k_atomics<false><<<ctas, 256>>>(buf, lines, 64 * 4, K, kstep, sink); });
                    printf("%5d CTAs x %d atomics, %6d target lines, per-CTA spread %3d words: returning %6.1f us   fire-and-forget %6.1f us\n",
                           ctas, K, lines, kstep, r, n);
                }
            }
        }
    }
    return 0;
}
