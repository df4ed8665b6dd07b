// Random gather of 48-byte records (global, L2-resident) into shared memory and back out as a contiguous stream:
// which path moves the most records per microsecond?  (tile_sort's record gather: 0.68 M records per view.)
//   0  one thread per record: 3 x LDG.128, 3 x STG.128 (the round-2 gather)
//   1  three lanes per record: lane q loads / stores 16-byte word q (one L1TEX wavefront per record and direction)
//   2  one 48-byte cp.async.bulk global -> shared per record (issued by the record's thread), chunk leaves as ONE
//      shared -> global bulk copy
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_gather.cu -o /tmp/ubench_gather
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int THREADS = 256;
constexpr int CHUNK = 256;  // records per chunk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(THREADS) gather0(const float4* __restrict__ rec, const uint32_t* __restrict__ idx,
                                                   float4* __restrict__ out, int per_cta) {
    const uint32_t* my = idx + (size_t)blockIdx.x * per_cta;
    float4* o = out + (size_t)blockIdx.x * per_cta * 3;
    for (int i = threadIdx.x; i < per_cta; i += THREADS) {
        const float4* s = rec + (size_t)my[i] * 3;
        const float4 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2);
        o[(size_t)i * 3] = a;
        o[(size_t)i * 3 + 1] = b;
        o[(size_t)i * 3 + 2] = c;
    }
}

__global__ void __launch_bounds__(THREADS) gather1(const float4* __restrict__ rec, const uint32_t* __restrict__ idx,
                                                   float4* __restrict__ out, int per_cta) {
    const uint32_t* my = idx + (size_t)blockIdx.x * per_cta;
    float4* o = out + (size_t)blockIdx.x * per_cta * 3;
    for (int w = threadIdx.x; w < per_cta * 3; w += THREADS) {  // word w of the CTA's output
        const int r = w / 3, q = w - 3 * r;
        o[w] = __ldg(rec + (size_t)my[r] * 3 + q);
    }
}

__global__ void __launch_bounds__(THREADS) gather2(const float4* __restrict__ rec, const uint32_t* __restrict__ idx,
                                                   float4* __restrict__ out, int per_cta) {
    __shared__ __align__(128) float4 buf[2][CHUNK * 3];
    __shared__ uint64_t full[2];
    const uint32_t* my = idx + (size_t)blockIdx.x * per_cta;
    float4* o = out + (size_t)blockIdx.x * per_cta * 3;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[b])), "r"(THREADS));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const int n_chunks = per_cta / CHUNK;
    for (int c = 0; c < n_chunks; c++) {
        const int b = c & 1;
        if (c >= 2) {  // the bulk store that read buf[b] two chunks ago must have finished reading
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        const int i = c * CHUNK + threadIdx.x;
        const float4* s = rec + (size_t)my[i] * 3;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[b])), "r"(48u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(&buf[b][threadIdx.x * 3])),
                     "l"(s), "r"(48u), "r"(smem_u32(&full[b]))
                     : "memory");
        // wait for the whole chunk
        const uint32_t parity = (c >> 1) & 1;
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(smem_u32(&full[b])), "r"(parity)
                : "memory");
        }
        // (the rasterizer would patch the region mask into the staged records here)
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o + (size_t)c * CHUNK * 3),
                         "r"(smem_u32(&buf[b][0])), "r"((uint32_t)(CHUNK * 48))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_us(F f, int reps = 10) {
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
    const int P = 200000;                  // records (9.6 MB: L2-resident)
    const int ctas = 2664, per_cta = 256;  // 0.68 M gathered records, one chunk per CTA (like a 256-instance tile)
    const size_t R = (size_t)ctas * per_cta;
    std::vector<uint32_t> h(R);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < R; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        h[i] = (uint32_t)(s % P);
    }
    float4 *rec, *out;
    uint32_t* idx;
    cudaMalloc(&rec, (size_t)P * 48);
    cudaMalloc(&out, R * 48);
    cudaMalloc(&idx, R * 4);
    cudaMemset(rec, 1, (size_t)P * 48);
    cudaMemcpy(idx, h.data(), R * 4, cudaMemcpyHostToDevice);
    printf("# %zu records of 48 B gathered from %d (random), %d CTAs x %d, best of 10, us\n", R, P, ctas, per_cta);
    for (int shape = 0; shape < 2; shape++) {
        const int c = shape == 0 ? ctas : ctas / 8, pc = shape == 0 ? per_cta : per_cta * 8;
        printf("%d CTAs x %d records:\n", c, pc);
        printf("  thread per record (3 LDG.128 + 3 STG.128)  %7.1f us\n", time_us([&] { gather0<<<c, THREADS>>>(rec, idx, out, pc); }));
        printf("  three lanes per record                     %7.1f us\n", time_us([&] { gather1<<<c, THREADS>>>(rec, idx, out, pc); }));
        printf("  48-byte bulk copies + one bulk store/chunk %7.1f us\n", time_us([&] { gather2<<<c, THREADS>>>(rec, idx, out, pc); }));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
