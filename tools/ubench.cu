// Instruction-throughput microbenchmarks for sm_100a (B200): warp-instructions per clock per SM for the
// instruction classes the blend kernels are made of.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// tools/ubench.cu -o tools/ubench.bin ; run on the GPU box.  Evidence for DESIGN.md section 4 (issue model).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
#define CHAINS 8

enum Kind { K_FFMA, K_FFMA2, K_FADD, K_FMUL, K_SHFL, K_SEL, K_MUFU, K_LDS128, K_MIX_FFMA_IADD, K_MIX_FFMA2_IADD, K_MIX_FFMA_FADD,
            K_FADD2, K_FMUL2, K_LOP3, K_MIX_FFMA2_FFMA, K_MIX_SHFL_FFMA, K_NUM };
const char* kind_name[] = {"FFMA", "FFMA2", "FADD", "FMUL", "SHFL.BFLY", "SEL", "MUFU.EX2", "LDS.128(bcast)", "FFMA+IADD3 (1:1)", "FFMA2+IADD3 (1:1)",
                           "FFMA+FADD (1:1)", "FADD2", "FMUL2", "LOP3", "FFMA2+FFMA (1:1)", "SHFL+FFMA (1:1)"};
// instructions per chain-iteration for each kind
const int kind_inst[] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 2, 2};

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, int n_iter, float seed) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float a[CHAINS], b[CHAINS];
    int ia[CHAINS];
    unsigned long long p[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
        a[c] = seed + c + threadIdx.x;
        b[c] = seed * 0.5f + c;
        ia[c] = c + threadIdx.x;
        float2 t = make_float2(a[c], b[c]);
        p[c] = *reinterpret_cast<unsigned long long*>(&t);
    }
    const float m = seed * 0.999f, d = seed * 0.001f;
    float2 m2f = make_float2(m, m), d2f = make_float2(d, d);
    unsigned long long m2 = *reinterpret_cast<unsigned long long*>(&m2f), d2 = *reinterpret_cast<unsigned long long*>(&d2f);
    for (int it = 0; it < n_iter; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m), "f"(d));
            if (KIND == K_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(m2), "l"(d2));
            if (KIND == K_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(d2));
            if (KIND == K_FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(m2));
            if (KIND == K_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(d));
            if (KIND == K_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m));
            if (KIND == K_SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 16, 31, 0xffffffff;" : "+f"(a[c]));
            if (KIND == K_SEL) asm volatile("{.reg .pred q; setp.gt.f32 q, %1, 0f00000000; selp.f32 %0, %0, %2, q;}" : "+f"(a[c]) : "f"(b[c]), "f"(d));
            if (KIND == K_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
            if (KIND == K_LOP3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(ia[c]) : "r"(it));
            if (KIND == K_LDS128) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(ia[c] + it) & 63])));
                a[c] += v.x; ia[c] += 0;
            }
            if (KIND == K_MIX_FFMA_IADD) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m), "f"(d));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[c]) : "r"(it));
            }
            if (KIND == K_MIX_FFMA2_IADD) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(m2), "l"(d2));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[c]) : "r"(it));
            }
            if (KIND == K_MIX_FFMA_FADD) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m), "f"(d));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[c]) : "f"(d));
            }
            if (KIND == K_MIX_FFMA2_FFMA) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(m2), "l"(d2));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m), "f"(d));
            }
            if (KIND == K_MIX_SHFL_FFMA) {
                asm volatile("shfl.sync.bfly.b32 %0, %0, 16, 31, 0xffffffff;" : "+f"(b[c]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m), "f"(d));
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
        float2 t = *reinterpret_cast<float2*>(&p[c]);
        acc += a[c] + b[c] + (float)ia[c] + t.x + t.y;
    }
    if (acc == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int KIND>
void run(float* out, int sms, double clk_ghz) {
    const int ctas = sms * 4;  // 4 CTAs x 8 warps = 32 warps per SM
    bench<KIND><<<ctas, 256>>>(out, 64, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        bench<KIND><<<ctas, 256>>>(out, ITERS, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warp_inst = (double)ctas * 8 * ITERS * CHAINS * kind_inst[KIND];
    const double per_clk_sm = warp_inst / (best * 1e-3 * clk_ghz * 1e9) / sms;
    printf("%-22s %8.3f ms  %6.3f warp-inst/clk/SM (at %.3f GHz nominal)  %7.1f G warp-inst/s\n", kind_name[KIND], best, per_clk_sm,
           clk_ghz, warp_inst / (best * 1e-3) * 1e-9);
}

template <int K>
struct RunAll {
    static void go(float* out, int sms, double clk) {
        run<K>(out, sms, clk);
        RunAll<K + 1>::go(out, sms, clk);
    }
};
template <>
struct RunAll<K_NUM> {
    static void go(float*, int, double) {}
};

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("# %s, %d SMs, max clock %.3f GHz\n", prop.name, prop.multiProcessorCount, clk_khz * 1e-6);
    float* out;
    cudaMalloc(&out, sizeof(float) * 148 * 4 * 256 * 4);
    RunAll<0>::go(out, prop.multiProcessorCount, clk_khz * 1e-6);
    return 0;
}
