// ffma2_issue.cu -- does the issue port stay busy during the second cycle of an FFMA2?
// Per iteration each warp runs NF packed FMAs (fma.rn.f32x2) interleaved with NA integer adds (alu pipe) or NL shared
// loads.  If the shadow cycle of an FFMA2 can issue another instruction, time(NF, NA) = max(2 NF, NF + NA) cycles per
// SMSP-warp-iteration at saturation; if not, time = 2 NF + NA.
#include <cstdio>
#include <cuda_runtime.h>

template <int NA, int MODE>
__global__ void __launch_bounds__(512) k(float *sink, int iters, float seed, int *isink)
{
    __shared__ float sm[1024];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 256] = seed; sm[threadIdx.x + 512] = seed; sm[threadIdx.x + 768] = seed;
    __syncthreads();
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0ull;
    int ia[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ia[i] = threadIdx.x + i;
    float fl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) fl[i] = seed;
    const float m = seed * 1.0000001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("{ .reg .b64 vb, sb; mov.b64 vb, {%1,%2}; mov.b64 sb, {%3,%3}; fma.rn.f32x2 %0, vb, sb, %0; }"
                         : "+l"(acc[i]) : "f"(fl[i]), "f"(fl[(i + 1) & 7]), "f"(m));
            if (MODE == 0) {
#pragma unroll
                for (int a = 0; a < NA; ++a) asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[(i + a) & 7]) : "r"(it));
            } else {
#pragma unroll
                for (int a = 0; a < NA; ++a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + ((threadIdx.x + 32 * a + i) & 1023)))); fl[(i + a) & 7] = v; }
            }
        }
    }
    float s = 0.f; int is = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32)) + fl[i]; is += ia[i]; }
    if (s == 123.456f) sink[0] = s;
    if (is == 12345) isink[0] = is;
}

template <int NA, int MODE>
void run(const char *name, float *sink, int *isink, int sms, int warps_per_smsp)
{
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 32 * 4 * warps_per_smsp;
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<NA, MODE><<<sms, threads>>>(sink, iters, 1.0f, isink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cyc = best * 1e-3 * clk_khz * 1e3;
    // per SMSP: warps_per_smsp warps x iters x 8 x (1 FFMA2 + NA other)
    const double per_group = cyc / ((double)iters * 8 * warps_per_smsp);
    printf("%s NA=%d warps/SMSP=%d: %.3f ms, %.2f cycles per (FFMA2 + %d other) per SMSP (nominal clock)\n", name, NA, warps_per_smsp, best, per_group, NA);
}

int main()
{
    float *sink; int *isink; cudaMalloc(&sink, 64); cudaMalloc(&isink, 64);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {2, 3, 4}) {
        run<0, 0>("iadd", sink, isink, sms, w);
        run<1, 0>("iadd", sink, isink, sms, w);
        run<2, 0>("iadd", sink, isink, sms, w);
        run<3, 0>("iadd", sink, isink, sms, w);
        run<1, 1>("lds ", sink, isink, sms, w);
        run<2, 1>("lds ", sink, isink, sms, w);
    }
    return 0;
}
