// Issue rate of the packed FP32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2, PTX *.f32x2) against
// their scalar forms, alone and interleaved with integer work. Each thread runs 8 independent
// chains; 148 x 8 CTAs x 256 threads. Prints lane-operations per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/f32x2_rate tools/f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 32768;
constexpr int kChains = 8;

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float* out, float seed, int iters) {
    float a[kChains], b[kChains];
    unsigned long long v[kChains];
    unsigned iacc[kChains];
    const float m = seed * 1.0000001f, c = seed * 1e-9f;
    const unsigned long long m2 = pk(m, m), c2 = pk(c, c);
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = seed + i + threadIdx.x;
        b[i] = seed - i;
        v[i] = pk(a[i], b[i]);
        iacc[i] = threadIdx.x * 7u + i;
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (MODE == 0) {   // scalar FFMA, 2 per chain step (same lane-operations as one FFMA2)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(c));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(m), "f"(c));
            } else if (MODE == 1) {   // FFMA2
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(m2), "l"(c2));
            } else if (MODE == 2) {   // FADD2
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(c2));
            } else if (MODE == 3) {   // FMUL2
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(m2));
            } else if (MODE == 4) {   // FFMA2 + one integer op per FFMA2
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(m2), "l"(c2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iacc[i]) : "r"(0x9e3779b9u), "r"((unsigned)it));
            } else if (MODE == 5) {   // 2 scalar FFMA + one integer op
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(c));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(m), "f"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iacc[i]) : "r"(0x9e3779b9u), "r"((unsigned)it));
            } else if (MODE == 6) {   // FFMA2 + FMNMX (the pair-loop mix: packed arithmetic, scalar min)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(m2), "l"(c2));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(m));
            } else if (MODE == 7) {   // 2 scalar FFMA + FMNMX
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(c));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(m), "f"(c));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(m));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
        s += a[i] + b[i] + lo + hi + (float)iacc[i];
    }
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
static void run(const char* name, float* out, double lane_ops_per_step, int clock_khz, int sms) {
    const int grid = sms * 8;
    for (int w = 0; w < 3; ++w) rate_kernel<MODE><<<grid, 256>>>(out, 1.0f, kIters);   // warm clocks
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        rate_kernel<MODE><<<grid, 256>>>(out, 1.0f, kIters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        ms = t < ms ? t : ms;
    }
    const double steps = (double)grid * 256 * kIters * kChains;
    const double clocks = ms * 1e-3 * clock_khz * 1e3;
    printf("%-44s %8.3f ms  %7.1f FP32 lane-ops/clk/SM  (%.2f issue slots/clk/SM for the whole mix)\n", name, ms,
           steps * lane_ops_per_step / clocks / sms, steps * (MODE == 0 ? 2 : MODE == 5 || MODE == 7 ? 3 : MODE == 4 || MODE == 6 ? 2 : 1) / 32.0 / clocks / sms);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, 4);
    printf("%s, %d SMs, clock attr %d kHz (rates are per nominal clock)\n", p.name, p.multiProcessorCount, khz);
    run<0>("2 x FFMA (scalar)", out, 2, khz, p.multiProcessorCount);
    run<1>("FFMA2", out, 2, khz, p.multiProcessorCount);
    run<2>("FADD2", out, 2, khz, p.multiProcessorCount);
    run<3>("FMUL2", out, 2, khz, p.multiProcessorCount);
    run<5>("2 x FFMA + LOP3", out, 2, khz, p.multiProcessorCount);
    run<4>("FFMA2 + LOP3", out, 2, khz, p.multiProcessorCount);
    run<7>("2 x FFMA + FMNMX", out, 2, khz, p.multiProcessorCount);
    run<6>("FFMA2 + FMNMX", out, 2, khz, p.multiProcessorCount);
    return 0;
}
