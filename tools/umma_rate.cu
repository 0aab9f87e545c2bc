// Issue rate of small tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as a function of N and of the shared-memory
// operand layout (no swizzle = the interleaved core-matrix layout the fused policy kernels use, vs 128-byte swizzle).
// One CTA per SM; one thread issues `chain` MMAs into the same accumulator, commits, waits; clock64 around it.
// Results are garbage (operands are zeros): only the time matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= 1ull << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int layout, int chain, int rounds, int distinct, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
        const uint32_t a0 = base, b0 = base + 16 * 1024;
        long long best = 1ll << 60;
        for (int r = 0; r < rounds; ++r) {
            const long long t0 = clock64();
            for (int c = 0; c < chain; ++c) {
                const uint32_t off = distinct ? (uint32_t)(c & 3) * (layout == 0 ? 4096u : 32u) : 0u;
                // no swizzle: LBO = 128 rows * 16 B, SBO = 128 B; 128B swizzle: SBO = 8 rows * 128 B, K advance = 32 B inside the row
                const uint64_t da = layout == 0 ? desc(a0 + off, 2048, 128, 0) : desc(a0 + off, 16, 1024, 2);
                const uint64_t db = layout == 0 ? desc(b0 + off, (uint32_t)N * 16, 128, 0) : desc(b0 + off, 16, 1024, 2);
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                    "l"(da), "l"(db), "r"(idesc), "r"(c == 0 ? 0u : 1u)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile(
                "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
                    smem_u32(&bar)),
                "r"((uint32_t)(r & 1))
                : "memory");
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        out[blockIdx.x] = best;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    printf("tcgen05.mma kind::f16 M=128 K=16, cycles for a chain of MMAs + commit + wait (min over rounds, mean over SMs)\n");
    for (int layout = 0; layout <= 2; layout += 2)
        for (int N : {16, 32, 64, 96, 128, 256})
            for (int chain : {1, 8, 32}) {
                rate_kernel<<<148, 128, 50 * 1024>>>(N, layout, chain, 20, 1, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("error %s\n", cudaGetErrorString(e));
                    return 1;
                }
                long long h[148];
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double acc = 0;
                for (int i = 0; i < 148; ++i) acc += (double)h[i];
                printf("%-12s N=%3d chain=%2d: %8.0f cycles  (%6.1f per MMA)\n", layout == 0 ? "no-swizzle" : "swizzle-128B", N, chain, acc / 148,
                       acc / 148 / chain);
            }
    return 0;
}
