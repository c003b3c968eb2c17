// tools/umma_bench.cu -- microbenchmark: cycles per tcgen05.mma (M=128, K=16, fp16) as a function of
// the shared-memory operand layout (swizzle mode, LBO/SBO strides, start-address misalignment) and N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu
// Data content is irrelevant for timing (smem is zero-filled).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}
struct Cfg { int N; uint32_t a_lbo, a_sbo, a_layout, a_shift, b_lbo, b_sbo, b_layout, kstep_a, kstep_b; int nmma; int conv_like; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem) + c.a_shift, b0 = smem_u32(smem) + 32 * 1024;
        long long t0 = clock64();
        for (int i = 0; i < c.nmma; ++i) {
            const int k = i & 3;
            uint32_t aoff = k * c.kstep_a, boff = k * c.kstep_b;
            if (c.conv_like) {  // 9 taps x 4 k-steps: tap shifts of (dy*32+dx)*16 B on A, a fresh weight block per (tap, k) on B
                const int j = i % 36, tap = j >> 2;
                aoff = ((tap / 3) * 32 + tap % 3) * 16 + (j & 3) * c.kstep_a;
                boff = (uint32_t)j * (c.N * 32);
                if (c.conv_like == 2) boff = (j & 3) * (c.N * 32);      // A distinct, B reused
                if (c.conv_like == 3) aoff = (j & 3) * c.kstep_a;       // B distinct, A reused
            }
            const uint64_t ad = make_desc(a0 + aoff, c.a_lbo, c.a_sbo, c.a_layout, 0);
            const uint64_t bd = make_desc(b0 + boff, c.b_lbo, c.b_sbo, c.b_layout, 0);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
    long long *out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { const char *name; Cfg c; } cases[] = {
        // no-swizzle, pixel-major planes (what conv_tc uses): LBO = 3072 plane stride, SBO = 128
        {"noswz N64  LBO3072 aligned    ", {64, 3072, 128, 0, 0, 1024, 128, 0, 6144, 2048, 4096}},
        {"noswz N64  LBO3072 shift16    ", {64, 3072, 128, 0, 16, 1024, 128, 0, 6144, 2048, 4096}},
        {"noswz N64  LBO3072 shift528   ", {64, 3072, 128, 0, 528, 1024, 128, 0, 6144, 2048, 4096}},
        {"noswz N64  LBO3136(+64)       ", {64, 3136, 128, 0, 0, 1024, 128, 0, 6272, 2048, 4096}},
        {"noswz N64  LBO128 SBO256 (std)", {64, 128, 256, 0, 0, 128, 256, 0, 4096, 2048, 4096}},
        {"noswz N128 LBO3072 aligned    ", {128, 3072, 128, 0, 0, 2048, 128, 0, 6144, 4096, 4096}},
        {"noswz N256 LBO3072 aligned    ", {256, 3072, 128, 0, 0, 4096, 128, 0, 6144, 8192, 4096}},
        {"noswz N64  conv-like A+B distinct", {64, 3072, 128, 0, 0, 1024, 128, 0, 6144, 2048, 3600, 1}},
        {"noswz N64  conv-like A distinct  ", {64, 3072, 128, 0, 0, 1024, 128, 0, 6144, 2048, 3600, 2}},
        {"noswz N64  conv-like B distinct  ", {64, 3072, 128, 0, 0, 1024, 128, 0, 6144, 2048, 3600, 3}},
        {"noswz N128 conv-like A+B distinct", {128, 3072, 128, 0, 0, 2048, 128, 0, 6144, 4096, 3600, 1}},
        {"noswz N128 conv-like A distinct  ", {128, 3072, 128, 0, 0, 2048, 128, 0, 6144, 4096, 3600, 2}},
        {"noswz N32  conv-like A+B distinct", {32, 3072, 128, 0, 0, 512, 128, 0, 6144, 1024, 3600, 1}},
        {"noswz N16  conv-like A+B distinct", {16, 3072, 128, 0, 0, 256, 128, 0, 6144, 512, 3600, 1}},
        // 128B swizzle, K-major, rows of 128 B (64 fp16), 8-row atoms of 1024 B; k-step = +32 B
        {"sw128 N64  aligned            ", {64, 16, 1024, 2, 0, 16, 1024, 2, 32, 32, 4096}},
        {"sw128 N64  shift128 (1 row)   ", {64, 16, 1024, 2, 128, 16, 1024, 2, 32, 32, 4096}},
        {"sw128 N128 aligned            ", {128, 16, 1024, 2, 0, 16, 1024, 2, 32, 32, 4096}},
        {"sw128 N256 aligned            ", {256, 16, 1024, 2, 0, 16, 1024, 2, 32, 32, 4096}},
        // 32B swizzle (rows of 32 B = 16 fp16): one K=16 step per row; SBO = 256
        {"sw32  N64  aligned            ", {64, 16, 256, 6, 0, 16, 256, 6, 4096, 2048, 4096}},
        {"sw64  N64  aligned            ", {64, 16, 512, 4, 0, 16, 512, 4, 32, 32, 4096}},
    };
    for (auto &cs : cases) {
        bench<<<148, 128, 200 * 1024>>>(cs.c, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
        printf("%s  %s  cycles/MMA = %.1f  (ideal N/2 = %d)\n", cs.name, e == cudaSuccess ? "ok " : cudaGetErrorString(e),
               (double)cyc / cs.c.nmma, cs.c.N / 2);
        if (e != cudaSuccess) break;
    }
    return 0;
}
