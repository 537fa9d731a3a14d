// Microbenchmark: how fast does one CTA retire back-to-back cp.async.bulk (UBLKCP) copies of a given size?
// Prints ns per copy and GB/s per SM for N copies in flight (one mbarrier per batch), sizes 512 B .. 16 KB, all 148 SMs active.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// mode 0: one thread issues all n copies of a batch; mode 1: n different warps issue one copy each; mode 2: n lanes of one warp
__global__ void k(const uint8_t* src, size_t src_bytes, int n, int bytes, int rounds, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < rounds; ++r) {
        if (threadIdx.x == 0) mbar_expect(smem_u32(&bar), (uint32_t)n * bytes);
        __syncthreads();
        if (mode == 0) {
            if (threadIdx.x == 0)
                for (int i = 0; i < n; ++i) {
                    size_t off = ((size_t)(blockIdx.x * 7919 + r * 131 + i * 17) * 16384) % (src_bytes - 16384);
                    bulk(smem_u32(smem) + i * bytes, src + off, bytes, smem_u32(&bar));
                }
        } else if (mode == 2) {
            if (warp == 0 && lane < n) {                      // one instruction, n active lanes
                size_t off = ((size_t)(blockIdx.x * 7919 + r * 131 + lane * 17) * 16384) % (src_bytes - 16384);
                bulk(smem_u32(smem) + lane * bytes, src + off, bytes, smem_u32(&bar));
            }
        } else if (warp < n && lane == 0) {
            size_t off = ((size_t)(blockIdx.x * 7919 + r * 131 + warp * 17) * 16384) % (src_bytes - 16384);
            bulk(smem_u32(smem) + warp * bytes, src + off, bytes, smem_u32(&bar));
        }
        if (threadIdx.x == 0) while (!mbar_try(smem_u32(&bar), ph)) {}
        ph ^= 1;
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
    const size_t src_bytes = 64ull << 20;    // L2 resident
    uint8_t* src; long long* out;
    cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("mode n bytes  ns_per_batch  ns_per_copy  GB/s_per_SM (SM clock %d kHz nominal)\n", clk_khz);
    for (int mode = 0; mode < 3; ++mode)
        for (int bytes : {512, 2048, 16384})
            for (int n : {1, 2, 4, 8, 12}) {
                if ((size_t)n * bytes > 196 * 1024) continue;
                const int rounds = 2000;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k<<<148, 512, 200 * 1024>>>(src, src_bytes, n, bytes, 50, mode, out);
                cudaEventRecord(e0);
                k<<<148, 512, 200 * 1024>>>(src, src_bytes, n, bytes, rounds, mode, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                double ns_batch = ms * 1e6 / rounds;
                printf("%d %2d %5d  %9.1f  %9.1f  %8.2f\n", mode, n, bytes, ns_batch, ns_batch / n, (double)n * bytes / ns_batch);
            }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
