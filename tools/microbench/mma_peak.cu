// Microbenchmark: tensor-pipe ceiling of the MMA form the sparse-conv kernel uses -- tcgen05.mma cta_group::1, M = 128,
// A operand in tensor memory (".ts"), B operand a K-major SWIZZLE_128B shared-memory tile -- for kind::f16 and kind::i8 and
// N = 16 .. 256.  One CTA per SM issues `iters` back-to-back MMAs of one 32-byte k-step each (K = 16 fp16 / 32 int8) on
// fixed operands (no loads, no epilogue), committing every 64.  Prints TFLOP/s (TOPS) over all SMs: the denominator for the
// "INT8 tensor-pipe utilisation" the bench reports (MEASURED_PEAKS.json has no INT8 entry).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../quantization-on-3d-object-detection_b200/csrc \
//        -o mma_peak mma_peak.cu && ./mma_peak
#include "ql_common.cuh"
#include <stdio.h>

template <bool kInt8>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (kInt8)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <bool kInt8>
__global__ void __launch_bounds__(128, 1) k_peak(int n, int iters) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar[4];                            // batch b commits to bar[b % 4]: one outstanding arrival per barrier
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (ql_smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ql_mbar_init(ql_smem_u32(&bar[i]), 1); ql_fence_mbar_init(); }
    if (threadIdx.x < 32) { ql_tmem_alloc(ql_smem_u32(&tmem_base_s), 512); ql_tmem_relinquish(); }
    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        uint32_t idesc = 0;
        if (kInt8) idesc |= (2u << 4) | (1u << 7) | (1u << 10); else idesc |= 1u << 4;
        idesc |= (uint32_t)(n >> 3) << 17;
        idesc |= (uint32_t)(128 >> 4) << 24;
        uint64_t bd = 0;                                   // [n rows x 128 B] K-major SWIZZLE_128B tile at `base`
        bd |= (uint64_t)((base & 0x3FFFFu) >> 4);
        bd |= (uint64_t)1 << 16;
        bd |= (uint64_t)(1024 >> 4) << 32;
        bd |= (uint64_t)1 << 46;
        bd |= (uint64_t)2 << 61;
        const int batches = iters >> 6;
        for (int b = 0; b < batches; ++b) {
            if (b >= 4) ql_mbar_wait(ql_smem_u32(&bar[b & 3]), (uint32_t)(((b >> 2) - 1) & 1));    // batch b - 4 has retired
#pragma unroll 8
            for (int j = 0; j < 64; ++j)                   // k-step j%4 of the 128-byte row: +2 in the descriptor, +8 A columns
                mma_ts<kInt8>(tmem, tmem + 256u + (uint32_t)((j & 3) * 8), bd + (uint64_t)((j & 3) * 2), idesc, 1u);
            ql_tc_commit(ql_smem_u32(&bar[b & 3]));
        }
        for (int b = batches > 4 ? batches - 4 : 0; b < batches; ++b) ql_mbar_wait(ql_smem_u32(&bar[b & 3]), (uint32_t)((b >> 2) & 1));
    }
    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    if (threadIdx.x < 32) ql_tmem_dealloc(tmem, 512);
}

// The conv kernel's issue pattern: per UNIT one mbarrier wait (already complete), tcgen05.fence::after_thread_sync, 4 MMAs and one
// tcgen05.commit -- what does the single issuing thread pay per unit beyond the 4 MMAs?
template <bool kInt8>
__global__ void __launch_bounds__(128, 1) k_unit_pattern(int n, int units, int with_wait, int with_commit, int issuers, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_done, bar_sink[16], bar_final[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (ql_smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        ql_mbar_init(ql_smem_u32(&bar_done), 1);
        for (int i = 0; i < 16; ++i) ql_mbar_init(ql_smem_u32(&bar_sink[i]), 1);
        ql_mbar_init(ql_smem_u32(&bar_final[0]), 1);
        ql_mbar_init(ql_smem_u32(&bar_final[1]), 1);
        ql_fence_mbar_init();
    }
    if (threadIdx.x < 32) { ql_tmem_alloc(ql_smem_u32(&tmem_base_s), 512); ql_tmem_relinquish(); }
    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int w = threadIdx.x >> 5;                        // issuer w (warp w, lane 0) accumulates into columns w * 128 ..
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        uint32_t idesc = 0;
        if (kInt8) idesc |= (2u << 4) | (1u << 7) | (1u << 10); else idesc |= 1u << 4;
        idesc |= (uint32_t)(n >> 3) << 17;
        idesc |= (uint32_t)(128 >> 4) << 24;
        uint64_t bd = 0;
        bd |= (uint64_t)((base & 0x3FFFFu) >> 4);
        bd |= (uint64_t)1 << 16;
        bd |= (uint64_t)(1024 >> 4) << 32;
        bd |= (uint64_t)1 << 46;
        bd |= (uint64_t)2 << 61;
        const uint32_t d = tmem + (uint32_t)(w * 128), a0 = tmem + 256u + (uint32_t)(w * 32);
        const long long t0 = clock64();
        for (int u = 0; u < units; ++u) {
            if (with_wait) { ql_mbar_wait(ql_smem_u32(&bar_done), 1u); ql_tc_fence_after(); }   // parity 1 of a fresh barrier: complete
#pragma unroll
            for (int j = 0; j < 4; ++j)
                mma_ts<kInt8>(d, a0 + (uint32_t)(j * 8), bd + (uint64_t)(j * 2), idesc, 1u);
            if (with_commit) ql_tc_commit(ql_smem_u32(&bar_sink[(u & 7) + 8 * w]));                // arrivals just flip phases
        }
        const long long t1 = clock64();
        ql_tc_commit(ql_smem_u32(&bar_final[w]));
        ql_mbar_wait(ql_smem_u32(&bar_final[w]), 0u);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && w == 0) { cycles_out[0] = t1 - t0; cycles_out[1] = t2 - t0; }
    }
    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    if (threadIdx.x < 32) ql_tmem_dealloc(tmem, 512);
}

template <bool kInt8>
static double run(int n, int iters, int sms) {
    const size_t smem = 1024 + 256 * 128;
    cudaFuncSetAttribute(k_peak<kInt8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_peak<kInt8><<<sms, 128, smem>>>(n, 1024);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_peak<kInt8><<<sms, 128, smem>>>(n, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double kelems = kInt8 ? 32.0 : 16.0;
    return 2.0 * 128.0 * n * kelems * (double)iters * sms / (ms * 1e-3) / 1e12;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 1 << 18;
    printf("tcgen05.mma cta_group::1 M=128, A in TMEM, B in smem (SWIZZLE_128B), %d SMs, %d k-steps per CTA\n", sms, iters);
    printf("%6s %14s %14s\n", "N", "f16 TFLOP/s", "i8 TOPS");
    for (int n : {16, 32, 64, 128, 256}) {
        const double f = run<false>(n, iters, sms), i = run<true>(n, iters, sms);
        printf("%6d %14.1f %14.1f\n", n, f, i);
    }
    // sustained: the same N = 256 loop for >= 3 s of back-to-back launches (power / clock settling), against the burst figure above
    for (int kind = 0; kind < 2; ++kind) {
        const size_t smem = 1024 + 256 * 128;
        const int it2 = 1 << 20;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        float ms = 0.f;
        int launches = 0;
        do {
            for (int r = 0; r < 8; ++r) {
                if (kind) k_peak<true><<<sms, 128, smem>>>(256, it2); else k_peak<false><<<sms, 128, smem>>>(256, it2);
                ++launches;
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        } while (ms < 3000.f);
        const double kelems = kind ? 32.0 : 16.0;
        printf("sustained %s N=256: %.1f %s over %.2f s (%d launches)\n", kind ? "kind::i8 " : "kind::f16", 2.0 * 128.0 * 256 * kelems * (double)it2 * sms * launches / (ms * 1e-3) / 1e12,
               kind ? "TOPS" : "TFLOP/s", ms * 1e-3, launches);
    }
    long long* cyc;
    cudaMalloc(&cyc, 16);
    const size_t smem = 1024 + 256 * 128;
    cudaFuncSetAttribute(k_unit_pattern<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("\nper-UNIT cost for the issuing thread (4 MMAs per unit, kind::f16), cycles per unit: issue loop / until retired\n");
    printf("%6s %22s %22s %22s\n", "N", "4 MMAs only", "+ wait + fence", "+ wait + fence + commit");
    for (int n : {16, 64, 128}) {
        long long h[3][2];
        for (int v = 0; v < 3; ++v) {
            k_unit_pattern<false><<<sms, 128, smem>>>(n, 16384, v >= 1, v >= 2, 1, cyc);
            cudaMemcpy(h[v], cyc, 16, cudaMemcpyDeviceToHost);
        }
        printf("%6d %10.1f /%10.1f %10.1f /%10.1f %10.1f /%10.1f\n", n, h[0][0] / 16384.0, h[0][1] / 16384.0, h[1][0] / 16384.0, h[1][1] / 16384.0,
               h[2][0] / 16384.0, h[2][1] / 16384.0);
    }
    printf("\nTWO issuing threads (warps 0 and 1, separate accumulators), each 16384 units of 4 MMAs + wait + fence + commit: cycles per unit per thread\n");
    for (int n : {16, 64, 128}) {
        long long h1[2], h2[2];
        k_unit_pattern<false><<<sms, 128, smem>>>(n, 16384, 1, 1, 1, cyc);
        cudaMemcpy(h1, cyc, 16, cudaMemcpyDeviceToHost);
        k_unit_pattern<false><<<sms, 128, smem>>>(n, 16384, 1, 1, 2, cyc);
        cudaMemcpy(h2, cyc, 16, cudaMemcpyDeviceToHost);
        printf("%6d   one issuer %8.1f   two issuers %8.1f (x%.2f MMA rate)\n", n, h1[1] / 16384.0, h2[1] / 16384.0, 2.0 * h1[1] / h2[1]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
