// Microbenchmark: what does ONE small tcgen05.mma cost on sm_100a?
//
// The sparse-conv kernel issues M=128, N=c_out (16..128), K=16 (f16) / K=32 (i8) MMAs whose A operand sits in tensor memory
// (the ".ts" form).  The role trace (profiles/r02_conv_role_trace.md) shows the issuing thread ~110 cycles per MMA at N=16 --
// against an arithmetic floor of 128*N/256 = 8 cycles.  This program measures the instruction in isolation: one CTA, one
// issuing thread, R back-to-back MMAs, then one commit; cycles = clock64 around issue (`issue`) and around issue + completion
// (`total`).  Swept: operand form (A in TMEM / A in shared memory), kind (f16 / i8), N, and the number of accumulators the
// stream rotates over (1 = every MMA depends on the one before it).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue mma_issue.cu && ./mma_issue
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));          \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of the fully active warp: ptxas recognises the pattern and issues UTCxMMA straight from uniform registers.  (An MMA
// issued under `threadIdx.x == 0` instead is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop -- ~150 cycles per MMA.)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <bool kInt8, bool kTs>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if constexpr (kTs) {
        if constexpr (kInt8)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                         "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                         "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
                         : "memory");
    } else {
        if constexpr (kInt8)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
                         : "memory");
    }
}

// K-major, SWIZZLE_32B descriptor of a [rows x 32 bytes] operand: 8-row atoms 256 bytes apart
__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}

struct Out {
    long long issue, total;
};

template <bool kInt8, bool kTs, bool kFixed>
__global__ void __launch_bounds__(128) k_bench(int n, int n_acc, int issuers, int reps, Out* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // zero the A slots (4 slots x 8 columns at column 448) so the arithmetic is defined
    {
        const uint32_t lane_base = (uint32_t)(threadIdx.x & ~31) << 16;
        for (int c = 0; c < 32; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_base + 448u + (uint32_t)c), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int wid = threadIdx.x >> 5;
    if (wid < issuers) {                                                    // the whole warp runs the loop; one elected lane issues
        const uint32_t tmem_d = tmem + (uint32_t)(wid * n_acc * n);        // this issuer's private accumulators
        uint32_t idesc = 0;
        if (kInt8) idesc |= (2u << 4) | (1u << 7) | (1u << 10);
        else idesc |= 1u << 4;
        idesc |= (uint32_t)(n >> 3) << 17;
        idesc |= (128u >> 4) << 24;
        const uint64_t a_desc = desc_sw32(sbase);                 // 128 rows x 32 B = 4 KB
        const uint64_t b_desc0 = desc_sw32(sbase + 8192u);        // up to 256 rows x 32 B = 8 KB per weight slab, 4 slabs
        const long long t0 = clock64();
        if constexpr (kFixed) {
            // the same operands every time, unrolled: nothing but the instruction itself on the issuing thread
            if (elect_one()) mma<kInt8, kTs>(tmem_d, tmem + 448u, a_desc, b_desc0, idesc, 0u);
#pragma unroll 16
            for (int r = 1; r < reps; ++r)
                if (elect_one()) mma<kInt8, kTs>(tmem_d, tmem + 448u, a_desc, b_desc0, idesc, 1u);
        } else {
            // the conv kernel's pattern: operand addresses change per MMA (A slot, weight slab, accumulator)
            uint32_t acc_i = 0;
            for (int r = 0; r < reps; ++r) {
                const uint32_t d = tmem_d + acc_i * (uint32_t)n;
                acc_i = (acc_i + 1u == (uint32_t)n_acc) ? 0u : acc_i + 1u;
                const uint32_t a = tmem + 448u + (uint32_t)((r & 3) * 8);
                const uint64_t b = b_desc0 + (uint64_t)(((r & 3) * 8192) >> 4);
                if (elect_one()) mma<kInt8, kTs>(d, a, a_desc, b, idesc, r >= n_acc ? 1u : 0u);
            }
        }
        const long long t1 = clock64();
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[wid])) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar[wid])), "r"(0u)
                         : "memory");
        const long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) {
            out[wid].issue = t1 - t0;
            out[wid].total = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <bool kInt8, bool kTs, bool kFixed>
void run(const char* name, Out* d_out) {
    const int reps = 4096;
    CK(cudaFuncSetAttribute(k_bench<kInt8, kTs, kFixed>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int n : {16, 32, 64, 128, 256}) {
        for (int issuers : {1, 2, 4}) {
            for (int n_acc : {1, 2}) {
                if (n * n_acc * issuers > 448) continue;
                if (kFixed && n_acc > 1) continue;
                Out h[4] = {};
                for (int it = 0; it < 2; ++it) {                             // the second launch is the measurement
                    k_bench<kInt8, kTs, kFixed><<<1, 128, 64 * 1024>>>(n, n_acc, issuers, reps, d_out);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
                long long issue = 0, total = 0;
                for (int i = 0; i < issuers; ++i) {
                    issue = h[i].issue > issue ? h[i].issue : issue;
                    total = h[i].total > total ? h[i].total : total;
                }
                // cycles per MMA of the WHOLE SM: the slowest issuer's time over all the MMAs issued
                printf("%-8s %-7s N=%3d issuers=%d accumulators/issuer=%d  issue %6.1f  total %6.1f cyc per MMA (SM-wide)  floor %5.1f\n", name,
                       kFixed ? "fixed" : "varying", n, issuers, n_acc, (double)issue / (reps * issuers), (double)total / (reps * issuers),
                       128.0 * n / 256.0);
            }
        }
    }
}

int main() {
    Out* d_out;
    CK(cudaMalloc(&d_out, 4 * sizeof(Out)));
    printf("tcgen05.mma cta_group::1, M=128, K=16 (f16) / 32 (i8); each issuer = lane 0 of its own warp, private accumulators,\n"
           "4096 back-to-back MMAs then one commit; 'fixed' = same operands, unrolled x16; 'varying' = A slot / weight slab / accumulator change per MMA\n");
    run<false, true, true>("f16 .ts", d_out);
    run<false, true, false>("f16 .ts", d_out);
    run<false, false, true>("f16 .ss", d_out);
    run<true, true, true>("i8  .ts", d_out);
    run<true, true, false>("i8  .ts", d_out);
    return 0;
}
