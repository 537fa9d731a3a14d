// Rate of the warp-level tensor path (mma.sync: HMMA / IMMA) on sm_100a -- the ceiling of csrc/spconv_warp.cu, the
// register-gather conv for narrow layers.  Every warp keeps NACC independent accumulators and issues mma.sync back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int NACC, bool kInt8>
__global__ void __launch_bounds__(1024, 1) k_rate(int iters, float* sink) {
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
    uint32_t b[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
    float c[NACC][4];
    int ci[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[i][j] = 0.f; ci[i][j] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if constexpr (kInt8) {
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(ci[i][0]), "+r"(ci[i][1]), "+r"(ci[i][2]), "+r"(ci[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            } else {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j] + (float)ci[i][j];
    if (s == 12345.678f) sink[0] = s;
}

template <int NACC, bool kInt8>
void run(int threads, const char* name) {
    float* sink;
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k_rate<NACC, kInt8><<<148, threads>>>(100, sink);
    cudaEventRecord(e0);
    k_rate<NACC, kInt8><<<148, threads>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double macs = (double)148 * (threads / 32) * iters * NACC * 16.0 * 8.0 * (kInt8 ? 32.0 : 16.0);
    printf("%-28s warps/SM %2d  acc %d : %8.1f T%s/s  (%.3f ms, %s)\n", name, threads / 32, NACC, 2.0 * macs / ms * 1e-9,
           kInt8 ? "OP" : "FLOP", ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    for (int threads : {128, 256, 512, 1024}) {
        run<4, false>(threads, "m16n8k16 f16->f32");
        run<8, false>(threads, "m16n8k16 f16->f32");
        run<4, true>(threads, "m16n8k32 s8->s32");
        run<8, true>(threads, "m16n8k32 s8->s32");
    }
    return 0;
}
