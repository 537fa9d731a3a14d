// A/B microbenchmark: two ways of bringing gathered feature rows next to the tensor core on sm_100a.
//
//   A  "ldg_sttm"     the sparse-conv kernel's current producer: 16 warps read rulebook ids, load the rows with LDG.256
//                     (one lane per row up to 64-byte rows, four lanes per row above) and write them to tensor memory with
//                     tcgen05.st -- the A operand of a ".ts" MMA.  Missing neighbours (-1) read the all-zero row in front of the
//                     feature matrix.
//   B  "tma_gather4"  cp.async.bulk.tensor.2d.tile::gather4: one warp reads the ids (4 per lane) and hands them to the TMA unit,
//                     which writes 4 rows per instruction into a swizzled shared-memory slab -- the A operand of a ".ss" MMA.
//                     Missing neighbours are out-of-bounds coordinates: the TMA zero-fills them without touching memory.
//
// Neither leg issues MMAs or stores results: what is measured is the gather front end alone, all 148 SMs, persistent CTAs over
// tiles of 128 output rows x 27 kernel offsets with a synthetic neighbour table (40 % of the 27 neighbours present, present
// neighbours at row + small strides as in a key-sorted voxel list).  Output: time, gathered rows/s (all 27*128 per tile) and
// the bytes/s of rows that exist.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather_ab gather_ab.cu && ./gather_ab
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int kOffsets = 27;
constexpr int kTileM = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// bounded: a fill that never lands (bad tensor map) traps after ~2 s instead of hanging the box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    uint64_t t0 = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (!ok && (++spins & 255u) == 0) {
            uint64_t t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
}

// ------------------------------------------------------------------ A: LDG -> tcgen05.st
__device__ __forceinline__ void ldg32(const uint8_t* p, uint32_t* v) {
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void sttm8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// ROWB = bytes per feature row (32 .. 256).  16 producer warps = 4 teams x 4 warps; team t takes offsets t, t+4, ...
template <int ROWB>
__global__ void __launch_bounds__(512, 1) k_ldg_sttm(const uint8_t* __restrict__ feats, const int* __restrict__ nbr, int n_tiles) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int team = warp >> 2, q = warp & 3;                      // q = TMEM lane quarter of this warp
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    constexpr int kCols = ROWB / 4;                                // TMEM columns per gathered row
    constexpr bool kQuad = ROWB >= 128;
    uint32_t slot = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int k = team; k < kOffsets; k += 4) {
            const int* ids = nbr + ((int64_t)tile * kOffsets + k) * kTileM + q * 32;
            const uint32_t dst = tmem + lane_base + (uint32_t)(team * 128 + (slot & 1u) * 64u);   // two 64-column slots per team
            if constexpr (!kQuad) {
                const int id = __ldg(ids + lane);
                const uint8_t* src = feats + (int64_t)id * ROWB;
                uint32_t v[ROWB / 4];
#pragma unroll
                for (int c = 0; c < ROWB / 32; ++c) ldg32(src + c * 32, v + c * 8);
#pragma unroll
                for (int c = 0; c < ROWB / 32; ++c) sttm8(dst + (uint32_t)(c * 8), v + c * 8);
            } else {
                // four lanes per row: lane = 4 * (row % 8) + 32-byte piece; 4 passes of 8 rows per 128-byte chunk
#pragma unroll
                for (int ch = 0; ch < ROWB / 128; ++ch) {
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int id = __ldg(ids + i * 8 + (lane >> 2));
                        ldg32(feats + (int64_t)id * ROWB + ch * 128 + (lane & 3) * 32, v + i * 8);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) sttm8(dst + (uint32_t)(ch * 32 + i * 8), v + i * 8);   // same TMEM write volume as 16x256b
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            ++slot;
            (void)kCols;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------ B: TMA gather4 -> shared memory
// One issuing warp per `kIssuers`; each (tile, offset) slab = 128 rows x ROWB bytes = 32 gather4 per 128-byte column chunk,
// lane l fetching rows 4l .. 4l+3.  Ring of kDepth slabs; a slab is reused when its previous fill has landed.
template <int ROWB, int kDepth, int kIssuers>
__global__ void __launch_bounds__(32 * kIssuers, 1) k_tma_gather4(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ nbr,
                                                                  int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[kIssuers][kDepth];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr int kSlab = kTileM * ROWB;
    constexpr int kChunk = ROWB > 128 ? 128 : ROWB;                 // bytes per gather4 row piece (= the tensor map's box width)
    constexpr int kChunks = ROWB / kChunk;
    if (lane == 0) {
        for (int i = 0; i < kDepth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[warp][i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t my = base + (uint32_t)(warp * kDepth * kSlab);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int k = warp; k < kOffsets; k += kIssuers, ++it) {
            const uint32_t s = it % kDepth, use = it / kDepth;
            const uint32_t bar = smem_u32(&full[warp][s]);
            if (use > 0) mbar_wait(bar, (use - 1) & 1u);             // the slab's previous fill has landed ("consumed" at once)
            const int4 id = __ldg(reinterpret_cast<const int4*>(nbr + ((int64_t)tile * kOffsets + k) * kTileM) + lane);
            if (lane == 0)
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"((uint32_t)kSlab)
                             : "memory");
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                const uint32_t dst = my + s * (uint32_t)kSlab + (uint32_t)(ch * kTileM * kChunk) + (uint32_t)(lane * 4 * kChunk);
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
                    "l"(&tmap), "r"(ch * (kChunk / 2)), "r"(id.x), "r"(id.y), "r"(id.z), "r"(id.w), "r"(bar)
                    : "memory");
            }
        }
    }
    // drain
    for (uint32_t s = 0; s < (uint32_t)kDepth; ++s) {
        const uint32_t uses = it / kDepth + (s < it % kDepth ? 1u : 0u);
        if (uses > 0) mbar_wait(smem_u32(&full[warp][s]), (uses - 1) & 1u);
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWB>
void run(EncodeTiled encode, int n_rows, int sms) {
    const int n_tiles = (n_rows + kTileM - 1) / kTileM;
    // feature matrix with one leading zero row (leg A's -1 target)
    uint8_t* feats_alloc;
    CK(cudaMalloc(&feats_alloc, (size_t)(n_rows + 1) * ROWB));
    CK(cudaMemset(feats_alloc, 1, (size_t)(n_rows + 1) * ROWB));
    CK(cudaMemset(feats_alloc, 0, ROWB));
    uint8_t* feats = feats_alloc + ROWB;
    // neighbour table: offset k of row r is present with p = 0.4, at r + stride[k]
    static const int strides[kOffsets] = {-3901, -3850, -3799, -52, -1, 50, 3797, 3848, 3899, -3900, -3849, -3798, -51, 0,
                                          51,    3798,  3849,  3900, -3899, -3848, -3797, -50, 1,  52, 3799, 3850, 3901};
    std::vector<int> h((size_t)n_tiles * kOffsets * kTileM);
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    size_t present = 0;
    for (int t = 0; t < n_tiles; ++t)
        for (int k = 0; k < kOffsets; ++k)
            for (int r = 0; r < kTileM; ++r) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const int row = t * kTileM + r, src = row + strides[k];
                const bool on = row < n_rows && src >= 0 && src < n_rows && (k == 13 || (rng >> 40) % 100 < 38);
                h[((size_t)t * kOffsets + k) * kTileM + r] = on ? src : -1;
                present += on;
            }
    int* nbr;
    CK(cudaMalloc(&nbr, h.size() * 4));
    CK(cudaMemcpy(nbr, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double rows_total = (double)n_tiles * kOffsets * kTileM;
    auto report = [&](const char* name, float ms) {
        printf("  %-34s %8.1f us  %7.2f G rows/s  %7.1f GB/s of present rows (%.0f %% present)\n", name, ms * 1e3, rows_total / ms * 1e-6,
               (double)present * ROWB / ms * 1e-6, 100.0 * present / rows_total);
    };
    printf("row = %3d bytes (%d fp16 channels), %d rows, %d tiles x 27 offsets\n", ROWB, ROWB / 2, n_rows, n_tiles);
    {
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            CK(cudaEventRecord(e0));
            k_ldg_sttm<ROWB><<<sms, 512>>>(feats, nbr, n_tiles);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        report("A ldg_sttm (16 producer warps)", best);
    }
    {
        constexpr int kChunk = ROWB > 128 ? 128 : ROWB;
        CUtensorMap tmap;
        const cuuint64_t dims[2] = {(cuuint64_t)(ROWB / 2), (cuuint64_t)n_rows};
        const cuuint64_t gstrides[1] = {(cuuint64_t)ROWB};
        const cuuint32_t box[2] = {(cuuint32_t)(kChunk / 2), 1u};
        const cuuint32_t estr[2] = {1u, 1u};
        const CUtensorMapSwizzle sw = kChunk == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (kChunk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, feats, dims, gstrides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("  cuTensorMapEncodeTiled failed: %d\n", (int)r);
            return;
        }
        auto leg = [&](auto kern, int issuers, int depth, const char* name) {
            const size_t smem = (size_t)issuers * depth * kTileM * ROWB + 1024;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float best = 1e30f;
            for (int it = 0; it < 4; ++it) {
                CK(cudaEventRecord(e0));
                kern<<<sms, 32 * issuers, smem>>>(tmap, nbr, n_tiles);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it > 0 && ms < best) best = ms;
            }
            CK(cudaGetLastError());
            report(name, best);
        };
        constexpr int kD1 = (160 * 1024) / (kTileM * ROWB) < 16 ? (160 * 1024) / (kTileM * ROWB) : 16;
        constexpr int kD2 = kD1 / 2 < 1 ? 1 : kD1 / 2;
        constexpr int kD4 = kD1 / 4 < 1 ? 1 : kD1 / 4;
        leg(k_tma_gather4<ROWB, kD1, 1>, 1, kD1, "B tma_gather4, 1 issuing warp");
        leg(k_tma_gather4<ROWB, kD2, 2>, 2, kD2, "B tma_gather4, 2 issuing warps");
        leg(k_tma_gather4<ROWB, kD4, 4>, 4, kD4, "B tma_gather4, 4 issuing warps");
    }
    CK(cudaFree(nbr));
    CK(cudaFree(feats_alloc));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode || q != cudaDriverEntryPointSuccess) {
        fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    printf("%s, %d SMs; persistent CTAs, one per SM\n", prop.name, prop.multiProcessorCount);
    const int n_rows = 600000;
    run<32>(encode, n_rows, prop.multiProcessorCount);
    run<64>(encode, n_rows, prop.multiProcessorCount);
    run<128>(encode, n_rows / 2, prop.multiProcessorCount);
    run<256>(encode, n_rows / 4, prop.multiProcessorCount);
    return 0;
}
