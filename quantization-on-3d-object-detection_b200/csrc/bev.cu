// BEV densify: sparse (N, C) rows at [b, d, y, x] -> dense out[b, c*D + d, y, x], zero filled in the same pass.
//
// Replaces HeightCompression.forward -> [EXT] SparseConvTensor.dense() + permute + view
// (pcdet/models/backbones_2d/map_to_bev/height_compression.py:20-24), i.e. memset + scatter + permute copy.
// Here every output byte is written exactly once, coalesced along x: a CTA owns one (b, y, x-range) strip,
// resolves its D*XT cells through the coordinate hash, stages the present feature rows in shared memory
// (coalesced row reads) and then streams all C*D channel planes of the strip.
#include "ql_common.cuh"

namespace {

constexpr int kBevThreads = 256;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kBevThreads) k_bev_densify(const TIn* __restrict__ feats, int C, const uint2* __restrict__ table,
                                                             uint32_t cap_mask, QlGrid g, int XT, int row_pitch_f, TOut* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    int* s_slot = reinterpret_cast<int*>(smem_raw);                       // [D*XT] -> staged row slot or -1
    float* s_rows = reinterpret_cast<float*>(smem_raw + ((g.D * XT * 4 + 15) & ~15));   // [m][row_pitch_f]
    __shared__ int s_count;
    const int tiles_x = (g.W + XT - 1) / XT;
    const int xt = blockIdx.x % tiles_x;
    const int y = (blockIdx.x / tiles_x) % g.H;
    const int b = blockIdx.x / (tiles_x * g.H);
    const int x0 = xt * XT;
    const int xn = min(XT, g.W - x0);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // 1. resolve cells, compact the present ones
    for (int cell = threadIdx.x; cell < g.D * XT; cell += blockDim.x) {
        const int d = cell / XT, xi = cell % XT;
        int slot = -1;
        if (xi < xn) {
            const int idx = ql_hash_lookup(table, cap_mask, ql_key(g, b, d, y, x0 + xi));
            if (idx >= 0) {
                slot = atomicAdd(&s_count, 1);
                // remember the global row in the first float of the staged row; replaced by data below
                reinterpret_cast<int*>(s_rows + (size_t)slot * row_pitch_f)[0] = idx;
            }
        }
        s_slot[cell] = slot;
    }
    __syncthreads();
    const int m = s_count;
    // 2. stage rows (coalesced over channels); the row index sits in element 0 until overwritten, so read it first
    for (int j = threadIdx.x / 32; j < m; j += blockDim.x / 32) {
        float* dst = s_rows + (size_t)j * row_pitch_f;
        const int idx = reinterpret_cast<int*>(dst)[0];
        __syncwarp();
        const TIn* src = feats + (int64_t)idx * C;
        for (int c = threadIdx.x & 31; c < C; c += 32) dst[c] = (float)src[c];
    }
    __syncthreads();
    // 3. stream the C*D planes of this strip
    const int planes = C * g.D;
    for (int i = threadIdx.x; i < planes * xn; i += blockDim.x) {
        const int p = i / xn, xi = i % xn;
        const int c = p / g.D, d = p % g.D;
        const int slot = s_slot[d * XT + xi];
        const float v = slot >= 0 ? s_rows[(size_t)slot * row_pitch_f + c] : 0.f;
        out[(((int64_t)b * planes + p) * g.H + y) * g.W + x0 + xi] = (TOut)v;
    }
}

}  // namespace

extern "C" int ql_bev_densify(const void* feats, int32_t in_dtype, int32_t c, const uint64_t* table, int64_t table_cap, int32_t B,
                              int32_t D, int32_t H, int32_t W, void* out, int32_t out_dtype, ql_stream_t stream_) {
    if (!feats || !table || !out || c <= 0 || B <= 0 || D <= 0 || H <= 0 || W <= 0) return QL_ERR_INVALID;
    if (table_cap <= 0 || (table_cap & (table_cap - 1))) return QL_ERR_INVALID;
    if ((in_dtype != QL_F16 && in_dtype != QL_F32) || (out_dtype != QL_F16 && out_dtype != QL_F32)) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    QlGrid g{B, D, H, W};
    const int row_pitch_f = c + 1;                      // odd pitch (C even): conflict-free column reads
    // strip width: as wide as shared memory allows (worst case every cell present)
    int XT = W;
    auto smem_for = [&](int xt) { return (size_t)((D * xt * 4 + 15) & ~15) + (size_t)D * xt * row_pitch_f * 4; };
    while (XT > 1 && smem_for(XT) > 200 * 1024) XT = (XT + 1) / 2;
    if (smem_for(XT) > 200 * 1024) return QL_ERR_UNSUPPORTED;
    size_t smem = smem_for(XT);
    int tiles_x = (W + XT - 1) / XT;
    unsigned grid = (unsigned)((int64_t)B * H * tiles_x);
    cudaStream_t st = (cudaStream_t)stream_;
    uint32_t mask = (uint32_t)(table_cap - 1);
#define QL_BEV_LAUNCH(TI, TO)                                                                                             \
    do {                                                                                                                  \
        if (cudaFuncSetAttribute(k_bev_densify<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
            return QL_ERR_CUDA;                                                                                           \
        k_bev_densify<TI, TO><<<grid, kBevThreads, smem, st>>>((const TI*)feats, c, (const uint2*)table, mask, g, XT,     \
                                                               row_pitch_f, (TO*)out);                                    \
    } while (0)
    if (in_dtype == QL_F16 && out_dtype == QL_F16) QL_BEV_LAUNCH(__half, __half);
    else if (in_dtype == QL_F16 && out_dtype == QL_F32) QL_BEV_LAUNCH(__half, float);
    else if (in_dtype == QL_F32 && out_dtype == QL_F16) QL_BEV_LAUNCH(float, __half);
    else QL_BEV_LAUNCH(float, float);
#undef QL_BEV_LAUNCH
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
