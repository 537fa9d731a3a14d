// BEV densify: sparse (N, C) rows at [b, d, y, x] -> dense out[b, c*D + d, y, x], zero filled in the same pass.
//
// Replaces HeightCompression.forward -> [EXT] SparseConvTensor.dense() + permute + view
// (pcdet/models/backbones_2d/map_to_bev/height_compression.py:20-24), i.e. memset + scatter + permute copy.
// Here every output byte is written exactly once with vector stores that are contiguous along x:
//   pass 1  k_bev_index : one hash lookup per grid cell -> idxmap[b][d][y][x] (row or -1), B*D*H*W*4 bytes
//   pass 2  k_bev_write : a thread owns VEC consecutive x of one (b, d, y) row and CC consecutive channels; it reads
//                         VEC cell indices (one vector load), one 16/32-byte feature segment per present cell, transposes
//                         in registers and writes CC plane rows (VEC elements each).  A warp's stores for one plane are
//                         32*VEC contiguous elements.
#include "ql_common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_bev_index(const uint2* __restrict__ table, uint32_t cap_mask, QlGrid g,
                                                   int* __restrict__ idxmap) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)g.B * g.D * g.H * g.W;
    if (cell >= n) return;
    idxmap[cell] = ql_hash_lookup(table, cap_mask, (uint32_t)cell);       // the linear cell index IS the key
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// kRanked: the cell -> row lookup is fused into the writer (no idxmap pass, no second launch): the VEC cells of a thread are VEC
// consecutive keys, VEC-aligned (VEC divides 32 and W), so they are bits of ONE bitmap word: row = word_prefix + popc(bits below).
template <typename TIn, typename TOut, int VEC, int CC, bool kRanked>
__global__ void __launch_bounds__(256) k_bev_write(const TIn* __restrict__ feats, int C, const int* __restrict__ idxmap, QlGrid g,
                                                   TOut* __restrict__ out, const uint32_t* __restrict__ bitmap,
                                                   const uint32_t* __restrict__ word_prefix, int64_t n_rows, const int* __restrict__ n_dev) {
    const int xgroups = g.W / VEC;
    const int cchunks = C / CC;
    const int64_t total = (int64_t)cchunks * g.B * g.D * g.H * xgroups;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int xg = (int)(t % xgroups); t /= xgroups;
    const int y = (int)(t % g.H); t /= g.H;
    const int d = (int)(t % g.D); t /= g.D;
    const int b = (int)(t % g.B); t /= g.B;
    const int c0 = (int)t * CC;
    const int64_t cell0 = (((int64_t)b * g.D + d) * g.H + y) * g.W + (int64_t)xg * VEC;
    int idx[VEC];
    if constexpr (kRanked) {
        const int64_t n = n_dev ? min((int64_t)*n_dev, n_rows) : n_rows;
        const uint32_t w = (uint32_t)(cell0 >> 5), sh = (uint32_t)(cell0 & 31);
        const uint32_t bits = __ldg(bitmap + w);
        const uint32_t mine = (bits >> sh) & ((VEC == 32) ? 0xFFFFFFFFu : ((1u << VEC) - 1u));
        uint32_t rank = mine ? __ldg(word_prefix + w) + (uint32_t)__popc(bits & ((1u << sh) - 1u)) : 0u;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const bool on = (mine >> v) & 1u;
            idx[v] = (on && (int64_t)rank < n) ? (int)rank : -1;
            rank += on ? 1u : 0u;
        }
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) idx[v] = idxmap[cell0 + v];
    }
    TOut vals[CC][VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (idx[v] >= 0) {
            const TIn* src = feats + (int64_t)idx[v] * C + c0;
            TIn seg[CC];
            if constexpr (CC * sizeof(TIn) == 16) {
                *reinterpret_cast<uint4*>(seg) = *reinterpret_cast<const uint4*>(src);
            } else if constexpr (CC * sizeof(TIn) == 32) {
                reinterpret_cast<uint4*>(seg)[0] = reinterpret_cast<const uint4*>(src)[0];
                reinterpret_cast<uint4*>(seg)[1] = reinterpret_cast<const uint4*>(src)[1];
            } else {
#pragma unroll
                for (int c = 0; c < CC; ++c) seg[c] = src[c];
            }
#pragma unroll
            for (int c = 0; c < CC; ++c) vals[c][v] = from_f<TOut>(to_f<TIn>(seg[c]));
        } else {
#pragma unroll
            for (int c = 0; c < CC; ++c) vals[c][v] = from_f<TOut>(0.f);
        }
    }
    const int64_t plane_stride = (int64_t)g.H * g.W;
    TOut* o = out + (((int64_t)b * C + c0) * g.D + d) * plane_stride + (int64_t)y * g.W + (int64_t)xg * VEC;
#pragma unroll
    for (int c = 0; c < CC; ++c) {
        TOut* dst = o + (int64_t)c * g.D * plane_stride;
        if constexpr (VEC * sizeof(TOut) == 16) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals[c]);
        } else if constexpr (VEC * sizeof(TOut) == 8) {
            *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(vals[c]);
        } else if constexpr (VEC * sizeof(TOut) == 4) {
            *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(vals[c]);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) dst[v] = vals[c][v];
        }
    }
}

struct RankArgs {
    const uint32_t* bitmap;
    const uint32_t* word_prefix;
    int64_t n_rows;
    const int* n_dev;
};

template <typename TIn, typename TOut, int VEC, int CC>
int launch_write(const void* feats, int C, const int* idxmap, QlGrid g, void* out, cudaStream_t st, const RankArgs* ra) {
    const int64_t total = (int64_t)(C / CC) * g.B * g.D * g.H * (g.W / VEC);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (ra)
        k_bev_write<TIn, TOut, VEC, CC, true><<<blocks, 256, 0, st>>>((const TIn*)feats, C, nullptr, g, (TOut*)out, ra->bitmap, ra->word_prefix,
                                                                       ra->n_rows, ra->n_dev);
    else
        k_bev_write<TIn, TOut, VEC, CC, false><<<blocks, 256, 0, st>>>((const TIn*)feats, C, idxmap, g, (TOut*)out, nullptr, nullptr, 0, nullptr);
    return 0;
}

template <typename TIn, typename TOut>
int dispatch_write(const void* feats, int C, const int* idxmap, QlGrid g, void* out, cudaStream_t st, const RankArgs* ra = nullptr) {
    constexpr int kMaxVec = 16 / (int)sizeof(TOut);                    // 8 for fp16, 4 for fp32
    const bool feat_vec = (C % 8 == 0);
    if (feat_vec) {
        // 16 channels (one 32-byte feature segment) per thread when they divide C: half the threads, twice the stores in flight each
        if (C % 16 == 0 && sizeof(TIn) == 2) {
            if (g.W % kMaxVec == 0) return launch_write<TIn, TOut, kMaxVec, 16>(feats, C, idxmap, g, out, st, ra);
            if (g.W % 4 == 0) return launch_write<TIn, TOut, 4, 16>(feats, C, idxmap, g, out, st, ra);
        }
        if (g.W % kMaxVec == 0) return launch_write<TIn, TOut, kMaxVec, 8>(feats, C, idxmap, g, out, st, ra);
        if (g.W % 4 == 0) return launch_write<TIn, TOut, 4, 8>(feats, C, idxmap, g, out, st, ra);
        if (g.W % 2 == 0) return launch_write<TIn, TOut, 2, 8>(feats, C, idxmap, g, out, st, ra);
        return launch_write<TIn, TOut, 1, 8>(feats, C, idxmap, g, out, st, ra);
    }
    if (g.W % 4 == 0) return launch_write<TIn, TOut, 4, 1>(feats, C, idxmap, g, out, st, ra);
    return launch_write<TIn, TOut, 1, 1>(feats, C, idxmap, g, out, st, ra);
}

}  // namespace

extern "C" size_t ql_bev_densify_workspace_bytes(int32_t B, int32_t D, int32_t H, int32_t W) {
    return (size_t)B * D * H * W * 4;
}

extern "C" int ql_bev_densify(const void* feats, int32_t in_dtype, int32_t c, const uint64_t* table, int64_t table_cap, int32_t B,
                              int32_t D, int32_t H, int32_t W, void* out, int32_t out_dtype, void* workspace, size_t workspace_bytes,
                              ql_stream_t stream_) {
    if (!feats || !table || !out || !workspace || c <= 0 || B <= 0 || D <= 0 || H <= 0 || W <= 0) return QL_ERR_INVALID;
    if (table_cap <= 0 || (table_cap & (table_cap - 1))) return QL_ERR_INVALID;
    if ((in_dtype != QL_F16 && in_dtype != QL_F32) || (out_dtype != QL_F16 && out_dtype != QL_F32)) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    if (workspace_bytes < ql_bev_densify_workspace_bytes(B, D, H, W)) return QL_ERR_WORKSPACE;
    QlGrid g{B, D, H, W};
    cudaStream_t st = (cudaStream_t)stream_;
    int* idxmap = (int*)workspace;
    const int64_t cells = (int64_t)B * D * H * W;
    k_bev_index<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>((const uint2*)table, (uint32_t)(table_cap - 1), g, idxmap);
    if (in_dtype == QL_F16 && out_dtype == QL_F16) dispatch_write<__half, __half>(feats, c, idxmap, g, out, st);
    else if (in_dtype == QL_F16 && out_dtype == QL_F32) dispatch_write<__half, float>(feats, c, idxmap, g, out, st);
    else if (in_dtype == QL_F32 && out_dtype == QL_F16) dispatch_write<float, __half>(feats, c, idxmap, g, out, st);
    else dispatch_write<float, float>(feats, c, idxmap, g, out, st);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_bev_densify_ranked(const void* feats, int32_t in_dtype, int32_t c, const uint32_t* bitmap, const uint32_t* word_prefix,
                                     int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D, int32_t H, int32_t W, void* out,
                                     int32_t out_dtype, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    if (!feats || !bitmap || !word_prefix || !out || !workspace || c <= 0 || B <= 0 || D <= 0 || H <= 0 || W <= 0 || n_cap < 0)
        return QL_ERR_INVALID;
    if ((in_dtype != QL_F16 && in_dtype != QL_F32) || (out_dtype != QL_F16 && out_dtype != QL_F32)) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    if (workspace_bytes < ql_bev_densify_workspace_bytes(B, D, H, W)) return QL_ERR_WORKSPACE;
    QlGrid g{B, D, H, W};
    cudaStream_t st = (cudaStream_t)stream_;
    // one kernel: the rank lookup is fused into the writer (round 1 ran an index pass into `workspace` first: two launches for a
    // 30 us stage); the workspace argument is kept for ABI stability and unused
    (void)workspace;
    const RankArgs ra{bitmap, word_prefix, n_cap, n_dev};
    if (in_dtype == QL_F16 && out_dtype == QL_F16) dispatch_write<__half, __half>(feats, c, nullptr, g, out, st, &ra);
    else if (in_dtype == QL_F16 && out_dtype == QL_F32) dispatch_write<__half, float>(feats, c, nullptr, g, out, st, &ra);
    else if (in_dtype == QL_F32 && out_dtype == QL_F16) dispatch_write<float, __half>(feats, c, nullptr, g, out, st, &ra);
    else dispatch_write<float, float>(feats, c, nullptr, g, out, st, &ra);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
