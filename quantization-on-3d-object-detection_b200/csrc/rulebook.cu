// Rulebook (indice-pair) generation for submanifold and strided sparse convolution.
//
// Replaces the [EXT] spconv indice-pair generation that runs inside SubMConv3d / SparseConv3d.forward (call sites
// pcdet/models/backbones_3d/spconv_backbone.py:12-17,194-231; one build per indice_key, 9 per VoxelResBackBone8x
// forward).  Output format is the MaskImplicitGemm-style table the tcgen05 conv kernel consumes directly:
//     nbr[tile][k][128]  (tile = out_row / 128)  =  input row feeding out_row through kernel offset k, or -1
// so that one 128-row MMA tile's rulebook is a single contiguous K*512-byte TMA bulk copy.
//
// Probing is warp-wide: the 32 lanes of a warp own 32 consecutive output rows (coalesced int4 coord loads and
// coalesced 128-byte rulebook stores per offset) and issue their open-addressing probes for a batch of kernel
// offsets before resolving any of them, so ~9 independent 8-byte table loads are in flight per lane.
//
// Strided conv output numbering is first-touch order over the enumeration (input row i, offset k) -> i*K+k
// (deterministic, reproducible by the oracle): insert (atomicMin of the sequence number) / count owners /
// scan / assign, then the same pair-gather kernel as the submanifold case.
#include "ql_common.cuh"
#include "ql_scan.cuh"

namespace {

struct ConvGeom {
    int kd, kh, kw;       // kernel (z, y, x)
    int sd, sh, sw;       // stride
    int pd, ph, pw;       // padding
};

constexpr int kProbeBatch = 9;
#define QL_ID_FLAG 0x80000000u

// nbr for output rows: in = out*stride - pad + offset, looked up in the input table.
__global__ void __launch_bounds__(QL_TILE_M) k_rb_pairs(const int4* __restrict__ out_coords, int64_t n_cap,
                                                        const int* __restrict__ n_dev, QlGrid gin, ConvGeom cg,
                                                        const uint2* __restrict__ table, uint32_t cap_mask,
                                                        int* __restrict__ nbr) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_cap;
    const int64_t tile = blockIdx.x;
    if (tile * QL_TILE_M >= n) return;                       // tiles past the device-side row count are never read
    const int r = threadIdx.x;
    const int64_t row = tile * QL_TILE_M + r;
    const int K = cg.kd * cg.kh * cg.kw;
    int* dst = nbr + tile * (int64_t)K * QL_TILE_M + r;
    if (row >= n) {
        for (int k = 0; k < K; ++k) dst[(int64_t)k * QL_TILE_M] = -1;
        return;
    }
    const int4 c = out_coords[row];                          // b, z, y, x
    const int bz = c.y * cg.sd - cg.pd, by = c.z * cg.sh - cg.ph, bx = c.w * cg.sw - cg.pw;
    for (int k0 = 0; k0 < K; k0 += kProbeBatch) {
        uint32_t key[kProbeBatch], slot[kProbeBatch];
        uint2 e[kProbeBatch];
        bool ok[kProbeBatch];
#pragma unroll
        for (int j = 0; j < kProbeBatch; ++j) {
            int k = k0 + j;
            int kz = k / (cg.kh * cg.kw), ky = (k / cg.kw) % cg.kh, kx = k % cg.kw;
            int z = bz + kz, y = by + ky, x = bx + kx;
            ok[j] = k < K && z >= 0 && z < gin.D && y >= 0 && y < gin.H && x >= 0 && x < gin.W;
            key[j] = ok[j] ? ql_key(gin, c.x, z, y, x) : 0u;
            slot[j] = ql_hash_slot(key[j], cap_mask);
        }
#pragma unroll
        for (int j = 0; j < kProbeBatch; ++j)
            if (ok[j]) e[j] = __ldg(&table[slot[j]]);
#pragma unroll
        for (int j = 0; j < kProbeBatch; ++j) {
            int k = k0 + j;
            if (k >= K) break;
            int res = -1;
            if (ok[j]) {
                uint2 ee = e[j];
                uint32_t s = slot[j];
                while (ee.x != key[j] && ee.x != QL_HASH_EMPTY) {
                    s = (s + 1) & cap_mask;
                    ee = __ldg(&table[s]);
                }
                if (ee.x == key[j]) res = (int)ee.y;
            }
            dst[(int64_t)k * QL_TILE_M] = res;
        }
    }
}

// candidate output of input coord c through offset k (c + pad - k must be divisible by the stride and in range)
__device__ __forceinline__ bool out_candidate(const int4& c, int k, const ConvGeom& cg, const QlGrid& gout, uint32_t& key) {
    int kz = k / (cg.kh * cg.kw), ky = (k / cg.kw) % cg.kh, kx = k % cg.kw;
    int nz = c.y + cg.pd - kz, ny = c.z + cg.ph - ky, nx = c.w + cg.pw - kx;
    if (nz < 0 || ny < 0 || nx < 0) return false;
    if (nz % cg.sd || ny % cg.sh || nx % cg.sw) return false;
    int oz = nz / cg.sd, oy = ny / cg.sh, ox = nx / cg.sw;
    if (oz >= gout.D || oy >= gout.H || ox >= gout.W) return false;
    key = ql_key(gout, c.x, oz, oy, ox);
    return true;
}

__global__ void k_rb_insert(const int4* __restrict__ in_coords, int64_t n_cap, const int* __restrict__ n_dev, ConvGeom cg,
                            QlGrid gout, uint2* out_table, uint32_t cap_mask) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_cap;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = in_coords[i];
    const int K = cg.kd * cg.kh * cg.kw;
    for (int k = 0; k < K; ++k) {
        uint32_t key;
        if (!out_candidate(c, k, cg, gout, key)) continue;
        uint32_t s = ql_hash_insert(out_table, cap_mask, key);
        atomicMin(&out_table[s].y, (uint32_t)(i * K + k));
    }
}

// mode 0: count owned candidates per block; mode 1: assign ids, write coords, flag the table value with the id
template <int kMode>
__global__ void __launch_bounds__(QL_SCAN_THREADS) k_rb_number(const int4* __restrict__ in_coords, int64_t n_cap,
                                                              const int* __restrict__ n_dev, ConvGeom cg, QlGrid gout,
                                                              uint2* out_table, uint32_t cap_mask, int* block_counts,
                                                              int4* out_coords, int64_t n_out_cap) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_cap;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int K = cg.kd * cg.kh * cg.kw;
    int4 c = make_int4(0, 0, 0, 0);
    int owned = 0;
    if (i < n) {
        c = in_coords[i];
        for (int k = 0; k < K; ++k) {
            uint32_t key;
            if (!out_candidate(c, k, cg, gout, key)) continue;
            uint32_t s = ql_hash_find_slot(out_table, cap_mask, key);
            if (s != 0xFFFFFFFFu && out_table[s].y == (uint32_t)(i * K + k)) ++owned;
        }
    }
    int total;
    int ex = block_exclusive_scan(owned, total);
    if (kMode == 0) {
        if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
        return;
    }
    if (i >= n || owned == 0) return;
    int id = ex + block_counts[blockIdx.x];
    for (int k = 0; k < K; ++k) {
        uint32_t key;
        if (!out_candidate(c, k, cg, gout, key)) continue;
        uint32_t s = ql_hash_find_slot(out_table, cap_mask, key);
        if (s == 0xFFFFFFFFu || out_table[s].y != (uint32_t)(i * K + k)) continue;
        // sequence numbers are < 2^31, ids carry bit 31: a concurrent owner test on this slot can never match an id
        if ((int64_t)id < n_out_cap) {
            out_coords[id] = ql_unkey(gout, key);
            out_table[s].y = QL_ID_FLAG | (uint32_t)id;
        } else {
            out_table[s].y = 0xFFFFFFFFu;                    // overflow: dropped, lookups miss
        }
        ++id;
    }
}

__global__ void k_rb_strip_flag(uint2* table, int64_t cap) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    uint32_t y = table[s].y;
    if (y != 0xFFFFFFFFu && (y & QL_ID_FLAG)) table[s].y = y & ~QL_ID_FLAG;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

inline bool geom_from_host(const int32_t* k, const int32_t* s, const int32_t* p, ConvGeom& cg) {
    if (!k) return false;
    cg.kd = k[0]; cg.kh = k[1]; cg.kw = k[2];
    cg.sd = s ? s[0] : 1; cg.sh = s ? s[1] : 1; cg.sw = s ? s[2] : 1;
    cg.pd = p ? p[0] : k[0] / 2; cg.ph = p ? p[1] : k[1] / 2; cg.pw = p ? p[2] : k[2] / 2;
    if (cg.kd <= 0 || cg.kh <= 0 || cg.kw <= 0 || cg.sd <= 0 || cg.sh <= 0 || cg.sw <= 0) return false;
    if (cg.pd < 0 || cg.ph < 0 || cg.pw < 0) return false;
    if ((int64_t)cg.kd * cg.kh * cg.kw > 343) return false;
    return true;
}

}  // namespace

extern "C" int64_t ql_rulebook_num_tiles(int64_t n_out_cap) { return (n_out_cap + QL_TILE_M - 1) / QL_TILE_M; }

extern "C" int ql_rulebook_subm(const int32_t* coords, int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D, int32_t H,
                                int32_t W, const int32_t* ksize, const uint64_t* table, int64_t table_cap, int32_t* nbr_out,
                                ql_stream_t stream_) {
    ConvGeom cg;
    if (!coords || !table || !nbr_out || !geom_from_host(ksize, nullptr, nullptr, cg)) return QL_ERR_INVALID;
    if (!(cg.kd & 1) || !(cg.kh & 1) || !(cg.kw & 1)) return QL_ERR_INVALID;   // submanifold needs odd kernels
    if (table_cap <= 0 || (table_cap & (table_cap - 1))) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    if (n_cap <= 0) return QL_OK;
    QlGrid g{B, D, H, W};
    unsigned tiles = (unsigned)ql_rulebook_num_tiles(n_cap);
    k_rb_pairs<<<tiles, QL_TILE_M, 0, (cudaStream_t)stream_>>>((const int4*)coords, n_cap, n_dev, g, cg, (const uint2*)table,
                                                               (uint32_t)(table_cap - 1), nbr_out);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" size_t ql_rulebook_strided_workspace_bytes(int64_t n_in_cap, int32_t kvol) {
    (void)kvol;
    return align256((size_t)((n_in_cap + QL_SCAN_THREADS - 1) / QL_SCAN_THREADS + 1) * 4) + 256;
}

extern "C" int ql_rulebook_strided(const int32_t* in_coords, int64_t n_in_cap, const int32_t* n_in_dev, int32_t B, int32_t D,
                                   int32_t H, int32_t W, const int32_t* ksize, const int32_t* stride, const int32_t* pad,
                                   const uint64_t* in_table, int64_t in_table_cap, int32_t* out_coords, int64_t n_out_cap,
                                   int32_t* n_out_dev, uint64_t* out_table, int64_t out_table_cap, int32_t* nbr_out,
                                   void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    ConvGeom cg;
    if (!in_coords || !in_table || !out_coords || !n_out_dev || !out_table || !nbr_out || !workspace || !stride || !pad ||
        !geom_from_host(ksize, stride, pad, cg))
        return QL_ERR_INVALID;
    if (in_table_cap <= 0 || (in_table_cap & (in_table_cap - 1)) || out_table_cap <= 0 ||
        (out_table_cap & (out_table_cap - 1)) || out_table_cap < 2 * n_out_cap || n_out_cap <= 0)
        return QL_ERR_INVALID;
    const int K = cg.kd * cg.kh * cg.kw;
    if ((double)n_in_cap * K >= 2147483648.0) return QL_ERR_UNSUPPORTED;
    if (workspace_bytes < ql_rulebook_strided_workspace_bytes(n_in_cap, K)) return QL_ERR_WORKSPACE;
    QlGrid gin{B, D, H, W};
    QlGrid gout{B, (D + 2 * cg.pd - cg.kd) / cg.sd + 1, (H + 2 * cg.ph - cg.kh) / cg.sh + 1, (W + 2 * cg.pw - cg.kw) / cg.sw + 1};
    if (gout.D <= 0 || gout.H <= 0 || gout.W <= 0) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0 || (double)B * gout.D * gout.H * gout.W >= 4294967295.0)
        return QL_ERR_GRID_TOO_LARGE;

    int* block_counts = (int*)workspace;
    if (cudaMemsetAsync(out_table, 0xFF, (size_t)out_table_cap * 8, st) != cudaSuccess) return QL_ERR_CUDA;
    if (n_in_cap == 0) {
        if (cudaMemsetAsync(n_out_dev, 0, 8, st) != cudaSuccess) return QL_ERR_CUDA;
        return QL_OK;
    }
    uint32_t omask = (uint32_t)(out_table_cap - 1);
    unsigned nb = (unsigned)((n_in_cap + QL_SCAN_THREADS - 1) / QL_SCAN_THREADS);
    k_rb_insert<<<nb, QL_SCAN_THREADS, 0, st>>>((const int4*)in_coords, n_in_cap, n_in_dev, cg, gout, (uint2*)out_table, omask);
    k_rb_number<0><<<nb, QL_SCAN_THREADS, 0, st>>>((const int4*)in_coords, n_in_cap, n_in_dev, cg, gout, (uint2*)out_table, omask,
                                                   block_counts, (int4*)out_coords, n_out_cap);
    k_scan_blocks<<<1, QL_SCAN_THREADS, 0, st>>>(block_counts, (int)nb, n_out_dev + 1, n_out_dev, n_out_cap);
    k_rb_number<1><<<nb, QL_SCAN_THREADS, 0, st>>>((const int4*)in_coords, n_in_cap, n_in_dev, cg, gout, (uint2*)out_table, omask,
                                                   block_counts, (int4*)out_coords, n_out_cap);
    k_rb_strip_flag<<<(unsigned)((out_table_cap + 255) / 256), 256, 0, st>>>((uint2*)out_table, out_table_cap);
    unsigned tiles = (unsigned)ql_rulebook_num_tiles(n_out_cap);
    k_rb_pairs<<<tiles, QL_TILE_M, 0, st>>>((const int4*)out_coords, n_out_cap, n_out_dev, gin, cg, (const uint2*)in_table,
                                            (uint32_t)(in_table_cap - 1), nbr_out);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
