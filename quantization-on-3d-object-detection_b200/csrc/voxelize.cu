// GPU-hash voxelization with fused mean VFE, first-touch voxel order (== the reference's CPU voxelizer order).
//
// Replaces: VoxelGeneratorWrapper.generate -> [EXT] spconv Point2VoxelCPU3d (data_processor.py:45-61,151-153),
//           collate_batch's batch column (dataset.py:237-244) and MeanVFE.forward (mean_vfe.py:25-29);
//           with max_pts == 0 the cap-free DynamicMeanVFE.forward (dynamic_mean_vfe.py:53-72).
//
// Passes (all stream ordered, no host sync):
//   1 insert   : point -> cell key -> hash slot (atomicCAS), slot.value = min point index (atomicMin)
//   2 count    : flag(p) = [slot.value == p]  (p is the first point of its voxel); per-block flag counts
//   3 scan     : exclusive scan of the block counts (one CTA)
//   4 assign   : voxel id = #flags before p (first-touch order); write coords, vox_first
//   5 gather   : hard mode: cascade atomicMin keeps the max_pts smallest point indices per voxel (deterministic);
//                dynamic mode: atomicAdd sums / counts
//   6 finalize : mean, num_points; hash value := voxel id (or -1 if dropped by the voxel cap)
#include "ql_common.cuh"
#include "ql_scan.cuh"

namespace {

constexpr int kThreads = QL_SCAN_THREADS;
constexpr int kItems = 4;                       // points per thread in the scan passes
constexpr int kBlockPts = kThreads * kItems;
constexpr int kMaxFrames = 4096;                // frames per batch the per-frame voxel cap can track

struct VoxParams {
    const float* points;
    int64_t n_points;
    int stride, has_b, n_feat;
    float mnx, mny, mnz, vsx, vsy, vsz;
    int gx, gy, gz;                             // grid size x,y,z
    QlGrid grid;                                // B, D=gz (caller's sparse_shape may be gz+1; key uses gz... see host)
    int max_pts;
    int64_t max_voxels;
    int out_stride;                             // floats per out_feats row (>= n_feat; the pad columns are written as zeros)
    int64_t frame_cap;                          // > 0: at most this many voxels per frame (MAX_NUMBER_OF_VOXELS, first-touch order)
    int* frames;                                // [3][B] ints: voxels found per frame, first provisional id, id shift
};

__device__ __forceinline__ bool point_cell(const VoxParams& P, int64_t p, int& b, int& cz, int& cy, int& cx) {
    const float* row = P.points + p * P.stride;
    int o = P.has_b ? 1 : 0;
    b = P.has_b ? (int)row[0] : 0;
    // fp32 floor((p - min) / vs) with IEEE division: identical to numpy/torch fp32 (dynamic_mean_vfe.py:53)
    float fx = floorf(__fdiv_rn(__fsub_rn(row[o + 0], P.mnx), P.vsx));
    float fy = floorf(__fdiv_rn(__fsub_rn(row[o + 1], P.mny), P.vsy));
    float fz = floorf(__fdiv_rn(__fsub_rn(row[o + 2], P.mnz), P.vsz));
    bool ok = fx >= 0.f && fx < (float)P.gx && fy >= 0.f && fy < (float)P.gy && fz >= 0.f && fz < (float)P.gz &&
              b >= 0 && b < P.grid.B;
    cx = (int)fx; cy = (int)fy; cz = (int)fz;
    return ok;
}

__global__ void k_vox_insert(VoxParams P, uint2* table, uint32_t cap_mask, uint32_t* pt_slot) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_points) return;
    int b, cz, cy, cx;
    uint32_t slot = 0xFFFFFFFFu;
    if (point_cell(P, p, b, cz, cy, cx)) {
        uint32_t key = ql_key(P.grid, b, cz, cy, cx);
        slot = ql_hash_insert(table, cap_mask, key);
        atomicMin(&table[slot].y, (uint32_t)p);
    }
    pt_slot[p] = slot;
}

__device__ __forceinline__ bool is_first(const uint2* table, const uint32_t* pt_slot, int64_t p, int64_t n) {
    if (p >= n) return false;
    uint32_t s = pt_slot[p];
    return s != 0xFFFFFFFFu && table[s].y == (uint32_t)p;
}

__global__ void k_vox_count(VoxParams P, const uint2* table, const uint32_t* pt_slot, int* block_counts) {
    const int64_t n = P.n_points;
    int64_t base = (int64_t)blockIdx.x * kBlockPts + threadIdx.x * kItems;
    int c = 0;
    bool f[kItems];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        f[i] = is_first(table, pt_slot, base + i, n);
        c += f[i] ? 1 : 0;
    }
    if (P.frame_cap > 0) {
        // voxels found per frame: runs of equal frame index are merged per thread, then per CTA in shared memory, so the
        // global counters see a handful of atomics per CTA (a CTA's points belong to one or two frames)
        constexpr int kHist = 64;
        __shared__ int s_hist[kHist];
        const int B = P.grid.B;
        const bool use_smem = B <= kHist;
        if (use_smem) {
            for (int j = threadIdx.x; j < B; j += blockDim.x) s_hist[j] = 0;
            __syncthreads();
        }
        int run_b = -1, run_c = 0;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            if (!f[i]) continue;
            const int b = P.has_b ? (int)P.points[(base + i) * P.stride] : 0;
            if (b != run_b) {
                if (run_c) atomicAdd(use_smem ? &s_hist[run_b] : &P.frames[run_b], run_c);
                run_b = b; run_c = 0;
            }
            ++run_c;
        }
        if (run_c) atomicAdd(use_smem ? &s_hist[run_b] : &P.frames[run_b], run_c);
        if (use_smem) {
            __syncthreads();
            for (int j = threadIdx.x; j < B; j += blockDim.x)
                if (s_hist[j]) atomicAdd(&P.frames[j], s_hist[j]);
        }
    }
    int total;
    block_exclusive_scan(c, total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// per-frame cap: frame b's voxels hold the provisional (global first-touch) ids [F_b, F_b + cnt_b) when the points arrive
// frame by frame; it keeps the first min(cnt_b, cap) of them and every kept id moves down by the voxels dropped before it
__global__ void k_vox_frames(VoxParams P, int* n_voxels_dev) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int B = P.grid.B;
    int* cnt = P.frames;
    int* first = P.frames + B;
    int* shift = P.frames + 2 * B;
    int64_t run = 0, dropped = 0;
    for (int b = 0; b < B; ++b) {
        first[b] = (int)run;
        shift[b] = (int)dropped;
        const int64_t c = cnt[b];
        run += c;
        dropped += c > P.frame_cap ? c - P.frame_cap : 0;
    }
    const int64_t kept = run - dropped;
    n_voxels_dev[0] = (int)(kept < P.max_voxels ? kept : P.max_voxels);
}

__global__ void k_vox_assign(VoxParams P, const uint2* table, const uint32_t* pt_slot, const int* block_offsets,
                             int* pt_vid, uint32_t* vox_first, int32_t* out_coords) {
    int64_t base = (int64_t)blockIdx.x * kBlockPts + threadIdx.x * kItems;
    bool f[kItems];
    int c = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        f[i] = is_first(table, pt_slot, base + i, P.n_points);
        c += f[i] ? 1 : 0;
    }
    int total;
    int ex = block_exclusive_scan(c, total) + block_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        if (!f[i]) continue;
        int64_t p = base + i;
        int vid = ex++;
        if (P.frame_cap > 0) {
            const int b = P.has_b ? (int)P.points[p * P.stride] : 0;
            const int64_t r = (int64_t)vid - P.frames[P.grid.B + b];
            vid = (r >= 0 && r < P.frame_cap) ? vid - P.frames[2 * P.grid.B + b] : 0x7FFFFFFF;   // beyond the frame's cap: dropped
        }
        pt_vid[p] = vid;
        if ((int64_t)vid < P.max_voxels) {
            vox_first[vid] = (uint32_t)p;
            int b, cz, cy, cx;
            point_cell(P, p, b, cz, cy, cx);
            reinterpret_cast<int4*>(out_coords)[vid] = make_int4(b, cz, cy, cx);
        }
    }
}

__global__ void k_vox_gather(VoxParams P, const uint2* table, const uint32_t* pt_slot, const int* pt_vid,
                             uint32_t* tmin, float* sums, int* cnt) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_points) return;
    uint32_t s = pt_slot[p];
    if (s == 0xFFFFFFFFu) return;
    int vid = pt_vid[table[s].y];
    if ((int64_t)vid >= P.max_voxels) return;
    if (P.max_pts > 0) {
        // keep the max_pts smallest point indices: slot t ends up holding the (t+1)-th smallest regardless of order
        uint32_t v = (uint32_t)p;
        uint32_t* a = tmin + (int64_t)vid * P.max_pts;
        for (int t = 0; t < P.max_pts; ++t) {
            uint32_t old = atomicMin(&a[t], v);
            if (old == 0xFFFFFFFFu) break;
            v = old > v ? old : v;
        }
    } else {
        const float* row = P.points + p * P.stride + (P.has_b ? 1 : 0);
        for (int f = 0; f < P.n_feat; ++f) atomicAdd(&sums[(int64_t)vid * P.n_feat + f], row[f]);
        atomicAdd(&cnt[vid], 1);
    }
}

__global__ void k_vox_finalize(VoxParams P, uint2* table, const uint32_t* pt_slot, const uint32_t* vox_first,
                               const uint32_t* tmin, const float* sums, const int* cnt, const int* n_out,
                               float* out_feats, int32_t* out_npts) {
    int vid = blockIdx.x * blockDim.x + threadIdx.x;
    if (vid >= *n_out) return;
    int n = 0;
    if (P.max_pts > 0) {
        float acc[16];
#pragma unroll
        for (int f = 0; f < 16; ++f) acc[f] = 0.f;
        const uint32_t* a = tmin + (int64_t)vid * P.max_pts;
        for (int t = 0; t < P.max_pts; ++t) {
            uint32_t p = a[t];
            if (p == 0xFFFFFFFFu) break;
            const float* row = P.points + (int64_t)p * P.stride + (P.has_b ? 1 : 0);
#pragma unroll
            for (int f = 0; f < 16; ++f)
                if (f < P.n_feat) acc[f] = __fadd_rn(acc[f], row[f]);
            ++n;
        }
        float d = (float)(n > 0 ? n : 1);
#pragma unroll
        for (int f = 0; f < 16; ++f)
            if (f < P.n_feat) out_feats[(int64_t)vid * P.out_stride + f] = __fdiv_rn(acc[f], d);
    } else {
        n = cnt[vid];
        float d = (float)(n > 0 ? n : 1);
        for (int f = 0; f < P.n_feat; ++f)
            out_feats[(int64_t)vid * P.out_stride + f] = __fdiv_rn(sums[(int64_t)vid * P.n_feat + f], d);
    }
    for (int f = P.n_feat; f < P.out_stride; ++f) out_feats[(int64_t)vid * P.out_stride + f] = 0.f;
    out_npts[vid] = n;
    table[pt_slot[vox_first[vid]]].y = (uint32_t)vid;          // coords -> row for the stage-1 rulebook
}

// voxels dropped by the cap keep their key in the table; their value becomes -1 (lookup miss)
__global__ void k_vox_drop(VoxParams P, uint2* table, const uint32_t* pt_slot, const int* pt_vid) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_points) return;
    int vid = pt_vid[p];
    if (vid >= 0 && (int64_t)vid >= P.max_voxels) table[pt_slot[p]].y = 0xFFFFFFFFu;
}

__global__ void k_mean_vfe(const float* voxels, const void* num, int num_is_float, int64_t V, int T, int F, float* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V * F) return;
    int64_t v = i / F;
    int f = (int)(i % F);
    float s = 0.f;
    for (int t = 0; t < T; ++t) s = __fadd_rn(s, voxels[(v * T + t) * F + f]);
    float n = num_is_float ? ((const float*)num)[v] : (float)((const int*)num)[v];
    out[i] = __fdiv_rn(s, fmaxf(n, 1.0f));
}

__global__ void k_hash_build(const int4* coords, int64_t n_cap, const int* n_dev, QlGrid g, uint2* table, uint32_t cap_mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    if (i >= n) return;
    int4 c = coords[i];
    uint32_t s = ql_hash_insert(table, cap_mask, ql_key(g, c.x, c.y, c.z, c.w));
    table[s].y = (uint32_t)i;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct VoxWs {
    size_t pt_slot, pt_vid, blocks, vox_first, tmin, sums, cnt, ntotal, frames, total;
};

VoxWs vox_ws_layout(int64_t max_points, int64_t max_voxels, int n_feat, int max_pts) {
    VoxWs w;
    size_t o = 0;
    int64_t nb = (max_points + kBlockPts - 1) / kBlockPts + 1;
    w.pt_slot = o; o += align256((size_t)max_points * 4);
    w.pt_vid = o; o += align256((size_t)max_points * 4);
    w.blocks = o; o += align256((size_t)nb * 4);
    w.vox_first = o; o += align256((size_t)max_voxels * 4);
    w.tmin = o; o += align256((size_t)max_voxels * (size_t)(max_pts > 0 ? max_pts : 0) * 4);
    w.sums = o; o += align256(max_pts > 0 ? 0 : (size_t)max_voxels * n_feat * 4);
    w.cnt = o; o += align256(max_pts > 0 ? 0 : (size_t)max_voxels * 4);
    w.ntotal = o; o += 256;
    w.frames = o; o += align256((size_t)3 * kMaxFrames * 4);
    w.total = o;
    return w;
}

}  // namespace

extern "C" int64_t ql_hash_capacity(int64_t max_entries) {
    int64_t c = 1024;
    while (c < 2 * max_entries) c <<= 1;
    return c;
}

extern "C" int ql_hash_build(const int32_t* coords, int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D, int32_t H,
                             int32_t W, uint64_t* table, int64_t table_cap, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (!coords || !table || table_cap <= 0 || (table_cap & (table_cap - 1)) || table_cap < 2 * n_cap) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    if (cudaMemsetAsync(table, 0xFF, (size_t)table_cap * 8, st) != cudaSuccess) return QL_ERR_CUDA;
    if (n_cap > 0) {
        QlGrid g{B, D, H, W};
        k_hash_build<<<(unsigned)((n_cap + 255) / 256), 256, 0, st>>>((const int4*)coords, n_cap, n_dev, g, (uint2*)table,
                                                                    (uint32_t)(table_cap - 1));
        QL_CUDA_CHECK_LAST();
    }
    return QL_OK;
}

extern "C" size_t ql_voxelize_workspace_bytes(int64_t max_points, int64_t max_voxels, int32_t n_feat, int32_t max_pts) {
    return vox_ws_layout(max_points, max_voxels, n_feat, max_pts).total;
}

// phases: 1 = passes 1-4 (hash insert, first-touch numbering: out_coords and n_voxels_dev are final), 2 = passes 5-7 (point
// selection, means, num_points, table values), 3 = both.  Phase 2 must follow phase 1 with the same arguments and workspace.
static int voxelize_phases(int phases, const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col,
                           int32_t n_feat, const float* range_min, const float* vsize, const int32_t* grid_xyz,
                           int32_t batch_size, int32_t max_pts, int64_t max_voxels, int64_t max_voxels_per_frame,
                           float* out_feats, int32_t out_feat_stride, int32_t* out_coords, int32_t* out_npts, int32_t* n_voxels_dev, uint64_t* table,
                           int64_t table_cap, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if ((!points && n_points > 0) || !range_min || !vsize || !grid_xyz || !out_feats || !out_coords || !out_npts || !n_voxels_dev || !table ||
        !workspace)
        return QL_ERR_INVALID;
    if (n_feat < 3 || n_feat > 16 || point_stride < n_feat + (has_batch_col ? 1 : 0) || max_pts < 0 || max_voxels <= 0 ||
        batch_size <= 0 || n_points < 0 || n_points >= 2147483647LL || out_feat_stride < n_feat || max_voxels_per_frame < 0 ||
        (max_voxels_per_frame > 0 && batch_size > kMaxFrames))
        return QL_ERR_INVALID;
    if (table_cap <= 0 || (table_cap & (table_cap - 1)) || table_cap < 2 * (n_points < max_voxels ? n_points : n_points))
        return QL_ERR_INVALID;  // the table must hold every distinct cell the points touch (<= n_points)
    VoxWs w = vox_ws_layout(n_points, max_voxels, n_feat, max_pts);
    if (workspace_bytes < w.total) return QL_ERR_WORKSPACE;
    // the key uses the backbone's sparse_shape depth (grid z + 1, spconv_backbone.py:191) so that the table can be
    // handed to the stage-1 rulebook unchanged
    int D = grid_xyz[2] + 1;
    if ((double)batch_size * D * grid_xyz[1] * grid_xyz[0] >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;

    VoxParams P;
    P.points = points; P.n_points = n_points; P.stride = point_stride; P.has_b = has_batch_col; P.n_feat = n_feat;
    P.out_stride = out_feat_stride;
    P.mnx = range_min[0]; P.mny = range_min[1]; P.mnz = range_min[2];
    P.vsx = vsize[0]; P.vsy = vsize[1]; P.vsz = vsize[2];
    P.gx = grid_xyz[0]; P.gy = grid_xyz[1]; P.gz = grid_xyz[2];
    P.grid = QlGrid{batch_size, D, grid_xyz[1], grid_xyz[0]};
    P.max_pts = max_pts; P.max_voxels = max_voxels;
    P.frame_cap = max_voxels_per_frame;
    P.frames = (int*)((char*)workspace + w.frames);

    char* ws = (char*)workspace;
    uint32_t* pt_slot = (uint32_t*)(ws + w.pt_slot);
    int* pt_vid = (int*)(ws + w.pt_vid);
    int* blocks = (int*)(ws + w.blocks);
    uint32_t* vox_first = (uint32_t*)(ws + w.vox_first);
    uint32_t* tmin = (uint32_t*)(ws + w.tmin);
    float* sums = (float*)(ws + w.sums);
    int* cnt = (int*)(ws + w.cnt);

    if (phases & 1) {
        if (cudaMemsetAsync(table, 0xFF, (size_t)table_cap * 8, st) != cudaSuccess) return QL_ERR_CUDA;
        if (cudaMemsetAsync(pt_vid, 0xFF, (size_t)(n_points > 0 ? n_points : 1) * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    }
    if (phases & 2) {
        if (max_pts > 0) {
            if (cudaMemsetAsync(tmin, 0xFF, (size_t)max_voxels * max_pts * 4, st) != cudaSuccess) return QL_ERR_CUDA;
        } else {
            if (cudaMemsetAsync(sums, 0, (size_t)max_voxels * n_feat * 4, st) != cudaSuccess) return QL_ERR_CUDA;
            if (cudaMemsetAsync(cnt, 0, (size_t)max_voxels * 4, st) != cudaSuccess) return QL_ERR_CUDA;
        }
    }
    if (n_points == 0) {
        if ((phases & 1) && cudaMemsetAsync(n_voxels_dev, 0, 8, st) != cudaSuccess) return QL_ERR_CUDA;
        return QL_OK;
    }
    uint32_t cap_mask = (uint32_t)(table_cap - 1);
    unsigned gp = (unsigned)((n_points + kThreads - 1) / kThreads);
    unsigned nb = (unsigned)((n_points + kBlockPts - 1) / kBlockPts);
    if (phases & 1) {
        k_vox_insert<<<gp, kThreads, 0, st>>>(P, (uint2*)table, cap_mask, pt_slot);
        if (max_voxels_per_frame > 0 && cudaMemsetAsync(P.frames, 0, (size_t)3 * batch_size * 4, st) != cudaSuccess) return QL_ERR_CUDA;
        k_vox_count<<<nb, kThreads, 0, st>>>(P, (const uint2*)table, pt_slot, blocks);
        k_scan_blocks<<<1, kThreads, 0, st>>>(blocks, (int)nb, n_voxels_dev + 1, n_voxels_dev, max_voxels);
        if (max_voxels_per_frame > 0) k_vox_frames<<<1, 32, 0, st>>>(P, n_voxels_dev);
        k_vox_assign<<<nb, kThreads, 0, st>>>(P, (const uint2*)table, pt_slot, blocks, pt_vid, vox_first, out_coords);
    }
    if (phases & 2) {
        k_vox_gather<<<gp, kThreads, 0, st>>>(P, (const uint2*)table, pt_slot, pt_vid, tmin, sums, cnt);
        k_vox_drop<<<gp, kThreads, 0, st>>>(P, (uint2*)table, pt_slot, pt_vid);
        k_vox_finalize<<<(unsigned)((max_voxels + kThreads - 1) / kThreads), kThreads, 0, st>>>(
            P, (uint2*)table, pt_slot, vox_first, tmin, sums, cnt, n_voxels_dev, out_feats, out_npts);
    }
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

#define QL_VOX_ARGS points, n_points, point_stride, has_batch_col, n_feat, range_min, vsize, grid_xyz, batch_size, max_pts, max_voxels, \
                    max_voxels_per_frame, out_feats, out_feat_stride, out_coords, out_npts, n_voxels_dev, table, table_cap, workspace,     \
                    workspace_bytes, stream_
#define QL_VOX_PARAMS                                                                                                                     \
    const float *points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat, const float *range_min,           \
        const float *vsize, const int32_t *grid_xyz, int32_t batch_size, int32_t max_pts, int64_t max_voxels,                            \
        int64_t max_voxels_per_frame, float *out_feats, int32_t out_feat_stride, int32_t *out_coords, int32_t *out_npts,                 \
        int32_t *n_voxels_dev, uint64_t *table, int64_t table_cap, void *workspace, size_t workspace_bytes, ql_stream_t stream_
extern "C" int ql_voxelize_mean(QL_VOX_PARAMS) { return voxelize_phases(3, QL_VOX_ARGS); }
// the two halves of ql_voxelize_mean, for callers that overlap work needing only the coordinates (rulebooks) with the
// feature passes: ql_voxelize_coords leaves out_coords / n_voxels_dev final; ql_voxelize_features (same arguments, same
// workspace, stream-ordered after it) fills out_feats / out_npts and the table values
extern "C" int ql_voxelize_coords(QL_VOX_PARAMS) { return voxelize_phases(1, QL_VOX_ARGS); }
extern "C" int ql_voxelize_features(QL_VOX_PARAMS) { return voxelize_phases(2, QL_VOX_ARGS); }
#undef QL_VOX_ARGS
#undef QL_VOX_PARAMS

extern "C" int ql_mean_vfe(const float* voxels, const void* num_points, int32_t is_float, int64_t V, int32_t T, int32_t F,
                           float* out, ql_stream_t stream_) {
    if (!voxels || !num_points || !out || T <= 0 || F <= 0 || V < 0) return QL_ERR_INVALID;
    if (V == 0) return QL_OK;
    k_mean_vfe<<<(unsigned)((V * F + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(voxels, num_points, is_float, V, T, F, out);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
