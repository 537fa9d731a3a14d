// CenterHead post-processing on the device: heat-map top-K, box decode, range / score mask, rotated BEV NMS with the
// suppression sweep on the GPU -- no device->host copy anywhere on the path.
//
// Replaces:
//   CenterHead.generate_predicted_boxes              pcdet/models/dense_heads/center_head.py:297-365
//   centernet_utils._topk / decode_bbox_from_heatmap pcdet/models/model_utils/centernet_utils.py:155-241
//   model_nms_utils.class_agnostic_nms               pcdet/models/model_utils/model_nms_utils.py:6-25
//   iou3d_nms_utils.nms_gpu                          pcdet/ops/iou3d_nms/iou3d_nms_utils.py:120-135
//   nms_kernel + the host sweep of iou3d_nms.cpp     pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:295-339, iou3d_nms.cpp:137-183
//     (the reference cudaMallocs the mask, copies it to the host synchronously and sweeps it in a serial CPU loop per call)
//
// Kernels (4 launches per head, all stream ordered):
//   k_ch_candidates     every heat-map cell: score = sigmoid(logit); cells above the score threshold are appended to the
//                       frame's candidate list (the reference thresholds AFTER its top-K; a cell below the threshold can never
//                       be output, so dropping it first changes nothing but the amount of sorting)
//   k_ch_select_decode  one CTA per frame: exact top-K of the candidates (radix select on the score bits when the list is
//                       longer than the sort width, then a bitonic sort by (score desc, cell index asc)), gather of the
//                       regression maps, decode, range mask, order-preserving compaction
//   k_nms_mask          one thread per pair of the upper triangle of the pairwise rotated-IoU matrix (far-apart pairs rejected by a
//                       circle test), a warp ballot = half a 64-bit suppression word
//   k_nms_sweep         one CTA per frame: mask rows staged in shared memory, one warp runs the greedy sweep 64 boxes at a time
//                       (fixed-point iteration over the block's diagonal bit matrix), the CTA gathers the kept boxes / scores / labels
#include "ql_common.cuh"

namespace {

constexpr int kSelThreads = 1024;            // == the bitonic sort width; MAX_OBJ_PER_SAMPLE (K) must not exceed it
constexpr int kNmsBlock = 64;                // boxes per mask word
constexpr int kMaxBoxDim = 9;                // x y z dx dy dz heading (+ vx vy)

struct HeadMaps {
    const float* hm;        // [B, C, H, W] logits
    const float* center;    // [B, 2, H, W]
    const float* center_z;  // [B, 1, H, W]
    const float* dim;       // [B, 3, H, W] log-sizes
    const float* rot;       // [B, 2, H, W] (cos, sin)
    const float* vel;       // [B, 2, H, W] or null
    const float* iou;       // [B, 1, H, W] or null
    int B, C, H, W, K;
    float stride, vsx, vsy, pcx, pcy;
    float lim[6];
    float score_thresh;     // < 0: no score threshold
    const int* class_map;   // [C] or null: label = class_map[class]
    // SPARSE head (VoxelNeXtHead): the "maps" are row-major per-voxel arrays hm [N, C], center [N, 2], ... and a voxel's cell is
    // sp_indices[row] = (b, y, x); a candidate's index is cls * sp_n + row.  Null = the dense NCHW head above.
    const int* sp_indices;
    int sp_n;
    int64_t list_stride;    // candidate-list capacity per frame (C*H*W dense, C*N sparse)
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }   // torch: 1 / (1 + exp(-x))

__global__ void __launch_bounds__(256) k_ch_candidates(HeadMaps M, uint2* __restrict__ cand, int* __restrict__ cand_count) {
    const int64_t n = (int64_t)M.C * M.H * M.W;
    const int b = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    float s = 0.f;
    if (i < n) {
        s = sigmoidf_ref(M.hm[(int64_t)b * n + i]);
        pass = M.score_thresh < 0.f || s > M.score_thresh;
    }
    const uint32_t vote = __ballot_sync(0xffffffffu, pass);
    if (vote == 0u) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&cand_count[b], __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pass) cand[(int64_t)b * n + base + __popc(vote & ((1u << lane) - 1u))] = make_uint2(__float_as_uint(s), (uint32_t)i);
}

// sparse head: one thread per (voxel row, class); the frame comes from the row's batch index
__global__ void __launch_bounds__(256) k_vh_candidates(HeadMaps M, const int* __restrict__ n_dev, uint2* __restrict__ cand, int* __restrict__ cand_count) {
    const int n_rows = n_dev ? min(*n_dev, M.sp_n) : M.sp_n;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    float s = 0.f;
    int b = 0, row = 0, cls = 0;
    if (i < (int64_t)n_rows * M.C) {
        row = (int)(i / M.C); cls = (int)(i - (int64_t)row * M.C);
        b = M.sp_indices[3 * row];
        s = sigmoidf_ref(M.hm[i]);
        pass = b >= 0 && b < M.B && (M.score_thresh < 0.f || s > M.score_thresh);
    }
    if (!pass) return;
    // lanes of the same frame share one atomic (rows are ordered by frame almost everywhere)
    const uint32_t peers = __match_any_sync(__activemask(), b);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cand_count[b], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    cand[(int64_t)b * M.list_stride + base + __popc(peers & ((1u << lane) - 1u))] = make_uint2(__float_as_uint(s), (uint32_t)(cls * M.sp_n + row));
}

// descending by score bits (scores are positive floats: their bit patterns order like the values), ties by ascending cell index
__device__ __forceinline__ uint64_t sort_key(uint32_t score_bits, uint32_t idx) { return ((uint64_t)score_bits << 32) | (uint64_t)(~idx); }

__global__ void __launch_bounds__(kSelThreads) k_ch_select_decode(HeadMaps M, const uint2* __restrict__ cand, const int* __restrict__ cand_count,
                                                                   int box_dim, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                                   int* __restrict__ out_labels, float* __restrict__ out_iou, int* __restrict__ out_count) {
    __shared__ uint64_t keys[kSelThreads];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_mask, s_remaining, s_fill;
    __shared__ int warp_sums[kSelThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint2* list = cand + (int64_t)b * M.list_stride;
    const int n = cand_count[b];
    keys[tid] = 0ull;
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_remaining = (uint32_t)M.K; s_fill = 0u; }
    __syncthreads();
    if (n <= kSelThreads) {
        if (tid < n) { const uint2 e = list[tid]; keys[tid] = sort_key(e.x, e.y); }
    } else {
        // radix select, 8 bits per pass from the top: after 4 passes s_prefix is the K-th largest score exactly and s_remaining the
        // number of candidates EQUAL to it that belong to the top K
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int j = tid; j < 256; j += kSelThreads) hist[j] = 0u;
            __syncthreads();
            const uint32_t prefix = s_prefix, mask = s_mask;
            for (int j = tid; j < n; j += kSelThreads) {
                const uint32_t k = list[j].x;
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t rem = s_remaining, d = 255u;
                for (;; --d) {
                    const uint32_t h = hist[d];
                    if (h >= rem || d == 0u) break;
                    rem -= h;
                }
                s_remaining = rem;
                s_prefix = prefix | (d << shift);
                s_mask = mask | (255u << shift);
            }
            __syncthreads();
        }
        const uint32_t kth = s_prefix;
        // everything above the K-th score, then the ties (as many as the sort width holds; the sort orders them by cell index)
        for (int j = tid; j < n; j += kSelThreads) {
            const uint2 e = list[j];
            if (e.x > kth) keys[atomicAdd(&s_fill, 1u)] = sort_key(e.x, e.y);
        }
        __syncthreads();
        for (int j = tid; j < n; j += kSelThreads) {
            const uint2 e = list[j];
            if (e.x == kth) {
                const uint32_t pos = atomicAdd(&s_fill, 1u);
                if (pos < (uint32_t)kSelThreads) keys[pos] = sort_key(e.x, e.y);
            }
        }
    }
    __syncthreads();
    // bitonic sort, descending
    for (int size = 2; size <= kSelThreads; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int partner = tid ^ stride;
            if (partner > tid) {
                const uint64_t a = keys[tid], c = keys[partner];
                const bool desc = (tid & size) == 0;
                if (desc ? a < c : a > c) { keys[tid] = c; keys[partner] = a; }
            }
            __syncthreads();
        }
    }
    // decode candidate `tid` of the frame's top K
    const int n_top = n < M.K ? n : M.K;
    bool keep = false;
    float box[kMaxBoxDim];
    float score = 0.f, iou_v = 0.f;
    int label = 0;
    if (tid < n_top) {
        const uint64_t k = keys[tid];
        const uint32_t idx = ~(uint32_t)k;
        score = __uint_as_float((uint32_t)(k >> 32));
        // dense: plane p of a [B, P, H, W] map at (f * P + p * hw) + cell; sparse: column p of a [N, P] array at cell * P + p
        const bool sp = M.sp_indices != nullptr;
        const int hw = sp ? 1 : M.H * M.W;
        const uint32_t per_cls = sp ? (uint32_t)M.sp_n : (uint32_t)hw;
        const int cls = (int)(idx / per_cls);
        const int64_t cell = (int64_t)(idx % per_cls);
        const int y = sp ? M.sp_indices[3 * cell + 1] : (int)(cell / M.W), x = sp ? M.sp_indices[3 * cell + 2] : (int)(cell % M.W);
        const int64_t f = sp ? 0 : (int64_t)b * hw;                         // frame offset in units of one H*W plane
        #define QL_HEAD_AT(ptr, P, p) (sp ? (ptr)[cell * (P) + (p)] : (ptr)[(f * (P) + (int64_t)(p) * hw) + cell])
        const float cx = QL_HEAD_AT(M.center, 2, 0), cy = QL_HEAD_AT(M.center, 2, 1);
        // xs = (x + center_x) * stride * voxel_x + pc_min_x, one fp32 rounding per torch op (centernet_utils.py:188-193)
        box[0] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn((float)x, cx), M.stride), M.vsx), M.pcx);
        box[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn((float)y, cy), M.stride), M.vsy), M.pcy);
        box[2] = QL_HEAD_AT(M.center_z, 1, 0);
#pragma unroll
        for (int j = 0; j < 3; ++j) box[3 + j] = expf(QL_HEAD_AT(M.dim, 3, j));
        const float rc = QL_HEAD_AT(M.rot, 2, 0), rs = QL_HEAD_AT(M.rot, 2, 1);
        box[6] = atan2f(rs, rc);
        box[7] = box[8] = 0.f;
        if (M.vel) { box[7] = QL_HEAD_AT(M.vel, 2, 0); box[8] = QL_HEAD_AT(M.vel, 2, 1); }
        if (M.iou) {
            iou_v = __fmul_rn(__fadd_rn(QL_HEAD_AT(M.iou, 1, 0), 1.0f), 0.5f);
            if (sp) iou_v = fminf(fmaxf(iou_v, 0.f), 1.f);                  // decode_bbox_from_voxels_nuscenes clamps (centernet_utils.py:324)
        }
        #undef QL_HEAD_AT
        label = M.class_map ? M.class_map[cls] : cls;
        keep = box[0] >= M.lim[0] && box[1] >= M.lim[1] && box[2] >= M.lim[2] && box[0] <= M.lim[3] && box[1] <= M.lim[4] && box[2] <= M.lim[5];
        if (M.score_thresh >= 0.f) keep = keep && score > M.score_thresh;
    }
    // order-preserving compaction over the CTA
    const uint32_t vote = __ballot_sync(0xffffffffu, keep);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) warp_sums[wid] = __popc(vote);
    __syncthreads();
    if (wid == 0) {
        int v = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        warp_sums[lane] = v;                                                // inclusive
    }
    __syncthreads();
    if (keep) {
        const int pos = (wid ? warp_sums[wid - 1] : 0) + __popc(vote & ((1u << lane) - 1u));
        float* o = out_boxes + ((int64_t)b * M.K + pos) * box_dim;
        for (int j = 0; j < box_dim; ++j) o[j] = box[j];
        out_scores[(int64_t)b * M.K + pos] = score;
        out_labels[(int64_t)b * M.K + pos] = label;
        if (out_iou) out_iou[(int64_t)b * M.K + pos] = iou_v;
    }
    if (tid == 0) out_count[b] = warp_sums[kSelThreads / 32 - 1];
}

// ---------------------------------------------------------------------------------------------- rotated BEV IoU
// Two rectangles [x, y, z, dx, dy, dz, heading]: the intersection polygon's vertices are the edge-edge crossings plus the corners
// of one rectangle inside the other; they are ordered by angle around their centroid and the area is a triangle fan.  The
// arithmetic follows iou3d_nms_kernel.cu:35-235 operation for operation (same tolerances: EPS 1e-8, containment margin 1e-2)
// so that a pair's "IoU > threshold" bit agrees with the reference's.
struct P2 {
    float x, y;
};
constexpr float kEps = 1e-8f;

__device__ __forceinline__ float cross3(const P2& p1, const P2& p2, const P2& p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

__device__ __forceinline__ bool seg_cross(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2& out) {
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) && fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) &&
          fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return false;
    const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0.f && s3 * s4 > 0.f)) return false;
    const float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > kEps) {
        out.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        out.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        out.x = (b0 * c1 - b1 * c0) / D;
        out.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

__device__ __forceinline__ bool inside_rect(const float* box, const P2& p) {
    const float margin = 1e-2f;
    const float c = cosf(-box[6]), s = sinf(-box[6]);
    const float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    const float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + margin && fabsf(ry) < box[4] / 2 + margin;
}

__device__ __forceinline__ void rect_corners(const float* box, P2 (&c)[5]) {
    const float hx = box[3] / 2, hy = box[4] / 2;
    const float x1 = box[0] - hx, y1 = box[1] - hy, x2 = box[0] + hx, y2 = box[1] + hy;
    const float ca = cosf(box[6]), sa = sinf(box[6]);
    const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k].x = (px[k] - box[0]) * ca + (py[k] - box[1]) * (-sa) + box[0];
        c[k].y = (px[k] - box[0]) * sa + (py[k] - box[1]) * ca + box[1];
    }
    c[4] = c[0];
}

__device__ float rect_overlap(const float* a, const float* b) {
    P2 ca[5], cb[5];
    rect_corners(a, ca);
    rect_corners(b, cb);
    P2 pts[16];
    float ang[16];
    P2 centre{0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_cross(ca[i + 1], ca[i], cb[j + 1], cb[j], pts[cnt])) {
                centre.x = centre.x + pts[cnt].x;
                centre.y = centre.y + pts[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (inside_rect(a, cb[k])) {
            centre.x = centre.x + cb[k].x;
            centre.y = centre.y + cb[k].y;
            pts[cnt++] = cb[k];
        }
        if (inside_rect(b, ca[k])) {
            centre.x = centre.x + ca[k].x;
            centre.y = centre.y + ca[k].y;
            pts[cnt++] = ca[k];
        }
    }
    if (cnt == 0) return 0.f;
    centre.x /= cnt;
    centre.y /= cnt;
    // ascending angle around the centroid; a stable insertion sort gives the order of the reference's bubble sort
    for (int k = 0; k < cnt; ++k) ang[k] = atan2f(pts[k].y - centre.y, pts[k].x - centre.x);
    for (int k = 1; k < cnt; ++k) {
        const P2 p = pts[k];
        const float t = ang[k];
        int j = k - 1;
        while (j >= 0 && ang[j] > t) {
            pts[j + 1] = pts[j];
            ang[j + 1] = ang[j];
            --j;
        }
        pts[j + 1] = p;
        ang[j + 1] = t;
    }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ux = pts[k].x - pts[0].x, uy = pts[k].y - pts[0].y;
        const float vx = pts[k + 1].x - pts[0].x, vy = pts[k + 1].y - pts[0].y;
        area += ux * vy - uy * vx;
    }
    return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float rect_iou(const float* a, const float* b) {
    const float sa = a[3] * a[4], sb = b[3] * b[4];
    const float ov = rect_overlap(a, b);
    return ov / fmaxf(sa + sb - ov, kEps);
}

// One thread per (row, column) pair of the upper triangle: grid (col block, row group, frame), block = 64 columns x kMaskRows rows.
// A warp covers 32 consecutive columns of one row, so its ballot is one half of the row's 64-bit mask word.  Pairs whose
// circumscribed circles (+ the containment margin) do not meet have no crossing and no contained corner: IoU = 0 without the
// polygon code -- the common case by far, and what keeps the divergent slow path rare.
constexpr int kMaskRows = 4;
__global__ void __launch_bounds__(kNmsBlock * kMaskRows) k_nms_mask(const float* __restrict__ boxes, int box_stride, const int* __restrict__ counts,
                                                                    int n_cap, int pre_max, float thresh, uint32_t* __restrict__ mask32, int col_blocks,
                                                                    float* __restrict__ iou_out) {
    const int b = blockIdx.z, cbk = blockIdx.x;
    int n = counts ? counts[b] : n_cap;
    n = n < n_cap ? n : n_cap;
    n = n < pre_max ? n : pre_max;
    const int c = threadIdx.x & (kNmsBlock - 1), r = threadIdx.x / kNmsBlock;
    const int row = blockIdx.y * kMaskRows + r, col = cbk * kNmsBlock + c;
    const int row0 = blockIdx.y * kMaskRows;
    if (row0 >= n || cbk * kNmsBlock >= n || cbk < row0 / kNmsBlock) return;       // CTA-uniform: nothing of the upper triangle here
    __shared__ float scol[kNmsBlock * 7];
    __shared__ float srow[kMaskRows * 7];
    const float* fb = boxes + (int64_t)b * n_cap * box_stride;
    if (r == 0 && col < n) {
#pragma unroll
        for (int j = 0; j < 7; ++j) scol[c * 7 + j] = fb[(int64_t)col * box_stride + j];
    }
    if (threadIdx.x < kMaskRows * 7) {
        const int rr = threadIdx.x / 7, j = threadIdx.x % 7;
        if (row0 + rr < n) srow[threadIdx.x] = fb[(int64_t)(row0 + rr) * box_stride + j];
    }
    __syncthreads();
    bool hit = false;
    if (row < n && col < n && col > row) {
        const float* A = srow + r * 7;
        const float* Bx = scol + c * 7;
        const float dx = A[0] - Bx[0], dy = A[1] - Bx[1];
        const float ra = 0.5f * sqrtf(A[3] * A[3] + A[4] * A[4]), rb = 0.5f * sqrtf(Bx[3] * Bx[3] + Bx[4] * Bx[4]);
        const float reach = ra + rb + 0.1f;
        float v = 0.f;
        if (dx * dx + dy * dy <= reach * reach) v = rect_iou(A, Bx);
        if (iou_out) iou_out[((int64_t)b * n_cap + row) * n_cap + col] = v;
        hit = v > thresh;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && row < n) mask32[(((int64_t)b * n_cap + row) * col_blocks + cbk) * 2 + ((c >> 5) & 1)] = bits;
}

// one CTA per frame; dynamic shared memory holds the frame's mask rows [n][col_blocks].  The greedy sweep is sequential only inside a
// 64-box block (through the block's diagonal mask words); one warp resolves a block by fixed-point iteration, then ORs the kept rows
// into the removed-words of the later blocks, lane c owning word c.
__global__ void __launch_bounds__(256) k_nms_sweep(const float* __restrict__ boxes, int box_stride, int box_dim, const float* __restrict__ scores,
                                                   const int* __restrict__ labels, const int* __restrict__ counts, int n_cap, int pre_max,
                                                   int post_max, const unsigned long long* __restrict__ mask, int col_blocks, int label_offset,
                                                   int* __restrict__ keep, int* __restrict__ keep_count, float* __restrict__ out_boxes,
                                                   float* __restrict__ out_scores, int* __restrict__ out_labels) {
    extern __shared__ unsigned long long smask[];
    __shared__ int s_nk;
    int* skeep = reinterpret_cast<int*>(smask + (size_t)n_cap * col_blocks);
    const int b = blockIdx.x;
    int n = counts ? counts[b] : n_cap;
    n = n < n_cap ? n : n_cap;
    n = n < pre_max ? n : pre_max;
    const int cb_n = (n + kNmsBlock - 1) / kNmsBlock;
    const unsigned long long* gm = mask + (int64_t)b * n_cap * col_blocks;
    // only the upper triangle was written: word c of row r exists for c >= r / 64
    for (int i = threadIdx.x; i < n * cb_n; i += blockDim.x) {
        const int r = i / cb_n, c = i % cb_n;
        smask[r * col_blocks + c] = c >= (r >> 6) ? gm[(int64_t)r * col_blocks + c] : 0ull;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned long long remv = 0ull;                                     // lane c: removed bits of block c
        int nk = 0;
        for (int w = 0; w < cb_n && nk < post_max; ++w) {
            const int base = w * kNmsBlock;
            const int in_block = min(n - base, kNmsBlock);
            unsigned long long dead = __shfl_sync(0xffffffffu, remv, w);
            if (in_block < 64) dead |= ~0ull << in_block;
            // Inside the block box j survives iff it is alive and no SURVIVING earlier box of the block suppresses it.  Lane L owns
            // boxes L and L + 32 and the columns of the diagonal 64 x 64 bit matrix that belong to them; the survivor set is the fixed
            // point of kept[j] = alive[j] & !(col[j] & kept), reached level by level of the suppression chains (a handful of
            // ballots instead of 64 dependent steps).
            const int j0 = lane, j1 = lane + 32;
            unsigned long long col0 = 0ull, col1 = 0ull;
#pragma unroll 8
            for (int i = 0; i < kNmsBlock; ++i) {
                if (i < in_block) {
                    const unsigned long long word = smask[(base + i) * col_blocks + w];
                    col0 |= ((word >> j0) & 1ull) << i;
                    col1 |= ((word >> j1) & 1ull) << i;
                }
            }
            const bool alive0 = !((dead >> j0) & 1ull), alive1 = !((dead >> j1) & 1ull);
            unsigned long long kept = (unsigned long long)__ballot_sync(0xffffffffu, alive0) | ((unsigned long long)__ballot_sync(0xffffffffu, alive1) << 32);
            for (int it = 0; it < kNmsBlock; ++it) {
                const bool k0 = alive0 && !(col0 & kept), k1 = alive1 && !(col1 & kept);
                const unsigned long long nxt = (unsigned long long)__ballot_sync(0xffffffffu, k0) | ((unsigned long long)__ballot_sync(0xffffffffu, k1) << 32);
                if (nxt == kept) break;
                kept = nxt;
            }
            const int p0 = nk + __popcll(kept & ((1ull << j0) - 1ull)), p1 = nk + __popcll(kept & ((1ull << j1) - 1ull));
            if (((kept >> j0) & 1ull) && p0 < post_max) skeep[p0] = base + j0;
            if (((kept >> j1) & 1ull) && p1 < post_max) skeep[p1] = base + j1;
            nk = min(nk + __popcll(kept), post_max);
            // the kept rows of this block suppress boxes of the later blocks
            if (lane > w && lane < cb_n) {
                unsigned long long acc = 0ull;
#pragma unroll 8
                for (int j = 0; j < kNmsBlock; ++j)
                    if ((kept >> j) & 1ull) acc |= smask[(base + j) * col_blocks + lane];
                remv |= acc;
            }
        }
        __syncwarp();
        if (lane == 0) { s_nk = nk; keep_count[b] = nk; }
    }
    __syncthreads();
    const int nk = s_nk;
    for (int k = threadIdx.x; k < post_max; k += blockDim.x) keep[(int64_t)b * post_max + k] = k < nk ? skeep[k] : -1;
    if (out_boxes) {
        for (int e = threadIdx.x; e < nk * box_dim; e += blockDim.x) {
            const int k = e / box_dim, j = e % box_dim;
            out_boxes[((int64_t)b * post_max + k) * box_dim + j] = boxes[((int64_t)b * n_cap + skeep[k]) * box_stride + j];
        }
    }
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
        if (out_scores && scores) out_scores[(int64_t)b * post_max + k] = scores[(int64_t)b * n_cap + skeep[k]];
        if (out_labels && labels) out_labels[(int64_t)b * post_max + k] = labels[(int64_t)b * n_cap + skeep[k]] + label_offset;
    }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t ql_centerhead_decode_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
    return align256((size_t)B * C * H * W * sizeof(uint2)) + align256((size_t)B * sizeof(int));
}

extern "C" int ql_centerhead_decode(const float* hm, const float* center, const float* center_z, const float* dim, const float* rot, const float* vel,
                                    const float* iou, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K, float feature_map_stride,
                                    const float* voxel_size_xy, const float* pc_min_xy, const float* center_limit_range, float score_thresh,
                                    const int32_t* class_map, float* out_boxes, float* out_scores, int32_t* out_labels, float* out_iou,
                                    int32_t* out_count, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    if (!hm || !center || !center_z || !dim || !rot || !voxel_size_xy || !pc_min_xy || !center_limit_range || !out_boxes || !out_scores ||
        !out_labels || !out_count || !workspace)
        return QL_ERR_INVALID;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0 || K > kSelThreads || (iou && !out_iou)) return QL_ERR_INVALID;
    if ((double)C * H * W >= 2147483647.0) return QL_ERR_GRID_TOO_LARGE;
    if (workspace_bytes < ql_centerhead_decode_workspace_bytes(B, C, H, W)) return QL_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream_;
    HeadMaps M;
    M.hm = hm; M.center = center; M.center_z = center_z; M.dim = dim; M.rot = rot; M.vel = vel; M.iou = iou;
    M.B = B; M.C = C; M.H = H; M.W = W; M.K = K;
    M.stride = feature_map_stride; M.vsx = voxel_size_xy[0]; M.vsy = voxel_size_xy[1]; M.pcx = pc_min_xy[0]; M.pcy = pc_min_xy[1];
    for (int i = 0; i < 6; ++i) M.lim[i] = center_limit_range[i];
    M.score_thresh = score_thresh;
    M.class_map = class_map;
    M.sp_indices = nullptr; M.sp_n = 0; M.list_stride = (int64_t)C * H * W;
    uint2* cand = (uint2*)workspace;
    int* cand_count = (int*)((char*)workspace + align256((size_t)B * C * H * W * sizeof(uint2)));
    if (cudaMemsetAsync(cand_count, 0, (size_t)B * sizeof(int), st) != cudaSuccess) return QL_ERR_CUDA;
    const int64_t n = (int64_t)C * H * W;
    k_ch_candidates<<<dim3((unsigned)((n + 255) / 256), (unsigned)B), 256, 0, st>>>(M, cand, cand_count);
    k_ch_select_decode<<<(unsigned)B, kSelThreads, 0, st>>>(M, cand, cand_count, vel ? 9 : 7, out_boxes, out_scores, out_labels, out_iou, out_count);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

// ---------------------------------------------------------------------------------------------- VoxelNeXt sparse head
extern "C" size_t ql_voxelhead_decode_workspace_bytes(int32_t B, int32_t C, int64_t n_cap) {
    return align256((size_t)B * C * (size_t)n_cap * sizeof(uint2)) + align256((size_t)B * sizeof(int));
}

extern "C" int ql_voxelhead_decode(const float* hm, const float* center, const float* center_z, const float* dim, const float* rot, const float* vel,
                                   const float* iou, const int32_t* indices_byx, int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t C,
                                   int32_t K, float feature_map_stride, const float* voxel_size_xy, const float* pc_min_xy,
                                   const float* center_limit_range, float score_thresh, const int32_t* class_map, float* out_boxes,
                                   float* out_scores, int32_t* out_labels, float* out_iou, int32_t* out_count, void* workspace,
                                   size_t workspace_bytes, ql_stream_t stream_) {
    if (!hm || !center || !center_z || !dim || !rot || !indices_byx || !voxel_size_xy || !pc_min_xy || !center_limit_range || !out_boxes ||
        !out_scores || !out_labels || !out_count || !workspace)
        return QL_ERR_INVALID;
    if (B <= 0 || C <= 0 || n_cap < 0 || K <= 0 || K > kSelThreads || (iou && !out_iou)) return QL_ERR_INVALID;
    if ((double)C * (double)n_cap >= 2147483647.0) return QL_ERR_GRID_TOO_LARGE;
    if (workspace_bytes < ql_voxelhead_decode_workspace_bytes(B, C, n_cap)) return QL_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream_;
    HeadMaps M;
    M.hm = hm; M.center = center; M.center_z = center_z; M.dim = dim; M.rot = rot; M.vel = vel; M.iou = iou;
    M.B = B; M.C = C; M.H = 1; M.W = 1; M.K = K;
    M.stride = feature_map_stride; M.vsx = voxel_size_xy[0]; M.vsy = voxel_size_xy[1]; M.pcx = pc_min_xy[0]; M.pcy = pc_min_xy[1];
    for (int i = 0; i < 6; ++i) M.lim[i] = center_limit_range[i];
    M.score_thresh = score_thresh;
    M.class_map = class_map;
    M.sp_indices = indices_byx; M.sp_n = (int)n_cap; M.list_stride = (int64_t)C * n_cap;
    uint2* cand = (uint2*)workspace;
    int* cand_count = (int*)((char*)workspace + align256((size_t)B * C * (size_t)n_cap * sizeof(uint2)));
    if (cudaMemsetAsync(cand_count, 0, (size_t)B * sizeof(int), st) != cudaSuccess) return QL_ERR_CUDA;
    const int64_t n = (int64_t)C * n_cap;
    if (n > 0) k_vh_candidates<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(M, n_dev, cand, cand_count);
    k_ch_select_decode<<<(unsigned)B, kSelThreads, 0, st>>>(M, cand, cand_count, vel ? 9 : 7, out_boxes, out_scores, out_labels, out_iou, out_count);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

namespace {
// One CTA per (class, frame): the frame's decoded boxes of that class, re-scored score^(1-r) * iou^r and sorted by the new score
// (descending, ties by position) -- the per-class input of rotate_class_specific_nms_iou (voxelnext_head.py:308-331).
__global__ void __launch_bounds__(kSelThreads) k_vh_class_split(const float* __restrict__ boxes, int box_dim, const float* __restrict__ scores,
                                                                 const int* __restrict__ labels, const float* __restrict__ ious,
                                                                 const int* __restrict__ counts, int K, int B, const float* __restrict__ rectifier,
                                                                 float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                                 int* __restrict__ out_labels, int* __restrict__ out_counts) {
    __shared__ uint64_t keys[kSelThreads];
    __shared__ int s_n;
    const int b = blockIdx.x, cls = blockIdx.y, tid = threadIdx.x;
    const int n = min(counts[b], K);
    if (tid == 0) s_n = 0;
    __syncthreads();
    uint64_t key = 0ull;
    if (tid < n && labels[(int64_t)b * K + tid] == cls) {
        const float r = rectifier[cls];
        const float sc = __fmul_rn(powf(scores[(int64_t)b * K + tid], __fsub_rn(1.0f, r)), powf(ious[(int64_t)b * K + tid], r));
        key = ((uint64_t)__float_as_uint(fmaxf(sc, 0.f)) << 32) | (uint64_t)(~(uint32_t)tid) | (1ull << 63);   // bit 63: a real entry
        atomicAdd(&s_n, 1);
    }
    keys[tid] = key;
    __syncthreads();
    for (int size = 2; size <= kSelThreads; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int partner = tid ^ stride;
            if (partner > tid) {
                const uint64_t a = keys[tid], c = keys[partner];
                const bool desc = (tid & size) == 0;
                if (desc ? a < c : a > c) { keys[tid] = c; keys[partner] = a; }
            }
            __syncthreads();
        }
    }
    const int m = s_n;
    const int64_t slot = ((int64_t)cls * B + b) * K;
    if (tid < m) {
        const uint64_t k = keys[tid];
        const int src = (int)(~(uint32_t)k);
        const float* in = boxes + ((int64_t)b * K + src) * box_dim;
        float* o = out_boxes + (slot + tid) * box_dim;
        for (int j = 0; j < box_dim; ++j) o[j] = in[j];
        out_scores[slot + tid] = __uint_as_float((uint32_t)(k >> 32) & 0x7FFFFFFFu);
        out_labels[slot + tid] = cls;
    }
    if (tid == 0) out_counts[cls * B + b] = m;
}
}  // namespace

extern "C" int ql_voxelhead_class_split(const float* boxes, int32_t box_dim, const float* scores, const int32_t* labels, const float* ious,
                                        const int32_t* counts, int32_t B, int32_t K, int32_t num_class, const float* rectifier_dev,
                                        float* out_boxes, float* out_scores, int32_t* out_labels, int32_t* out_counts, ql_stream_t stream_) {
    if (!boxes || !scores || !labels || !ious || !counts || !rectifier_dev || !out_boxes || !out_scores || !out_labels || !out_counts)
        return QL_ERR_INVALID;
    if (B <= 0 || K <= 0 || K > kSelThreads || num_class <= 0 || box_dim < 7 || box_dim > kMaxBoxDim) return QL_ERR_INVALID;
    k_vh_class_split<<<dim3((unsigned)B, (unsigned)num_class), kSelThreads, 0, (cudaStream_t)stream_>>>(
        boxes, box_dim, scores, labels, ious, counts, K, B, rectifier_dev, out_boxes, out_scores, out_labels, out_counts);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" size_t ql_nms_rotated_workspace_bytes(int32_t B, int32_t n_cap) {
    const size_t col_blocks = ((size_t)n_cap + kNmsBlock - 1) / kNmsBlock;
    return align256((size_t)B * n_cap * col_blocks * sizeof(unsigned long long));
}

extern "C" int ql_nms_rotated(const float* boxes, int32_t box_stride, int32_t box_dim, const float* scores, const int32_t* labels,
                              const int32_t* counts, int32_t B, int32_t n_cap, float thresh, int32_t pre_max, int32_t post_max,
                              int32_t label_offset, int32_t* keep, int32_t* keep_count, float* out_boxes, float* out_scores,
                              int32_t* out_labels, float* iou_out, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    if (!boxes || !keep || !keep_count || !workspace || B <= 0 || n_cap <= 0 || box_stride < 7 || box_dim < 7 || box_dim > box_stride ||
        pre_max <= 0 || post_max <= 0)
        return QL_ERR_INVALID;
    const int col_blocks = (n_cap + kNmsBlock - 1) / kNmsBlock;
    if (col_blocks > 32) return QL_ERR_UNSUPPORTED;                        // the sweep keeps one removed-word per lane: n_cap <= 2048
    if (workspace_bytes < ql_nms_rotated_workspace_bytes(B, n_cap)) return QL_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream_;
    unsigned long long* mask = (unsigned long long*)workspace;
    // rows past a frame's count are never written or read; words left of the diagonal are never read
    k_nms_mask<<<dim3((unsigned)col_blocks, (unsigned)((n_cap + kMaskRows - 1) / kMaskRows), (unsigned)B), kNmsBlock * kMaskRows, 0, st>>>(
        boxes, box_stride, counts, n_cap, pre_max, thresh, (uint32_t*)mask, col_blocks, iou_out);
    const size_t smem = (size_t)n_cap * col_blocks * sizeof(unsigned long long) + (size_t)post_max * sizeof(int);
    if (smem > 200 * 1024) return QL_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(k_nms_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return QL_ERR_CUDA;
    k_nms_sweep<<<(unsigned)B, 256, smem, st>>>(boxes, box_stride, box_dim, scores, labels, counts, n_cap, pre_max, post_max, mask, col_blocks,
                                                label_offset, keep, keep_count, out_boxes, out_scores, out_labels);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
