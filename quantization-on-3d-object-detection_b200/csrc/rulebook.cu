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
// Every rulebook also emits a per-tile offset mask (bit k of kmask[tile] set iff some row of the tile has a
// neighbour through offset k), and with a mask the rulebook is COMPACT: a tile's live slabs are stored first, in
// ascending k ([tile][j][128], j = rank of k among the mask's set bits; the tile's block keeps its K*512-byte stride).
// The empty (tile, offset) slabs -- 60 % of them on the backbone's key-sorted stages -- are neither written here nor
// read by the conv kernel, whose loader fetches a unit's slabs with one contiguous bulk copy.  Without a mask
// (tile_kmask == NULL) the layout is the dense nbr[tile][k][128].
#include "ql_common.cuh"
#include "ql_scan.cuh"

namespace {

struct ConvGeom {
    int kd, kh, kw;       // kernel (z, y, x)
    int sd, sh, sw;       // stride
    int pd, ph, pw;       // padding
};

constexpr int kProbeBatch = 9;

// Compact write-out of one tile: thread r holds its column of the dense [K][128] block in the shared-memory stash
// s_res[k * 128 + r]; the live offsets (bits of s_mask, complete after the barrier) go to slabs 0, 1, ... of the tile.
__device__ __forceinline__ void write_compact_tile(const int* s_res, const uint32_t* s_mask, int mask_words, int r, int* dst,
                                                   uint32_t* kmask_tile) {
    __syncthreads();
    int j = 0;
    for (int w = 0; w < mask_words; ++w) {
        uint32_t m = s_mask[w];
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1u;
            dst[(int64_t)j * QL_TILE_M] = s_res[(w * 32 + b) * QL_TILE_M + r];
            ++j;
        }
    }
    if (r < mask_words) kmask_tile[r] = s_mask[r];
}

// nbr for output rows: in = out*stride - pad + offset, looked up in the input table.
__global__ void __launch_bounds__(QL_TILE_M) k_rb_pairs(const int4* __restrict__ out_coords, int64_t n_cap,
                                                        const int* __restrict__ n_dev, QlGrid gin, ConvGeom cg,
                                                        const uint2* __restrict__ table, uint32_t cap_mask,
                                                        int* __restrict__ nbr, uint32_t* __restrict__ kmask) {
    __shared__ uint32_t s_mask[QL_MASK_WORDS_MAX];
    extern __shared__ int s_res[];                           // [K][128] stash of the tile when the output is compact (kmask)
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int64_t tile = blockIdx.x;
    if (tile * QL_TILE_M >= n) return;                       // tiles past the device-side row count are never read
    const int r = threadIdx.x;
    const int64_t row = tile * QL_TILE_M + r;
    const int K = cg.kd * cg.kh * cg.kw;
    int* dst = nbr + tile * (int64_t)K * QL_TILE_M + r;
    const int mask_words = (K + 31) >> 5;
    if (r < mask_words) s_mask[r] = 0u;
    __syncthreads();
    const bool live = row < n;
    const int4 c = live ? out_coords[row] : make_int4(0, 0, 0, 0);      // b, z, y, x
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
            ok[j] = live && k < K && z >= 0 && z < gin.D && y >= 0 && y < gin.H && x >= 0 && x < gin.W;
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
                    s = ql_hash_next(s, cap_mask);
                    ee = __ldg(&table[s]);
                }
                if (ee.x == key[j]) res = (int)ee.y;
            }
            if (kmask) s_res[k * QL_TILE_M + r] = res;
            else dst[(int64_t)k * QL_TILE_M] = res;
            const uint32_t any = __ballot_sync(0xffffffffu, res >= 0);
            if (any && (threadIdx.x & 31) == 0) atomicOr(&s_mask[k >> 5], 1u << (k & 31));
        }
    }
    if (kmask) write_compact_tile(s_res, s_mask, mask_words, r, dst, kmask + tile * mask_words);
}

// Submanifold rulebook over a KEY-SORTED site list (every stage a strided conv produced): the neighbour lookup is a rank
// query on the stage's bitmap -- row(key) = word_prefix[key / 32] + popc(bitmap[key / 32] below the bit) -- instead of a hash
// probe.  The kw cells of one (kz, ky) line are adjacent bits, mostly of ONE bitmap word, so an output row costs ~kd*kh
// bitmap/prefix word pairs (dense, L1/L2 resident) rather than K random 8-byte table probes.
// Also the pair pass of a STRIDED conv whose input stage is key-sorted: rows = output sites, in = out * stride - pad + k is
// looked up in the INPUT stage's rank index (n_in_* bound the valid input rows); no -1 fill, no scatter, no mask pass.
__global__ void __launch_bounds__(QL_TILE_M) k_rb_pairs_ranked(const int4* __restrict__ coords, int64_t n_cap,
                                                               const int* __restrict__ n_dev, QlGrid g, ConvGeom cg,
                                                               const uint32_t* __restrict__ bitmap,
                                                               const uint32_t* __restrict__ word_prefix, int* __restrict__ nbr,
                                                               uint32_t* __restrict__ kmask, int64_t n_in_cap,
                                                               const int* __restrict__ n_in_dev,
                                                               const int* __restrict__ row_perm) {
    __shared__ uint32_t s_mask[QL_MASK_WORDS_MAX];
    extern __shared__ int s_res[];                                   // [K][128] stash of the tile when the output is compact (kmask)
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int64_t n_in = n_in_dev ? min((int64_t)*n_in_dev, n_in_cap) : n_in_cap;
    const int64_t tile = blockIdx.x;
    if (tile * QL_TILE_M >= n) return;
    const int r = threadIdx.x;
    int64_t row = tile * QL_TILE_M + r;                              // tile slot; a grouped rulebook maps it to its output row
    bool live = row < n;
    if (row_perm && live) { row = row_perm[row]; live = row >= 0 && row < n; }
    const int K = cg.kd * cg.kh * cg.kw;
    int* dst = nbr + tile * (int64_t)K * QL_TILE_M + r;
    const int mask_words = (K + 31) >> 5;
    if (r < mask_words) s_mask[r] = 0u;
    __syncthreads();
    const int4 c = live ? coords[row] : make_int4(0, 0, 0, 0);      // b, z, y, x
    const int bz = c.y * cg.sd - cg.pd, by = c.z * cg.sh - cg.ph, bx = c.w * cg.sw - cg.pw;
    for (int kz = 0; kz < cg.kd; ++kz) {
        for (int ky = 0; ky < cg.kh; ++ky) {
            const int z = bz + kz, y = by + ky;
            const bool line_ok = live && z >= 0 && z < g.D && y >= 0 && y < g.H;
            // key of the line's first cell (x = bx, possibly negative: only in-range cells are looked at)
            const uint32_t key0 = ql_key(g, c.x, line_ok ? z : 0, line_ok ? y : 0, 0) + (uint32_t)bx;
            uint32_t w_cached = 0xFFFFFFFFu, bits = 0u, pre = 0u;
            for (int kx = 0; kx < cg.kw; ++kx) {
                const int x = bx + kx;
                int res = -1;
                if (line_ok && x >= 0 && x < g.W) {
                    const uint32_t key = key0 + (uint32_t)kx;
                    const uint32_t w = key >> 5;
                    if (w != w_cached) { bits = __ldg(bitmap + w); w_cached = w; pre = 0xFFFFFFFFu; }
                    const uint32_t bit = 1u << (key & 31u);
                    if (bits & bit) {
                        if (pre == 0xFFFFFFFFu) pre = __ldg(word_prefix + w);
                        const uint32_t rank = pre + (uint32_t)__popc(bits & (bit - 1u));
                        if ((int64_t)rank < n_in) res = (int)rank;
                    }
                }
                const int k = (kz * cg.kh + ky) * cg.kw + kx;
                if (kmask) s_res[k * QL_TILE_M + r] = res;
                else dst[(int64_t)k * QL_TILE_M] = res;
                const uint32_t any = __ballot_sync(0xffffffffu, res >= 0);
                if (any && (threadIdx.x & 31) == 0) atomicOr(&s_mask[k >> 5], 1u << (k & 31));
            }
        }
    }
    if (kmask) write_compact_tile(s_res, s_mask, mask_words, r, dst, kmask + tile * mask_words);
}

// ------------------------------------------------------------------------------------------------
// Strided (regular) sparse conv.  The active output set is built in a bitmap over the output grid
// (one bit per cell, <= B*Do*Ho*Wo/8 bytes, L2 resident for every backbone stage), numbered by a popcount scan
// -- so output rows come out SORTED BY LINEAR KEY ((b*Do+z)*Ho+y)*Wo+x: deterministic, reproducible by the oracle
// with np.unique, and spatially coherent (a 128-row MMA tile is a run along x) -- and the pairs are scattered
// from the INPUT side: an input row has only prod(ceil(k/s)) candidate outputs (3.4 on average for k=3, s=2)
// against K probes per output row.  The output hash table (out coords -> row) is filled in the same pass for the
// submanifold rulebooks / BEV densify that follow on this stage.
// ------------------------------------------------------------------------------------------------
constexpr int kWordsPerBlock = QL_SCAN_THREADS;                      // one bitmap word (32 cells) per thread

// calls f(k, out_key) for every kernel offset k through which input coord c feeds an in-range output site:
// out = (c + pad - k) / stride, exact division (so that in = out*stride - pad + k)
// Per axis the kernel taps k through which input coordinate c reaches an output are k == (c + pad) mod stride, i.e.
// k = r, r + s, r + 2s, ... < K with out = (c + pad - k) / s: walk only those (1-2 per axis for k=3, s=2) instead of testing
// all K taps for divisibility.  Strides 1 and 2 (every conv of the backbones) avoid the integer-division sequence.
__device__ __forceinline__ void axis_first(int n, int s, int& k0, int& o0) {   // n = c + pad >= 0
    if (s == 1) { k0 = 0; o0 = n; }
    else if (s == 2) { k0 = n & 1; o0 = n >> 1; }
    else { k0 = n % s; o0 = n / s; }
}

template <class F>
__device__ __forceinline__ void for_each_candidate(const int4& c, const ConvGeom& cg, const QlGrid& gout, F&& f) {
    int kz0, oz0, ky0, oy0, kx0, ox0;
    axis_first(c.y + cg.pd, cg.sd, kz0, oz0);
    axis_first(c.z + cg.ph, cg.sh, ky0, oy0);
    axis_first(c.w + cg.pw, cg.sw, kx0, ox0);
    for (int kz = kz0, oz = oz0; kz < cg.kd && oz >= 0; kz += cg.sd, --oz) {
        if (oz >= gout.D) continue;
        for (int ky = ky0, oy = oy0; ky < cg.kh && oy >= 0; ky += cg.sh, --oy) {
            if (oy >= gout.H) continue;
            for (int kx = kx0, ox = ox0; kx < cg.kw && ox >= 0; kx += cg.sw, --ox) {
                if (ox >= gout.W) continue;
                f((kz * cg.kh + ky) * cg.kw + kx, ql_key(gout, c.x, oz, oy, ox));
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_rb_mark(const int4* __restrict__ in_coords, int64_t n_cap, const int* __restrict__ n_dev,
                                                 ConvGeom cg, QlGrid gout, uint32_t* __restrict__ bitmap) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;        // clamped: a producer that found more rows than its capacity kept only n_cap
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = in_coords[i];
    for_each_candidate(c, cg, gout, [&](int, uint32_t key) {
        const uint32_t bit = 1u << (key & 31u);
        // most sites are marked by several inputs: test before the atomic
        if (!(__ldcg(&bitmap[key >> 5]) & bit)) atomicOr(&bitmap[key >> 5], bit);
    });
}

__global__ void __launch_bounds__(QL_SCAN_THREADS) k_rb_popc(const uint32_t* __restrict__ bitmap, int64_t n_words, int* block_counts) {
    const int64_t w = (int64_t)blockIdx.x * kWordsPerBlock + threadIdx.x;
    const int c = w < n_words ? __popc(bitmap[w]) : 0;
    int total;
    block_exclusive_scan(c, total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// rank of every set bit = its output row: write coords, insert (key -> row) into the output hash table, and keep the
// per-word exclusive prefix for the rank lookups of k_rb_scatter.  The set bits of a warp's 32 words are dealt out
// evenly to its lanes (item i -> lane i % 32), so dense words do not serialise on one thread.
__global__ void __launch_bounds__(QL_SCAN_THREADS) k_rb_emit(const uint32_t* __restrict__ bitmap, int64_t n_words,
                                                            const int* __restrict__ block_offsets, QlGrid gout,
                                                            uint32_t* __restrict__ word_prefix, int4* __restrict__ out_coords,
                                                            int64_t n_out_cap, uint2* out_table, uint32_t cap_mask) {
    const int64_t w = (int64_t)blockIdx.x * kWordsPerBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t bits = w < n_words ? bitmap[w] : 0u;
    const int c = __popc(bits);
    int total;
    const uint32_t run = (uint32_t)(block_exclusive_scan(c, total) + block_offsets[blockIdx.x]);
    if (w < n_words) word_prefix[w] = run;
    const uint32_t warp_run0 = __shfl_sync(0xffffffffu, run, 0);
    const uint32_t e = run - warp_run0;                                   // exclusive prefix inside the warp
    const uint32_t warp_total = __shfl_sync(0xffffffffu, e + (uint32_t)c, 31);
    const int64_t warp_w0 = w - lane;
    for (uint32_t i0 = 0; i0 < warp_total; i0 += 32) {
        const uint32_t i = i0 + (uint32_t)lane;
        // largest j with e_j <= i (words with no bits share their successor's prefix and are skipped by the search)
        int j = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t ej = __shfl_sync(0xffffffffu, e, (j + step) & 31);
            if (j + step < 32 && ej <= i) j += step;
        }
        const uint32_t bj = __shfl_sync(0xffffffffu, bits, j);
        const uint32_t ej = __shfl_sync(0xffffffffu, e, j);
        if (i < warp_total) {
            const uint32_t rank = warp_run0 + i;
            if ((int64_t)rank < n_out_cap) {
                const uint32_t pos = __fns(bj, 0, (int)(i - ej) + 1);
                const uint32_t key = (uint32_t)(warp_w0 + j) * 32u + pos;
                out_coords[rank] = ql_unkey(gout, key);
                if (out_table) {
                    const uint32_t s = ql_hash_insert(out_table, cap_mask, key);
                    out_table[s].y = rank;
                }
            }
        }
    }
}

// ---- renumbering of a site list by ascending key (the stage-1 grid: 371 M cells = 11.6 M bitmap words for ~0.6 M sites) ----
// The bitmap is huge and almost empty, and the coordinates are already known per input row, so nothing is decoded from the
// bits: four words per thread (16-byte loads) for the popcount and prefix passes, then every input row looks its rank up.
__global__ void __launch_bounds__(256) k_rn_mark(const int4* __restrict__ coords, int64_t n_cap, const int* __restrict__ n_dev, QlGrid g,
                                                 uint32_t* __restrict__ bitmap) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = coords[i];
    const uint32_t key = ql_key(g, c.x, c.y, c.z, c.w);
    atomicOr(&bitmap[key >> 5], 1u << (key & 31u));
}

__global__ void __launch_bounds__(QL_SCAN_THREADS) k_rn_popc4(const uint4* __restrict__ bitmap4, int64_t n_words4, int* block_counts) {
    const int64_t i = (int64_t)blockIdx.x * QL_SCAN_THREADS + threadIdx.x;
    int c = 0;
    if (i < n_words4) {
        const uint4 w = bitmap4[i];
        c = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
    int total;
    block_exclusive_scan(c, total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(QL_SCAN_THREADS) k_rn_prefix4(const uint4* __restrict__ bitmap4, int64_t n_words4,
                                                               const int* __restrict__ block_offsets, uint4* __restrict__ word_prefix4) {
    const int64_t i = (int64_t)blockIdx.x * QL_SCAN_THREADS + threadIdx.x;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (i < n_words4) w = bitmap4[i];
    const int c0 = __popc(w.x), c1 = __popc(w.y), c2 = __popc(w.z), c3 = __popc(w.w);
    int total;
    const uint32_t run = (uint32_t)(block_exclusive_scan(c0 + c1 + c2 + c3, total) + block_offsets[blockIdx.x]);
    if (i < n_words4) word_prefix4[i] = make_uint4(run, run + (uint32_t)c0, run + (uint32_t)(c0 + c1), run + (uint32_t)(c0 + c1 + c2));
}

// every input row takes its rank: coordinates, source row and (optionally) a payload row of chunks x 16 bytes move there
__global__ void __launch_bounds__(256) k_rn_assign(const int4* __restrict__ coords, int64_t n_cap, const int* __restrict__ n_dev, QlGrid g,
                                                   const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_prefix,
                                                   int4* __restrict__ out_coords, int* __restrict__ src_row,
                                                   const uint4* __restrict__ rows_in, uint4* __restrict__ rows_out, int chunks) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = coords[i];
    const uint32_t key = ql_key(g, c.x, c.y, c.z, c.w);
    const uint32_t w = key >> 5;
    const uint32_t rank = __ldg(word_prefix + w) + (uint32_t)__popc(__ldg(bitmap + w) & ((1u << (key & 31u)) - 1u));
    out_coords[rank] = c;
    if (src_row) src_row[rank] = (int)i;
    if (rows_in)
        for (int j = 0; j < chunks; ++j) rows_out[(int64_t)rank * chunks + j] = __ldg(rows_in + i * chunks + j);
}

// nbr := -1 for the live tiles (device-side row count)
__global__ void __launch_bounds__(256) k_rb_fill(int4* __restrict__ nbr4, int K, const int* __restrict__ n_out_dev,
                                                 int64_t n_out_cap) {
    const int64_t n = min((int64_t)*n_out_dev, n_out_cap);
    const int64_t tiles = (n + QL_TILE_M - 1) / QL_TILE_M;
    const int64_t total4 = tiles * K * (QL_TILE_M / 4);
    const int4 m1 = make_int4(-1, -1, -1, -1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) nbr4[i] = m1;
}

__global__ void __launch_bounds__(256) k_rb_scatter(const int4* __restrict__ in_coords, int64_t n_cap, const int* __restrict__ n_dev,
                                                    ConvGeom cg, QlGrid gout, const uint32_t* __restrict__ bitmap,
                                                    const uint32_t* __restrict__ word_prefix, int64_t n_out_cap,
                                                    int* __restrict__ nbr) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = in_coords[i];
    const int K = cg.kd * cg.kh * cg.kw;
    for_each_candidate(c, cg, gout, [&](int k, uint32_t key) {
        const uint32_t w = key >> 5;
        const uint32_t rank = __ldg(&word_prefix[w]) + (uint32_t)__popc(__ldg(&bitmap[w]) & ((1u << (key & 31u)) - 1u));
        if ((int64_t)rank >= n_out_cap) return;                      // site dropped by the caller's capacity
        nbr[((int64_t)(rank / QL_TILE_M) * K + k) * QL_TILE_M + (rank % QL_TILE_M)] = (int)i;
    });
}

// per-tile offset mask of a finished DENSE rulebook, which is then compacted in place: one CTA per live tile, thread r
// reads nbr[tile][k][r] for every k; afterwards it moves its column's live entries to slabs 0, 1, ... (slab j <= k is
// written after slab k was read and is never read again, and a thread only touches its own column: in place is safe)
__global__ void __launch_bounds__(QL_TILE_M) k_rb_kmask(int* nbr, int K, const int* __restrict__ n_out_dev,
                                                        int64_t n_out_cap, uint32_t* __restrict__ kmask, int mask_words) {
    __shared__ uint32_t s_mask[QL_MASK_WORDS_MAX];
    const int64_t n = min((int64_t)*n_out_dev, n_out_cap);
    const int64_t tile = blockIdx.x;
    if (tile * QL_TILE_M >= n) return;
    if (threadIdx.x < mask_words) s_mask[threadIdx.x] = 0u;
    __syncthreads();
    int* src = nbr + tile * (int64_t)K * QL_TILE_M + threadIdx.x;
    for (int k0 = 0; k0 < K; k0 += 9) {
        int v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = (k0 + j < K) ? __ldcg(src + (int64_t)(k0 + j) * QL_TILE_M) : -1;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const uint32_t any = __ballot_sync(0xffffffffu, v[j] >= 0);
            if (any && (threadIdx.x & 31) == 0) atomicOr(&s_mask[(k0 + j) >> 5], 1u << ((k0 + j) & 31));
        }
    }
    __syncthreads();
    if (threadIdx.x < mask_words) kmask[tile * mask_words + threadIdx.x] = s_mask[threadIdx.x];
    int j = 0;
    for (int w = 0; w < mask_words; ++w) {
        uint32_t m = s_mask[w];
        while (m) {
            const int k = w * 32 + __ffs((int)m) - 1;
            m &= m - 1u;
            if (j != k) src[(int64_t)j * QL_TILE_M] = __ldcg(src + (int64_t)k * QL_TILE_M);
            ++j;
        }
    }
}

// dynamic shared memory of the pair kernels: the [K][128] stash of a compact build (opt-in above 48 KB)
template <class Kern>
inline bool stash_smem(Kern kern, int K, bool compact, size_t& bytes) {
    bytes = compact ? (size_t)K * QL_TILE_M * sizeof(int) : 0;
    if (bytes > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return false;
    return true;
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
extern "C" int32_t ql_rulebook_mask_words(int32_t kvol) { return (kvol + 31) / 32; }

extern "C" int ql_rulebook_subm(const int32_t* coords, int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D, int32_t H,
                                int32_t W, const int32_t* ksize, const uint64_t* table, int64_t table_cap, int32_t* nbr_out,
                                uint32_t* tile_kmask, ql_stream_t stream_) {
    ConvGeom cg;
    if (!coords || !table || !nbr_out || !geom_from_host(ksize, nullptr, nullptr, cg)) return QL_ERR_INVALID;
    if (!(cg.kd & 1) || !(cg.kh & 1) || !(cg.kw & 1)) return QL_ERR_INVALID;   // submanifold needs odd kernels
    if (table_cap <= 0 || (table_cap & (table_cap - 1))) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    if (n_cap <= 0) return QL_OK;
    QlGrid g{B, D, H, W};
    unsigned tiles = (unsigned)ql_rulebook_num_tiles(n_cap);
    size_t stash;
    if (!stash_smem(k_rb_pairs, cg.kd * cg.kh * cg.kw, tile_kmask != nullptr, stash)) return QL_ERR_CUDA;
    k_rb_pairs<<<tiles, QL_TILE_M, stash, (cudaStream_t)stream_>>>((const int4*)coords, n_cap, n_dev, g, cg, (const uint2*)table,
                                                               (uint32_t)(table_cap - 1), nbr_out, tile_kmask);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

namespace {
struct StridedWs {
    size_t bitmap, prefix, blocks, total;
    int64_t n_words, n_blocks;
};
StridedWs strided_ws_layout(const QlGrid& gout) {
    StridedWs w;
    const int64_t cells = (int64_t)gout.B * gout.D * gout.H * gout.W;
    w.n_words = ((cells + 31) / 32 + 3) & ~(int64_t)3;                 // multiple of 4 words: the large-bitmap passes load 16 bytes
    w.n_blocks = (w.n_words + kWordsPerBlock - 1) / kWordsPerBlock;
    size_t o = 0;
    w.bitmap = o; o += align256((size_t)w.n_words * 4);
    w.prefix = o; o += align256((size_t)w.n_words * 4);
    w.blocks = o; o += align256((size_t)(w.n_blocks + 1) * 4);
    w.total = o;
    return w;
}
bool out_grid(int32_t B, int32_t D, int32_t H, int32_t W, const ConvGeom& cg, QlGrid& gout) {
    gout = QlGrid{B, (D + 2 * cg.pd - cg.kd) / cg.sd + 1, (H + 2 * cg.ph - cg.kh) / cg.sh + 1, (W + 2 * cg.pw - cg.kw) / cg.sw + 1};
    return B > 0 && gout.D > 0 && gout.H > 0 && gout.W > 0;
}
}  // namespace

extern "C" size_t ql_rulebook_strided_workspace_bytes(int32_t B, int32_t D, int32_t H, int32_t W, const int32_t* ksize,
                                                      const int32_t* stride, const int32_t* pad) {
    ConvGeom cg;
    QlGrid gout;
    if (!stride || !pad || !geom_from_host(ksize, stride, pad, cg) || !out_grid(B, D, H, W, cg, gout)) return 0;
    return strided_ws_layout(gout).total;
}

extern "C" int ql_rulebook_strided(const int32_t* in_coords, int64_t n_in_cap, const int32_t* n_in_dev, int32_t B, int32_t D,
                                   int32_t H, int32_t W, const int32_t* ksize, const int32_t* stride, const int32_t* pad,
                                   int32_t* out_coords, int64_t n_out_cap, int32_t* n_out_dev, uint64_t* out_table,
                                   int64_t out_table_cap, int32_t* nbr_out, uint32_t* tile_kmask, void* workspace,
                                   size_t workspace_bytes, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    ConvGeom cg;
    if (!in_coords || !out_coords || !n_out_dev || !nbr_out || !workspace || !stride || !pad ||
        !geom_from_host(ksize, stride, pad, cg))
        return QL_ERR_INVALID;
    if (out_table && (out_table_cap <= 0 || (out_table_cap & (out_table_cap - 1)) || out_table_cap < 2 * n_out_cap))
        return QL_ERR_INVALID;
    if (n_out_cap <= 0 || n_out_cap >= 2147483647LL || n_in_cap < 0 || n_in_cap >= 2147483647LL) return QL_ERR_INVALID;
    const int K = cg.kd * cg.kh * cg.kw;
    QlGrid gout;
    if (!out_grid(B, D, H, W, cg, gout)) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0 || (double)B * gout.D * gout.H * gout.W >= 4294967295.0)
        return QL_ERR_GRID_TOO_LARGE;
    const StridedWs w = strided_ws_layout(gout);
    if (workspace_bytes < w.total) return QL_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    uint32_t* bitmap = (uint32_t*)(ws + w.bitmap);
    uint32_t* prefix = (uint32_t*)(ws + w.prefix);
    int* blocks = (int*)(ws + w.blocks);
    const int mask_words = (K + 31) / 32;

    if (out_table && cudaMemsetAsync(out_table, 0xFF, (size_t)out_table_cap * 8, st) != cudaSuccess) return QL_ERR_CUDA;
    if (cudaMemsetAsync(bitmap, 0, (size_t)w.n_words * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    const unsigned gin = (unsigned)((n_in_cap + 255) / 256);
    if (gin) k_rb_mark<<<gin, 256, 0, st>>>((const int4*)in_coords, n_in_cap, n_in_dev, cg, gout, bitmap);
    k_rb_popc<<<(unsigned)w.n_blocks, QL_SCAN_THREADS, 0, st>>>(bitmap, w.n_words, blocks);
    k_scan_blocks<<<1, QL_SCAN_THREADS, 0, st>>>(blocks, (int)w.n_blocks, n_out_dev + 1, n_out_dev, n_out_cap);
    k_rb_emit<<<(unsigned)w.n_blocks, QL_SCAN_THREADS, 0, st>>>(bitmap, w.n_words, blocks, gout, prefix, (int4*)out_coords, n_out_cap,
                                                              (uint2*)out_table, out_table ? (uint32_t)(out_table_cap - 1) : 0u);
    k_rb_fill<<<4 * ql_num_sms(), 256, 0, st>>>((int4*)nbr_out, K, n_out_dev, n_out_cap);
    if (gin)
        k_rb_scatter<<<gin, 256, 0, st>>>((const int4*)in_coords, n_in_cap, n_in_dev, cg, gout, bitmap, prefix, n_out_cap, nbr_out);
    if (tile_kmask)
        k_rb_kmask<<<(unsigned)ql_rulebook_num_tiles(n_out_cap), QL_TILE_M, 0, st>>>(nbr_out, K, n_out_dev, n_out_cap, tile_kmask,
                                                                                      mask_words);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_rulebook_strided_index(int32_t B, int32_t D, int32_t H, int32_t W, const int32_t* ksize, const int32_t* stride,
                                         const int32_t* pad, void* workspace, const uint32_t** bitmap, const uint32_t** word_prefix,
                                         int64_t* n_words) {
    ConvGeom cg;
    QlGrid gout;
    if (!workspace || !bitmap || !word_prefix || !stride || !pad || !geom_from_host(ksize, stride, pad, cg) ||
        !out_grid(B, D, H, W, cg, gout))
        return QL_ERR_INVALID;
    const StridedWs w = strided_ws_layout(gout);
    *bitmap = (const uint32_t*)((char*)workspace + w.bitmap);
    *word_prefix = (const uint32_t*)((char*)workspace + w.prefix);
    if (n_words) *n_words = w.n_words;
    return QL_OK;
}

// Renumber a site list (distinct coordinates, any order) by ascending linear key and leave its rank index -- the same
// (bitmap, word_prefix) pair, in the same workspace layout, that a 1x1x1 stride-1 ql_rulebook_strided over the grid would
// leave (ql_rulebook_strided_index(B, D, H, W, {1,1,1}, {1,1,1}, {0,0,0}, workspace) returns the pointers).
extern "C" int ql_renumber_by_key(const int32_t* in_coords, int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D, int32_t H,
                                  int32_t W, int32_t* out_coords, int32_t* n_out_dev, int32_t* src_row, const void* rows_in,
                                  void* rows_out, int32_t row_bytes, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (!in_coords || !out_coords || !n_out_dev || !workspace || n_cap <= 0 || n_cap >= 2147483647LL || in_coords == out_coords)
        return QL_ERR_INVALID;
    if ((rows_in != nullptr) != (rows_out != nullptr) || (rows_in && (row_bytes <= 0 || row_bytes % 16 != 0 || rows_in == rows_out)))
        return QL_ERR_INVALID;
    if (rows_in && ((((uintptr_t)rows_in) | ((uintptr_t)rows_out)) & 15)) return QL_ERR_INVALID;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    const QlGrid g{B, D, H, W};
    const StridedWs w = strided_ws_layout(g);
    if (workspace_bytes < w.total) return QL_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    uint32_t* bitmap = (uint32_t*)(ws + w.bitmap);
    uint32_t* prefix = (uint32_t*)(ws + w.prefix);
    int* blocks = (int*)(ws + w.blocks);
    const int64_t n4 = w.n_words / 4;                                    // the layout pads the word count to a multiple of 4
    const unsigned nb4 = (unsigned)((n4 + QL_SCAN_THREADS - 1) / QL_SCAN_THREADS);
    const unsigned gin = (unsigned)((n_cap + 255) / 256);
    if (cudaMemsetAsync(bitmap, 0, (size_t)w.n_words * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    k_rn_mark<<<gin, 256, 0, st>>>((const int4*)in_coords, n_cap, n_dev, g, bitmap);
    k_rn_popc4<<<nb4, QL_SCAN_THREADS, 0, st>>>((const uint4*)bitmap, n4, blocks);
    k_scan_blocks<<<1, QL_SCAN_THREADS, 0, st>>>(blocks, (int)nb4, n_out_dev + 1, n_out_dev, n_cap);
    k_rn_prefix4<<<nb4, QL_SCAN_THREADS, 0, st>>>((const uint4*)bitmap, n4, blocks, (uint4*)prefix);
    k_rn_assign<<<gin, 256, 0, st>>>((const int4*)in_coords, n_cap, n_dev, g, bitmap, prefix, (int4*)out_coords, src_row,
                                     (const uint4*)rows_in, (uint4*)rows_out, rows_in ? row_bytes / 16 : 0);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

namespace {
// ------------------------------------------------------------------------------------------------
// KEY-SORTED voxelisation straight from the points (the engine's front end, round 2).
// The hash voxeliser (voxelize.cu) numbers voxels in first-touch order like the reference's CPU voxeliser and the engine then
// renumbered them by key: hash insert + first-touch numbering (two scans) + renumbering = 13 launches, 0.16 ms on the critical path
// in front of the first rulebook.  When the order is going to be ascending-key anyway, the voxel id of a point IS the rank of its
// cell in the stage-1 bitmap: mark -> popcount / scan / prefix (the rank index the rulebooks use) -> every point looks its rank up.
// No hash table, no first-touch pass, no renumbering, no row move.  The per-frame voxel cap (MAX_NUMBER_OF_VOXELS, first-touch
// order in the reference) cannot be reproduced in this order: frame_counts reports the voxels found per frame and the caller falls
// back to the hash voxeliser when a frame exceeds its cap (qlidar/engine.py).  Features: the same deterministic first-T-points
// selection and left-to-right sums as voxelize.cu, so the means are bit-identical to the hash path's.
// ------------------------------------------------------------------------------------------------
struct PtGeom {
    const float* points;
    int64_t n_points;
    int stride, has_b, n_feat;
    float mnx, mny, mnz, vsx, vsy, vsz;
    int gx, gy, gz;
    QlGrid grid;
};

__device__ __forceinline__ uint32_t point_key(const PtGeom& P, int64_t p) {
    const float* row = P.points + p * P.stride;
    const int o = P.has_b ? 1 : 0;
    const int b = P.has_b ? (int)row[0] : 0;
    // fp32 floor((p - min) / vs) with IEEE division: identical to numpy/torch fp32 (dynamic_mean_vfe.py:53) and to voxelize.cu
    const float fx = floorf(__fdiv_rn(__fsub_rn(row[o + 0], P.mnx), P.vsx));
    const float fy = floorf(__fdiv_rn(__fsub_rn(row[o + 1], P.mny), P.vsy));
    const float fz = floorf(__fdiv_rn(__fsub_rn(row[o + 2], P.mnz), P.vsz));
    const bool ok = fx >= 0.f && fx < (float)P.gx && fy >= 0.f && fy < (float)P.gy && fz >= 0.f && fz < (float)P.gz && b >= 0 && b < P.grid.B;
    return ok ? ql_key(P.grid, b, (int)fz, (int)fy, (int)fx) : 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(256) k_vs_mark(PtGeom P, uint32_t* __restrict__ bitmap, uint32_t* __restrict__ pt_key) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_points) return;
    const uint32_t key = point_key(P, p);
    pt_key[p] = key;
    if (key == 0xFFFFFFFFu) return;
    const uint32_t bit = 1u << (key & 31u);
    if (!(__ldcg(&bitmap[key >> 5]) & bit)) atomicOr(&bitmap[key >> 5], bit);      // ~3 points per voxel: test before the atomic
}

// point -> rank of its cell (= voxel id, ascending key); the voxel's coordinates are written by every one of its points (same value)
__global__ void __launch_bounds__(256) k_vs_rank(int64_t n_points, QlGrid g, const uint32_t* __restrict__ bitmap,
                                                 const uint32_t* __restrict__ word_prefix, int64_t max_voxels, uint32_t* __restrict__ pt_key_rank,
                                                 int4* __restrict__ out_coords) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    const uint32_t key = pt_key_rank[p];
    if (key == 0xFFFFFFFFu) return;
    const uint32_t w = key >> 5;
    const uint32_t rank = __ldg(word_prefix + w) + (uint32_t)__popc(__ldg(bitmap + w) & ((1u << (key & 31u)) - 1u));
    if ((int64_t)rank >= max_voxels) { pt_key_rank[p] = 0xFFFFFFFFu; return; }    // beyond the batch capacity: dropped (reported as found > kept)
    pt_key_rank[p] = rank;
    out_coords[rank] = ql_unkey(g, key);
}

// voxels found per frame = rank of the frame's first cell in the next frame minus its own
__global__ void k_vs_frames(QlGrid g, const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_prefix, const int* __restrict__ n_dev,
                            int* __restrict__ frame_counts) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.B) return;
    const int64_t per = (int64_t)g.D * g.H * g.W;
    auto rank_of = [&](int64_t cell) -> int64_t {
        if (cell >= per * g.B) return (int64_t)n_dev[1];                          // total found
        const uint32_t w = (uint32_t)(cell >> 5), sh = (uint32_t)(cell & 31);
        return (int64_t)word_prefix[w] + __popc(bitmap[w] & ((1u << sh) - 1u));
    };
    frame_counts[b] = (int)(rank_of(per * (b + 1)) - rank_of(per * b));
}

// the max_pts smallest point indices of every voxel: slot t ends up holding the (t+1)-th smallest regardless of arrival order
__global__ void __launch_bounds__(256) k_vs_select(int64_t n_points, const uint32_t* __restrict__ pt_rank, int max_pts, uint32_t* __restrict__ tmin) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    const uint32_t r = pt_rank[p];
    if (r == 0xFFFFFFFFu) return;
    uint32_t v = (uint32_t)p;
    uint32_t* a = tmin + (int64_t)r * max_pts;
    for (int t = 0; t < max_pts; ++t) {
        const uint32_t old = atomicMin(&a[t], v);
        if (old == 0xFFFFFFFFu) break;
        v = old > v ? old : v;
    }
}

__global__ void __launch_bounds__(256) k_vs_mean(PtGeom P, const uint32_t* __restrict__ tmin, int max_pts, const int* __restrict__ n_dev,
                                                 float* __restrict__ out_feats, int out_stride, int* __restrict__ out_npts) {
    const int vid = blockIdx.x * blockDim.x + threadIdx.x;
    if (vid >= n_dev[0]) return;
    float acc[16];
#pragma unroll
    for (int f = 0; f < 16; ++f) acc[f] = 0.f;
    const uint32_t* a = tmin + (int64_t)vid * max_pts;
    int n = 0;
    for (int t = 0; t < max_pts; ++t) {
        const uint32_t p = a[t];
        if (p == 0xFFFFFFFFu) break;
        const float* row = P.points + (int64_t)p * P.stride + (P.has_b ? 1 : 0);
#pragma unroll
        for (int f = 0; f < 16; ++f)
            if (f < P.n_feat) acc[f] = __fadd_rn(acc[f], row[f]);
        ++n;
    }
    const float d = (float)(n > 0 ? n : 1);
#pragma unroll
    for (int f = 0; f < 16; ++f)
        if (f < P.n_feat) out_feats[(int64_t)vid * out_stride + f] = __fdiv_rn(acc[f], d);
    for (int f = P.n_feat; f < out_stride; ++f) out_feats[(int64_t)vid * out_stride + f] = 0.f;
    out_npts[vid] = n;
}

inline size_t vs_align(size_t x) { return (x + 255) & ~(size_t)255; }

static bool pt_geom(PtGeom& P, const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                    const float* range_min, const float* vsize, const int32_t* grid_xyz, int32_t batch_size) {
    if ((!points && n_points > 0) || !range_min || !vsize || !grid_xyz || n_feat < 3 || n_feat > 16 ||
        point_stride < n_feat + (has_batch_col ? 1 : 0) || batch_size <= 0 || n_points < 0 || n_points >= 2147483647LL)
        return false;
    P.points = points; P.n_points = n_points; P.stride = point_stride; P.has_b = has_batch_col; P.n_feat = n_feat;
    P.mnx = range_min[0]; P.mny = range_min[1]; P.mnz = range_min[2];
    P.vsx = vsize[0]; P.vsy = vsize[1]; P.vsz = vsize[2];
    P.gx = grid_xyz[0]; P.gy = grid_xyz[1]; P.gz = grid_xyz[2];
    P.grid = QlGrid{batch_size, grid_xyz[2] + 1, grid_xyz[1], grid_xyz[0]};      // the backbone's sparse_shape depth (spconv_backbone.py:191)
    return true;
}

}  // namespace

extern "C" size_t ql_voxelize_sorted_workspace_bytes(int64_t max_points, int64_t max_voxels, int32_t max_pts) {
    return vs_align((size_t)max_points * 4) + vs_align((size_t)max_voxels * (size_t)(max_pts > 0 ? max_pts : 1) * 4);
}

// coordinates half: out_coords [max_voxels, 4] (b, z, y, x) ascending by key, n_voxels_dev {kept, found}, frame_counts [B] (optional),
// and in rank_workspace the stage's rank index (the layout of ql_rulebook_strided_index(B, D, H, W, 1, 1, 0))
extern "C" int ql_voxelize_sorted_coords(const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                                         const float* range_min, const float* vsize, const int32_t* grid_xyz, int32_t batch_size,
                                         int64_t max_voxels, int32_t* out_coords, int32_t* n_voxels_dev, int32_t* frame_counts,
                                         void* rank_workspace, size_t rank_workspace_bytes, void* workspace, size_t workspace_bytes,
                                         ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PtGeom P;
    if (!pt_geom(P, points, n_points, point_stride, has_batch_col, n_feat, range_min, vsize, grid_xyz, batch_size) || !out_coords ||
        !n_voxels_dev || !rank_workspace || !workspace || max_voxels <= 0)
        return QL_ERR_INVALID;
    const QlGrid g = P.grid;
    if ((double)g.B * g.D * g.H * g.W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    const StridedWs w = strided_ws_layout(g);
    if (rank_workspace_bytes < w.total || workspace_bytes < ql_voxelize_sorted_workspace_bytes(n_points, max_voxels, 1)) return QL_ERR_WORKSPACE;
    char* ws = (char*)rank_workspace;
    uint32_t* bitmap = (uint32_t*)(ws + w.bitmap);
    uint32_t* prefix = (uint32_t*)(ws + w.prefix);
    int* blocks = (int*)(ws + w.blocks);
    uint32_t* pt_key = (uint32_t*)workspace;
    const int64_t n4 = w.n_words / 4;
    const unsigned nb4 = (unsigned)((n4 + QL_SCAN_THREADS - 1) / QL_SCAN_THREADS);
    if (cudaMemsetAsync(bitmap, 0, (size_t)w.n_words * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    const unsigned gp = (unsigned)((n_points + 255) / 256);
    if (n_points > 0) k_vs_mark<<<gp, 256, 0, st>>>(P, bitmap, pt_key);
    k_rn_popc4<<<nb4, QL_SCAN_THREADS, 0, st>>>((const uint4*)bitmap, n4, blocks);
    k_scan_blocks<<<1, QL_SCAN_THREADS, 0, st>>>(blocks, (int)nb4, n_voxels_dev + 1, n_voxels_dev, max_voxels);
    k_rn_prefix4<<<nb4, QL_SCAN_THREADS, 0, st>>>((const uint4*)bitmap, n4, blocks, (uint4*)prefix);
    if (n_points > 0) k_vs_rank<<<gp, 256, 0, st>>>(n_points, g, bitmap, prefix, max_voxels, pt_key, (int4*)out_coords);
    if (frame_counts) k_vs_frames<<<(unsigned)((g.B + 63) / 64), 64, 0, st>>>(g, bitmap, prefix, n_voxels_dev, frame_counts);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

// features half (same points, same workspace, stream-ordered after the coordinates half): out_feats [max_voxels, out_feat_stride]
// = the mean of each voxel's first max_pts points (by point index), out_npts = how many that was
extern "C" int ql_voxelize_sorted_features(const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                                           const float* range_min, const float* vsize, const int32_t* grid_xyz, int32_t batch_size,
                                           int32_t max_pts, int64_t max_voxels, const int32_t* n_voxels_dev, float* out_feats,
                                           int32_t out_feat_stride, int32_t* out_npts, void* workspace, size_t workspace_bytes,
                                           ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PtGeom P;
    if (!pt_geom(P, points, n_points, point_stride, has_batch_col, n_feat, range_min, vsize, grid_xyz, batch_size) || !n_voxels_dev ||
        !out_feats || !out_npts || !workspace || max_voxels <= 0 || out_feat_stride < n_feat)
        return QL_ERR_INVALID;
    if (max_pts <= 0) return QL_ERR_UNSUPPORTED;                            // the cap-free (dynamic) mean stays on the hash voxeliser
    if (workspace_bytes < ql_voxelize_sorted_workspace_bytes(n_points, max_voxels, max_pts)) return QL_ERR_WORKSPACE;
    const uint32_t* pt_rank = (const uint32_t*)workspace;
    uint32_t* tmin = (uint32_t*)((char*)workspace + vs_align((size_t)n_points * 4));
    if (cudaMemsetAsync(tmin, 0xFF, (size_t)max_voxels * max_pts * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    if (n_points > 0) k_vs_select<<<(unsigned)((n_points + 255) / 256), 256, 0, st>>>(n_points, pt_rank, max_pts, tmin);
    k_vs_mean<<<(unsigned)((max_voxels + 255) / 256), 256, 0, st>>>(P, tmin, max_pts, n_voxels_dev, out_feats, out_feat_stride, out_npts);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_rulebook_subm_ranked(const int32_t* coords, int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D, int32_t H,
                                       int32_t W, const int32_t* ksize, const uint32_t* bitmap, const uint32_t* word_prefix,
                                       int32_t* nbr_out, uint32_t* tile_kmask, ql_stream_t stream_) {
    ConvGeom cg;
    if (!coords || !bitmap || !word_prefix || !nbr_out || !geom_from_host(ksize, nullptr, nullptr, cg)) return QL_ERR_INVALID;
    if (!(cg.kd & 1) || !(cg.kh & 1) || !(cg.kw & 1)) return QL_ERR_INVALID;   // submanifold needs odd kernels
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    if (n_cap <= 0) return QL_OK;
    QlGrid g{B, D, H, W};
    unsigned tiles = (unsigned)ql_rulebook_num_tiles(n_cap);
    size_t stash;
    if (!stash_smem(k_rb_pairs_ranked, cg.kd * cg.kh * cg.kw, tile_kmask != nullptr, stash)) return QL_ERR_CUDA;
    k_rb_pairs_ranked<<<tiles, QL_TILE_M, stash, (cudaStream_t)stream_>>>((const int4*)coords, n_cap, n_dev, g, cg, bitmap, word_prefix,
                                                                      nbr_out, tile_kmask, n_cap, n_dev, nullptr);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

// ------------------------------------------------------------------------------------------------
// GROUPED submanifold rulebook.  The conv kernel's cost is the number of non-empty (tile, kernel offset) slabs, and with
// 128 consecutive rows per tile nearly every offset is live in nearly every tile although a row has only 7-15 of its 27
// neighbours.  spconv's MaskImplicitGemm sorts the outputs by their neighbour mask for the same reason.  Here the rows
// are binned (counting sort, 512 bins) by a 9-bit LINE key -- bit (dz, dy) set iff any cell of the kernel's x-line at
// (z + dz, y + dy) is active, dz, dy in {-1, 0, 1} -- which on LiDAR surfaces recovers almost all of what a full 27-bit
// mask sort gives (synthetic Waymo frame, live offsets per tile: stage 1 27 -> 10.8 (full sort 9.5), stage 2 21.3 -> 17.7
// (15.9)), costs 9 bitmap words per row, and needs no multi-pass radix sort.  Tile slot -> output row goes to row_perm;
// the feature / coordinate arrays keep their order, only the tiling of the output rows changes, so results are
// bit-identical to the ungrouped rulebook's (skipped slabs only ever added exact zeros).
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kGroupBins = 512;

__device__ __forceinline__ uint32_t line_key(const int4& c, const QlGrid& g, const ConvGeom& cg, const uint32_t* __restrict__ bitmap) {
    const int hx = cg.kw >> 1;
    uint32_t key = 0u;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            if ((dz != 0 && cg.kd < 3) || (dy != 0 && cg.kh < 3)) continue;
            const int z = c.y + dz, y = c.z + dy;
            if (z < 0 || z >= g.D || y < 0 || y >= g.H) continue;
            const int x0 = max(c.w - hx, 0), x1 = min(c.w + hx, g.W - 1);
            const uint32_t k0 = ql_key(g, c.x, z, y, x0), k1 = k0 + (uint32_t)(x1 - x0);
            // cells k0..k1 of the bitmap (at most two words for kw <= 32)
            const uint32_t w0 = k0 >> 5, w1 = k1 >> 5;
            uint32_t any;
            if (w0 == w1) {
                const uint32_t m = (0xFFFFFFFFu >> (31u - (k1 & 31u))) & (0xFFFFFFFFu << (k0 & 31u));
                any = __ldg(bitmap + w0) & m;
            } else {
                any = (__ldg(bitmap + w0) & (0xFFFFFFFFu << (k0 & 31u))) | (__ldg(bitmap + w1) & (0xFFFFFFFFu >> (31u - (k1 & 31u))));
            }
            if (any) key |= 1u << ((dz + 1) * 3 + (dy + 1));
        }
    }
    return key;
}

__global__ void __launch_bounds__(256) k_rb_linekey(const int4* __restrict__ coords, int64_t n_cap, const int* __restrict__ n_dev,
                                                    QlGrid g, ConvGeom cg, const uint32_t* __restrict__ bitmap,
                                                    uint16_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[kGroupBins];
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    if ((int64_t)blockIdx.x * blockDim.x >= n) return;
    for (int i = threadIdx.x; i < kGroupBins; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row < n) {
        const uint32_t key = line_key(coords[row], g, cg, bitmap);
        keys[row] = (uint16_t)key;
        // neighbouring rows mostly share a key: one shared-memory atomic per distinct key of the warp
        const uint32_t peers = __match_any_sync(__activemask(), key);
        if ((threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&s_hist[key], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kGroupBins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// hist holds the bins' first slots after the scan; every row takes the next slot of its bin
__global__ void __launch_bounds__(256) k_rb_group_scatter(const uint16_t* __restrict__ keys, int64_t n_cap, const int* __restrict__ n_dev,
                                                          uint32_t* __restrict__ cursor, int* __restrict__ row_perm) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slots = (n + QL_TILE_M - 1) / QL_TILE_M * QL_TILE_M;
    if (row >= n) {
        if (row < slots) row_perm[row] = -1;                         // padding of the last tile
        return;
    }
    const uint32_t key = keys[row];
    const uint32_t peers = __match_any_sync(__activemask(), key);
    const int leader = __ffs((int)peers) - 1, lane = threadIdx.x & 31;
    uint32_t base = 0u;
    if (lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    row_perm[base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = (int)row;
}
}  // namespace

extern "C" size_t ql_rulebook_group_workspace_bytes(int64_t n_cap) {
    if (n_cap <= 0) return 0;
    return align256((size_t)n_cap * 2) + align256((size_t)kGroupBins * 4);
}

extern "C" int ql_rulebook_subm_ranked_grouped(const int32_t* coords, int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D,
                                               int32_t H, int32_t W, const int32_t* ksize, const uint32_t* bitmap,
                                               const uint32_t* word_prefix, int32_t* nbr_out, uint32_t* tile_kmask,
                                               int32_t* row_perm_out, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    ConvGeom cg;
    if (!coords || !bitmap || !word_prefix || !nbr_out || !row_perm_out || !workspace || !geom_from_host(ksize, nullptr, nullptr, cg))
        return QL_ERR_INVALID;
    if (!(cg.kd & 1) || !(cg.kh & 1) || !(cg.kw & 1) || cg.kw > 31) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    if (n_cap <= 0) return QL_OK;
    if (n_cap >= 2147483647LL) return QL_ERR_INVALID;
    if (workspace_bytes < ql_rulebook_group_workspace_bytes(n_cap)) return QL_ERR_WORKSPACE;
    uint16_t* keys = (uint16_t*)workspace;
    uint32_t* hist = (uint32_t*)((char*)workspace + align256((size_t)n_cap * 2));
    QlGrid g{B, D, H, W};
    const unsigned tiles = (unsigned)ql_rulebook_num_tiles(n_cap);
    const unsigned blocks = (unsigned)((tiles * (int64_t)QL_TILE_M + 255) / 256);       // covers the padding slots of the last tile
    if (cudaMemsetAsync(hist, 0, (size_t)kGroupBins * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    k_rb_linekey<<<blocks, 256, 0, st>>>((const int4*)coords, n_cap, n_dev, g, cg, bitmap, keys, hist);
    k_scan_blocks<<<1, QL_SCAN_THREADS, 0, st>>>((int*)hist, kGroupBins, nullptr, nullptr, 0);
    k_rb_group_scatter<<<blocks, 256, 0, st>>>(keys, n_cap, n_dev, hist, row_perm_out);
    size_t stash;
    if (!stash_smem(k_rb_pairs_ranked, cg.kd * cg.kh * cg.kw, tile_kmask != nullptr, stash)) return QL_ERR_CUDA;
    k_rb_pairs_ranked<<<tiles, QL_TILE_M, stash, st>>>((const int4*)coords, n_cap, n_dev, g, cg, bitmap, word_prefix, nbr_out, tile_kmask,
                                                   n_cap, n_dev, row_perm_out);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

// ------------------------------------------------------------------------------------------------
// VoxelNeXt 2-D merge (VoxelResBackBone8xVoxelNeXt.bev_out, spconv_backbone_voxelnext.py:149-164): drop z, keep the
// unique (b, y, x) rows in ascending order (== torch.unique(dim=0)) and sum the features of the rows that collapse
// onto the same site (index_add_).  Same bitmap + popcount-scan numbering as the strided rulebook; the sums are fp32
// atomics (like index_add_ on the GPU, the order of the <= D additions per site is not fixed).
// ------------------------------------------------------------------------------------------------
namespace {
struct Merge2dWs {
    size_t bitmap, prefix, blocks, acc, total;
    int64_t n_words, n_blocks;
};
Merge2dWs merge2d_ws_layout(int B, int H, int W, int64_t n_out_cap, int c, bool need_acc) {
    Merge2dWs w;
    const int64_t cells = (int64_t)B * H * W;
    w.n_words = (cells + 31) / 32;
    w.n_blocks = (w.n_words + kWordsPerBlock - 1) / kWordsPerBlock;
    size_t o = 0;
    w.bitmap = o; o += align256((size_t)w.n_words * 4);
    w.prefix = o; o += align256((size_t)w.n_words * 4);
    w.blocks = o; o += align256((size_t)(w.n_blocks + 1) * 4);
    w.acc = o; o += need_acc ? align256((size_t)n_out_cap * c * 4) : 0;
    w.total = o;
    return w;
}

__global__ void __launch_bounds__(256) k_m2d_mark(const int4* __restrict__ coords, int64_t n_cap, const int* __restrict__ n_dev, int B,
                                                  int H, int W, int cscale, uint32_t* __restrict__ bitmap) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = coords[i];                                                   // b, z, y, x
    c.z *= cscale; c.w *= cscale;                                         // a coarser stage's site on the target grid (voxelnext :194-195)
    if (c.x < 0 || c.x >= B || c.z < 0 || c.z >= H || c.w < 0 || c.w >= W) return;
    const uint32_t key = (uint32_t)((c.x * H + c.z) * W + c.w);
    const uint32_t bit = 1u << (key & 31u);
    if (!(__ldcg(&bitmap[key >> 5]) & bit)) atomicOr(&bitmap[key >> 5], bit);
}

__global__ void __launch_bounds__(QL_SCAN_THREADS) k_m2d_emit(const uint32_t* __restrict__ bitmap, int64_t n_words,
                                                             const int* __restrict__ block_offsets, int H, int W,
                                                             uint32_t* __restrict__ word_prefix, int* __restrict__ out_coords,
                                                             int64_t n_out_cap, int out_cols) {
    const int64_t w = (int64_t)blockIdx.x * kWordsPerBlock + threadIdx.x;
    uint32_t bits = w < n_words ? bitmap[w] : 0u;
    int total;
    uint32_t rank = (uint32_t)(block_exclusive_scan(__popc(bits), total) + block_offsets[blockIdx.x]);
    if (w < n_words) word_prefix[w] = rank;
    while (bits) {
        const uint32_t pos = (uint32_t)__ffs((int)bits) - 1u;
        bits &= bits - 1u;
        if ((int64_t)rank < n_out_cap) {
            uint32_t key = (uint32_t)w * 32u + pos;
            const int x = (int)(key % (uint32_t)W); key /= (uint32_t)W;
            const int y = (int)(key % (uint32_t)H); key /= (uint32_t)H;
            if (out_cols == 4) {                                          // [b, 0, y, x]: the form the 2-D tail's rulebooks take
                reinterpret_cast<int4*>(out_coords)[rank] = make_int4((int)key, 0, y, x);
            } else {
                out_coords[3 * (int64_t)rank] = (int)key;
                out_coords[3 * (int64_t)rank + 1] = y;
                out_coords[3 * (int64_t)rank + 2] = x;
            }
        }
        ++rank;
    }
}

template <typename TIn>
__global__ void __launch_bounds__(256) k_m2d_add(const TIn* __restrict__ feats, int c, const int4* __restrict__ coords, int64_t n_cap,
                                                 const int* __restrict__ n_dev, int B, int H, int W, int cscale,
                                                 const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_prefix,
                                                 int64_t n_out_cap, float* __restrict__ acc) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int groups = c >> 2;                                            // c % 4 == 0 (checked on the host)
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * groups) return;
    const int64_t row = t / groups;
    const int g4 = (int)(t - row * groups) * 4;
    int4 cc = coords[row];
    cc.z *= cscale; cc.w *= cscale;
    if (cc.x < 0 || cc.x >= B || cc.z < 0 || cc.z >= H || cc.w < 0 || cc.w >= W) return;
    const uint32_t key = (uint32_t)((cc.x * H + cc.z) * W + cc.w);
    const uint32_t wd = key >> 5;
    const uint32_t rank = __ldg(word_prefix + wd) + (uint32_t)__popc(__ldg(bitmap + wd) & ((1u << (key & 31u)) - 1u));
    if ((int64_t)rank >= n_out_cap) return;
    float v[4];
    if constexpr (sizeof(TIn) == 2) {
        const uint2 raw = *reinterpret_cast<const uint2*>(feats + row * c + g4);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
        const float4 f = *reinterpret_cast<const float4*>(feats + row * c + g4);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
    float* dst = acc + (int64_t)rank * c + g4;
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dst + j, v[j]);
}

// acc[0 .. min(n_out, cap) * c) = 0: only the rows the merge produced (the capacity is a sum of stage capacities, hundreds of MB)
__global__ void __launch_bounds__(256) k_m2d_zero(float4* __restrict__ acc4, int64_t n_out_cap, int c, const int* __restrict__ n_out_dev) {
    const int64_t total4 = min((int64_t)n_out_dev[0], n_out_cap) * c / 4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) acc4[i] = z;
}

__global__ void __launch_bounds__(256) k_m2d_to_half(const float* __restrict__ acc, int64_t n_elems_cap, int c, const int* __restrict__ n_out_dev,
                                                     __half* __restrict__ out) {
    const int64_t total = min((int64_t)n_out_dev[0] * c, n_elems_cap);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __float2half_rn(acc[i]);
}
}  // namespace

extern "C" size_t ql_bev_merge2d_workspace_bytes(int32_t B, int32_t H, int32_t W, int64_t n_out_cap, int32_t c, int32_t out_dtype) {
    if (B <= 0 || H <= 0 || W <= 0 || n_out_cap <= 0 || c <= 0) return 0;
    return merge2d_ws_layout(B, H, W, n_out_cap, c, out_dtype != QL_F32).total;
}

extern "C" int ql_bev_merge2d(const void* feats, int32_t in_dtype, int32_t c, const int32_t* coords, int64_t n_cap, const int32_t* n_dev,
                              int32_t B, int32_t H, int32_t W, void* out_feats, int32_t out_dtype, int32_t* out_coords,
                              int64_t n_out_cap, int32_t* n_out_dev, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    const int32_t one = 1;
    return ql_bev_merge2d_multi(1, &feats, in_dtype, c, &coords, &n_cap, &n_dev, &one, B, H, W, out_feats, out_dtype, out_coords, 3, n_out_cap,
                                n_out_dev, workspace, workspace_bytes, stream_);
}

extern "C" int ql_bev_merge2d_multi(int32_t n_seg, const void* const* feats, int32_t in_dtype, int32_t c, const int32_t* const* coords,
                                    const int64_t* n_cap, const int32_t* const* n_dev, const int32_t* coord_scale, int32_t B, int32_t H,
                                    int32_t W, void* out_feats, int32_t out_dtype, int32_t* out_coords, int32_t out_coord_cols,
                                    int64_t n_out_cap, int32_t* n_out_dev, void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_seg <= 0 || n_seg > 8 || !feats || !coords || !n_cap || !n_dev || !coord_scale || !out_feats || !out_coords || !n_out_dev || !workspace)
        return QL_ERR_INVALID;
    if ((in_dtype != QL_F16 && in_dtype != QL_F32) || (out_dtype != QL_F16 && out_dtype != QL_F32)) return QL_ERR_INVALID;
    if (out_coord_cols != 3 && out_coord_cols != 4) return QL_ERR_INVALID;
    if (c <= 0 || c % 4 != 0 || B <= 0 || H <= 0 || W <= 0 || n_out_cap <= 0) return QL_ERR_INVALID;
    for (int i = 0; i < n_seg; ++i)
        if (!feats[i] || !coords[i] || n_cap[i] < 0 || n_cap[i] >= 2147483647LL || coord_scale[i] <= 0) return QL_ERR_INVALID;
    if ((double)B * H * W >= 4294967295.0) return QL_ERR_GRID_TOO_LARGE;
    const Merge2dWs w = merge2d_ws_layout(B, H, W, n_out_cap, c, out_dtype != QL_F32);
    if (workspace_bytes < w.total) return QL_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    uint32_t* bitmap = (uint32_t*)(ws + w.bitmap);
    uint32_t* prefix = (uint32_t*)(ws + w.prefix);
    int* blocks = (int*)(ws + w.blocks);
    float* acc = out_dtype == QL_F32 ? (float*)out_feats : (float*)(ws + w.acc);
    if (cudaMemsetAsync(bitmap, 0, (size_t)w.n_words * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    for (int i = 0; i < n_seg; ++i)
        if (n_cap[i] > 0)
            k_m2d_mark<<<(unsigned)((n_cap[i] + 255) / 256), 256, 0, st>>>((const int4*)coords[i], n_cap[i], n_dev[i], B, H, W, coord_scale[i], bitmap);
    k_rb_popc<<<(unsigned)w.n_blocks, QL_SCAN_THREADS, 0, st>>>(bitmap, w.n_words, blocks);
    k_scan_blocks<<<1, QL_SCAN_THREADS, 0, st>>>(blocks, (int)w.n_blocks, n_out_dev + 1, n_out_dev, n_out_cap);
    k_m2d_emit<<<(unsigned)w.n_blocks, QL_SCAN_THREADS, 0, st>>>(bitmap, w.n_words, blocks, H, W, prefix, out_coords, n_out_cap, out_coord_cols);
    k_m2d_zero<<<4 * ql_num_sms(), 256, 0, st>>>((float4*)acc, n_out_cap, c, n_out_dev);
    for (int i = 0; i < n_seg; ++i) {
        if (n_cap[i] <= 0) continue;
        const int64_t threads = n_cap[i] * (c / 4);
        if (in_dtype == QL_F16)
            k_m2d_add<__half><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const __half*)feats[i], c, (const int4*)coords[i], n_cap[i], n_dev[i],
                                                                                 B, H, W, coord_scale[i], bitmap, prefix, n_out_cap, acc);
        else
            k_m2d_add<float><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const float*)feats[i], c, (const int4*)coords[i], n_cap[i], n_dev[i],
                                                                                B, H, W, coord_scale[i], bitmap, prefix, n_out_cap, acc);
    }
    if (out_dtype == QL_F16)
        k_m2d_to_half<<<4 * ql_num_sms(), 256, 0, st>>>(acc, n_out_cap * (int64_t)c, c, n_out_dev, (__half*)out_feats);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_rulebook_strided_ranked(const int32_t* in_coords, int64_t n_in_cap, const int32_t* n_in_dev, int32_t B, int32_t D,
                                          int32_t H, int32_t W, const int32_t* ksize, const int32_t* stride, const int32_t* pad,
                                          const uint32_t* in_bitmap, const uint32_t* in_word_prefix, int32_t* out_coords,
                                          int64_t n_out_cap, int32_t* n_out_dev, int32_t* nbr_out, uint32_t* tile_kmask,
                                          void* workspace, size_t workspace_bytes, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    ConvGeom cg;
    if (!in_coords || !in_bitmap || !in_word_prefix || !out_coords || !n_out_dev || !nbr_out || !workspace || !stride || !pad ||
        !geom_from_host(ksize, stride, pad, cg))
        return QL_ERR_INVALID;
    if (n_out_cap <= 0 || n_out_cap >= 2147483647LL || n_in_cap < 0 || n_in_cap >= 2147483647LL) return QL_ERR_INVALID;
    QlGrid gout;
    if (!out_grid(B, D, H, W, cg, gout)) return QL_ERR_INVALID;
    if ((double)B * D * H * W >= 4294967295.0 || (double)B * gout.D * gout.H * gout.W >= 4294967295.0)
        return QL_ERR_GRID_TOO_LARGE;
    const StridedWs w = strided_ws_layout(gout);
    if (workspace_bytes < w.total) return QL_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    uint32_t* bitmap = (uint32_t*)(ws + w.bitmap);
    uint32_t* prefix = (uint32_t*)(ws + w.prefix);
    int* blocks = (int*)(ws + w.blocks);
    if (cudaMemsetAsync(bitmap, 0, (size_t)w.n_words * 4, st) != cudaSuccess) return QL_ERR_CUDA;
    const unsigned gin = (unsigned)((n_in_cap + 255) / 256);
    if (gin) k_rb_mark<<<gin, 256, 0, st>>>((const int4*)in_coords, n_in_cap, n_in_dev, cg, gout, bitmap);
    k_rb_popc<<<(unsigned)w.n_blocks, QL_SCAN_THREADS, 0, st>>>(bitmap, w.n_words, blocks);
    k_scan_blocks<<<1, QL_SCAN_THREADS, 0, st>>>(blocks, (int)w.n_blocks, n_out_dev + 1, n_out_dev, n_out_cap);
    k_rb_emit<<<(unsigned)w.n_blocks, QL_SCAN_THREADS, 0, st>>>(bitmap, w.n_words, blocks, gout, prefix, (int4*)out_coords, n_out_cap,
                                                              nullptr, 0u);
    // pairs from the output side through the INPUT stage's rank index: every (tile, offset) slab is written exactly once
    QlGrid gin_grid{B, D, H, W};
    size_t stash;
    if (!stash_smem(k_rb_pairs_ranked, cg.kd * cg.kh * cg.kw, tile_kmask != nullptr, stash)) return QL_ERR_CUDA;
    k_rb_pairs_ranked<<<(unsigned)ql_rulebook_num_tiles(n_out_cap), QL_TILE_M, stash, st>>>((const int4*)out_coords, n_out_cap, n_out_dev,
                                                                                        gin_grid, cg, in_bitmap, in_word_prefix, nbr_out,
                                                                                        tile_kmask, n_in_cap, n_in_dev, nullptr);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
