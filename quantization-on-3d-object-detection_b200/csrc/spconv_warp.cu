// Register-gather sparse convolution for NARROW layers (rows of 16..64 bytes, c_out <= 64): warp-level tensor MMAs
// (mma.sync m16n8k16 f16 / m16n8k32 s8), no shared-memory A tile, no tensor memory, no inter-warp protocol.
//
// Same contract as k_spconv_ts (spconv_mma.cu; QConvNd.forward -> [EXT] spconv conv forward, quant/quant.py:36-58, plus the
// BatchNorm1d / ReLU / residual of spconv_backbone.py:8-27,51-67): same rulebook (compact slabs + per-tile offset mask), same
// packed weights, same fused epilogue, same zero-row contract.  ql_spconv_mma_rows routes a layer here when it qualifies.
//
// Why a second kernel: on the C = 16 / 32 layers of the backbone a 128-row tcgen05 tile does 15-42 tiny MMAs (N = 16 / 32 uses
// 1/16 - 1/8 of the tensor pipe) and the time goes into the four-role hand-off pipeline around them, all roles 60-80 % busy
// (profiles/r02_conv_role_trace.md).  These layers are gather-bound, not tensor-bound; what they need is many independent
// gathers in flight.  Here every WARP owns 16 output rows: it reads its rows' neighbour indices of a live offset, loads the
// neighbours' feature rows straight into the A fragments of mma.sync (a quad of lanes reads one row's contiguous bytes; the K
// order this leaves is undone when the weights are laid out in shared memory), multiplies with B fragments from shared memory and
// keeps the 16 x c_out accumulators in registers.  32 resident warps per SM hide the latency; an offset slab none of whose 16 rows
// has a neighbour is skipped (the tcgen05 kernel can only skip per 128 rows).  Output channels are permuted across the MMA's
// n index so that every lane ends up with 2*c_out/8 CONSECUTIVE channels of its two rows: residual loads and stores are 8-32
// contiguous bytes per lane.
#include "ql_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kThreads = 512;                       // 16 warps: two 128-row rulebook tiles per CTA iteration

struct WarpConvParams {
    const uint8_t* feats;   // [n_in][RB], preceded by one all-zero row (index -1)
    const int* nbr;         // compact rulebook [tiles][kvol][128]
    const uint32_t* kmask;  // [tiles][mask_words] or null (dense rulebook)
    const int* row_perm;    // grouped rulebook: tile slot -> output row, or null
    const int* n_out_dev;
    int64_t n_out_cap;
    int kvol, mask_words;
    int ch;                 // chunk width of the packed weight image (32 / 64 / 128)
    const uint8_t* w_packed;
    const float* scale;
    const float* shift;
    const float* act_scale_dev;
    const __half* residual;
    int relu;
    void* out;
    int out_dtype;
    int8_t* out_q;
    const float* out_qscale;
    float* absmax;
};

__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_s8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    int* ci = reinterpret_cast<int*>(c);
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(ci[0]), "+r"(ci[1]), "+r"(ci[2]), "+r"(ci[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// the row of neighbour `id` (-1 = the zero row), this lane's BL bytes: base + id * RB in one IMAD.WIDE
template <int AW>
__device__ __forceinline__ void gather_row(const uint8_t* base, int id, int rb, uint32_t (&v)[AW]) {
    uint64_t a;
    asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(a) : "r"(id), "r"(rb), "l"(base));
    if constexpr (AW == 1) {
        asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v[0]) : "l"(a));
    } else if constexpr (AW == 2) {
        asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "l"(a));
    } else if constexpr (AW == 4) {
        asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(a));
    } else {
        asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(a));
        asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(a + 16));
    }
}

template <int N>
__device__ __forceinline__ void store_words(void* dst, const uint32_t* w) {
    if constexpr (N == 1) *reinterpret_cast<uint32_t*>(dst) = w[0];
    else if constexpr (N == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    else {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
}
template <int N>
__device__ __forceinline__ void load_words(const void* src, uint32_t* w) {
    if constexpr (N == 1) w[0] = __ldg(reinterpret_cast<const uint32_t*>(src));
    else if constexpr (N == 2) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(src)); w[0] = v.x; w[1] = v.y; }
    else {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
}

// byte offset, inside the packed weight image of spconv_mma.cu (ql_pack_weights_host), of source K word `w` (4 bytes) of
// output channel `oc` in the chunk of kernel offset k (one chunk per offset: rows <= 128 bytes)
__device__ __forceinline__ uint32_t packed_word_offset(int ch, int c_out, int k, int oc, int w) {
    int c = w;                                                       // TMEM column that holds source word w (inverse of k_word_src)
    if (ch == 64) c = 8 * ((w >> 1) & 1) + 2 * (w >> 2) + (w & 1);
    else if (ch == 128) c = 8 * ((w >> 1) & 3) + 2 * (w >> 3) + (w & 1);
    const uint32_t r = (uint32_t)oc, c16 = (uint32_t)(c >> 2);
    const uint32_t x = ch == 128 ? (r & 7u) : (ch == 64 ? ((r >> 1) & 3u) : ((r >> 2) & 1u));
    return (uint32_t)k * (uint32_t)(c_out * ch) + (r >> 3) * (uint32_t)(8 * ch) + (r & 7u) * (uint32_t)ch + ((c16 ^ x) << 4) + 4u * (uint32_t)(c & 3);
}

// RB = bytes per input row (16: int8 x 16 channels, 32, 64, 128), COUT = 16 / 32 / 64
template <bool kInt8, int RB, int COUT, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_spconv_warp(const WarpConvParams p) {
    constexpr int BL = RB / 4;                         // bytes of a row per lane of the quad
    constexpr int AW = BL / 4;                         // 32-bit words per row per lane
    constexpr bool kHalf = BL == 4;                    // a row is half a k-step: two live offsets share one MMA
    constexpr int KS = kHalf ? 1 : BL / 8;             // MMA k-steps per offset
    constexpr int NT = COUT / 8;                       // n-tiles
    constexpr int NB = AW <= 2 ? 4 : (AW == 4 ? 2 : 1);   // offsets gathered per batch (loads in flight per warp: 2 * NB)
    constexpr int kWOff = kHalf ? (NT / 2) * 256 : KS * (NT / 2) * 512;   // shared-memory weight bytes per kernel offset
    extern __shared__ uint4 smem_dyn[];
    uint8_t* s_w = reinterpret_cast<uint8_t*>(smem_dyn);
    float* s_scale = reinterpret_cast<float*>(s_w + (size_t)p.kvol * kWOff);
    float* s_shift = s_scale + COUT;
    float* s_qscale = s_shift + COUT;
    uint32_t* s_absmax = reinterpret_cast<uint32_t*>(s_qscale + COUT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // ---- prologue: weights from the packed (tcgen05) image into B-fragment order; epilogue constants
    {
        constexpr int kRowWords = RB / 4;
        const int total = p.kvol * COUT * kRowWords;
        for (int i = tid; i < total; i += kThreads) {
            const int w = i % kRowWords, oc = (i / kRowWords) % COUT, k = i / (kRowWords * COUT);
            const uint32_t val = __ldg(reinterpret_cast<const uint32_t*>(p.w_packed + packed_word_offset(p.ch, COUT, k, oc, w)));
            const int tq = oc / (2 * NT), rem = oc % (2 * NT), j = rem >> 1, e = rem & 1;
            uint32_t dst;
            if constexpr (kHalf) {
                const int ln = (2 * tq + e) * 4 + w;                                     // word w of the 16-byte row belongs to quad lane w
                dst = (uint32_t)k * kWOff + (uint32_t)(((j >> 1) * 32 + ln) * 8 + (j & 1) * 4);
            } else {
                const int b = 4 * w, tl = b / BL, r = b % BL, s = r >> 3, h = (r >> 2) & 1;
                const int ln = (2 * tq + e) * 4 + tl;
                dst = (uint32_t)k * kWOff + (uint32_t)(((s * (NT / 2) + (j >> 1)) * 32 + ln) * 16 + (j & 1) * 8 + h * 4);
            }
            *reinterpret_cast<uint32_t*>(s_w + dst) = val;
        }
        const float act = p.act_scale_dev ? *p.act_scale_dev : 1.0f;
        for (int c = tid; c < COUT; c += kThreads) {
            s_scale[c] = p.scale[c] * act;
            s_shift[c] = p.shift[c];
            s_qscale[c] = p.out_qscale ? p.out_qscale[c] : 0.f;
            s_absmax[c] = 0u;
        }
    }
    __syncthreads();

    const int64_t n_out = p.n_out_dev ? min((int64_t)*p.n_out_dev, p.n_out_cap) : p.n_out_cap;
    const int64_t n_groups = (n_out + 15) >> 4;
    const uint8_t* const fbase = p.feats + t * BL;
    const uint8_t* const wlane = s_w + lane * (kHalf ? 8 : 16);
    const int cb = t * 2 * NT;                            // this lane's first output channel

    for (int64_t grp = (int64_t)blockIdx.x * (kThreads / 32) + warp; grp < n_groups; grp += (int64_t)gridDim.x * (kThreads / 32)) {
        const int64_t tile = grp >> 3;
        const int* slab = p.nbr + tile * (int64_t)p.kvol * QL_TILE_M + ((int)(grp & 7) * 16 + g);
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

        // live offsets of the tile, walked in ascending order: kk[] = kernel offset, slab ordinal = running count
        uint32_t mword = 0u;
        int mw = -1, ord = 0;
        auto next_offset = [&]() -> int {                    // next live kernel offset or -1
            while (mword == 0u) {
                if (++mw >= p.mask_words) return -1;
                if (p.kmask) mword = __ldg(p.kmask + tile * p.mask_words + mw);
                else { const int rem = p.kvol - 32 * mw; mword = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u); }
            }
            const int b = __ffs(mword) - 1;
            mword &= mword - 1u;
            return mw * 32 + b;
        };
        int kk[NB], idA[NB], idB[NB];
        auto fetch_ids = [&]() {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                kk[i] = next_offset();
                idA[i] = -1; idB[i] = -1;
                if (kk[i] >= 0) {
                    const int* s = slab + (int64_t)(p.kmask ? ord : kk[i]) * QL_TILE_M;
                    idA[i] = __ldg(s);
                    idB[i] = __ldg(s + 8);
                    ++ord;
                }
            }
        };
        fetch_ids();
        while (kk[0] >= 0) {
            int ck[NB];
            bool live[NB];
            uint32_t a[NB][2][AW];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                ck[i] = kk[i];
                live[i] = ck[i] >= 0 && __any_sync(0xffffffffu, (idA[i] & idB[i]) >= 0);      // some row of the 16 has this neighbour
                if (live[i]) {
                    gather_row<AW>(fbase, idA[i], RB, a[i][0]);
                    gather_row<AW>(fbase, idB[i], RB, a[i][1]);
                } else {
#pragma unroll
                    for (int w = 0; w < AW; ++w) { a[i][0][w] = 0u; a[i][1][w] = 0u; }
                }
            }
            fetch_ids();                                      // the next batch's indices travel under this batch's rows
            if constexpr (kHalf) {
#pragma unroll
                for (int i = 0; i < NB; i += 2) {
                    if (live[i] || live[i + 1]) {
                        const int k0 = live[i] ? ck[i] : ck[i + 1], k1 = live[i + 1] ? ck[i + 1] : ck[i];     // a dead half multiplies zeros
#pragma unroll
                        for (int jj = 0; jj < NT / 2; ++jj) {
                            const uint2 b0 = *reinterpret_cast<const uint2*>(wlane + (size_t)k0 * kWOff + jj * 256);
                            const uint2 b1 = *reinterpret_cast<const uint2*>(wlane + (size_t)k1 * kWOff + jj * 256);
                            mma_s8(acc[2 * jj], a[i][0][0], a[i][1][0], a[i + 1][0][0], a[i + 1][1][0], b0.x, b1.x);
                            mma_s8(acc[2 * jj + 1], a[i][0][0], a[i][1][0], a[i + 1][0][0], a[i + 1][1][0], b0.y, b1.y);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (live[i]) {
                        const uint8_t* wk = wlane + (size_t)ck[i] * kWOff;
#pragma unroll
                        for (int s = 0; s < KS; ++s) {
#pragma unroll
                            for (int jj = 0; jj < NT / 2; ++jj) {
                                const uint4 b = *reinterpret_cast<const uint4*>(wk + (s * (NT / 2) + jj) * 512);
                                if constexpr (kInt8) {
                                    mma_s8(acc[2 * jj], a[i][0][2 * s], a[i][1][2 * s], a[i][0][2 * s + 1], a[i][1][2 * s + 1], b.x, b.y);
                                    mma_s8(acc[2 * jj + 1], a[i][0][2 * s], a[i][1][2 * s], a[i][0][2 * s + 1], a[i][1][2 * s + 1], b.z, b.w);
                                } else {
                                    mma_f16(acc[2 * jj], a[i][0][2 * s], a[i][1][2 * s], a[i][0][2 * s + 1], a[i][1][2 * s + 1], b.x, b.y);
                                    mma_f16(acc[2 * jj + 1], a[i][0][2 * s], a[i][1][2 * s], a[i][0][2 * s + 1], a[i][1][2 * s + 1], b.z, b.w);
                                }
                            }
                        }
                    }
                }
            }
        }

        // ---- epilogue: lane (g, t) holds rows g and g + 8 of the group, channels cb .. cb + 2 NT - 1
        uint32_t amax[2 * NT];
#pragma unroll
        for (int c = 0; c < 2 * NT; ++c) amax[c] = 0u;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int64_t slot = grp * 16 + g + 8 * hh;
            int64_t row = slot;
            if (p.row_perm) row = slot < n_out ? (int64_t)__ldg(p.row_perm + slot) : -1;
            const bool row_ok = row >= 0 && row < n_out;
            if (!row_ok) continue;
            if (p.out_dtype == QL_S32) {
                uint32_t v[2 * NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) { v[2 * j] = __float_as_uint(acc[j][2 * hh]); v[2 * j + 1] = __float_as_uint(acc[j][2 * hh + 1]); }
                store_words<2 * NT>(reinterpret_cast<uint32_t*>(p.out) + row * COUT + cb, v);
                continue;
            }
            float y[2 * NT];
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float av = kInt8 ? (float)__float_as_int(acc[j][2 * hh + e]) : acc[j][2 * hh + e];
                    y[2 * j + e] = fmaf(av, s_scale[cb + 2 * j + e], s_shift[cb + 2 * j + e]);
                }
            if (p.residual) {
                uint32_t r[NT];
                load_words<NT>(p.residual + row * COUT + cb, r);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r[j]));
                    y[2 * j] += f.x; y[2 * j + 1] += f.y;
                }
            }
            if (p.relu) {
#pragma unroll
                for (int c = 0; c < 2 * NT; ++c) y[c] = fmaxf(y[c], 0.f);
            }
            if (p.out_dtype == QL_F16) {
                uint32_t h[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const __half2 hv = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                    h[j] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                store_words<NT>(reinterpret_cast<__half*>(p.out) + row * COUT + cb, h);
            } else {
                uint32_t v[2 * NT];
#pragma unroll
                for (int c = 0; c < 2 * NT; ++c) v[c] = __float_as_uint(y[c]);
                store_words<2 * NT>(reinterpret_cast<float*>(p.out) + row * COUT + cb, v);
            }
            if (p.out_q) {
                uint32_t q[NT / 2];
#pragma unroll
                for (int wq = 0; wq < NT / 2; ++wq) {
                    uint32_t word = 0u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float tq = rintf(y[4 * wq + c] * s_qscale[cb + 4 * wq + c]);
                        tq = fminf(fmaxf(tq, -127.f), 127.f);
                        word |= ((uint32_t)(uint8_t)(int8_t)(int)tq) << (8 * c);
                    }
                    q[wq] = word;
                }
                store_words<NT / 2>(p.out_q + row * COUT + cb, q);
            }
            if (p.absmax) {
#pragma unroll
                for (int c = 0; c < 2 * NT; ++c) amax[c] = max(amax[c], __float_as_uint(fabsf(y[c])));
            }
        }
        if (p.absmax && p.out_dtype != QL_S32) {
            // max over the 8 lanes that share t (their 16 rows), then one shared-memory atomic per channel
#pragma unroll
            for (int c = 0; c < 2 * NT; ++c) {
                uint32_t m = amax[c];
                m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
                m = max(m, __shfl_xor_sync(0xffffffffu, m, 8));
                m = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
                if (g == 0 && m) atomicMax(&s_absmax[cb + c], m);
            }
        }
    }

    if (p.absmax) {
        __syncthreads();
        for (int c = tid; c < COUT; c += kThreads) {
            const uint32_t v = s_absmax[c];
            if (v) atomicMax(reinterpret_cast<unsigned int*>(p.absmax) + c, v);
        }
    }
}

template <bool kInt8, int RB, int COUT>
cudaError_t launch_warp(const WarpConvParams& p, size_t smem_bytes, cudaStream_t st) {
    // two CTAs (32 warps, <= 64 registers) per SM when the weights fit twice and the accumulators leave room (c_out <= 32), else one
    const bool two = COUT <= 32 && 2 * (smem_bytes + 1024) <= 228 * 1024;
    int64_t groups = (p.n_out_cap + 15) / 16;
    int64_t want = (groups + kThreads / 32 - 1) / (kThreads / 32);
    cudaError_t e;
    if (two) {
        auto kern = k_spconv_warp<kInt8, RB, COUT, 2>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        const int64_t grid = want < 2 * ql_num_sms() ? want : 2 * ql_num_sms();
        kern<<<(int)grid, kThreads, smem_bytes, st>>>(p);
    } else {
        auto kern = k_spconv_warp<kInt8, RB, COUT, 1>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        const int64_t grid = want < ql_num_sms() ? want : ql_num_sms();
        kern<<<(int)grid, kThreads, smem_bytes, st>>>(p);
    }
    return cudaPeekAtLastError();
}

}  // namespace

// Launches the register-gather kernel when the layer qualifies.  Returns QL_OK / QL_ERR_CUDA when it did, 1 when the layer does
// not qualify (the caller then takes the tcgen05 kernel).
int ql_spconv_warp_try(const void* feats, int32_t in_dtype, const int32_t* nbr, const uint32_t* tile_kmask, const int32_t* row_perm,
                       int64_t n_out_cap, const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t kvol, const void* w_packed,
                       const float* scale, const float* shift, const float* act_scale_dev, const void* residual_f16, int32_t relu,
                       void* out, int32_t out_dtype, int8_t* out_q, const float* out_qscale, float* absmax, cudaStream_t st) {
    // OPT-IN (read per call so that one process can A/B): 0 / unset = never -- on the B200 the tcgen05 kernel is 2x faster on every
    // layer of the bench (profiles/r02_warp_kernel_ab.md) --, 1 = narrow layers, 2 = also 64-byte rows x 64 outputs
    const char* mode_s = getenv("QL_SPCONV_WARP");
    const int mode = (mode_s && *mode_s) ? atoi(mode_s) : 0;
    if (mode == 0) return 1;
    const bool i8 = in_dtype == QL_S8;
    const int rb = c_in * (i8 ? 1 : 2);
    if (!(rb == 32 || rb == 64 || (i8 && rb == 16))) return 1;
    if (!(c_out == 16 || c_out == 32 || c_out == 64)) return 1;
    if (((uintptr_t)feats & (uintptr_t)(rb / 4 - 1)) != 0) return 1;
    const size_t smem_bytes = (size_t)kvol * rb * c_out + 4 * (size_t)c_out * 4;
    if (smem_bytes > 220 * 1024) return 1;
    // wide outputs from wide rows keep the tcgen05 kernel unless asked for (QL_SPCONV_WARP=2): 64 x 64 is tensor work
    if (mode < 2 && rb * c_out > 64 * 32) return 1;

    WarpConvParams p;
    memset(&p, 0, sizeof(p));
    p.feats = (const uint8_t*)feats; p.nbr = nbr; p.kmask = tile_kmask; p.row_perm = row_perm; p.n_out_dev = n_out_dev; p.n_out_cap = n_out_cap;
    p.kvol = kvol; p.mask_words = (kvol + 31) / 32; p.ch = rb <= 32 ? 32 : (rb <= 64 ? 64 : 128);
    p.w_packed = (const uint8_t*)w_packed; p.scale = scale; p.shift = shift; p.act_scale_dev = act_scale_dev;
    p.residual = (const __half*)residual_f16; p.relu = relu; p.out = out; p.out_dtype = out_dtype;
    p.out_q = out_q; p.out_qscale = out_qscale; p.absmax = absmax;

    cudaError_t e = cudaErrorInvalidValue;
#define QL_WARP_CASE(I8, RBv, CO) if (i8 == I8 && rb == RBv && c_out == CO) e = launch_warp<I8, RBv, CO>(p, smem_bytes, st);
    QL_WARP_CASE(false, 32, 16) QL_WARP_CASE(false, 32, 32) QL_WARP_CASE(false, 32, 64)
    QL_WARP_CASE(false, 64, 16) QL_WARP_CASE(false, 64, 32) QL_WARP_CASE(false, 64, 64)
    QL_WARP_CASE(true, 16, 16) QL_WARP_CASE(true, 16, 32) QL_WARP_CASE(true, 16, 64)
    QL_WARP_CASE(true, 32, 16) QL_WARP_CASE(true, 32, 32) QL_WARP_CASE(true, 32, 64)
    QL_WARP_CASE(true, 64, 16) QL_WARP_CASE(true, 64, 32) QL_WARP_CASE(true, 64, 64)
#undef QL_WARP_CASE
    return e == cudaSuccess ? QL_OK : QL_ERR_CUDA;
}
