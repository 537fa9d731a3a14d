// Implicit gather-GEMM-scatter sparse convolution on tcgen05 tensor cores (sm_100a), fused epilogue.
//
// Replaces QConvNd.forward -> [EXT] spconv SubMConv3d/SparseConv3d forward (quant/quant.py:36-58) together with
// the BatchNorm1d / ReLU / residual-add that follow it (spconv_backbone.py:8-27,51-67).
//
// One persistent CTA per SM walks 128-row output tiles.  Per tile the conv is ONE GEMM
//     D[128, C_out] = A[128, K*C_in] . W[C_out, K*C_in]^T ,   K = kernel volume
// whose A operand never exists in memory: the K dimension is a flat byte string per row (offset-major,
// channel-minor), cut into 128-byte pipeline stages.  Roles (448 threads):
//   warps 0-7  gather producers : cp.async 16 B chunks feats[nbr[k][row]] -> SWIZZLE_128B K-major smem
//                                 (zero fill for missing neighbours), cp.async.mbarrier.arrive.noinc on the stage barrier
//   warps 8-11 epilogue         : tcgen05.ld accumulators -> dequant*BN scale/shift (+residual) (+ReLU)
//                                 -> fp16/fp32 rows (+ int8 re-quantised rows, + per-channel absmax)
//   warp  12   MMA issuer       : one lane issues tcgen05.mma (kind::f16 or kind::i8), accumulators in TMEM
//                                 (double buffered: tile i+1 accumulates while tile i drains)
//   warp  13   TMA loader       : cp.async.bulk of the tile's rulebook slab and of each stage's weight slab
//                                 (pre-swizzled image from ql_pack_weights_host) onto the stage's mbarrier
#include "ql_common.cuh"
#include <string.h>

namespace {

constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;   // 256
constexpr int kEpilogueThreads = 128;
constexpr int kEpilogueWarp0 = kProducerWarps;            // warps 8..11 (warp % 4 == TMEM lane quarter)
constexpr int kMmaWarp = kEpilogueWarp0 + 4;              // 12
constexpr int kLoaderWarp = kMmaWarp + 1;                 // 13
constexpr int kThreadsTotal = (kLoaderWarp + 1) * 32;     // 448
constexpr int kChunksPerThread = QL_TILE_M * 8 / kProducerThreads;   // 4 x 16-byte chunks per stage
constexpr int kStageABytes = QL_TILE_M * 128;      // 16 KB
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 232448;                // 227 KB opt-in maximum per CTA

struct ConvParams {
    const uint8_t* feats;
    const int* nbr;
    const int* n_out_dev;
    int64_t n_out_cap;
    int row_bytes;          // c_in * elem size
    int c_out, kvol;
    int n_kstages;          // ceil(kvol*row_bytes / 128)
    int last_ksteps;        // 32-byte MMA k-steps in the last stage
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
    int n_stages;           // pipeline depth
    int lag;                // producer arrive lag (cp.async groups in flight)
    int tmem_cols;          // allocated TMEM columns (power of two >= 2*c_out)
    // smem offsets from the 1024-aligned base
    int off_b, off_nbr, off_misc;
};

struct MiscSmem {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t nbr_full[2];
    uint64_t nbr_empty[2];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    // followed by: float scale[c_out], float shift[c_out], uint32 absmax[c_out], float qscale[c_out]
};

template <bool kInt8>
__device__ __forceinline__ uint32_t make_idesc(int n) {
    uint32_t d = 0;
    if (kInt8) {
        d |= 2u << 4;        // D format S32
        d |= 1u << 7;        // A signed int8
        d |= 1u << 10;       // B signed int8
    } else {
        d |= 1u << 4;        // D format F32;  A,B formats 0 = F16
    }
    // a_major = b_major = 0 (K-major), no negate, dense
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(QL_TILE_M >> 4) << 24;
    return d;
}

template <bool kInt8>
__global__ void __launch_bounds__(kThreadsTotal, 1) k_spconv_mma(const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte alignment
    const uint32_t smem_base_u32 = (ql_smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base_u32 - ql_smem_u32(smem_raw));
    MiscSmem* misc = reinterpret_cast<MiscSmem*>(smem + p.off_misc);
    float* s_scale = reinterpret_cast<float*>(misc + 1);
    float* s_shift = s_scale + p.c_out;
    uint32_t* s_absmax = reinterpret_cast<uint32_t*>(s_shift + p.c_out);
    float* s_qscale = reinterpret_cast<float*>(s_absmax + p.c_out);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int S = p.n_stages;

    const int64_t n_out = p.n_out_dev ? (int64_t)*p.n_out_dev : p.n_out_cap;
    const int64_t n_tiles = (n_out + QL_TILE_M - 1) / QL_TILE_M;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            ql_mbar_init(ql_smem_u32(&misc->full[s]), kProducerThreads + 1);
            ql_mbar_init(ql_smem_u32(&misc->empty[s]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ql_mbar_init(ql_smem_u32(&misc->nbr_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->nbr_empty[i]), kProducerThreads);
            ql_mbar_init(ql_smem_u32(&misc->acc_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->acc_empty[i]), kEpilogueThreads);
        }
        ql_fence_mbar_init();
    }
    {
        const float act = p.act_scale_dev ? *p.act_scale_dev : 1.0f;
        for (int c = tid; c < p.c_out; c += kThreadsTotal) {
            s_scale[c] = p.scale[c] * act;
            s_shift[c] = p.shift[c];
            s_absmax[c] = 0u;
            s_qscale[c] = p.out_qscale ? p.out_qscale[c] : 0.f;
        }
    }
    if (warp == kMmaWarp) {
        ql_tmem_alloc(ql_smem_u32(&misc->tmem_base), (uint32_t)p.tmem_cols);
        ql_tmem_relinquish();
    }
    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    const uint32_t tmem_base = misc->tmem_base;

    const uint32_t a_base = smem_base_u32;
    const uint32_t b_base = smem_base_u32 + p.off_b;
    const uint32_t b_stage_bytes = (uint32_t)p.c_out * 128u;
    const uint32_t nbr_bytes = (uint32_t)p.kvol * QL_TILE_M * 4u;

    if (warp < kProducerWarps) {
        // ============================ gather producers ============================
        const int c16 = tid & 7;
        const int rsub = tid >> 3;                           // rows rsub + 32*i
        const uint32_t dst_thread = ql_sw128_offset((uint32_t)rsub, (uint32_t)c16);   // + i*4096 for row rsub+32i
        uint32_t s = 0, ph = 0;                              // ring slot and its phase
        uint32_t it = 0;
        const uint32_t nbr_base_u32 = smem_base_u32 + (uint32_t)p.off_nbr + (uint32_t)rsub * 4u;
        const int64_t row_bytes = p.row_bytes;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int nb = it & 1;
            ql_mbar_wait(ql_smem_u32(&misc->nbr_full[nb]), (it >> 1) & 1);
            const uint32_t nbrs = nbr_base_u32 + (uint32_t)nb * nbr_bytes;
            // (offset, channel byte) of this thread's 16-byte chunk, advanced by 128 bytes of K per stage
            int koff = 0, ch = c16 * 16;
            while (ch >= p.row_bytes) { ch -= p.row_bytes; ++koff; }
            for (int ks = 0; ks < p.n_kstages; ++ks) {
                const bool kvalid = koff < p.kvol;
                // all neighbour indices first (independent LDS), then the copies
                int idx[kChunksPerThread];
                const uint32_t nrow = nbrs + (uint32_t)(kvalid ? koff : 0) * (QL_TILE_M * 4u);
#pragma unroll
                for (int i = 0; i < kChunksPerThread; ++i) idx[i] = ql_lds_s32(nrow + (uint32_t)i * 128u);
                ql_mbar_wait(ql_smem_u32(&misc->empty[s]), ph ^ 1u);
                const uint32_t dst0 = a_base + s * kStageABytes + dst_thread;
                const uint8_t* src0 = p.feats + ch;
#pragma unroll
                for (int i = 0; i < kChunksPerThread; ++i) {
                    const bool valid = kvalid && idx[i] >= 0;
                    ql_cp_async16(dst0 + (uint32_t)i * 4096u, src0 + (valid ? idx[i] : 0) * row_bytes, valid);
                }
                // completion of this thread's copies arrives on the stage barrier asynchronously: the producer never
                // blocks on memory latency, so up to n_stages gathers (n_stages * 16 KB) are in flight per SM
                ql_cp_async_mbar_arrive_noinc(ql_smem_u32(&misc->full[s]));
                ch += 128;
                while (ch >= p.row_bytes) { ch -= p.row_bytes; ++koff; }
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
            ql_mbar_arrive(ql_smem_u32(&misc->nbr_empty[nb]));
        }
    } else if (warp < kMmaWarp) {
        // ================================ epilogue ================================
        const int w = warp - kEpilogueWarp0;                              // TMEM lane quarter (warp id % 4)
        const int et = tid - kProducerThreads;               // 0..127 == row in tile
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            ql_mbar_wait(ql_smem_u32(&misc->acc_full[a]), (it >> 1) & 1);
            ql_tc_fence_after();
            const int64_t row = tile * QL_TILE_M + et;
            const bool row_ok = row < n_out;
            const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(a * p.c_out);
            for (int c0 = 0; c0 < p.c_out; c0 += 16) {
                uint32_t v[16];
                ql_tmem_ld16(taddr + (uint32_t)c0, v);
                ql_tmem_ld_wait();
                if (p.out_dtype == QL_S32) {
                    if (row_ok) {
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(p.out) + row * p.c_out + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) o[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                    continue;
                }
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float acc = kInt8 ? (float)(int)v[j] : __uint_as_float(v[j]);
                    y[j] = fmaf(acc, s_scale[c0 + j], s_shift[c0 + j]);
                }
                if (p.residual && row_ok) {
                    const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + row * p.c_out + c0);
                    uint4 ra = r4[0], rb = r4[1];
                    const __half2* h = reinterpret_cast<const __half2*>(&ra);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = __half22float2(h[j]);
                        y[2 * j] += f.x; y[2 * j + 1] += f.y;
                    }
                    h = reinterpret_cast<const __half2*>(&rb);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = __half22float2(h[j]);
                        y[8 + 2 * j] += f.x; y[8 + 2 * j + 1] += f.y;
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
                }
                if (row_ok) {
                    if (p.out_dtype == QL_F16) {
                        uint32_t h[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            __half2 hh = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                            h[j] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + row * p.c_out + c0);
                        o[0] = make_uint4(h[0], h[1], h[2], h[3]);
                        o[1] = make_uint4(h[4], h[5], h[6], h[7]);
                    } else {
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.c_out + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) o[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
                    }
                    if (p.out_q) {
                        uint32_t qq[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t word = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float t = rintf(y[4 * q + j] * s_qscale[c0 + 4 * q + j]);
                                t = fminf(fmaxf(t, -127.f), 127.f);
                                word |= ((uint32_t)(uint8_t)(int8_t)(int)t) << (8 * j);
                            }
                            qq[q] = word;
                        }
                        *reinterpret_cast<uint4*>(p.out_q + row * p.c_out + c0) = make_uint4(qq[0], qq[1], qq[2], qq[3]);
                    }
                }
                if (p.absmax) {
                    // warp-wide max per channel (redux.sync), then one shared-memory atomic per channel
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t m = __reduce_max_sync(0xffffffffu, row_ok ? __float_as_uint(fabsf(y[j])) : 0u);
                        if (lane == j) atomicMax(&s_absmax[c0 + j], m);
                    }
                }
            }
            ql_tc_fence_before();
            ql_mbar_arrive(ql_smem_u32(&misc->acc_empty[a]));
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer ===============================
        // The whole warp walks the loop (warp-uniform control flow, so the uniform-datapath tcgen05 instructions need no
        // per-lane election loops); one elected lane issues the MMAs and the commits.
        const bool leader = ql_elect_one();
        const uint32_t idesc = make_idesc<kInt8>(p.c_out);
        const uint64_t adesc0 = ql_umma_desc_sw128(a_base);
        const uint64_t bdesc0 = ql_umma_desc_sw128(b_base);
        const uint32_t b_step16 = b_stage_bytes >> 4;
        uint32_t s = 0, ph = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            ql_mbar_wait(ql_smem_u32(&misc->acc_empty[a]), ((it >> 1) & 1) ^ 1u);
            ql_tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(a * p.c_out);
            for (int ks = 0; ks < p.n_kstages; ++ks) {
                ql_mbar_wait(ql_smem_u32(&misc->full[s]), ph);
                ql_fence_proxy_async();              // cp.async (generic proxy) writes -> tcgen05.mma (async proxy) reads
                ql_tc_fence_after();
                if (leader) {
                    const uint64_t adesc = adesc0 + (uint64_t)(s * (kStageABytes >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t)(s * b_step16);
                    const int nk = (ks == p.n_kstages - 1) ? p.last_ksteps : 4;
                    // advance 32 bytes along K inside the 128-byte swizzle span: +2 in the (addr >> 4) field
                    ql_tc_mma<kInt8>(d_tmem, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
                    if (nk > 1) ql_tc_mma<kInt8>(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                    if (nk > 2) ql_tc_mma<kInt8>(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                    if (nk > 3) ql_tc_mma<kInt8>(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                    ql_tc_commit(ql_smem_u32(&misc->empty[s]));
                    if (ks == p.n_kstages - 1) ql_tc_commit(ql_smem_u32(&misc->acc_full[a]));
                }
                __syncwarp();
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // =============================== TMA loader ===============================
        const bool leader = ql_elect_one();
        uint32_t s = 0, ph = 0, it = 0;
        if (leader && (int64_t)blockIdx.x < n_tiles) {
            ql_mbar_arrive_expect_tx(ql_smem_u32(&misc->nbr_full[0]), nbr_bytes);
            ql_bulk_g2s(smem_base_u32 + (uint32_t)p.off_nbr, p.nbr + (int64_t)blockIdx.x * p.kvol * QL_TILE_M, nbr_bytes,
                        ql_smem_u32(&misc->nbr_full[0]));
        }
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int64_t next = tile + gridDim.x;
            if (next < n_tiles) {
                const uint32_t itn = it + 1;
                const int nb = itn & 1;
                ql_mbar_wait(ql_smem_u32(&misc->nbr_empty[nb]), ((itn >> 1) & 1) ^ 1u);
                if (leader) {
                    ql_mbar_arrive_expect_tx(ql_smem_u32(&misc->nbr_full[nb]), nbr_bytes);
                    ql_bulk_g2s(smem_base_u32 + (uint32_t)p.off_nbr + (uint32_t)nb * nbr_bytes, p.nbr + next * p.kvol * QL_TILE_M,
                                nbr_bytes, ql_smem_u32(&misc->nbr_full[nb]));
                }
            }
            for (int ks = 0; ks < p.n_kstages; ++ks) {
                ql_mbar_wait(ql_smem_u32(&misc->empty[s]), ph ^ 1u);
                if (leader) {
                    ql_mbar_arrive_expect_tx(ql_smem_u32(&misc->full[s]), b_stage_bytes);
                    ql_bulk_g2s(b_base + s * b_stage_bytes, p.w_packed + (int64_t)ks * b_stage_bytes, b_stage_bytes,
                                ql_smem_u32(&misc->full[s]));
                }
                __syncwarp();
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        }
    }

    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    if (p.absmax) {
        for (int c = tid; c < p.c_out; c += kThreadsTotal) {
            uint32_t v = s_absmax[c];
            if (v) atomicMax(reinterpret_cast<unsigned int*>(p.absmax) + c, v);
        }
    }
    if (warp == kMmaWarp) ql_tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

inline int elem_size(int dtype) { return dtype == QL_S8 ? 1 : (dtype == QL_F16 ? 2 : 0); }

}  // namespace

extern "C" size_t ql_packed_weight_bytes(int32_t c_in, int32_t c_out, int32_t kvol, int32_t elem_dtype) {
    int es = elem_size(elem_dtype);
    if (es == 0 || c_in <= 0 || c_out <= 0 || kvol <= 0) return 0;
    size_t kbytes = (size_t)kvol * c_in * es;
    size_t stages = (kbytes + 127) / 128;
    return stages * (size_t)c_out * 128;
}

// w_host: [c_out][kvol][c_in] elements (== the reference layout (oc, kd, kh, kw, ic) flattened, quant/quant.py:37-39).
// packed: per 128-byte K stage one [c_out x 128 B] K-major SWIZZLE_128B image, zero padded -- exactly what the
// loader warp bulk-copies into shared memory.
extern "C" int ql_pack_weights_host(const void* w_host, int32_t elem_dtype, int32_t c_in, int32_t c_out, int32_t kvol,
                                    void* packed_host) {
    int es = elem_size(elem_dtype);
    if (!w_host || !packed_host || es == 0 || c_in <= 0 || c_out <= 0 || kvol <= 0) return QL_ERR_INVALID;
    if ((c_in * es) % 16 != 0 || c_out % 16 != 0 || c_out > 256) return QL_ERR_UNSUPPORTED;
    size_t kbytes = (size_t)kvol * c_in * es;
    size_t total = ql_packed_weight_bytes(c_in, c_out, kvol, elem_dtype);
    memset(packed_host, 0, total);
    const uint8_t* src = (const uint8_t*)w_host;
    uint8_t* dst = (uint8_t*)packed_host;
    for (int oc = 0; oc < c_out; ++oc) {
        for (size_t kb = 0; kb < kbytes; kb += 16) {
            size_t stage = kb / 128;
            uint32_t c16 = (uint32_t)((kb % 128) / 16);
            memcpy(dst + stage * (size_t)c_out * 128 + ql_sw128_offset((uint32_t)oc, c16), src + (size_t)oc * kbytes + kb, 16);
        }
    }
    return QL_OK;
}

extern "C" int ql_spconv_mma(const void* feats, int32_t in_dtype, const int32_t* nbr, int64_t n_out_cap,
                             const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t kvol, const void* w_packed,
                             const float* scale, const float* shift, const float* act_scale_dev, const void* residual_f16,
                             int32_t relu, void* out, int32_t out_dtype, int8_t* out_q, const float* out_qscale,
                             float* absmax, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (!feats || !nbr || !w_packed || !scale || !shift || !out) return QL_ERR_INVALID;
    int es = elem_size(in_dtype);
    if (es == 0) return QL_ERR_INVALID;
    if (out_dtype != QL_F16 && out_dtype != QL_F32 && out_dtype != QL_S32) return QL_ERR_INVALID;
    if (out_q && !out_qscale) return QL_ERR_INVALID;
    if (c_in <= 0 || (c_in * es) % 16 != 0 || c_out < 16 || c_out % 16 != 0 || c_out > 256 || kvol <= 0 || kvol > 343)
        return QL_ERR_UNSUPPORTED;
    if (n_out_cap <= 0) return QL_OK;

    ConvParams p;
    p.feats = (const uint8_t*)feats; p.nbr = nbr; p.n_out_dev = n_out_dev; p.n_out_cap = n_out_cap;
    p.row_bytes = c_in * es; p.c_out = c_out; p.kvol = kvol;
    int kbytes = kvol * p.row_bytes;
    p.n_kstages = (kbytes + 127) / 128;
    int last_bytes = kbytes - (p.n_kstages - 1) * 128;
    p.last_ksteps = (last_bytes + 31) / 32;
    p.w_packed = (const uint8_t*)w_packed; p.scale = scale; p.shift = shift; p.act_scale_dev = act_scale_dev;
    p.residual = (const __half*)residual_f16; p.relu = relu; p.out = out; p.out_dtype = out_dtype;
    p.out_q = out_q; p.out_qscale = out_qscale; p.absmax = absmax;

    int tm = 32;
    while (tm < 2 * c_out) tm <<= 1;
    p.tmem_cols = tm;
    const int nbr_bytes = 2 * kvol * QL_TILE_M * 4;
    const int misc_bytes = (int)sizeof(MiscSmem) + 4 * c_out * 4;
    const int stage_bytes = kStageABytes + c_out * 128;
    int avail = kSmemBudget - 1024 /*alignment slack*/ - nbr_bytes - ((misc_bytes + 127) & ~127);
    int S = avail / stage_bytes;
    if (S > kMaxStages) S = kMaxStages;
    if (S < 2) return QL_ERR_UNSUPPORTED;
    p.n_stages = S;
    p.lag = S - 1 < 3 ? S - 1 : 3;
    if (p.lag < 1) p.lag = 1;
    p.off_b = S * kStageABytes;
    p.off_nbr = p.off_b + S * c_out * 128;
    p.off_misc = (p.off_nbr + nbr_bytes + 127) & ~127;
    size_t smem_bytes = 1024 + (size_t)p.off_misc + misc_bytes;

    int64_t tiles = (n_out_cap + QL_TILE_M - 1) / QL_TILE_M;
    int grid = (int)(tiles < ql_num_sms() ? tiles : ql_num_sms());
    cudaError_t e;
    if (in_dtype == QL_S8) {
        e = cudaFuncSetAttribute(k_spconv_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return QL_ERR_CUDA;
        k_spconv_mma<true><<<grid, kThreadsTotal, smem_bytes, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_spconv_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return QL_ERR_CUDA;
        k_spconv_mma<false><<<grid, kThreadsTotal, smem_bytes, st>>>(p);
    }
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
