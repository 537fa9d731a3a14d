// Activation quantizer kernels ([EXT] pytorch_quantization TensorQuantizer as configured at quant/quant.py:14-32),
// the fp32 SIMT stem conv (un-quantized conv_input, quant_centerpoint.py:24-26) and library-level helpers.
#include "ql_common.cuh"
#include <stdio.h>

namespace {

__device__ __forceinline__ float load_as_float(const void* x, int dtype, int64_t i) {
    return dtype == QL_F16 ? __half2float(((const __half*)x)[i]) : ((const float*)x)[i];
}

// absmax[c] = max(absmax[c], max over rows |x[:, c]|); fp32 bit patterns of non-negative floats order like uints
__global__ void __launch_bounds__(256) k_absmax_cols(const void* x, int dtype, int64_t n_cap, const int* n_dev, int c,
                                                     float* absmax) {
    extern __shared__ uint32_t s_max[];
    const int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    for (int j = threadIdx.x; j < c; j += blockDim.x) s_max[j] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    if (256 % c == 0) {
        // thread <-> fixed channel: rows advance by 256/c per step, the block covers a contiguous row range
        const int ch = threadIdx.x % c;
        const int rows_per_step = 256 / c;
        const int64_t rows_per_block = (n + gridDim.x - 1) / gridDim.x;
        const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
        const int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
        float m = 0.f;
        for (int64_t r = r0 + threadIdx.x / c; r < r1; r += rows_per_step) m = fmaxf(m, fabsf(load_as_float(x, dtype, r * c + ch)));
        atomicMax(&s_max[ch], __float_as_uint(m));
    } else {
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
            for (int j = 0; j < c; ++j) {
                const int jj = (j + lane) % c;
                atomicMax(&s_max[jj], __float_as_uint(fabsf(load_as_float(x, dtype, r * c + jj))));
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c; j += blockDim.x)
        if (s_max[j]) atomicMax(reinterpret_cast<unsigned int*>(absmax) + j, s_max[j]);
}

// fp16 rows, c % 8 == 0, 256 % (c / 8) == 0: a thread owns one 8-channel group (one 16-byte load per row) and walks rows
__global__ void __launch_bounds__(256) k_absmax_cols_h8(const __half* __restrict__ x, int64_t n_cap, const int* __restrict__ n_dev, int c,
                                                        float* absmax) {
    extern __shared__ uint32_t s_max[];
    const int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    for (int j = threadIdx.x; j < c; j += blockDim.x) s_max[j] = 0u;
    __syncthreads();
    const int groups = c >> 3;
    const int g = threadIdx.x % groups;
    const int rows_per_step = 256 / groups;
    const int64_t stride = (int64_t)gridDim.x * rows_per_step;
    __half2 m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = __float2half2_rn(0.f);
    for (int64_t r = (int64_t)blockIdx.x * rows_per_step + threadIdx.x / groups; r < n; r += stride) {
        const uint4 v = *reinterpret_cast<const uint4*>(x + r * c + g * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], __habs2(h[j]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(m[j]);
        atomicMax(&s_max[g * 8 + 2 * j], __float_as_uint(f.x));
        atomicMax(&s_max[g * 8 + 2 * j + 1], __float_as_uint(f.y));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c; j += blockDim.x)
        if (s_max[j]) atomicMax(reinterpret_cast<unsigned int*>(absmax) + j, s_max[j]);
}

__device__ __forceinline__ float quant_scale_of(float amax, float bound) {
    // scale = bound / amax; amax <= 2^-24 -> 0 ([EXT] fake_tensor_quant epsilon rule)
    return amax <= (1.0f / 16777216.0f) ? 0.f : __fdiv_rn(bound, amax);
}

__device__ __forceinline__ float quant_code(float v, float scale, float bound) {
    float q = rintf(__fmul_rn(v, scale));              // round half to even, like torch.round
    return fminf(fmaxf(q, -bound), bound);
}

// one thread per 8 consecutive channels of a row (c % 8 == 0) or per element otherwise
__global__ void __launch_bounds__(256) k_quantize_rows(const void* x, int in_dtype, int64_t n_cap, const int* n_dev, int c,
                                                       const float* absmax, const float* smooth, float bound, int mode,
                                                       void* out, float* act_scale_out) {
    extern __shared__ float s_par[];                   // [c] scale (quantise), [c] inverse (de-quantise), [c] smooth
    float* s_scale = s_par;
    float* s_inv = s_par + c;
    float* s_smooth = s_par + 2 * c;
    const int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    __shared__ float s_tensor_amax;
    if (threadIdx.x == 0) {
        float m = 0.f;
        if (mode == QL_Q_CODES_PER_TENSOR || mode == QL_Q_FAKE_PER_TENSOR)
            for (int j = 0; j < c; ++j) m = fmaxf(m, smooth ? __fdiv_rn(absmax[j], smooth[j]) : absmax[j]);
        s_tensor_amax = m;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c; j += blockDim.x) {
        float am = (mode == QL_Q_FAKE_PER_CHANNEL) ? (smooth ? __fdiv_rn(absmax[j], smooth[j]) : absmax[j]) : s_tensor_amax;
        float sc = quant_scale_of(am, bound);
        s_scale[j] = sc;
        s_inv[j] = sc == 0.f ? 0.f : 1.0f;             // placeholder, de-quantisation divides by sc (see below)
        s_smooth[j] = smooth ? smooth[j] : 1.0f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && act_scale_out) act_scale_out[0] = __fdiv_rn(s_tensor_amax, bound);
    __syncthreads();
    const int64_t total = n * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        float v = load_as_float(x, in_dtype, i);
        if (smooth) v = __fdiv_rn(v, s_smooth[ch]);
        const float sc = s_scale[ch];
        const float q = quant_code(v, sc, bound);
        if (mode == QL_Q_CODES_PER_TENSOR) {
            ((int8_t*)out)[i] = (int8_t)(int)q;
        } else {
            ((__half*)out)[i] = __float2half_rn(sc == 0.f ? 0.f : __fdiv_rn(q, sc));
        }
    }
}

// fp16 rows, c % 8 == 0: one thread per 8 consecutive channels (16-byte load, 8- or 16-byte store); same arithmetic
__global__ void __launch_bounds__(256) k_quantize_rows_h8(const __half* __restrict__ x, int64_t n_cap, const int* __restrict__ n_dev, int c,
                                                          const float* absmax, const float* smooth, float bound, int mode,
                                                          void* out, float* act_scale_out) {
    extern __shared__ float s_par[];                   // [c] scale, [c] smooth
    float* s_scale = s_par;
    float* s_smooth = s_par + c;
    const int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    // per-tensor amax = max over the channels, taken by the whole block (one thread walking c dependent global loads cost every
    // CTA of the C = 128 launches ~10 us before its first row)
    __shared__ unsigned int s_tensor_amax_u;
    if (threadIdx.x == 0) s_tensor_amax_u = 0u;
    __syncthreads();
    {
        float m = 0.f;
        for (int j = threadIdx.x; j < c; j += blockDim.x) {
            const float am = smooth ? __fdiv_rn(absmax[j], smooth[j]) : absmax[j];
            s_scale[j] = am;                                   // parked: turned into the scale below
            s_smooth[j] = smooth ? smooth[j] : 1.0f;
            m = fmaxf(m, am);
        }
        const unsigned int mu = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(m, 0.f)));   // non-negative floats order like their bits
        if ((threadIdx.x & 31) == 0 && mu) atomicMax(&s_tensor_amax_u, mu);
    }
    __syncthreads();
    const float s_tensor_amax = (mode == QL_Q_CODES_PER_TENSOR || mode == QL_Q_FAKE_PER_TENSOR) ? __uint_as_float(s_tensor_amax_u) : 0.f;
    for (int j = threadIdx.x; j < c; j += blockDim.x)
        s_scale[j] = quant_scale_of(mode == QL_Q_FAKE_PER_CHANNEL ? s_scale[j] : s_tensor_amax, bound);
    if (blockIdx.x == 0 && threadIdx.x == 0 && act_scale_out) act_scale_out[0] = __fdiv_rn(s_tensor_amax, bound);
    __syncthreads();
    const int groups = c >> 3;
    const int64_t total = n * groups;
    const bool has_smooth = smooth != nullptr;
    const bool pow2 = (groups & (groups - 1)) == 0;
    // two 16-byte loads in flight per thread (one per half of the grid-stride step): with one, the resident threads of an SM hold
    // ~32 KB in flight, short of what the HBM latency needs
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * step) {
      const int64_t i1 = i0 + step;
      uint4 raw2[2];
      raw2[0] = *reinterpret_cast<const uint4*>(x + i0 * 8);
      raw2[1] = i1 < total ? *reinterpret_cast<const uint4*>(x + i1 * 8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = u == 0 ? i0 : i1;
        if (i >= total) break;
        const int ch0 = (pow2 ? (int)(i & (int64_t)(groups - 1)) : (int)(i % groups)) * 8;
        const uint4 raw = raw2[u];
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            v[2 * j] = f.x; v[2 * j + 1] = f.y;
        }
        float q[8], sc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (has_smooth) v[j] = __fdiv_rn(v[j], s_smooth[ch0 + j]);
            sc[j] = s_scale[ch0 + j];
            q[j] = quant_code(v[j], sc[j], bound);
        }
        if (mode == QL_Q_CODES_PER_TENSOR) {
            uint32_t w0 = 0, w1 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                w0 |= ((uint32_t)(uint8_t)(int8_t)(int)q[j]) << (8 * j);
                w1 |= ((uint32_t)(uint8_t)(int8_t)(int)q[4 + j]) << (8 * j);
            }
            *reinterpret_cast<uint2*>((int8_t*)out + i * 8) = make_uint2(w0, w1);
        } else {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = sc[2 * j] == 0.f ? 0.f : __fdiv_rn(q[2 * j], sc[2 * j]);
                const float b = sc[2 * j + 1] == 0.f ? 0.f : __fdiv_rn(q[2 * j + 1], sc[2 * j + 1]);
                const __half2 hh = __floats2half2_rn(a, b);
                o[j] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>((__half*)out + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
}

// per-row amax (GQConv3d, quant/quant_conv3d.py:112-131): one warp per row
__global__ void __launch_bounds__(256) k_fake_quant_per_row(const void* x, int in_dtype, int64_t n_cap, const int* n_dev, int c,
                                                            float bound, __half* out) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < n; r += n_warps) {
        float m = 0.f;
        for (int j = lane; j < c; j += 32) m = fmaxf(m, fabsf(load_as_float(x, in_dtype, r * c + j)));
        m = ql_warp_max(m);
        const float sc = quant_scale_of(m, bound);
        for (int j = lane; j < c; j += 32) {
            float q = quant_code(load_as_float(x, in_dtype, r * c + j), sc, bound);
            out[r * c + j] = __float2half_rn(sc == 0.f ? 0.f : __fdiv_rn(q, sc));
        }
    }
}

// fp32 SIMT stem conv.  A warp owns one 128-row rulebook tile and every thread FOUR of its rows (lane, lane+32, lane+64,
// lane+96), a CTA of 4 warps four tiles.  The first form of this kernel (one row per thread) was bound by the shared-memory
// pipe: every FMA group re-read its weights, 20 LDS.128 (80 scalar LDS) per kernel offset per warp against 80 FMAs.
// With 4 rows per thread a weight vector read once feeds 16 FMAs; an offset no row of the warp uses is skipped, a row that
// lacks it accumulates x = 0 (exact: fmaf(0, w, acc) == acc), so every row keeps the (k, ic) summation order of the
// reference loop.  kRow8: rows are 8 floats apart and 32-byte aligned (the engine's padded voxel features): one 256-bit
// load per neighbour instead of C_IN scalar loads.  C_IN == 0: generic width, scalar loads.
template <int C_OUT, int C_IN, bool kRow8>
__global__ void __launch_bounds__(QL_TILE_M) k_stem_conv(const float* __restrict__ feats, int fstride, int c_in, const int* __restrict__ nbr,
                                                         const uint32_t* __restrict__ kmask, int64_t n_cap, const int* __restrict__ n_dev, int kvol,
                                                         const float* __restrict__ w, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int relu, void* out, int out_dtype,
                                                         float* absmax) {
    constexpr int R = 4;
    constexpr int CI = C_IN > 0 ? C_IN : 16;           // register rows sized for the widest generic input
    extern __shared__ float s_w[];                     // [kvol][c_in][C_OUT] then uint32 absmax[C_OUT]
    uint32_t* s_absmax = reinterpret_cast<uint32_t*>(s_w + kvol * c_in * C_OUT);
    const int64_t n = n_dev ? min((int64_t)*n_dev, (int64_t)n_cap) : n_cap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * 4 + warp;
    if ((int64_t)blockIdx.x * 4 * QL_TILE_M >= n) return;             // whole CTA past the device-side row count
    for (int i = threadIdx.x; i < kvol * c_in * C_OUT; i += blockDim.x) s_w[i] = w[i];
    if (threadIdx.x < C_OUT) s_absmax[threadIdx.x] = 0u;
    __syncthreads();
    const bool tile_live = tile * QL_TILE_M < n;
    float acc[R][C_OUT];
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
        for (int c = 0; c < C_OUT; ++c) acc[j][c] = 0.f;
    if (tile_live) {
        // compact rulebook: the tile's live offsets (bits of its mask, ascending k) are slabs 0, 1, ...; no mask = every offset
        const int* nb = nbr + tile * (int64_t)kvol * QL_TILE_M + lane;
        const int mask_words = (kvol + 31) >> 5;
        uint32_t mw[QL_MASK_WORDS_MAX];
        int n_live = 0;
#pragma unroll
        for (int i = 0; i < QL_MASK_WORDS_MAX; ++i) {
            uint32_t w = 0u;
            if (i < mask_words) {
                const int rem = kvol - 32 * i;
                w = kmask ? __ldg(kmask + tile * mask_words + i) : (rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u));
            }
            mw[i] = w;
            n_live += __popc(w);
        }
        int idx_next[R];
#pragma unroll
        for (int j = 0; j < R; ++j) idx_next[j] = n_live > 0 ? __ldg(nb + j * 32) : -1;
        int wi = 0;
        uint32_t cur = mw[0];
        for (int s = 0; s < n_live; ++s) {
            while (cur == 0u) {
                ++wi;
                cur = 0u;
#pragma unroll
                for (int i = 1; i < QL_MASK_WORDS_MAX; ++i)
                    if (i == wi) cur = mw[i];
            }
            const int k = wi * 32 + __ffs((int)cur) - 1;
            cur &= cur - 1u;
            int idx[R];
            bool any = false;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                idx[j] = idx_next[j];
                any |= idx[j] >= 0;
            }
            if (s + 1 < n_live) {
#pragma unroll
                for (int j = 0; j < R; ++j) idx_next[j] = __ldg(nb + (s + 1) * QL_TILE_M + j * 32);
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            float x[R][kRow8 ? 8 : CI];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if constexpr (kRow8) {
                    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
                    if (idx[j] >= 0) {
                        asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                                     : "l"(feats + (int64_t)idx[j] * 8));
                    }
                    x[j][0] = lo.x; x[j][1] = lo.y; x[j][2] = lo.z; x[j][3] = lo.w;
                    x[j][4] = hi.x; x[j][5] = hi.y; x[j][6] = hi.z; x[j][7] = hi.w;
                } else {
#pragma unroll
                    for (int ic = 0; ic < CI; ++ic)
                        x[j][ic] = (idx[j] >= 0 && ic < c_in) ? __ldg(feats + (int64_t)idx[j] * fstride + ic) : 0.f;
                }
            }
            const float4* wk = reinterpret_cast<const float4*>(s_w + k * c_in * C_OUT);
#pragma unroll
            for (int ic = 0; ic < CI; ++ic) {
                if (C_IN == 0 && ic >= c_in) break;
#pragma unroll
                for (int c4 = 0; c4 < C_OUT / 4; ++c4) {
                    const float4 wv = wk[ic * (C_OUT / 4) + c4];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const float xv = x[j][ic];
                        acc[j][4 * c4] = fmaf(xv, wv.x, acc[j][4 * c4]);
                        acc[j][4 * c4 + 1] = fmaf(xv, wv.y, acc[j][4 * c4 + 1]);
                        acc[j][4 * c4 + 2] = fmaf(xv, wv.z, acc[j][4 * c4 + 2]);
                        acc[j][4 * c4 + 3] = fmaf(xv, wv.w, acc[j][4 * c4 + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t row = tile * QL_TILE_M + j * 32 + lane;
            const bool row_ok = row < n;
#pragma unroll
            for (int c = 0; c < C_OUT; ++c) {
                float y = fmaf(acc[j][c], scale[c], shift[c]);
                if (relu) y = fmaxf(y, 0.f);
                acc[j][c] = y;
            }
            if (row_ok) {
                if (out_dtype == QL_F16) {
                    __half2* o = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(out) + row * C_OUT);
#pragma unroll
                    for (int c = 0; c < C_OUT / 2; ++c) o[c] = __floats2half2_rn(acc[j][2 * c], acc[j][2 * c + 1]);
                } else {
                    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * C_OUT);
#pragma unroll
                    for (int c = 0; c < C_OUT / 4; ++c) o[c] = make_float4(acc[j][4 * c], acc[j][4 * c + 1], acc[j][4 * c + 2], acc[j][4 * c + 3]);
                }
            }
            if (absmax) {
#pragma unroll
                for (int c = 0; c < C_OUT; ++c) {
                    const uint32_t m = __reduce_max_sync(0xffffffffu, row_ok ? __float_as_uint(fabsf(acc[j][c])) : 0u);
                    if (lane == (c & 31)) atomicMax(&s_absmax[c], m);
                }
            }
        }
    }
    if (absmax) {
        __syncthreads();
        if (threadIdx.x < C_OUT && s_absmax[threadIdx.x])
            atomicMax(reinterpret_cast<unsigned int*>(absmax) + threadIdx.x, s_absmax[threadIdx.x]);
    }
}

thread_local char g_last_cuda_error[256] = "";

}  // namespace

extern "C" int ql_abi_version(void) { return 3; }

extern "C" const char* ql_error_string(int code) {
    switch (code) {
        case QL_OK: return "ok";
        case QL_ERR_INVALID: return "invalid argument";
        case QL_ERR_CUDA: return "CUDA runtime error";
        case QL_ERR_GRID_TOO_LARGE: return "B*D*H*W does not fit the 32-bit coordinate key";
        case QL_ERR_WORKSPACE: return "workspace too small";
        case QL_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}

extern "C" const char* ql_last_cuda_error(void) {
    cudaError_t e = cudaGetLastError();
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(e));
    return g_last_cuda_error;
}

extern "C" int ql_num_sms_on_device(void) { return ql_num_sms(); }

extern "C" int ql_absmax_cols(const void* x, int32_t dtype, int64_t n_cap, const int32_t* n_dev, int32_t c, float* absmax,
                              ql_stream_t stream_) {
    if (!x || !absmax || c <= 0 || c > 4096 || (dtype != QL_F16 && dtype != QL_F32)) return QL_ERR_INVALID;
    if (n_cap <= 0) return QL_OK;
    int64_t blocks = (n_cap * c + 256 * 16 - 1) / (256 * 16);
    int grid = (int)(blocks < 1 ? 1 : (blocks > 4 * ql_num_sms() ? 4 * ql_num_sms() : blocks));
    if (dtype == QL_F16 && c % 8 == 0 && 256 % (c / 8) == 0 && ((uintptr_t)x & 15) == 0) {
        const int rows_per_step = 256 / (c / 8);
        int64_t want = (n_cap + (int64_t)rows_per_step * 8 - 1) / ((int64_t)rows_per_step * 8);
        grid = (int)(want < 1 ? 1 : (want > 8 * ql_num_sms() ? 8 * ql_num_sms() : want));
        k_absmax_cols_h8<<<grid, 256, (size_t)c * 4, (cudaStream_t)stream_>>>((const __half*)x, n_cap, n_dev, c, absmax);
    } else {
        k_absmax_cols<<<grid, 256, (size_t)c * 4, (cudaStream_t)stream_>>>(x, dtype, n_cap, n_dev, c, absmax);
    }
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_quantize_rows(const void* x, int32_t in_dtype, int64_t n_cap, const int32_t* n_dev, int32_t c,
                                const float* absmax, const float* smooth, int32_t bits, int32_t mode, void* out,
                                float* act_scale_out, ql_stream_t stream_) {
    if (!x || !out || c <= 0 || c > 4096 || bits < 2 || bits > 16 || (in_dtype != QL_F16 && in_dtype != QL_F32))
        return QL_ERR_INVALID;
    if (mode < QL_Q_CODES_PER_TENSOR || mode > QL_Q_FAKE_PER_ROW) return QL_ERR_INVALID;
    if (mode != QL_Q_FAKE_PER_ROW && !absmax) return QL_ERR_INVALID;
    if (mode == QL_Q_CODES_PER_TENSOR && bits > 8) return QL_ERR_INVALID;
    if (n_cap <= 0) return QL_OK;
    const float bound = (float)((1 << (bits - 1)) - 1);
    int64_t blocks = (n_cap * c + 256 * 8 - 1) / (256 * 8);
    int grid = (int)(blocks < 1 ? 1 : (blocks > 8 * ql_num_sms() ? 8 * ql_num_sms() : blocks));
    if (mode == QL_Q_FAKE_PER_ROW) {
        k_fake_quant_per_row<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, in_dtype, n_cap, n_dev, c, bound, (__half*)out);
    } else if (in_dtype == QL_F16 && c % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0) {
        int64_t want = (n_cap * (c / 8) + 256 * 4 - 1) / (256 * 4);
        grid = (int)(want < 1 ? 1 : (want > 16 * ql_num_sms() ? 16 * ql_num_sms() : want));
        k_quantize_rows_h8<<<grid, 256, (size_t)c * 8, (cudaStream_t)stream_>>>((const __half*)x, n_cap, n_dev, c, absmax, smooth, bound,
                                                                               mode, out, act_scale_out);
    } else {
        k_quantize_rows<<<grid, 256, (size_t)c * 12, (cudaStream_t)stream_>>>(x, in_dtype, n_cap, n_dev, c, absmax, smooth, bound,
                                                                            mode, out, act_scale_out);
    }
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_stem_conv(const float* feats, int32_t feat_stride, int32_t c_in, const int32_t* nbr, const uint32_t* tile_kmask, int64_t n_out_cap, const int32_t* n_out_dev,
                            int32_t c_out, int32_t kvol, const float* w, const float* scale, const float* shift, int32_t relu,
                            void* out, int32_t out_dtype, float* absmax, ql_stream_t stream_) {
    if (!feats || !nbr || !w || !scale || !shift || !out) return QL_ERR_INVALID;
    if (c_in <= 0 || c_in > 16 || feat_stride < c_in || kvol <= 0 || kvol > 343 || (out_dtype != QL_F16 && out_dtype != QL_F32))
        return QL_ERR_INVALID;
    const bool row8 = feat_stride == 8 && ((uintptr_t)feats & 31) == 0;
    if (c_out != 16 && c_out != 32) return QL_ERR_UNSUPPORTED;
    if (n_out_cap <= 0) return QL_OK;
    unsigned tiles = (unsigned)((n_out_cap + 4 * QL_TILE_M - 1) / (4 * QL_TILE_M));       // 4 rulebook tiles per CTA
    size_t smem = (size_t)kvol * c_in * c_out * 4 + (size_t)c_out * 4;
    cudaStream_t st = (cudaStream_t)stream_;
#define QL_STEM_LAUNCH(CO, CI, R8)                                                                                                 \
    do {                                                                                                                           \
        if (smem > 48 * 1024 &&                                                                                                    \
            cudaFuncSetAttribute(k_stem_conv<CO, CI, R8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)  \
            return QL_ERR_CUDA;                                                                                                    \
        k_stem_conv<CO, CI, R8><<<tiles, QL_TILE_M, smem, st>>>(feats, feat_stride, c_in, nbr, tile_kmask, n_out_cap, n_out_dev, kvol, w, scale, \
                                                                shift, relu, out, out_dtype, absmax);                              \
    } while (0)
#define QL_STEM_CI(CO, CI)                     \
    do {                                       \
        if (row8) QL_STEM_LAUNCH(CO, CI, true); \
        else QL_STEM_LAUNCH(CO, CI, false);    \
    } while (0)
    if (c_out == 16) {
        if (c_in == 5) QL_STEM_CI(16, 5);
        else if (c_in == 4) QL_STEM_CI(16, 4);
        else QL_STEM_LAUNCH(16, 0, false);
    } else {
        if (c_in == 5) QL_STEM_CI(32, 5);
        else if (c_in == 4) QL_STEM_CI(32, 4);
        else QL_STEM_LAUNCH(32, 0, false);
    }
#undef QL_STEM_CI
#undef QL_STEM_LAUNCH
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}


// ------------------------------------------------------------------------------------------------
// Row permutation: out[r] = in[src_row[r]] for r < n (rows of row_bytes = 16 * m bytes; src_row < 0 -> zeros).  Used by the
// engine to bring the voxeliser's first-touch-ordered rows into ascending-key order (src_row = the 1x1x1 rulebook of the
// renumbering build), after which stage 1 is indexed by rank like every later stage.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_permute_rows(const uint4* __restrict__ in, uint4* __restrict__ out, int chunks,
                                                      const int* __restrict__ src_row, int64_t n_cap, const int* __restrict__ n_dev) {
    const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / chunks;
    if (r >= n) return;
    const int c = (int)(t - r * chunks);
    const int s = __ldg(src_row + r);
    out[r * chunks + c] = s >= 0 ? __ldg(in + (int64_t)s * chunks + c) : make_uint4(0u, 0u, 0u, 0u);
}
}  // namespace

extern "C" int ql_permute_rows(const void* in, void* out, int32_t row_bytes, const int32_t* src_row, int64_t n_cap,
                               const int32_t* n_dev, ql_stream_t stream_) {
    if (!in || !out || !src_row || row_bytes <= 0 || row_bytes % 16 != 0 || n_cap < 0 || in == out) return QL_ERR_INVALID;
    if (((uintptr_t)in | (uintptr_t)out) & 15) return QL_ERR_INVALID;
    if (n_cap == 0) return QL_OK;
    const int chunks = row_bytes / 16;
    const int64_t threads = n_cap * chunks;
    k_permute_rows<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream_>>>((const uint4*)in, (uint4*)out, chunks, src_row,
                                                                                         n_cap, n_dev);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
