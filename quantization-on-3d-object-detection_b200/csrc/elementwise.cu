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
    const int64_t n = n_dev ? (int64_t)*n_dev : n_cap;
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
    const int64_t n = n_dev ? (int64_t)*n_dev : n_cap;
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

// per-row amax (GQConv3d, quant/quant_conv3d.py:112-131): one warp per row
__global__ void __launch_bounds__(256) k_fake_quant_per_row(const void* x, int in_dtype, int64_t n_cap, const int* n_dev, int c,
                                                            float bound, __half* out) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_cap;
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

// fp32 SIMT stem conv: one thread per output row, C_OUT accumulators in registers, weights in smem.
template <int C_OUT>
__global__ void __launch_bounds__(QL_TILE_M) k_stem_conv(const float* __restrict__ feats, int c_in, const int* __restrict__ nbr,
                                                         int64_t n_cap, const int* __restrict__ n_dev, int kvol,
                                                         const float* __restrict__ w, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int relu, void* out, int out_dtype,
                                                         float* absmax) {
    extern __shared__ float s_w[];                     // [kvol][c_in][C_OUT] then uint32 absmax[C_OUT]
    uint32_t* s_absmax = reinterpret_cast<uint32_t*>(s_w + kvol * c_in * C_OUT);
    const int64_t n = n_dev ? (int64_t)*n_dev : n_cap;
    const int64_t tile = blockIdx.x;
    if (tile * QL_TILE_M >= n) return;
    for (int i = threadIdx.x; i < kvol * c_in * C_OUT; i += blockDim.x) s_w[i] = w[i];
    if (threadIdx.x < C_OUT) s_absmax[threadIdx.x] = 0u;
    __syncthreads();
    const int r = threadIdx.x;
    const int64_t row = tile * QL_TILE_M + r;
    float acc[C_OUT];
#pragma unroll
    for (int j = 0; j < C_OUT; ++j) acc[j] = 0.f;
    const int* nb = nbr + tile * (int64_t)kvol * QL_TILE_M + r;
    for (int k = 0; k < kvol; ++k) {
        const int idx = nb[k * QL_TILE_M];
        if (idx < 0) continue;
        const float* xr = feats + (int64_t)idx * c_in;
        const float* wk = s_w + k * c_in * C_OUT;
        for (int ic = 0; ic < c_in; ++ic) {
            const float xv = xr[ic];
#pragma unroll
            for (int j = 0; j < C_OUT; ++j) acc[j] = fmaf(xv, wk[ic * C_OUT + j], acc[j]);
        }
    }
    if (row < n) {
#pragma unroll
        for (int j = 0; j < C_OUT; ++j) {
            float y = fmaf(acc[j], scale[j], shift[j]);
            if (relu) y = fmaxf(y, 0.f);
            acc[j] = y;
        }
        if (out_dtype == QL_F16) {
            __half2* o = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(out) + row * C_OUT);
#pragma unroll
            for (int j = 0; j < C_OUT / 2; ++j) o[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        } else {
            float* o = reinterpret_cast<float*>(out) + row * C_OUT;
#pragma unroll
            for (int j = 0; j < C_OUT; ++j) o[j] = acc[j];
        }
    }
    if (absmax) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int j = 0; j < C_OUT; ++j) {
            const uint32_t m = __reduce_max_sync(0xffffffffu, row < n ? __float_as_uint(fabsf(acc[j])) : 0u);
            if (lane == (j & 31)) atomicMax(&s_absmax[j], m);
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

extern "C" int ql_abi_version(void) { return 1; }

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
    k_absmax_cols<<<grid, 256, (size_t)c * 4, (cudaStream_t)stream_>>>(x, dtype, n_cap, n_dev, c, absmax);
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
    } else {
        k_quantize_rows<<<grid, 256, (size_t)c * 12, (cudaStream_t)stream_>>>(x, in_dtype, n_cap, n_dev, c, absmax, smooth, bound,
                                                                            mode, out, act_scale_out);
    }
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_stem_conv(const float* feats, int32_t c_in, const int32_t* nbr, int64_t n_out_cap, const int32_t* n_out_dev,
                            int32_t c_out, int32_t kvol, const float* w, const float* scale, const float* shift, int32_t relu,
                            void* out, int32_t out_dtype, float* absmax, ql_stream_t stream_) {
    if (!feats || !nbr || !w || !scale || !shift || !out) return QL_ERR_INVALID;
    if (c_in <= 0 || c_in > 16 || kvol <= 0 || kvol > 343 || (out_dtype != QL_F16 && out_dtype != QL_F32)) return QL_ERR_INVALID;
    if (c_out != 16 && c_out != 32) return QL_ERR_UNSUPPORTED;
    if (n_out_cap <= 0) return QL_OK;
    unsigned tiles = (unsigned)((n_out_cap + QL_TILE_M - 1) / QL_TILE_M);
    size_t smem = (size_t)kvol * c_in * c_out * 4 + (size_t)c_out * 4;
    cudaStream_t st = (cudaStream_t)stream_;
    if (c_out == 16) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(k_stem_conv<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return QL_ERR_CUDA;
        k_stem_conv<16><<<tiles, QL_TILE_M, smem, st>>>(feats, c_in, nbr, n_out_cap, n_out_dev, kvol, w, scale, shift, relu, out,
                                                       out_dtype, absmax);
    } else {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(k_stem_conv<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return QL_ERR_CUDA;
        k_stem_conv<32><<<tiles, QL_TILE_M, smem, st>>>(feats, c_in, nbr, n_out_cap, n_out_dev, kvol, w, scale, shift, relu, out,
                                                       out_dtype, absmax);
    }
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
