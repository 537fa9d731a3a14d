// SmoothQuant weight preparation on the device: per-input-channel smoothing scale, smoothed per-output-channel int8 weight
// codes written straight into the conv kernel's packed shared-memory image, and the de-quantisation scale.
//
// Reference: quant/smoothquant.py:69-82 (the formula: s = amax_x^alpha / amax_w^(1-alpha) per input channel, zeros -> 1,
// x' = x / s, w' = w * s, then fake-quant of both) as applied to the sparse 3-D convs by the intent of SQConv3d
// (quant/quant_conv3d.py:141-236; non-functional as shipped).  With DYNAMIC activation statistics s changes with
// every forward, so the smoothed weights have to be re-quantised per call: round 1 did that on the host (a device -> host ->
// device round trip per layer, which kept SQConv3d out of the CUDA-graph engine); these two kernels keep it on the device
// (<= 0.9 M weights per layer), so that a whole W8A8-sq backbone forward is one graph replay.
#include "ql_common.cuh"

namespace {

// chunk geometry / swizzle of the packed weight image -- the device twins of chunk_geom, chunk_sw_offset and the inverse of
// k_word_src in spconv_mma.cu (tests/test_abi.py::test_pack_weights_host_layout pins the host versions; the GPU tests compare
// this kernel's image with ql_pack_weights_host's byte for byte)
__device__ __forceinline__ uint32_t sw_offset(int ch, uint32_t r, uint32_t c16) {
    const uint32_t x = ch == 128 ? (r & 7u) : (ch == 64 ? ((r >> 1) & 3u) : ((r >> 2) & 1u));
    return (r >> 3) * (uint32_t)(8 * ch) + (r & 7u) * (uint32_t)ch + ((c16 ^ x) << 4);
}
// TMEM column (4-byte K word slot) that holds source word ws of a row segment
__device__ __forceinline__ int k_word_dst(int ch, int ws) {
    if (ch < 64) return ws;
    const int krep = ch / 32;
    const int e = ws & 1, h = ws >> 1;
    return 8 * (h % krep) + 2 * (h / krep) + e;
}

__global__ void __launch_bounds__(256) k_sq_smooth(const float* __restrict__ act_absmax, const float* __restrict__ w_ic_absmax,
                                                   float alpha, int c_in, float* __restrict__ smooth) {
    for (int c = threadIdx.x; c < c_in; c += blockDim.x) {
        float s = __fdiv_rn(powf(act_absmax[c], alpha), powf(w_ic_absmax[c], 1.0f - alpha));
        if (s == 0.f || !isfinite(s)) s = 1.0f;                      // quant/smoothquant.py:76: zeros (and 0/0, x/0) -> 1
        smooth[c] = s;
    }
}

// one CTA per output channel: amax of the smoothed row, then its int8 codes into the packed image
__global__ void __launch_bounds__(256) k_sq_weights(const float* __restrict__ w, const float* __restrict__ smooth,
                                                    const float* __restrict__ bn_scale, int c_in, int c_out, int kvol, int ch, int nseg,
                                                    int8_t* __restrict__ packed, float* __restrict__ scale_out) {
    __shared__ float s_red[8];
    __shared__ float s_amax;
    const int oc = blockIdx.x;
    const int n = kvol * c_in;
    const float* row = w + (size_t)oc * n;
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(__fmul_rn(row[i], smooth[i % c_in])));
    m = ql_warp_max(m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t = fmaxf(t, s_red[i]);
        s_amax = t;
        scale_out[oc] = __fmul_rn(__fdiv_rn(t, 127.0f), bn_scale ? bn_scale[oc] : 1.0f);
    }
    __syncthreads();
    const float amax = s_amax;
    const float qs = amax <= (1.0f / 16777216.0f) ? 0.f : __fdiv_rn(127.0f, amax);
    const size_t chunk_bytes = (size_t)c_out * ch;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int k = i / c_in, ic = i - k * c_in;
        float q = rintf(__fmul_rn(__fmul_rn(row[i], smooth[ic]), qs));
        q = fminf(fmaxf(q, -127.f), 127.f);
        const int seg = ic / 128, bs = ic - seg * 128;               // int8: byte == channel; one segment = 128 bytes of the row
        const int c = k_word_dst(ch, bs >> 2);
        packed[(size_t)(k * nseg + seg) * chunk_bytes + sw_offset(ch, (uint32_t)oc, (uint32_t)(c >> 2)) + 4 * (c & 3) + (bs & 3)] = (int8_t)(int)q;
    }
}

}  // namespace

extern "C" int ql_sq_prepare_weights(const float* w, const float* w_ic_absmax, const float* act_absmax, float alpha, int32_t c_in,
                                     int32_t c_out, int32_t kvol, const float* bn_scale, float* smooth_out, int8_t* packed_out,
                                     float* scale_out, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (!w || !w_ic_absmax || !act_absmax || !smooth_out || !packed_out || !scale_out) return QL_ERR_INVALID;
    if (c_in <= 0 || c_in % 16 != 0 || c_out < 16 || c_out % 16 != 0 || c_out > 256 || kvol <= 0) return QL_ERR_UNSUPPORTED;
    const int ch = c_in > 128 ? 128 : (c_in <= 32 ? 32 : (c_in <= 64 ? 64 : 128));
    const int nseg = c_in > 128 ? (c_in + 127) / 128 : 1;
    // rows shorter than the chunk (c_in = 16 in a 32-byte chunk, 48 in 64 ...) leave zero padding: clear the image first
    const size_t bytes = (size_t)kvol * nseg * c_out * ch;
    if (c_in % ch != 0 && cudaMemsetAsync(packed_out, 0, bytes, st) != cudaSuccess) return QL_ERR_CUDA;
    k_sq_smooth<<<1, 256, 0, st>>>(act_absmax, w_ic_absmax, alpha, c_in, smooth_out);
    k_sq_weights<<<c_out, 256, 0, st>>>(w, smooth_out, bn_scale, c_in, c_out, kvol, ch, nseg, packed_out, scale_out);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

// ------------------------------------------------------------------------------------------------
// Dense SmoothQuant wrappers (quant/smoothquant.py: SQConv2d / SQConv1d / SQConvT2d / SQLinear; quant/SQSubM2d.py): the reference
// materialises F.unfold(x) in fp32 -- [B*L, ic*kh*kw] -- to take per-COLUMN activation maxima, divide by the smoothing scale and
// fake-quantise.  Here the unfolded matrix only ever exists as int8 codes: ql_unfold_absmax takes the per-column maxima straight
// from the NCHW map (no materialisation), ql_unfold_quantize writes the smoothed, quantised unfolded matrix once (1 byte per
// element, columns in F.unfold order c*kh*kw + ky*kw + kx, zero padded to a multiple of 16), and the product with the smoothed
// int8 weights is ONE ql_spconv_mma launch (kernel volume 1, identity rulebook, INT32 accumulate on tcgen05 kind::i8).
// ------------------------------------------------------------------------------------------------
namespace {

struct UnfoldGeom {
    int B, C, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo;
};

__device__ __forceinline__ float ld_in(const void* x, int dtype, int64_t i) {
    return dtype == QL_F16 ? __half2float(((const __half*)x)[i]) : ((const float*)x)[i];
}

// one CTA per (channel, batch) plane: every pixel updates the maxima of the kernel positions whose windows contain it
__global__ void __launch_bounds__(256) k_unfold_absmax(const void* __restrict__ x, int dtype, UnfoldGeom g, float* __restrict__ absmax) {
    extern __shared__ uint32_t s_m[];                    // [kh * kw]
    const int K = g.kh * g.kw;
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_m[i] = 0u;
    __syncthreads();
    const int c = blockIdx.x, b = blockIdx.y;
    const int64_t plane = ((int64_t)b * g.C + c) * g.H * g.W;
    for (int k = 0; k < K; ++k) {
        const int ky = k / g.kw, kx = k - ky * g.kw;
        float m = 0.f;
        // the window of column (c, ky, kx): rows oy*sh - ph + ky*dh, oy in [0, Ho); same for x
        for (int i = threadIdx.x; i < g.Ho * g.Wo; i += blockDim.x) {
            const int oy = i / g.Wo, ox = i - oy * g.Wo;
            const int y = oy * g.sh - g.ph + ky * g.dh, xx = ox * g.sw - g.pw + kx * g.dw;
            if (y >= 0 && y < g.H && xx >= 0 && xx < g.W) m = fmaxf(m, fabsf(ld_in(x, dtype, plane + (int64_t)y * g.W + xx)));
        }
        m = ql_warp_max(m);
        if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&s_m[k], __float_as_uint(m));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        if (s_m[i]) atomicMax(reinterpret_cast<unsigned int*>(absmax) + c * K + i, s_m[i]);
}

// tensor amax of the smoothed columns and the (de-)quantisation scales: amax_t = max_col absmax[col] / smooth[col]
__global__ void __launch_bounds__(256) k_unfold_scales(const float* __restrict__ absmax, const float* __restrict__ smooth, int n_cols,
                                                       float bound, float* __restrict__ scales /* {quant scale, act_scale = amax/bound} */) {
    __shared__ float s_red[8];
    float m = 0.f;
    for (int i = threadIdx.x; i < n_cols; i += blockDim.x) m = fmaxf(m, __fdiv_rn(absmax[i], smooth[i]));
    m = ql_warp_max(m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t = fmaxf(t, s_red[i]);
        scales[0] = t <= (1.0f / 16777216.0f) ? 0.f : __fdiv_rn(bound, t);
        scales[1] = __fdiv_rn(t, bound);
    }
}

// out[m][col] = clamp(rint((x / smooth[col]) * qscale)), m = (b, oy, ox), col = c*kh*kw + ky*kw + kx (F.unfold order); padding = 0
__global__ void __launch_bounds__(256) k_unfold_quantize(const void* __restrict__ x, int dtype, UnfoldGeom g, const float* __restrict__ smooth,
                                                         const float* __restrict__ scales, float bound, int n_cols, int col_stride,
                                                         int8_t* __restrict__ out) {
    const int64_t M = (int64_t)g.B * g.Ho * g.Wo;
    const float qs = scales[0];
    const int K = g.kh * g.kw;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M * col_stride; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = t / col_stride;
        const int col = (int)(t - m * col_stride);
        int8_t q = 0;
        if (col < n_cols) {
            const int c = col / K, k = col - c * K;
            const int ky = k / g.kw, kx = k - ky * g.kw;
            const int b = (int)(m / (g.Ho * g.Wo));
            const int r = (int)(m - (int64_t)b * g.Ho * g.Wo);
            const int oy = r / g.Wo, ox = r - oy * g.Wo;
            const int y = oy * g.sh - g.ph + ky * g.dh, xx = ox * g.sw - g.pw + kx * g.dw;
            if (y >= 0 && y < g.H && xx >= 0 && xx < g.W) {
                const float v = __fdiv_rn(ld_in(x, dtype, (((int64_t)b * g.C + c) * g.H + y) * g.W + xx), smooth[col]);
                q = (int8_t)(int)fminf(fmaxf(rintf(__fmul_rn(v, qs)), -bound), bound);
            }
        }
        out[t] = q;
    }
}

// k_unfold_absmax for 3 x 3 kernels in ONE pass over the plane: a pixel updates, in registers, the maxima of the (ky, kx) columns
// whose windows contain it (the general kernel above re-reads the plane once per kernel position: 282 us per BEV-backbone layer)
__global__ void __launch_bounds__(256) k_unfold_absmax_3x3(const void* __restrict__ x, int dtype, UnfoldGeom g, float* __restrict__ absmax) {
    __shared__ uint32_t s_m[9];
    extern __shared__ uint8_t s_tab[];                   // [H] which ky see row y, [W] which kx see column x (bit masks)
    uint8_t* s_vy = s_tab;
    uint8_t* s_vx = s_tab + g.H;
    if (threadIdx.x < 9) s_m[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < g.H + g.W; i += blockDim.x) {
        const bool is_y = i < g.H;
        const int pos = is_y ? i : i - g.H;
        const int pad = is_y ? g.ph : g.pw, dil = is_y ? g.dh : g.dw, st = is_y ? g.sh : g.sw, n_out = is_y ? g.Ho : g.Wo;
        uint32_t mask = 0u;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int t = pos + pad - k * dil;           // o * stride of the window that sees this row / column at kernel position k
            if (t >= 0 && t % st == 0 && t / st < n_out) mask |= 1u << k;
        }
        s_tab[i] = (uint8_t)mask;
    }
    __syncthreads();
    const int c = blockIdx.x, b = blockIdx.y;
    const int64_t plane = ((int64_t)b * g.C + c) * g.H * g.W;
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = 0.f;
    // a warp per image row (gridDim.z CTAs share the plane), lanes along x: no division in the loop
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int y = blockIdx.z * 8 + warp; y < g.H; y += gridDim.z * 8) {
        const uint32_t my = s_vy[y];
        if (my == 0u) continue;
        const int64_t row = plane + (int64_t)y * g.W;
#pragma unroll 2
        for (int xx = lane; xx < g.W; xx += 32) {
            const float v = fabsf(ld_in(x, dtype, row + xx));
            const uint32_t mx = s_vx[xx];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                    if (((my >> ky) & 1u) && ((mx >> kx) & 1u)) m[ky * 3 + kx] = fmaxf(m[ky * 3 + kx], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float w = ql_warp_max(m[k]);
        if ((threadIdx.x & 31) == 0 && w > 0.f) atomicMax(&s_m[k], __float_as_uint(w));
    }
    __syncthreads();
    if (threadIdx.x < 9 && s_m[threadIdx.x]) atomicMax(reinterpret_cast<unsigned int*>(absmax) + c * 9 + threadIdx.x, s_m[threadIdx.x]);
}

// k_unfold_quantize, tiled (C % 16 == 0, no padding columns): one CTA = 32 consecutive output pixels of one output row, all columns in
// groups of 16 channels.  The group's input patch is staged in shared memory with loads that are contiguous along x; every thread then
// produces 16 consecutive codes of one pixel's row (one 16-byte store; a pixel's 16 * kh * kw bytes of the group are contiguous).  The
// element-per-thread kernel above reads a different channel plane with every lane and stores single bytes: 1.02 ms per BEV-backbone layer.
constexpr int kUqPix = 32, kUqCg = 16;
__global__ void __launch_bounds__(256) k_unfold_quantize_tiled(const void* __restrict__ x, int dtype, UnfoldGeom g, const float* __restrict__ smooth,
                                                               const float* __restrict__ scales, float bound, int col_stride,
                                                               int8_t* __restrict__ out) {
    extern __shared__ float s_uq[];
    const int K = g.kh * g.kw, GK = kUqCg * K, SEG = GK / 16;
    const int Wp = (kUqPix - 1) * g.sw + (g.kw - 1) * g.dw + 1;
    float* s_patch = s_uq;                                   // [16][kh][Wp]
    float* s_smooth = s_patch + kUqCg * g.kh * Wp;           // [GK]
    int* s_off = reinterpret_cast<int*>(s_smooth + GK);      // [GK]: patch offset of column (c_local, ky, kx) for pixel 0
    const int tiles_x = (g.Wo + kUqPix - 1) / kUqPix;
    const int tx = blockIdx.x % tiles_x, oy = (blockIdx.x / tiles_x) % g.Ho, b = blockIdx.x / (tiles_x * g.Ho);
    const int ox0 = tx * kUqPix;
    const float qs = scales[0];
    for (int i = threadIdx.x; i < GK; i += blockDim.x) {
        const int cl = i / K, k = i - cl * K, ky = k / g.kw, kx = k - ky * g.kw;
        s_off[i] = (cl * g.kh + ky) * Wp + kx * g.dw;
    }
    const int64_t m0 = ((int64_t)b * g.Ho + oy) * g.Wo + ox0;
    const int x_start = ox0 * g.sw - g.pw;
    for (int c0 = 0; c0 < g.C; c0 += kUqCg) {
        __syncthreads();                                     // the previous group's readers are done (and s_off is visible)
        for (int i = threadIdx.x; i < kUqCg * g.kh * Wp; i += blockDim.x) {
            const int xi = i % Wp, r = i / Wp, ky = r % g.kh, cl = r / g.kh;
            const int y = oy * g.sh - g.ph + ky * g.dh, xg = x_start + xi;
            float v = 0.f;
            if (y >= 0 && y < g.H && xg >= 0 && xg < g.W) v = ld_in(x, dtype, (((int64_t)b * g.C + c0 + cl) * g.H + y) * g.W + xg);
            s_patch[i] = v;
        }
        for (int i = threadIdx.x; i < GK; i += blockDim.x) s_smooth[i] = smooth[c0 * K + i];
        __syncthreads();
        for (int item = threadIdx.x; item < kUqPix * SEG; item += blockDim.x) {
            const int p = item / SEG, j = item - p * SEG;
            if (ox0 + p >= g.Wo) continue;
            uint32_t w4[4];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                uint32_t word = 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int cl = 16 * j + 4 * q4 + e;
                    const float v = __fdiv_rn(s_patch[s_off[cl] + p * g.sw], s_smooth[cl]);
                    const float q = fminf(fmaxf(rintf(__fmul_rn(v, qs)), -bound), bound);
                    word |= ((uint32_t)(uint8_t)(int8_t)(int)q) << (8 * e);
                }
                w4[q4] = word;
            }
            *reinterpret_cast<uint4*>(out + (m0 + p) * col_stride + (int64_t)c0 * K + 16 * j) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
    }
}

bool unfold_geom(int32_t B, int32_t C, int32_t H, int32_t W, const int32_t* k, const int32_t* s, const int32_t* p, const int32_t* d, UnfoldGeom& g) {
    if (!k || !s || !p || !d || B <= 0 || C <= 0 || H <= 0 || W <= 0) return false;
    g = UnfoldGeom{B, C, H, W, k[0], k[1], s[0], s[1], p[0], p[1], d[0], d[1], 0, 0};
    if (g.kh <= 0 || g.kw <= 0 || g.sh <= 0 || g.sw <= 0 || g.ph < 0 || g.pw < 0 || g.dh <= 0 || g.dw <= 0) return false;
    g.Ho = (H + 2 * g.ph - g.dh * (g.kh - 1) - 1) / g.sh + 1;
    g.Wo = (W + 2 * g.pw - g.dw * (g.kw - 1) - 1) / g.sw + 1;
    return g.Ho > 0 && g.Wo > 0;
}

}  // namespace

extern "C" int ql_unfold_absmax(const void* x, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W, const int32_t* kernel_hw,
                                const int32_t* stride_hw, const int32_t* pad_hw, const int32_t* dil_hw, float* absmax_cols,
                                ql_stream_t stream_) {
    UnfoldGeom g;
    if (!x || !absmax_cols || (dtype != QL_F16 && dtype != QL_F32) || !unfold_geom(B, C, H, W, kernel_hw, stride_hw, pad_hw, dil_hw, g))
        return QL_ERR_INVALID;
    if (g.kh == 3 && g.kw == 3 && H + W <= 40000)
        k_unfold_absmax_3x3<<<dim3((unsigned)C, (unsigned)B, (unsigned)((H + 23) / 24 < 8 ? (H + 23) / 24 : 8)), 256, (size_t)(H + W),
                              (cudaStream_t)stream_>>>(x, dtype, g, absmax_cols);
    else
        k_unfold_absmax<<<dim3((unsigned)C, (unsigned)B), 256, (size_t)g.kh * g.kw * 4, (cudaStream_t)stream_>>>(x, dtype, g, absmax_cols);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}

extern "C" int ql_unfold_quantize(const void* x, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W, const int32_t* kernel_hw,
                                  const int32_t* stride_hw, const int32_t* pad_hw, const int32_t* dil_hw, const float* absmax_cols,
                                  const float* smooth_cols, int32_t bits, int32_t col_stride, int8_t* out, float* scales_out,
                                  ql_stream_t stream_) {
    UnfoldGeom g;
    if (!x || !absmax_cols || !smooth_cols || !out || !scales_out || (dtype != QL_F16 && dtype != QL_F32) || bits < 2 || bits > 8 ||
        !unfold_geom(B, C, H, W, kernel_hw, stride_hw, pad_hw, dil_hw, g))
        return QL_ERR_INVALID;
    const int n_cols = C * g.kh * g.kw;
    if (col_stride < n_cols || col_stride % 16 != 0) return QL_ERR_INVALID;
    const float bound = (float)((1 << (bits - 1)) - 1);
    cudaStream_t st = (cudaStream_t)stream_;
    k_unfold_scales<<<1, 256, 0, st>>>(absmax_cols, smooth_cols, n_cols, bound, scales_out);
    {
        const int K = g.kh * g.kw, Wp = (kUqPix - 1) * g.sw + (g.kw - 1) * g.dw + 1;
        const size_t smem = ((size_t)kUqCg * g.kh * Wp + 2 * (size_t)kUqCg * K) * 4;
        const int64_t ctas = (int64_t)g.B * g.Ho * ((g.Wo + kUqPix - 1) / kUqPix);
        if (C % kUqCg == 0 && col_stride == n_cols && smem <= 48 * 1024 && ctas < 2147483647LL && ((uintptr_t)out & 15) == 0) {
            k_unfold_quantize_tiled<<<(unsigned)ctas, 256, smem, st>>>(x, dtype, g, smooth_cols, scales_out, bound, col_stride, out);
            QL_CUDA_CHECK_LAST();
            return QL_OK;
        }
    }
    const int64_t total = (int64_t)g.B * g.Ho * g.Wo * col_stride;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 32 * ql_num_sms()) blocks = 32 * ql_num_sms();
    k_unfold_quantize<<<(unsigned)blocks, 256, 0, st>>>(x, dtype, g, smooth_cols, scales_out, bound, n_cols, col_stride, out);
    QL_CUDA_CHECK_LAST();
    return QL_OK;
}
