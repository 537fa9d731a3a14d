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
