/* TEST INFRASTRUCTURE (oracle) -- never linked into or called by the product library.
 *
 * Exact single-rounding restatements of the two fp32 device routines whose results feed an int8 quantiser, so that the
 * oracle can reproduce the kernels' int8 CODES bit for bit through a whole backbone (oracle/qlidar_oracle.py,
 * "kernel-numerics mirror").  numpy has no fused multiply-add; C's fmaf() is the IEEE single-rounding FMA the device's
 * FFMA implements.
 *
 *   qlo_stem_conv : the fp32 SIMT stem conv (csrc/elementwise.cu k_stem_conv) == the reference's un-quantised
 *                   conv_input (spconv_backbone.py:193-198; quant_centerpoint.py:24-26 no_list) + BatchNorm1d + ReLU
 *                   acc = fmaf(x[nbr[k][r]][ic], w[k][ic][oc], acc) over k ascending, ic ascending;  y = fmaf(acc, scale, shift)
 *   qlo_epilogue  : the conv kernel's epilogue (csrc/spconv_mma.cu): y = fmaf((float)acc, s[oc], shift[oc]) (+ residual) (ReLU)
 *                   == de-quantisation + folded BatchNorm1d + SparseBasicBlock's residual add (spconv_backbone.py:51-67)
 */
#include <math.h>
#include <stdint.h>

void qlo_stem_conv(const float* x, int64_t x_stride, int32_t c_in, const int32_t* nbr /* [K][n] */, int32_t K, int64_t n,
                   const float* w /* [K][c_in][c_out] */, int32_t c_out, const float* scale, const float* shift, int32_t relu,
                   float* y /* [n][c_out] */) {
    for (int64_t r = 0; r < n; ++r) {
        float* yr = y + r * c_out;
        for (int c = 0; c < c_out; ++c) yr[c] = 0.f;
        for (int k = 0; k < K; ++k) {
            const int32_t j = nbr[(int64_t)k * n + r];
            if (j < 0) continue;                               /* the kernel adds x = 0 here: fmaf(0, w, acc) == acc */
            const float* xr = x + (int64_t)j * x_stride;
            for (int ic = 0; ic < c_in; ++ic) {
                const float xv = xr[ic];
                const float* wk = w + ((int64_t)k * c_in + ic) * c_out;
                for (int c = 0; c < c_out; ++c) yr[c] = fmaf(xv, wk[c], yr[c]);
            }
        }
        for (int c = 0; c < c_out; ++c) {
            float v = fmaf(yr[c], scale[c], shift[c]);
            if (relu) v = fmaxf(v, 0.f);
            yr[c] = v;
        }
    }
}

void qlo_epilogue(const int32_t* acc, int64_t n, int32_t c, const float* s, const float* shift,
                  const float* residual /* [n][c] fp16 values widened to fp32, or NULL */, int32_t relu, float* y) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < c; ++j) {
            float v = fmaf((float)acc[i * c + j], s[j], shift[j]);
            if (residual) v += residual[i * c + j];
            if (relu) v = fmaxf(v, 0.f);
            y[i * c + j] = v;
        }
}
