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

/* ---------------------------------------------------------------------------------------------------------------------
 * Rotated BEV IoU of two boxes [x, y, z, dx, dy, dz, heading] -- restatement of box_overlap / iou_bev
 * (pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:35-235): the intersection polygon = edge-edge crossings + corners of one
 * rectangle inside the other (margin 1e-2), ordered by angle around their centroid (stable), area by a triangle fan.
 * fp32 with one rounding per operation (-ffp-contract=off); the device compilers fuse some multiply-adds, so a device IoU
 * can differ from this one in the last bits -- the suppression decision (IoU > thresh) only for pairs that close to thresh.
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct { float x, y; } qlo_p2;
#define QLO_EPS 1e-8f

static float qlo_cross3(qlo_p2 p1, qlo_p2 p2, qlo_p2 p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }

static int qlo_seg_cross(qlo_p2 p1, qlo_p2 p0, qlo_p2 q1, qlo_p2 q0, qlo_p2* out) {
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) && fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) &&
          fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return 0;
    const float s1 = qlo_cross3(q0, p1, p0), s2 = qlo_cross3(p1, q1, p0), s3 = qlo_cross3(p0, q1, q0), s4 = qlo_cross3(q1, p1, q0);
    if (!(s1 * s2 > 0.f && s3 * s4 > 0.f)) return 0;
    const float s5 = qlo_cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > QLO_EPS) {
        out->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        out->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        out->x = (b0 * c1 - b1 * c0) / D;
        out->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static int qlo_inside(const float* box, qlo_p2 p) {
    const float c = cosf(-box[6]), s = sinf(-box[6]);
    const float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    const float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + 1e-2f && fabsf(ry) < box[4] / 2 + 1e-2f;
}

static void qlo_corners(const float* box, qlo_p2* c) {
    const float hx = box[3] / 2, hy = box[4] / 2;
    const float px[4] = {box[0] - hx, box[0] + hx, box[0] + hx, box[0] - hx}, py[4] = {box[1] - hy, box[1] - hy, box[1] + hy, box[1] + hy};
    const float ca = cosf(box[6]), sa = sinf(box[6]);
    for (int k = 0; k < 4; ++k) {
        c[k].x = (px[k] - box[0]) * ca + (py[k] - box[1]) * (-sa) + box[0];
        c[k].y = (px[k] - box[0]) * sa + (py[k] - box[1]) * ca + box[1];
    }
    c[4] = c[0];
}

float qlo_rect_iou(const float* a, const float* b) {
    qlo_p2 ca[5], cb[5], pts[16], centre = {0.f, 0.f};
    float ang[16];
    int cnt = 0;
    qlo_corners(a, ca);
    qlo_corners(b, cb);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (qlo_seg_cross(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
                centre.x += pts[cnt].x;
                centre.y += pts[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (qlo_inside(a, cb[k])) { centre.x += cb[k].x; centre.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (qlo_inside(b, ca[k])) { centre.x += ca[k].x; centre.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    float ov = 0.f;
    if (cnt > 0) {
        centre.x /= cnt;
        centre.y /= cnt;
        for (int k = 0; k < cnt; ++k) ang[k] = atan2f(pts[k].y - centre.y, pts[k].x - centre.x);
        for (int k = 1; k < cnt; ++k) {                       /* stable ascending: the order of the reference's bubble sort */
            const qlo_p2 p = pts[k];
            const float t = ang[k];
            int j = k - 1;
            while (j >= 0 && ang[j] > t) { pts[j + 1] = pts[j]; ang[j + 1] = ang[j]; --j; }
            pts[j + 1] = p;
            ang[j + 1] = t;
        }
        float area = 0.f;
        for (int k = 0; k < cnt - 1; ++k)
            area += (pts[k].x - pts[0].x) * (pts[k + 1].y - pts[0].y) - (pts[k].y - pts[0].y) * (pts[k + 1].x - pts[0].x);
        ov = fabsf(area) / 2.0f;
    }
    const float sa = a[3] * a[4], sb = b[3] * b[4];
    return ov / fmaxf(sa + sb - ov, QLO_EPS);
}

/* pairwise IoU, upper triangle (j > i) of n boxes with row stride `stride` floats; out [n][n] */
void qlo_rect_iou_matrix(const float* boxes, int64_t n, int64_t stride, float* out) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) out[i * n + j] = qlo_rect_iou(boxes + i * stride, boxes + j * stride);
}
