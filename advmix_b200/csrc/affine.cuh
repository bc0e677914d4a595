// get_affine_transform as one device function, shared by the matrix kernel, the crop planners, the joints
// kernel and the heat-map decoder, so all of them see bit-identical matrices.
#pragma once
#include "common.cuh"

namespace advmix {

// get_affine_transform (transforms.py:69-101), shift=0: the reference's float32 point triples and
// cv2.getAffineTransform's float64 LU solve.  One definition, used by the matrix kernel, the crop planners and the
// joints kernel, so all of them see bit-identical matrices.
// inverse != 0 solves the opposite direction (output -> source, transform_preds at transforms.py:61-66).
__device__ __forceinline__ void affine_from_csr(float cx, float cy, double scale_x, int scale_f32, double rot_deg, int out_w,
                                                int out_h, double* m, int inverse = 0) {
    // scale_tmp = scale * 200.0 ; src_w * -0.5 : in the dtype numpy gives `scale`
    // (float32 under numpy<2 value-based casting, float64 under NEP 50 after `s * np.clip(...)`)
    double half;
    if (scale_f32) half = (double)__fmul_rn(__fmul_rn((float)scale_x, 200.0f), -0.5f);
    else half = __dmul_rn(__dmul_rn(scale_x, 200.0), -0.5);
    const double rot_rad = __ddiv_rn(__dmul_rn(3.141592653589793, rot_deg), 180.0);
    const double sn = sin(rot_rad), cs = cos(rot_rad);
    const double dirx = __dsub_rn(__dmul_rn(0.0, cs), __dmul_rn(half, sn));
    const double diry = __dadd_rn(__dmul_rn(0.0, sn), __dmul_rn(half, cs));
    float s[3][2], d[3][2];
    s[0][0] = cx; s[0][1] = cy;                                    // center + scale_tmp*shift(0)
    s[1][0] = (float)__dadd_rn(__dadd_rn((double)cx, dirx), 0.0);
    s[1][1] = (float)__dadd_rn(__dadd_rn((double)cy, diry), 0.0);
    d[0][0] = (float)(out_w * 0.5); d[0][1] = (float)(out_h * 0.5);
    const float ddy = (float)(out_w * -0.5);
    d[1][0] = (float)__dadd_rn(out_w * 0.5, 0.0);
    d[1][1] = (float)__dadd_rn(out_h * 0.5, (double)ddy);
    // get_3rd_point(a,b) = b + (-(a-b).y, (a-b).x), float32
    {
        float dx = __fsub_rn(s[0][0], s[1][0]), dy = __fsub_rn(s[0][1], s[1][1]);
        s[2][0] = __fadd_rn(s[1][0], -dy); s[2][1] = __fadd_rn(s[1][1], dx);
        dx = __fsub_rn(d[0][0], d[1][0]); dy = __fsub_rn(d[0][1], d[1][1]);
        d[2][0] = __fadd_rn(d[1][0], -dy); d[2][1] = __fadd_rn(d[1][1], dx);
    }
    if (inverse) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float t0 = s[k][0], t1 = s[k][1];
            s[k][0] = d[k][0]; s[k][1] = d[k][1];
            d[k][0] = t0; d[k][1] = t1;
        }
    }
    // cv2.getAffineTransform (transforms.py:95-99): the 6x6 system  [x y 1 0 0 0; 0 0 0 x y 1] * M = [u; v]  solved by
    // cv::solve(DECOMP_LU), i.e. OpenCV's generic LUImpl<double> (partial pivoting, `d = -1/pivot`, row updates
    // `a += alpha * b` as separate multiply and add, back substitution `s -= a * x` then `s / pivot`) - restated
    // operation by operation, so the matrix is bit-identical to cv2's (the closed form differs in the last ulp,
    // which flips rint() ties of the fixed-point warp terms for un-rotated crops).
    double A[6][6], rhs[6];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double x = (double)s[i][0], y = (double)s[i][1];
        A[2 * i][0] = x; A[2 * i][1] = y; A[2 * i][2] = 1.0; A[2 * i][3] = 0.0; A[2 * i][4] = 0.0; A[2 * i][5] = 0.0;
        A[2 * i + 1][0] = 0.0; A[2 * i + 1][1] = 0.0; A[2 * i + 1][2] = 0.0; A[2 * i + 1][3] = x; A[2 * i + 1][4] = y; A[2 * i + 1][5] = 1.0;
        rhs[2 * i] = (double)d[i][0]; rhs[2 * i + 1] = (double)d[i][1];
    }
    bool singular = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        int k = i;
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            double best = 0.0;                                  // |A[k][i]| with k a run-time row: select, do not index
#pragma unroll
            for (int r = i; r < 6; ++r) if (r == k) best = fabs(A[r][i]);
            if (fabs(A[j][i]) > best) k = j;
        }
        {
            double piv = 0.0;
#pragma unroll
            for (int r = i; r < 6; ++r) if (r == k) piv = fabs(A[r][i]);
            if (piv < 2.220446049250313e-14) singular = true;   // DBL_EPSILON * 100
        }
#pragma unroll
        for (int r = i + 1; r < 6; ++r) {
            if (r == k) {
#pragma unroll
                for (int j = i; j < 6; ++j) { const double t = A[i][j]; A[i][j] = A[r][j]; A[r][j] = t; }
                const double t = rhs[i]; rhs[i] = rhs[r]; rhs[r] = t;
            }
        }
        const double dd = __ddiv_rn(-1.0, A[i][i]);
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            const double alpha = __dmul_rn(A[j][i], dd);
#pragma unroll
            for (int c = i + 1; c < 6; ++c) A[j][c] = __dadd_rn(A[j][c], __dmul_rn(alpha, A[i][c]));
            rhs[j] = __dadd_rn(rhs[j], __dmul_rn(alpha, rhs[i]));
        }
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double sacc = rhs[i];
#pragma unroll
        for (int c = i + 1; c < 6; ++c) sacc = __dsub_rn(sacc, __dmul_rn(A[i][c], rhs[c]));
        rhs[i] = __ddiv_rn(sacc, A[i][i]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) m[i] = singular ? 0.0 : rhs[i];   // cv::solve leaves X zero-filled when LU fails
}

}  // namespace advmix
