// get_affine_transform as one device function, shared by the matrix kernel, the crop planners, the joints
// kernel and the heat-map decoder, so all of them see bit-identical matrices.
#pragma once
#include "common.cuh"

namespace advmix {

// get_affine_transform (transforms.py:69-101), shift=0: the reference's float32 point triples and a
// closed-form float64 3-point solve.  One definition, used by the matrix kernel, the crop planners and the
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
    // closed-form solve M*[p,1] = q  (float64)
    const double p0x = s[0][0], p0y = s[0][1];
    const double ax = (double)s[1][0] - p0x, ay = (double)s[1][1] - p0y;
    const double bx = (double)s[2][0] - p0x, by = (double)s[2][1] - p0y;
    const double det = ax * by - ay * bx;
    const double inv = det != 0.0 ? 1.0 / det : 0.0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double q0 = d[0][r], u = (double)d[1][r] - q0, v = (double)d[2][r] - q0;
        const double m0 = (u * by - v * ay) * inv;
        const double m1 = (v * ax - u * bx) * inv;
        m[3 * r + 0] = m0;
        m[3 * r + 1] = m1;
        m[3 * r + 2] = q0 - m0 * p0x - m1 * p0y;
    }
}

}  // namespace advmix
