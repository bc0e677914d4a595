// Shared pieces of the reference-actual AdvMix chains (lib/dataset/advaug.py): the per-image autoaug plan, PIL's
// SMOOTH / blend arithmetic and the closed-form gridmask.  Used by chains.cu (materialising kernels) and chainmix.cu
// (fused chain + mix kernels, which recompute the chains in registers).
#pragma once
#include "common.cuh"

namespace advmix {

enum { OP_NONE = 0, OP_EQUALIZE = 1, OP_POSTERIZE = 2, OP_SOLARIZE = 3, OP_INVERT = 4, OP_SHARPNESS = 5 };

struct AutoPlan {            // per image, lives in the workspace
    uint8_t pre[768];
    uint8_t post[768];
    float factor;            // sharpness blend factor
    int stencil;             // 1 if a sharpness stage is present
    int pad[2];
};

// PIL ImageFilter.SMOOTH at an interior pixel of the (pre-LUT mapped) image, then
// ImageEnhance.Sharpness blend.  p points at channel c of pixel (y,x); pitch in bytes.
__device__ __forceinline__ uint8_t sharpen_px(const uint8_t* __restrict__ p, int64_t pitch,
                                              const uint8_t* __restrict__ lut, float factor, bool interior) {
    const float v = (float)lut[p[0]];
    float smooth = v;
    if (interior) {
        const float k1 = __fdiv_rn(1.0f, 13.0f), k5 = __fdiv_rn(5.0f, 13.0f);
        float ss = 0.5f;
        const uint8_t* r = p + pitch;  // row y+1 first (PIL: in1, in0, in_1)
        ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn((float)lut[r[-3]], k1), __fmul_rn((float)lut[r[0]], k1)),
                                     __fmul_rn((float)lut[r[3]], k1)));
        r = p;
        ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn((float)lut[r[-3]], k1), __fmul_rn(v, k5)),
                                     __fmul_rn((float)lut[r[3]], k1)));
        r = p - pitch;
        ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn((float)lut[r[-3]], k1), __fmul_rn((float)lut[r[0]], k1)),
                                     __fmul_rn((float)lut[r[3]], k1)));
        smooth = ss <= 0.f ? 0.f : (ss >= 255.f ? 255.f : (float)(uint8_t)ss);
    }
    // Image.blend(smooth, img, factor)
    const float t = __fadd_rn(smooth, __fmul_rn(factor, __fsub_rn(v, smooth)));
    if (factor >= 0.f && factor <= 1.f) return (uint8_t)t;
    return t <= 0.f ? 0 : (t >= 255.f ? 255 : (uint8_t)t);
}

// ---- gridmask -----------------------------------------------------------------------
struct GridGeom { int l, hh, ww, oy, ox; };

__device__ __forceinline__ GridGeom grid_geom(int H, int W, int d) {
    GridGeom g;
    g.hh = (int)(1.5 * H);
    g.ww = (int)(1.5 * W);
    g.l = min(max((int)(d * 0.5 + 0.5), 1), d - 1);
    g.oy = (g.hh - H) / 2;
    g.ox = (g.ww - W) / 2;
    return g;
}

// mode=1 mask value at output pixel (y,x): 1 on the grid lines, 0 in the cells.
__device__ __forceinline__ float grid_mask_at(int y, int x, int d, int st_h, int st_w, const GridGeom& g) {
    bool line = false;
    int t = y + g.oy - st_h;
    if (t >= 0) { int i = t / d; line |= (i < g.hh / d) && (t - i * d < g.l); }
    t = x + g.ox - st_w;
    if (t >= 0) { int i = t / d; line |= (i < g.ww / d) && (t - i * d < g.l); }
    return line ? 1.0f : 0.0f;
}

// ---- softmax over K values for four pixels at once ----------------------------------------------------------------
// exp(x - max) * (1 / sum) with the SFU approximations ex2.approx (2^-22 relative) and rcp.approx (1 ulp): the weights are
// within 5e-7 of torch.softmax (the tests' bar is 2e-6) at 17 instructions per pixel for K = 3 instead of ~40 for
// expf + __frcp_rn - the mix kernels are instruction-issue bound, not HBM bound, on this part.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int K>
__device__ __forceinline__ void softmax4_sfu(float4 (&w)[K]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float m = (&w[0].x)[i];
#pragma unroll
        for (int k = 1; k < K; ++k) m = fmaxf(m, (&w[k].x)[i]);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float e = ex2_approx(__fmul_rn(__fsub_rn((&w[k].x)[i], m), 1.4426950408889634f));
            (&w[k].x)[i] = e;
            s = __fadd_rn(s, e);
        }
        const float r = rcp_approx(s);
#pragma unroll
        for (int k = 0; k < K; ++k) (&w[k].x)[i] = __fmul_rn((&w[k].x)[i], r);
    }
}

}  // namespace advmix
