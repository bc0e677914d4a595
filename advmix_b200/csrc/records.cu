// Row f4: the per-record scalar helpers of the dataset classes, batched - one thread per record.
//   advmix_xywh2cs        COCODataset._xywh2cs            lib/dataset/coco.py:205-220
//   advmix_half_body_cs   JointsDataset.half_body_transform lib/dataset/JointsDataset.py:69-111
//   advmix_select_data    JointsDataset.select_data        lib/dataset/JointsDataset.py:366-399
//   advmix_base_cs        centre / scale bookkeeping of get_base lib/dataset/JointsDataset.py:167-188
// Arithmetic follows numpy's dtype rules for the expressions in those functions under NEP 50 (numpy >= 2:
// a Python scalar adopts the dtype of the numpy value it meets), the numpy generation the fixtures were made with.
#include "common.cuh"

namespace advmix {

__global__ void xywh2cs_kernel(const double* __restrict__ box, float* __restrict__ center, float* __restrict__ scale, int B,
                               double aspect, double pixel_std) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double x = box[4 * b], y = box[4 * b + 1];
    double w = box[4 * b + 2], h = box[4 * b + 3];
    const float cx = (float)__dadd_rn(x, __dmul_rn(w, 0.5)), cy = (float)__dadd_rn(y, __dmul_rn(h, 0.5));
    if (w > __dmul_rn(aspect, h)) h = __ddiv_rn(__dmul_rn(w, 1.0), aspect);
    else if (w < __dmul_rn(aspect, h)) w = __dmul_rn(h, aspect);
    float sx = (float)__ddiv_rn(__dmul_rn(w, 1.0), pixel_std), sy = (float)__ddiv_rn(__dmul_rn(h, 1.0), pixel_std);
    if (cx != -1.0f) { sx = __fmul_rn(sx, 1.25f); sy = __fmul_rn(sy, 1.25f); }
    center[2 * b] = cx; center[2 * b + 1] = cy;
    scale[2 * b] = sx; scale[2 * b + 1] = sy;
}

// half_body_transform for one record: returns false where the reference returns (None, None)
__device__ __forceinline__ bool half_body_one(const double* __restrict__ jt, const double* __restrict__ vs, const uint8_t* __restrict__ upper,
                                              double randn, int J, float aspect, float pixel_std, float* c, float* sc) {
    int nu = 0, nl = 0;
    for (int j = 0; j < J; ++j)
        if (vs[3 * j] > 0) { if (upper[j]) ++nu; else ++nl; }
    // upper body if randn < 0.5 and it has > 2 joints, else the lower body if it has > 2, else the upper body
    const bool use_upper = (randn < 0.5 && nu > 2) || !(nl > 2);
    const int n = use_upper ? nu : nl;
    if (n < 2) return false;
    // np.array(selected, float32): mean over axis 0 adds the rows in order in float32, then divides by n
    float sx = 0.f, sy = 0.f, lox = 0.f, loy = 0.f, hix = 0.f, hiy = 0.f;
    bool first = true;
    for (int j = 0; j < J; ++j) {
        if (!(vs[3 * j] > 0) || (upper[j] != 0) != use_upper) continue;
        const float x = (float)jt[3 * j], y = (float)jt[3 * j + 1];
        if (first) { sx = x; sy = y; lox = hix = x; loy = hiy = y; first = false; }
        else {
            sx = __fadd_rn(sx, x); sy = __fadd_rn(sy, y);
            lox = fminf(lox, x); hix = fmaxf(hix, x); loy = fminf(loy, y); hiy = fmaxf(hiy, y);
        }
    }
    c[0] = __fdiv_rn(sx, (float)n); c[1] = __fdiv_rn(sy, (float)n);
    float w = __fsub_rn(hix, lox), h = __fsub_rn(hiy, loy);
    if (w > __fmul_rn(aspect, h)) h = __fdiv_rn(__fmul_rn(w, 1.0f), aspect);
    else if (w < __fmul_rn(aspect, h)) w = __fmul_rn(h, aspect);
    sc[0] = __fmul_rn(__fdiv_rn(__fmul_rn(w, 1.0f), pixel_std), 1.5f);
    sc[1] = __fmul_rn(__fdiv_rn(__fmul_rn(h, 1.0f), pixel_std), 1.5f);
    return true;
}

__global__ void half_body_kernel(const double* __restrict__ joints, const double* __restrict__ vis, const uint8_t* __restrict__ upper,
                                 const double* __restrict__ randn, float* __restrict__ center, float* __restrict__ scale,
                                 uint8_t* __restrict__ valid, int B, int J, float aspect, float pixel_std) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float c[2] = {0.f, 0.f}, sc[2] = {0.f, 0.f};
    const bool ok = half_body_one(joints + (int64_t)b * J * 3, vis + (int64_t)b * J * 3, upper, randn[b], J, aspect, pixel_std, c, sc);
    valid[b] = ok ? 1 : 0;
    center[2 * b] = ok ? c[0] : 0.f; center[2 * b + 1] = ok ? c[1] : 0.f;
    scale[2 * b] = ok ? sc[0] : 0.f; scale[2 * b + 1] = ok ? sc[1] : 0.f;
}

// The centre / scale bookkeeping of get_base / get_clean (JointsDataset.py:167-188) for a batch, given the draws:
// half-body box where the sample drew it (and it is valid), s = s * clip(randn*sf + 1, ...) (float32 array times a
// numpy float64 scalar: float64 under NEP 50), c[0] = width - c[0] - 1 for flipped samples (float32).
__global__ void base_cs_kernel(const float* __restrict__ rec_center, const float* __restrict__ rec_scale,
                               const double* __restrict__ joints, const double* __restrict__ vis, const uint8_t* __restrict__ upper,
                               const uint8_t* __restrict__ take_hb, const double* __restrict__ hb_randn,
                               const double* __restrict__ s_factor, const uint8_t* __restrict__ flip, const int32_t* __restrict__ src_w,
                               float* __restrict__ center, double* __restrict__ scale, int B, int J, float aspect, float pixel_std) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float c[2] = {rec_center[2 * b], rec_center[2 * b + 1]}, sc[2] = {rec_scale[2 * b], rec_scale[2 * b + 1]};
    if (take_hb && take_hb[b]) {
        float hc[2], hs[2];
        if (half_body_one(joints + (int64_t)b * J * 3, vis + (int64_t)b * J * 3, upper, hb_randn[b], J, aspect, pixel_std, hc, hs)) {
            c[0] = hc[0]; c[1] = hc[1]; sc[0] = hs[0]; sc[1] = hs[1];
        }
    }
    double s0 = (double)sc[0], s1 = (double)sc[1];
    if (s_factor) { s0 = __dmul_rn(s0, s_factor[b]); s1 = __dmul_rn(s1, s_factor[b]); }
    if (flip && flip[b]) c[0] = __fsub_rn(__fsub_rn((float)src_w[b], c[0]), 1.0f);
    center[2 * b] = c[0]; center[2 * b + 1] = c[1];
    scale[2 * b] = s0; scale[2 * b + 1] = s1;
}

__global__ void select_data_kernel(const double* __restrict__ joints, const double* __restrict__ vis, const float* __restrict__ center,
                                   const float* __restrict__ scale, uint8_t* __restrict__ keep, int B, int J, float pixel_std2) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* jt = joints + (int64_t)b * J * 3;
    const double* vs = vis + (int64_t)b * J * 3;
    int num_vis = 0;
    double jx = 0.0, jy = 0.0;
    for (int j = 0; j < J; ++j) {
        if (vs[3 * j] <= 0) continue;
        ++num_vis;
        jx = __dadd_rn(jx, jt[3 * j]); jy = __dadd_rn(jy, jt[3 * j + 1]);
    }
    if (num_vis == 0) { keep[b] = 0; return; }
    jx = __ddiv_rn(jx, (double)num_vis); jy = __ddiv_rn(jy, (double)num_vis);
    const float area = __fmul_rn(__fmul_rn(scale[2 * b], scale[2 * b + 1]), pixel_std2);      // float32 scalars x Python int
    const double dx = __dsub_rn(jx, (double)center[2 * b]), dy = __dsub_rn(jy, (double)center[2 * b + 1]);
    const double norm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    const float den = __fmul_rn((float)(0.2 * 0.2 * 2.0), area);                                 // Python float x float32 -> float32
    const double ks = exp(__ddiv_rn(__dmul_rn(-1.0, __dmul_rn(norm, norm)), (double)den));
    const double metric = (0.2 / 16) * num_vis + 0.45 - 0.2 / 16;
    keep[b] = ks > metric ? 1 : 0;
}

}  // namespace advmix

using namespace advmix;

extern "C" {

int advmix_xywh2cs(const double* boxes_xywh, float* center, float* scale, int B, double aspect_ratio, double pixel_std,
                   advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && aspect_ratio > 0 && pixel_std > 0, "xywh2cs: bad argument");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(boxes_xywh && center && scale, "xywh2cs: null argument");
    xywh2cs_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(boxes_xywh, center, scale, B, aspect_ratio, pixel_std);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_half_body_cs(const double* joints, const double* vis, const uint8_t* upper_body_mask, const double* randn_draw,
                        float* center, float* scale, uint8_t* valid, int B, int J, double aspect_ratio, double pixel_std,
                        advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0 && aspect_ratio > 0 && pixel_std > 0, "half_body_cs: bad argument");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(joints && vis && upper_body_mask && randn_draw && center && scale && valid, "half_body_cs: null argument");
    half_body_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(joints, vis, upper_body_mask, randn_draw, center, scale, valid,
                                                                     B, J, (float)aspect_ratio, (float)pixel_std);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_base_cs(const float* rec_center, const float* rec_scale, const double* joints, const double* vis,
                   const uint8_t* upper_body_mask, const uint8_t* take_half_body, const double* hb_randn,
                   const double* scale_factor, const uint8_t* flip_lr, const int32_t* src_w, float* center, double* scale,
                   int B, int J, double aspect_ratio, double pixel_std, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0 && aspect_ratio > 0 && pixel_std > 0, "base_cs: bad argument");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(rec_center && rec_scale && center && scale, "base_cs: null argument");
    ADVMIX_REQUIRE(!take_half_body || (joints && vis && upper_body_mask && hb_randn), "base_cs: half-body needs joints, vis, mask and draws");
    ADVMIX_REQUIRE(!flip_lr || src_w, "base_cs: flip needs src_w");
    base_cs_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(rec_center, rec_scale, joints, vis, upper_body_mask, take_half_body,
                                                                   hb_randn, scale_factor, flip_lr, src_w, center, scale, B, J,
                                                                   (float)aspect_ratio, (float)pixel_std);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_select_data(const double* joints, const double* vis, const float* center, const float* scale, uint8_t* keep, int B,
                       int J, double pixel_std, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0 && pixel_std > 0, "select_data: bad argument");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(joints && vis && center && scale && keep, "select_data: null argument");
    select_data_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(joints, vis, center, scale, keep, B, J,
                                                                       (float)(pixel_std * pixel_std));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"
