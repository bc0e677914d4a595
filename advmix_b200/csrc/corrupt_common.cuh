// Shared definitions for the 15 imagecorruptions operators.
#pragma once
#include "common.cuh"

namespace advmix {

enum CorruptOp {
    C_GAUSSIAN_NOISE = 0, C_SHOT_NOISE, C_IMPULSE_NOISE, C_DEFOCUS_BLUR, C_GLASS_BLUR, C_MOTION_BLUR,
    C_ZOOM_BLUR, C_SNOW, C_FROST, C_FOG, C_BRIGHTNESS, C_CONTRAST, C_ELASTIC, C_PIXELATE, C_JPEG,
    C_NUM_OPS
};

struct CorruptArgs {
    int op, severity;
    const uint8_t* in;
    uint8_t* out;
    int n;
    const int32_t* idx;
    int H, W;
    const void* rand_field;
    const double* rand_param;
    uint64_t seed;
    int64_t sample_base;
    const uint8_t* frost_bank;
    int frost_n, frost_h, frost_w;
    void* ws;
    size_t ws_bytes;
    cudaStream_t stream;
    size_t field_bytes;  // per-image stride of rand_field
};

// per-op launchers (each returns an ADVMIX_* code)
int run_gaussian_noise(const CorruptArgs&);
int run_shot_noise(const CorruptArgs&);
int run_impulse_noise(const CorruptArgs&);
int run_defocus_blur(const CorruptArgs&);
int run_glass_blur(const CorruptArgs&);
int run_motion_blur(const CorruptArgs&);
int run_zoom_blur(const CorruptArgs&);
int run_snow(const CorruptArgs&);
int run_frost(const CorruptArgs&);
int run_fog(const CorruptArgs&);
int run_brightness(const CorruptArgs&);
int run_contrast(const CorruptArgs&);
int run_elastic(const CorruptArgs&);
int run_pixelate(const CorruptArgs&);
int run_jpeg(const CorruptArgs&);

size_t ws_bytes_for(int op, int severity, int n, int H, int W);
size_t field_bytes_for(int op, int severity, int H, int W);

static inline int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// severity tables (index severity-1)
__host__ __device__ inline double sev_gaussian_noise(int s) { const double c[5] = {0.08, 0.12, 0.18, 0.26, 0.38}; return c[s - 1]; }
__host__ __device__ inline double sev_shot_noise(int s) { const double c[5] = {60, 25, 12, 5, 3}; return c[s - 1]; }
__host__ __device__ inline double sev_impulse_noise(int s) { const double c[5] = {0.03, 0.06, 0.09, 0.17, 0.27}; return c[s - 1]; }
__host__ __device__ inline int glass_delta(int s) { const int c[5] = {1, 2, 2, 3, 4}; return c[s - 1]; }
__host__ __device__ inline int glass_iters(int s) { const int c[5] = {2, 1, 3, 2, 2}; return c[s - 1]; }
__host__ __device__ inline double glass_sigma(int s) { const double c[5] = {0.7, 0.9, 1.0, 1.1, 1.5}; return c[s - 1]; }

// ---- image / slot addressing --------------------------------------------------------
__device__ __forceinline__ int slot_of(const int32_t* idx, int i) { return idx ? idx[i] : i; }

// ---- random draws: identical code path for the fill kernel and the fused kernels --------
struct SampleRng {
    Philox ph;
    uint32_t s_lo, s_hi;
    __device__ SampleRng(uint64_t seed, int64_t sample) : ph(seed), s_lo((uint32_t)sample), s_hi((uint32_t)((uint64_t)sample >> 32)) {}
    // 4 raw u32 for quad q of field `tag`
    __device__ __forceinline__ uint4 quad(uint32_t tag, uint64_t q) const {
        return ph((uint32_t)q, (uint32_t)(q >> 32) ^ (s_hi << 16), s_lo, tag);
    }
};

// field element accessors: element index e within the image's field
__device__ __forceinline__ float4 field_normal4(const float* inj, const SampleRng& r, uint32_t tag, uint64_t q) {
    if (inj) return *reinterpret_cast<const float4*>(inj + 4 * q);
    const uint4 u = r.quad(tag, q);
    const float2 a = box_muller(u.x, u.y), b = box_muller(u.z, u.w);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 field_uniform4(const float* inj, const SampleRng& r, uint32_t tag, uint64_t q) {
    if (inj) return *reinterpret_cast<const float4*>(inj + 4 * q);
    const uint4 u = r.quad(tag, q);
    return make_float4(u01(u.x), u01(u.y), u01(u.z), u01(u.w));
}
__device__ __forceinline__ float field_normal1(const float* inj, const SampleRng& r, uint32_t tag, uint64_t e) {
    if (inj) return inj[e];
    const uint4 u = r.quad(tag, e >> 2);
    const int l = (int)(e & 3);
    const float2 p = (l < 2) ? box_muller(u.x, u.y) : box_muller(u.z, u.w);
    return (l & 1) ? p.y : p.x;
}
__device__ __forceinline__ float field_uniform1(const float* inj, const SampleRng& r, uint32_t tag, uint64_t e) {
    if (inj) return inj[e];
    const uint4 u = r.quad(tag, e >> 2);
    const int l = (int)(e & 3);
    return u01(l == 0 ? u.x : l == 1 ? u.y : l == 2 ? u.z : u.w);
}
// glass-blur offsets for cell e = (iter*H + h)*W + w : (dx, dy) in [-delta, delta-1]
__device__ __forceinline__ int2 field_glass(const int8_t* inj, const SampleRng& r, uint64_t e, int delta) {
    if (inj) return make_int2(inj[2 * e], inj[2 * e + 1]);
    const uint4 u = r.quad(TAG_GLASS, e >> 1);
    const uint32_t a = (e & 1) ? u.z : u.x, b = (e & 1) ? u.w : u.y;
    return make_int2(-delta + (int)__umulhi(a, 2u * delta), -delta + (int)__umulhi(b, 2u * delta));
}
// per-sample scalar: uniform(lo, hi) = lo + (hi-lo)*u  (numpy's formula), float64
__device__ __forceinline__ double param_uniform(const double* inj, const SampleRng& r, double lo, double hi) {
    if (inj) return inj[0];
    const uint4 u = r.quad(TAG_PARAM, 0);
    return lo + (hi - lo) * (double)u01(u.x);
}

// uint8 <- float64 value already clipped to [0,255] (np.uint8() = C truncation)
__device__ __forceinline__ uint8_t trunc_u8(double v) { return (uint8_t)__double2int_rz(v); }
__device__ __forceinline__ double clip01(double v) { return fmin(fmax(v, 0.0), 1.0); }

// 256-entry table v/255. in float64 (each CTA fills its own copy: 256 divisions)
__device__ __forceinline__ void fill_div255(double* tab) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = __ddiv_rn((double)i, 255.0);
}

}  // namespace advmix
