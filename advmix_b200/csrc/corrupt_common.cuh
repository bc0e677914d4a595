// Shared definitions for the 15 common + 4 validation imagecorruptions operators.
#pragma once
#include "common.cuh"

namespace advmix {

enum CorruptOp {
    C_GAUSSIAN_NOISE = 0, C_SHOT_NOISE, C_IMPULSE_NOISE, C_DEFOCUS_BLUR, C_GLASS_BLUR, C_MOTION_BLUR,
    C_ZOOM_BLUR, C_SNOW, C_FROST, C_FOG, C_BRIGHTNESS, C_CONTRAST, C_ELASTIC, C_PIXELATE, C_JPEG,
    C_SPECKLE_NOISE, C_GAUSSIAN_BLUR, C_SPATTER, C_SATURATE,      // the 'validation' set
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
    bool fast;           // ADVMIX_CORRUPT_FAST: float32 arithmetic where an op has such a kernel
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
int run_speckle_noise(const CorruptArgs&);
int run_gaussian_blur(const CorruptArgs&);
int run_spatter(const CorruptArgs&);
int run_saturate(const CorruptArgs&);

// ADVMIX_CORRUPT_FAST variants (corrupt_fast32*.cu): return -1 when the shape has no float32 kernel (caller falls back)
int run_defocus_blur_fast(const CorruptArgs&);
int run_motion_blur_fast(const CorruptArgs&);
int run_zoom_blur_fast(const CorruptArgs&);
int run_snow_fast(const CorruptArgs&);
int run_fog_fast(const CorruptArgs&);
int run_elastic_fast(const CorruptArgs&);

// advmix_corrupt_sweep_u8c3 (corrupt_sweep.cu): all five severities of base.op from one read of the crops; outs[s] is the
// output of severity s + 1 (base.out / base.severity are unused).  The fused launchers return -1 when the op / shape / mode
// has no fused kernel; the caller then runs the five per-severity launches.
struct SweepArgs {
    CorruptArgs base;
    uint8_t* outs[5];
};
struct Sweep5Out { uint8_t* p[5]; };
struct Sweep5F { float v[5]; };
struct Sweep5D { double v[5]; };
struct Sweep5U { uint32_t v[5]; };
int run_zoom_blur_sweep_fast(const SweepArgs&);
int run_gaussian_noise_sweep(const SweepArgs&);
int run_impulse_noise_sweep(const SweepArgs&);
int run_contrast_sweep(const SweepArgs&);
int run_brightness_sweep(const SweepArgs&);
int run_frost_sweep(const SweepArgs&);
int run_fog_pair_fast(const CorruptArgs&, const float* field, uint8_t* out1, uint8_t* out2);   // severities 1 + 2 (same plasma map)
int run_elastic_sweep_fast(const SweepArgs&, const float* shared_field);   // base.field_bytes = per-image stride of shared_field

// float32 separable Gaussian on uint8 HWC images, radius 3 / 4 / 6 / 8, W % 4 == 0 (corrupt_fast32c.cu); -1 otherwise
int launch_gauss_u8_fast(const uint8_t* in, const int32_t* in_idx, uint8_t* out, const int32_t* out_idx, int n, int H, int W,
                         int radius, const double* d_w, float top255, cudaStream_t s);

// materialise the perf-mode draws of `a` (same values the in-register path would use)
int launch_fill_rand(const CorruptArgs& a, void* field, double* param);

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
// ---- perf-mode noise draws: 16 random bits per value, 8 values per Philox call ----------------------
// exact float of a 16-bit integer without the conversion pipe: 2^23 + k is exactly representable
__device__ __forceinline__ float u16_to_float(uint32_t k) { return __uint_as_float(0x4B000000u | k) - 8388608.0f; }
// float(byte k of w), k a compile-time constant after unrolling: ONE PRMT drops the byte into the mantissa of 2^23, one FADD removes 2^23
__device__ __forceinline__ float byte_of_word_f(uint32_t w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)k)) - 8388608.0f; }
// Box-Muller from one u32: low half -> radius, high half -> angle.  Approximate SFU intrinsics (lg2, rsq,
// sin, cos); the fill kernel calls the same function, so dumped and in-register normals are identical.
__device__ __forceinline__ float2 box_muller16(uint32_t w) {
    const float u1 = fmaf(u16_to_float(w & 0xFFFFu), 1.0f / 65536.0f, 0.5f / 65536.0f);   // (0,1)
    const float ang = u16_to_float(w >> 16) * (6.283185307179586f / 65536.0f);          // [0, 2pi)
    const float x = -1.3862943611198906f * __log2f(u1);                                 // -2 ln(u1) > 0
    const float r = x * rsqrtf(x);
    return make_float2(r * __cosf(ang), r * __sinf(ang));
}
// normals for elements 8b..8b+7 of the field
__device__ __forceinline__ void noise_normal8(const SampleRng& r, uint32_t tag, uint64_t b, float (&n)[8]) {
    const uint4 u = r.quad(tag, b);
    const float2 p0 = box_muller16(u.x), p1 = box_muller16(u.y), p2 = box_muller16(u.z), p3 = box_muller16(u.w);
    n[0] = p0.x; n[1] = p0.y; n[2] = p1.x; n[3] = p1.y; n[4] = p2.x; n[5] = p2.y; n[6] = p3.x; n[7] = p3.y;
}
// 16-bit draws for elements 8b..8b+7 (uniform = k/65536 for shot noise; impulse: low 15 bits flip, bit 15 salt)
__device__ __forceinline__ void noise_bits8(const SampleRng& r, uint32_t tag, uint64_t b, uint32_t (&k)[8]) {
    const uint4 u = r.quad(tag, b);
    k[0] = u.x & 0xFFFFu; k[1] = u.x >> 16; k[2] = u.y & 0xFFFFu; k[3] = u.y >> 16;
    k[4] = u.z & 0xFFFFu; k[5] = u.z >> 16; k[6] = u.w & 0xFFFFu; k[7] = u.w >> 16;
}

// glass-blur offsets for cell e = (iter*H + h)*W + w : (dx, dy) in [-delta, delta-1]
__device__ __forceinline__ int2 field_glass(const int8_t* inj, const SampleRng& r, uint64_t e, int delta) {
    if (inj) return make_int2(inj[2 * e], inj[2 * e + 1]);
    // one Philox block per four cells, 16 bits per draw: floor(u16 * 2*delta / 2^16)
    const uint4 u = r.quad(TAG_GLASS, e >> 2);
    const uint32_t k = (uint32_t)e & 3u;
    const uint32_t wd = k == 0 ? u.x : (k == 1 ? u.y : (k == 2 ? u.z : u.w));
    return make_int2(-delta + (int)(((wd & 0xFFFFu) * (2u * delta)) >> 16), -delta + (int)(((wd >> 16) * (2u * delta)) >> 16));
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
