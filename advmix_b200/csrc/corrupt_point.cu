// Point-wise and resampling corruptions: gaussian_noise, shot_noise, impulse_noise, frost,
// brightness, contrast, pixelate.  All are HBM-streaming kernels: one 12/16-byte group per
// thread iteration, 2 B of compulsory traffic per channel value.
#include "corrupt_common.cuh"

#include <cmath>
#include <vector>

namespace advmix {

constexpr int PT_THREADS = 256;

static inline dim3 point_grid(int64_t work_per_image, int n) {
    int64_t bx = (work_per_image + PT_THREADS - 1) / PT_THREADS;
    int64_t cap = std::max<int64_t>(1, ((int64_t)sm_count() * 8 + n - 1) / n);
    return dim3((unsigned)std::min(bx, cap), (unsigned)n);
}

__device__ __forceinline__ uint32_t pack4(uint8_t a, uint8_t b, uint8_t c, uint8_t d) {
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}

// ---- gaussian_noise: clip(x/255 + c*N(0,1), 0, 1)*255, float64 -------------------------------
// ---- speckle_noise (SPECKLE): clip(x + x*(c*N(0,1)), 0, 1)*255 --------------------------------
template <bool SPECKLE>
__global__ void __launch_bounds__(PT_THREADS)
gaussian_noise_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                      const float* __restrict__ field, size_t field_stride, uint64_t seed, int64_t sample_base,
                      int64_t quads, double c) {
    __shared__ double d255[256];
    fill_div255(d255);
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride) : nullptr;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)slot * quads * 4);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t)slot * quads * 4);
    for (int64_t q = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; q < quads; q += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w = __ldg(src + q);
        float nf[4];
        if (inj) {
            const float4 nz = *reinterpret_cast<const float4*>(inj + 4 * q);
            nf[0] = nz.x; nf[1] = nz.y; nf[2] = nz.z; nf[3] = nz.w;
        } else {
            float n8[8];
            noise_normal8(rng, TAG_FIELD0, (uint64_t)q >> 1, n8);
#pragma unroll
            for (int k = 0; k < 4; ++k) nf[k] = (q & 1) ? n8[4 + k] : n8[k];
        }
        uint8_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double x = d255[(w >> (8 * k)) & 255];
            const double nz = (double)nf[k] * c;
            o[k] = trunc_u8(clip01(SPECKLE ? x + x * nz : x + nz) * 255.0);
        }
        dst[q] = pack4(o[0], o[1], o[2], o[3]);
    }
}

// ---- shot_noise: inverse-CDF Poisson(lam = x/255*c) from a uniform --------------------------
constexpr int POISSON_KMAX = 128;

__global__ void __launch_bounds__(PT_THREADS)
shot_noise_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                  const float* __restrict__ field, size_t field_stride, uint64_t seed, int64_t sample_base,
                  int64_t quads, const double* __restrict__ cdf, const uint8_t* __restrict__ kout) {
    __shared__ uint8_t s_kout[POISSON_KMAX];
    if (threadIdx.x < POISSON_KMAX) s_kout[threadIdx.x] = kout[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride) : nullptr;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)slot * quads * 4);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t)slot * quads * 4);
    for (int64_t q = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; q < quads; q += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w = __ldg(src + q);
        float uf[4];
        if (inj) {
            const float4 uz = *reinterpret_cast<const float4*>(inj + 4 * q);
            uf[0] = uz.x; uf[1] = uz.y; uf[2] = uz.z; uf[3] = uz.w;
        } else {
            uint32_t k8[8];
            noise_bits8(rng, TAG_FIELD0, (uint64_t)q >> 1, k8);
#pragma unroll
            for (int k = 0; k < 4; ++k) uf[k] = (float)((q & 1) ? k8[4 + k] : k8[k]) * (1.0f / 65536.0f);
        }
        uint8_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double* row = cdf + ((w >> (8 * k)) & 255) * POISSON_KMAX;
            const double u = (double)uf[k];
            // smallest j with u < row[j]  (count of entries <= u), capped at KMAX-1
            int lo = 0, hi = POISSON_KMAX;
#pragma unroll
            for (int it = 0; it < 8; ++it) {   // 129 possible answers -> 8 halvings
                const int mid = min((lo + hi) >> 1, POISSON_KMAX - 1);
                if (lo < hi) { if (__ldg(row + mid) <= u) lo = mid + 1; else hi = mid; }
            }
            o[k] = s_kout[min(lo, POISSON_KMAX - 1)];
        }
        dst[q] = pack4(o[0], o[1], o[2], o[3]);
    }
}

// ---- impulse_noise: skimage random_noise 's&p' ---------------------------------------------
__global__ void __launch_bounds__(PT_THREADS)
impulse_noise_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                     const float* __restrict__ field, size_t field_stride, uint64_t seed, int64_t sample_base,
                     int64_t quads, double c) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const float* inj0 = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride) : nullptr;
    const float* inj1 = inj0 ? inj0 + quads * 4 : nullptr;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)slot * quads * 4);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t)slot * quads * 4);
    for (int64_t q = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; q < quads; q += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w = __ldg(src + q);
        float fa[4], fb[4];
        if (inj0) {
            const float4 a = *reinterpret_cast<const float4*>(inj0 + 4 * q), b = *reinterpret_cast<const float4*>(inj1 + 4 * q);
            fa[0] = a.x; fa[1] = a.y; fa[2] = a.z; fa[3] = a.w; fb[0] = b.x; fb[1] = b.y; fb[2] = b.z; fb[3] = b.w;
        } else {
            uint32_t k8[8];
            noise_bits8(rng, TAG_FIELD0, (uint64_t)q >> 1, k8);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t kk = (q & 1) ? k8[4 + k] : k8[k];
                fa[k] = (float)(kk & 0x7FFFu) * (1.0f / 32768.0f);
                fb[k] = (kk & 0x8000u) ? 0.25f : 0.75f;
            }
        }
        uint8_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t v = (w >> (8 * k)) & 255;   // (v/255.)*255 truncates back to v for all v
            o[k] = ((double)fa[k] < c) ? (fb[k] < 0.5f ? 255 : 0) : v;
        }
        dst[q] = pack4(o[0], o[1], o[2], o[3]);
    }
}

// ---- frost: clip(c0*img + c1*texture_crop, 0, 255) -------------------------------------------
__global__ void __launch_bounds__(PT_THREADS)
frost_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
             const double* __restrict__ param, uint64_t seed, int64_t sample_base, int H, int W,
             const uint8_t* __restrict__ bank, int fn, int fh, int fw, double c0, double c1) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    int tex, xs, ys;
    if (param) {
        tex = (int)param[4 * i]; xs = (int)param[4 * i + 1]; ys = (int)param[4 * i + 2];
    } else {
        const SampleRng rng(seed, sample_base + slot);
        const uint4 u = rng.quad(TAG_PARAM, 0);
        tex = (int)__umulhi(u.x, (uint32_t)min(5, fn));
        xs = fh > H ? (int)__umulhi(u.y, (uint32_t)(fh - H)) : 0;
        ys = fw > W ? (int)__umulhi(u.z, (uint32_t)(fw - W)) : 0;
    }
    const uint8_t* t = bank + ((int64_t)tex * fh + xs) * fw * 3 + (int64_t)ys * 3;
    const int64_t row = (int64_t)W * 3;
    const uint8_t* src = in + (int64_t)slot * H * row;
    uint8_t* dst = out + (int64_t)slot * H * row;
    const int64_t quads = H * row / 4;
    for (int64_t q = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; q < quads; q += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src) + q);
        uint8_t o[4];
        // row / column of the quad's first byte once (32-bit; the images are far below 2 GB), then step: one 64-bit
        // division per BYTE made this kernel divider-bound
        int y = (int)((uint32_t)(q * 4) / (uint32_t)row);
        int r = (int)((uint32_t)(q * 4) - (uint32_t)y * (uint32_t)row);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k > 0 && ++r == (int)row) { r = 0; ++y; }
            const double f = (double)__ldg(t + (int64_t)y * fw * 3 + r);
            const double v = c0 * (double)((w >> (8 * k)) & 255) + c1 * f;
            o[k] = trunc_u8(fmin(fmax(v, 0.0), 255.0));
        }
        reinterpret_cast<uint32_t*>(dst)[q] = pack4(o[0], o[1], o[2], o[3]);
    }
}

// ---- brightness: skimage rgb2hsv -> v += c -> hsv2rgb, float64 ---------------------------------
// ---- saturate (SAT): s = clip(s*c + c1, 0, 1) instead ------------------------------------------
// skimage rgb2hsv of one pixel (float64): value, saturation, hue in [0, 1)
__device__ __forceinline__ void rgb2hsv_px(double r, double g, double b, double& v, double& s, double& h) {
    v = fmax(fmax(r, g), b);
    const double delta = v - fmin(fmin(r, g), b);
    s = 0.0; h = 0.0;
    if (delta != 0.0) {
        s = delta / v;
        // skimage assigns the three cases one after the other, so on ties the LAST matching channel wins: pick the
        // numerator first and divide once (same operations on the selected branch, two float64 divisions fewer)
        double num = g - b, off = 0.0;
        if (g == v) { num = b - r; off = 2.0; }
        if (b == v) { num = r - g; off = 4.0; }
        h = num / delta;
        if (off != 0.0) h = off + h;
        {   // h / 6.0, correctly rounded without the division sequence: y = RN(1/6), q0 = RN(h*y), q = RN(q0 + (h - 6*q0)*y)
            const double y6 = 1.0 / 6.0, q0 = h * y6;
            h = fma(fma(-q0, 6.0, h), y6, q0);
        }
        h = h - trunc(h);            // fmod(h, 1.0)
        if (h < 0.0) h = h + 1.0;    // numpy's floored modulo
    }
}

template <bool SAT>
__device__ __forceinline__ void brightness_px(double r, double g, double b, double c, double c1, uint8_t* o) {
    double v, s, h;
    rgb2hsv_px(r, g, b, v, s, h);
    const double v2 = SAT ? v : clip01(v + c);
    if (SAT) s = clip01(s * c + c1);
    const double h6 = h * 6.0;
    const double hi = floor(h6);
    const double f = h6 - hi;
    const double p = v2 * (1.0 - s);
    const double q = v2 * (1.0 - f * s);
    const double t = v2 * (1.0 - (1.0 - f) * s);
    const int sel = ((int)hi) % 6;
    double R, G, B;
    switch (sel) {
        case 0: R = v2; G = t; B = p; break;
        case 1: R = q; G = v2; B = p; break;
        case 2: R = p; G = v2; B = t; break;
        case 3: R = p; G = q; B = v2; break;
        case 4: R = t; G = p; B = v2; break;
        default: R = v2; G = p; B = q; break;
    }
    o[0] = trunc_u8(clip01(R) * 255.0);
    o[1] = trunc_u8(clip01(G) * 255.0);
    o[2] = trunc_u8(clip01(B) * 255.0);
}

template <bool SAT>
__global__ void __launch_bounds__(PT_THREADS)
brightness_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                  int64_t groups, double c, double c1) {
    __shared__ double d255[256];
    fill_div255(d255);
    __syncthreads();
    const int slot = slot_of(idx, blockIdx.y);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)slot * groups * 12);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t)slot * groups * 12);
    for (int64_t g = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w0 = __ldg(src + 3 * g), w1 = __ldg(src + 3 * g + 1), w2 = __ldg(src + 3 * g + 2);
        const uint8_t v[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                               (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                               (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
        uint8_t o[12];
#pragma unroll
        for (int p = 0; p < 4; ++p) brightness_px<SAT>(d255[v[3 * p]], d255[v[3 * p + 1]], d255[v[3 * p + 2]], c, c1, o + 3 * p);
        dst[3 * g] = pack4(o[0], o[1], o[2], o[3]);
        dst[3 * g + 1] = pack4(o[4], o[5], o[6], o[7]);
        dst[3 * g + 2] = pack4(o[8], o[9], o[10], o[11]);
    }
}

// brightness, five severities (advmix_corrupt_sweep_u8c3): rgb2hsv and the severity-independent factors of hsv2rgb once per
// pixel; per severity the operations of brightness_px<false> on the same values.
__global__ void __launch_bounds__(PT_THREADS)
brightness_sweep_kernel(const uint8_t* __restrict__ in, Sweep5Out outs, const int32_t* __restrict__ idx, int64_t groups, Sweep5D cs) {
    __shared__ double d255[256];
    fill_div255(d255);
    __syncthreads();
    const int slot = slot_of(idx, blockIdx.y);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)slot * groups * 12);
    for (int64_t g = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w0 = __ldg(src + 3 * g), w1 = __ldg(src + 3 * g + 1), w2 = __ldg(src + 3 * g + 2);
        const uint8_t vb[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                                (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                                (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
        uint32_t ow[5][3] = {};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            double v, s, h;
            rgb2hsv_px(d255[vb[3 * p]], d255[vb[3 * p + 1]], d255[vb[3 * p + 2]], v, s, h);
            const double h6 = h * 6.0;
            const double hi = floor(h6);
            const double f = h6 - hi;
            const double ap = (1.0 - s), aq = (1.0 - f * s), at = (1.0 - (1.0 - f) * s);
            const int sel = ((int)hi) % 6;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const double v2 = clip01(v + cs.v[k]);
                const double pp = v2 * ap, q = v2 * aq, t = v2 * at;
                double R, G, B;
                switch (sel) {
                    case 0: R = v2; G = t; B = pp; break;
                    case 1: R = q; G = v2; B = pp; break;
                    case 2: R = pp; G = v2; B = t; break;
                    case 3: R = pp; G = q; B = v2; break;
                    case 4: R = t; G = pp; B = v2; break;
                    default: R = v2; G = pp; B = q; break;
                }
                ow[k][(3 * p) >> 2] |= (uint32_t)trunc_u8(clip01(R) * 255.0) << (8 * ((3 * p) & 3));
                ow[k][(3 * p + 1) >> 2] |= (uint32_t)trunc_u8(clip01(G) * 255.0) << (8 * ((3 * p + 1) & 3));
                ow[k][(3 * p + 2) >> 2] |= (uint32_t)trunc_u8(clip01(B) * 255.0) << (8 * ((3 * p + 2) & 3));
            }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(outs.p[k] + (int64_t)slot * groups * 12);
            dst[3 * g] = ow[k][0]; dst[3 * g + 1] = ow[k][1]; dst[3 * g + 2] = ow[k][2];
        }
    }
}

// frost, five severities: the texture crop (same draw for every severity) is read once
__global__ void __launch_bounds__(PT_THREADS)
frost_sweep_kernel(const uint8_t* __restrict__ in, Sweep5Out outs, const int32_t* __restrict__ idx, uint64_t seed, int64_t sample_base,
                   int H, int W, const uint8_t* __restrict__ bank, int fn, int fh, int fw, Sweep5D c0, Sweep5D c1) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const uint4 u = rng.quad(TAG_PARAM, 0);
    const int tex = (int)__umulhi(u.x, (uint32_t)min(5, fn));
    const int xs = fh > H ? (int)__umulhi(u.y, (uint32_t)(fh - H)) : 0;
    const int ys = fw > W ? (int)__umulhi(u.z, (uint32_t)(fw - W)) : 0;
    const uint8_t* t = bank + ((int64_t)tex * fh + xs) * fw * 3 + (int64_t)ys * 3;
    const int64_t row = (int64_t)W * 3;
    const uint8_t* src = in + (int64_t)slot * H * row;
    const int64_t quads = H * row / 4;
    for (int64_t q = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; q < quads; q += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src) + q);
        int y = (int)((uint32_t)(q * 4) / (uint32_t)row);
        int r = (int)((uint32_t)(q * 4) - (uint32_t)y * (uint32_t)row);
        double f[4], x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k > 0 && ++r == (int)row) { r = 0; ++y; }
            f[k] = (double)__ldg(t + (int64_t)y * fw * 3 + r);
            x[k] = (double)((w >> (8 * k)) & 255);
        }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            uint8_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = trunc_u8(fmin(fmax(c0.v[s] * x[k] + c1.v[s] * f[k], 0.0), 255.0));
            reinterpret_cast<uint32_t*>(outs.p[s] + (int64_t)slot * H * row)[q] = pack4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- contrast: per-image channel means (pass 1) then (x-m)*c+m (pass 2) --------------------------
__global__ void __launch_bounds__(PT_THREADS)
channel_sum_kernel(const uint8_t* __restrict__ in, const int32_t* __restrict__ idx, int64_t groups,
                   unsigned long long* __restrict__ sums) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)slot * groups * 12);
    uint32_t s0 = 0, s1 = 0, s2 = 0;
    for (int64_t g = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * PT_THREADS) {
        const uint32_t w0 = __ldg(src + 3 * g), w1 = __ldg(src + 3 * g + 1), w2 = __ldg(src + 3 * g + 2);
        s0 += (w0 & 255) + (w0 >> 24) + ((w1 >> 16) & 255) + ((w2 >> 8) & 255);
        s1 += ((w0 >> 8) & 255) + (w1 & 255) + (w1 >> 24) + ((w2 >> 16) & 255);
        s2 += ((w0 >> 16) & 255) + ((w1 >> 8) & 255) + (w2 & 255) + (w2 >> 24);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[3 * i], (unsigned long long)s0);
        atomicAdd(&sums[3 * i + 1], (unsigned long long)s1);
        atomicAdd(&sums[3 * i + 2], (unsigned long long)s2);
    }
}

__global__ void __launch_bounds__(PT_THREADS)
contrast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                int64_t groups, const unsigned long long* __restrict__ sums, double npix, double c) {
    __shared__ double d255[256];
    __shared__ double mean[3];
    fill_div255(d255);
    const int i = blockIdx.y, slot = slot_of(idx, i);
    if (threadIdx.x < 3) mean[threadIdx.x] = ((double)sums[3 * i + threadIdx.x] / 255.0) / npix;
    __syncthreads();
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)slot * groups * 12);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t)slot * groups * 12);
    const double m0 = mean[0], m1 = mean[1], m2 = mean[2];
    for (int64_t g = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * PT_THREADS) {
        uint32_t w[3] = {__ldg(src + 3 * g), __ldg(src + 3 * g + 1), __ldg(src + 3 * g + 2)};
        uint32_t r[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint8_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = (4 * k + j) % 3;
                const double m = ch == 0 ? m0 : (ch == 1 ? m1 : m2);
                const double x = d255[(w[k] >> (8 * j)) & 255];
                o[j] = trunc_u8(clip01((x - m) * c + m) * 255.0);
            }
            r[k] = pack4(o[0], o[1], o[2], o[3]);
        }
        dst[3 * g] = r[0]; dst[3 * g + 1] = r[1]; dst[3 * g + 2] = r[2];
    }
}

// ---- pixelate: PIL resize BOX (8bpc fixed point, horizontal then vertical) + NEAREST up ------
struct ResampleTab {          // device layout: bounds[2*out], then kk[out*ksize]
    std::vector<int32_t> data;
    int out, ksize;
};

static ResampleTab box_coeffs(int in_size, int out_size) {
    // Pillow precompute_coeffs() + normalize_coeffs_8bpc(), box filter (support 0.5)
    const int PRECISION_BITS = 32 - 8 - 2;
    double scale = (double)((float)in_size - 0.0f) / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 0.5 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    ResampleTab t;
    t.out = out_size;
    t.ksize = ksize;
    t.data.assign((size_t)2 * out_size + (size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double a = (x + xmin - center + 0.5) * ss;
            const double w = (a > -0.5 && a <= 0.5) ? 1.0 : 0.0;
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        t.data[2 * xx] = xmin;
        t.data[2 * xx + 1] = xmax;
        for (int x = 0; x < xmax; ++x) {
            const double v = k[x] * (1 << PRECISION_BITS);
            t.data[(size_t)2 * out_size + (size_t)xx * ksize + x] = (int)(v < 0 ? -0.5 + v : 0.5 + v);
        }
    }
    return t;
}

static std::vector<int32_t> nearest_tab(int in_size, int out_size) {
    // Pillow ImagingScaleAffine: xo = a0*0.5; xin = (int)xo; xo += a0 (incremental, float64)
    std::vector<int32_t> t(out_size);
    const double a0 = (double)in_size / out_size;
    double xo = 0.0 + a0 * 0.5;
    for (int x = 0; x < out_size; ++x) {
        int xin = xo < 0.0 ? -1 : (int)xo;
        if (xin < 0) xin = 0;
        if (xin >= in_size) xin = in_size - 1;
        t[x] = xin;
        xo += a0;
    }
    return t;
}

__device__ __forceinline__ uint8_t clip8_22(int v) {
    v >>= 22;
    return (uint8_t)max(0, min(255, v));
}

// horizontal: in [H][W][3] -> tmp [H][w2][3]
__global__ void __launch_bounds__(PT_THREADS)
pix_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ tmp, const int32_t* __restrict__ idx,
             const int32_t* __restrict__ tab, int H, int W, int w2, int ksize) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = tmp + (int64_t)i * H * w2 * 3;
    const int64_t total = (int64_t)H * w2;
    for (int64_t t = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * PT_THREADS) {
        const int y = (int)((uint32_t)t / (uint32_t)w2), xx = (int)((uint32_t)t - (uint32_t)y * (uint32_t)w2);
        const int xmin = tab[2 * xx], xmax = tab[2 * xx + 1];
        const int32_t* k = tab + 2 * w2 + xx * ksize;
        int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
        const uint8_t* p = src + ((int64_t)y * W + xmin) * 3;
        for (int x = 0; x < xmax; ++x) {
            const int kv = k[x];
            s0 += p[3 * x] * kv; s1 += p[3 * x + 1] * kv; s2 += p[3 * x + 2] * kv;
        }
        uint8_t* o = dst + t * 3;
        o[0] = clip8_22(s0); o[1] = clip8_22(s1); o[2] = clip8_22(s2);
    }
}

// vertical: tmp [H][w2][3] -> small [h2][w2][3]
__global__ void __launch_bounds__(PT_THREADS)
pix_v_kernel(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ small, const int32_t* __restrict__ tab,
             int H, int w2, int h2, int ksize) {
    const int i = blockIdx.y;
    const uint8_t* src = tmp + (int64_t)i * H * w2 * 3;
    uint8_t* dst = small + (int64_t)i * h2 * w2 * 3;
    const int64_t total = (int64_t)h2 * w2 * 3;
    const int64_t row = (int64_t)w2 * 3;
    for (int64_t t = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * PT_THREADS) {
        const int yy = (int)((uint32_t)t / (uint32_t)row);
        const int64_t r = (int64_t)((uint32_t)t - (uint32_t)yy * (uint32_t)row);
        const int ymin = tab[2 * yy], ymax = tab[2 * yy + 1];
        const int32_t* k = tab + 2 * h2 + yy * ksize;
        int s = 1 << 21;
        for (int y = 0; y < ymax; ++y) s += src[(int64_t)(y + ymin) * row + r] * k[y];
        dst[t] = clip8_22(s);
    }
}

// nearest upsample: out[y][x] = small[yin[y]][xin[x]]
__global__ void __launch_bounds__(PT_THREADS)
pix_up_kernel(const uint8_t* __restrict__ small, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
              const int32_t* __restrict__ xin, const int32_t* __restrict__ yin, int H, int W, int h2, int w2) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = small + (int64_t)i * h2 * w2 * 3;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t)slot * H * W * 3);
    const int64_t groups = (int64_t)H * W / 4;
    for (int64_t g = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * PT_THREADS) {
        uint8_t o[12];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t pix = g * 4 + p;
            const int y = (int)((uint32_t)pix / (uint32_t)W), x = (int)((uint32_t)pix - (uint32_t)y * (uint32_t)W);
            const uint8_t* s = src + ((int64_t)yin[y] * w2 + xin[x]) * 3;
            o[3 * p] = s[0]; o[3 * p + 1] = s[1]; o[3 * p + 2] = s[2];
        }
        dst[3 * g] = pack4(o[0], o[1], o[2], o[3]);
        dst[3 * g + 1] = pack4(o[4], o[5], o[6], o[7]);
        dst[3 * g + 2] = pack4(o[8], o[9], o[10], o[11]);
    }
}

// ------------------------------------------------------------------------------- launchers
static const float* inj_field(const CorruptArgs& a) { return reinterpret_cast<const float*>(a.rand_field); }

// perf-mode / table kernels (corrupt_fast.cu)
bool fast_ok(const CorruptArgs& a);
int run_gaussian_noise_fast(const CorruptArgs& a);
int run_impulse_noise_fast(const CorruptArgs& a);
int run_shot_noise_table(const CorruptArgs& a);
int run_contrast_fast(const CorruptArgs& a, unsigned long long* sums);
bool fast48_ok(const CorruptArgs& a);

int run_gaussian_noise(const CorruptArgs& a) {
    if (a.fast && !a.rand_field && fast_ok(a)) return run_gaussian_noise_fast(a);   // float32, in-register draws
    const int64_t quads = (int64_t)a.H * a.W * 3 / 4;
    gaussian_noise_kernel<false><<<point_grid(quads, a.n), PT_THREADS, 0, a.stream>>>(
        a.in, a.out, a.idx, inj_field(a), a.field_bytes, a.seed, a.sample_base, quads, sev_gaussian_noise(a.severity));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_speckle_noise(const CorruptArgs& a) {
    const double c[5] = {0.15, 0.2, 0.35, 0.45, 0.6};
    const int64_t quads = (int64_t)a.H * a.W * 3 / 4;
    gaussian_noise_kernel<true><<<point_grid(quads, a.n), PT_THREADS, 0, a.stream>>>(
        a.in, a.out, a.idx, inj_field(a), a.field_bytes, a.seed, a.sample_base, quads, c[a.severity - 1]);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_shot_noise(const CorruptArgs& a) {
    if (fast_ok(a) && (!a.rand_field || (reinterpret_cast<uintptr_t>(a.rand_field) & 15) == 0) && a.field_bytes % 16 == 0)
        return run_shot_noise_table(a);                                     // integer tables in shared memory (exact)
    const double c = sev_shot_noise(a.severity);
    // CDF rows for lam = (v/255)*c, recurrence p_k = p_{k-1}*lam/k (fixed op order, float64)
    std::vector<double> cdf((size_t)256 * POISSON_KMAX);
    for (int v = 0; v < 256; ++v) {
        const double lam = ((double)v / 255.0) * c;
        double p = std::exp(-lam), acc = p;
        cdf[(size_t)v * POISSON_KMAX] = acc;
        for (int k = 1; k < POISSON_KMAX; ++k) {
            p = p * lam / k;
            acc = acc + p;
            cdf[(size_t)v * POISSON_KMAX + k] = acc;
        }
    }
    std::vector<uint8_t> kout(POISSON_KMAX);
    for (int k = 0; k < POISSON_KMAX; ++k) {
        double v = (double)k / c;
        v = std::min(std::max(v, 0.0), 1.0) * 255.0;
        kout[k] = (uint8_t)(int)v;
    }
    const std::string key = "poisson_cdf_" + std::to_string(a.severity);
    const double* d_cdf = reinterpret_cast<const double*>(cached_table(key, cdf.data(), cdf.size() * sizeof(double)));
    const uint8_t* d_kout = reinterpret_cast<const uint8_t*>(cached_table(key + "_out", kout.data(), kout.size()));
    if (!d_cdf || !d_kout) return ADVMIX_ERR_CUDA;
    const int64_t quads = (int64_t)a.H * a.W * 3 / 4;
    shot_noise_kernel<<<point_grid(quads, a.n), PT_THREADS, 0, a.stream>>>(
        a.in, a.out, a.idx, inj_field(a), a.field_bytes, a.seed, a.sample_base, quads, d_cdf, d_kout);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_impulse_noise(const CorruptArgs& a) {
    if (!a.rand_field && fast_ok(a)) return run_impulse_noise_fast(a);
    const int64_t quads = (int64_t)a.H * a.W * 3 / 4;
    impulse_noise_kernel<<<point_grid(quads, a.n), PT_THREADS, 0, a.stream>>>(
        a.in, a.out, a.idx, inj_field(a), a.field_bytes, a.seed, a.sample_base, quads, sev_impulse_noise(a.severity));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_frost(const CorruptArgs& a) {
    ADVMIX_REQUIRE(a.frost_bank && a.frost_n > 0, "frost: texture bank required");
    ADVMIX_REQUIRE(a.frost_h >= a.H && a.frost_w >= a.W, "frost: textures (%dx%d) must cover the image (%dx%d)", a.frost_h, a.frost_w, a.H, a.W);
    const double c0[5] = {1, 0.8, 0.7, 0.65, 0.6}, c1[5] = {0.4, 0.6, 0.7, 0.7, 0.75};
    const int64_t quads = (int64_t)a.H * a.W * 3 / 4;
    ADVMIX_REQUIRE(quads < (int64_t)1 << 29, "frost: image too large");
    frost_kernel<<<point_grid(quads, a.n), PT_THREADS, 0, a.stream>>>(
        a.in, a.out, a.idx, a.rand_param, a.seed, a.sample_base, a.H, a.W, a.frost_bank, a.frost_n, a.frost_h,
        a.frost_w, c0[a.severity - 1], c1[a.severity - 1]);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_brightness(const CorruptArgs& a) {
    const double c[5] = {0.1, 0.2, 0.3, 0.4, 0.5};
    const int64_t groups = (int64_t)a.H * a.W / 4;
    brightness_kernel<false><<<point_grid(groups, a.n), PT_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, groups, c[a.severity - 1], 0.0);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_brightness_sweep(const SweepArgs& sw) {
    const CorruptArgs& a = sw.base;
    if (((int64_t)a.H * a.W) % 4 != 0) return -1;
    Sweep5D c{{0.1, 0.2, 0.3, 0.4, 0.5}};
    Sweep5Out o;
    for (int s = 0; s < 5; ++s) o.p[s] = sw.outs[s];
    const int64_t groups = (int64_t)a.H * a.W / 4;
    brightness_sweep_kernel<<<point_grid(groups, a.n), PT_THREADS, 0, a.stream>>>(a.in, o, a.idx, groups, c);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_frost_sweep(const SweepArgs& sw) {
    const CorruptArgs& a = sw.base;
    if (a.rand_param || !a.frost_bank || a.frost_n <= 0 || a.frost_h < a.H || a.frost_w < a.W || ((int64_t)a.H * a.W * 3) % 4 != 0) return -1;
    Sweep5D c0{{1, 0.8, 0.7, 0.65, 0.6}}, c1{{0.4, 0.6, 0.7, 0.7, 0.75}};
    Sweep5Out o;
    for (int s = 0; s < 5; ++s) o.p[s] = sw.outs[s];
    const int64_t quads = (int64_t)a.H * a.W * 3 / 4;
    if (quads >= (int64_t)1 << 29) return -1;
    frost_sweep_kernel<<<point_grid(quads, a.n), PT_THREADS, 0, a.stream>>>(a.in, o, a.idx, a.seed, a.sample_base, a.H, a.W, a.frost_bank,
                                                                          a.frost_n, a.frost_h, a.frost_w, c0, c1);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_saturate(const CorruptArgs& a) {
    const double c0[5] = {0.3, 0.1, 2, 5, 20}, c1[5] = {0, 0, 0, 0.1, 0.2};
    const int64_t groups = (int64_t)a.H * a.W / 4;
    brightness_kernel<true><<<point_grid(groups, a.n), PT_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, groups, c0[a.severity - 1],
                                                                                c1[a.severity - 1]);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_contrast(const CorruptArgs& a) {
    const double c[5] = {0.4, 0.3, 0.2, 0.1, 0.05};
    const int64_t groups = (int64_t)a.H * a.W / 4;
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(a.ws);
    if (a.fast && fast48_ok(a)) return run_contrast_fast(a, sums);
    ADVMIX_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)a.n * 3 * sizeof(unsigned long long), a.stream));
    channel_sum_kernel<<<point_grid(groups, a.n), PT_THREADS, 0, a.stream>>>(a.in, a.idx, groups, sums);
    ADVMIX_LAUNCH_OK();
    contrast_kernel<<<point_grid(groups, a.n), PT_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, groups, sums,
                                                                         (double)a.H * a.W, c[a.severity - 1]);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

void pixelate_dims(int severity, int H, int W, int* h2, int* w2) {
    const double c[5] = {0.6, 0.5, 0.4, 0.3, 0.25};
    *w2 = (int)(W * c[severity - 1]);
    *h2 = (int)(H * c[severity - 1]);
}

int run_pixelate(const CorruptArgs& a) {
    int h2, w2;
    pixelate_dims(a.severity, a.H, a.W, &h2, &w2);
    ADVMIX_REQUIRE(h2 > 0 && w2 > 0, "pixelate: image too small");
    const std::string key = "pix_" + std::to_string(a.H) + "x" + std::to_string(a.W) + "_" + std::to_string(a.severity);
    ResampleTab th = box_coeffs(a.W, w2), tv = box_coeffs(a.H, h2);
    std::vector<int32_t> xin = nearest_tab(w2, a.W), yin = nearest_tab(h2, a.H);
    const int32_t* d_th = reinterpret_cast<const int32_t*>(cached_table(key + "_h", th.data.data(), th.data.size() * 4));
    const int32_t* d_tv = reinterpret_cast<const int32_t*>(cached_table(key + "_v", tv.data.data(), tv.data.size() * 4));
    const int32_t* d_xin = reinterpret_cast<const int32_t*>(cached_table(key + "_x", xin.data(), xin.size() * 4));
    const int32_t* d_yin = reinterpret_cast<const int32_t*>(cached_table(key + "_y", yin.data(), yin.size() * 4));
    if (!d_th || !d_tv || !d_xin || !d_yin) return ADVMIX_ERR_CUDA;
    uint8_t* tmp = reinterpret_cast<uint8_t*>(a.ws);
    uint8_t* small = tmp + (size_t)a.n * a.H * w2 * 3;
    pix_h_kernel<<<point_grid((int64_t)a.H * w2, a.n), PT_THREADS, 0, a.stream>>>(a.in, tmp, a.idx, d_th, a.H, a.W, w2, th.ksize);
    ADVMIX_LAUNCH_OK();
    pix_v_kernel<<<point_grid((int64_t)h2 * w2 * 3, a.n), PT_THREADS, 0, a.stream>>>(tmp, small, d_tv, a.H, w2, h2, tv.ksize);
    ADVMIX_LAUNCH_OK();
    pix_up_kernel<<<point_grid((int64_t)a.H * a.W / 4, a.n), PT_THREADS, 0, a.stream>>>(small, a.out, a.idx, d_xin, d_yin, a.H, a.W, h2, w2);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // namespace advmix
