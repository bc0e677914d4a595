// ADVMIX_CORRUPT_FAST variants, second part (see corrupt_fast32.cu): snow, fog, elastic_transform.
#include "stencil_common.cuh"

#include <climits>

namespace advmix {

__device__ __forceinline__ uint32_t f32_to_u8_255b(float v) {       // clip(v, 0, 1) * 255 truncated
    v = fminf(fmaxf(v, 0.f), 1.f);
    return (uint32_t)__float2int_rz(__fmul_rn(v, 255.0f));
}

// ======================================================================== snow
// The zoomed / thresholded layer keeps its float64 bilinear sum (the threshold `layer < c3 -> 0` is a discontinuity: a
// float32 sum landing on the other side of it would move a flake by up to 255 * k_0 LSB), but is stored as float32; the
// 21..25-tap line blur - the expensive part - runs in float32.
struct ZoomLayerF { int top0, in0, out0, top1, in1, out1; double z0, z1; };

constexpr int SNOW_PADX = 24;        // |dx| <= 24 * |cos(angle)| <= 17 for angles in (-135, -45): replicated columns instead of x clamps

// layer [oh][ow + 2 * SNOW_PADX]: column xp holds the layer value at x = clamp(xp - SNOW_PADX, 0, ow - 1).
// The zoom geometry (source index pair + weight per output row / column: floor and fraction of o * z in float64) depends on the
// shape and the severity only and comes from a small host-built table; the kernel that evaluated it per output spent 122
// instructions per layer value (ncu: 70 % issue-active - integer division, two float64 coordinate evaluations, four 64-bit
// addresses, float64 blend).  The bilinear sum itself now runs in float32 and is redone in float64 only within 2e-5 of the
// threshold, where the two could disagree about `layer < c3 -> 0`.
struct SnowTap { int i0, i1; float t; int valid; };      // field row / column pair, weight of i1, inside the zoomed layer

__global__ void __launch_bounds__(ST_THREADS)
snow_layer_fast_kernel(float* __restrict__ layer, const float* __restrict__ field, size_t field_stride, int W, int oh, int ow,
                       const SnowTap* __restrict__ rows, const SnowTap* __restrict__ cols, float c0, float c1, float c3, double c0d, double c1d,
                       double c3d) {
    const int i = blockIdx.y;
    const float* f = reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride);
    const int owp = ow + 2 * SNOW_PADX;
    float* dst = layer + (int64_t)i * oh * owp;
    const int64_t total = (int64_t)oh * owp;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < total; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)owp), xp = (int)((uint32_t)p - (uint32_t)y * (uint32_t)owp);
        const SnowTap ry = rows[y];
        const SnowTap cx = cols[clampi(xp - SNOW_PADX, 0, ow - 1)];
        float t = 0.f;
        if (ry.valid && cx.valid) {
            const float* f0 = f + (size_t)ry.i0 * W;
            const float* f1 = f + (size_t)ry.i1 * W;
            const float a00 = __ldg(f0 + cx.i0), a01 = __ldg(f0 + cx.i1), a10 = __ldg(f1 + cx.i0), a11 = __ldg(f1 + cx.i1);
            const float v00 = fmaf(c1, a00, c0), v01 = fmaf(c1, a01, c0), v10 = fmaf(c1, a10, c0), v11 = fmaf(c1, a11, c0);
            const float top = fmaf(cx.t, v01 - v00, v00), bot = fmaf(cx.t, v11 - v10, v10);
            t = fmaf(ry.t, bot - top, top);
            if (fabsf(t - c3) < 2e-5f) {                 // too close to the threshold for float32: the float64 sum decides
                const double ty = (double)ry.t, tx = (double)cx.t, wy0 = 1.0 - ty, wx0 = 1.0 - tx;
                double td = 0.0;
                td = td + ((c0d + c1d * (double)a00) * wy0) * wx0;
                td = td + ((c0d + c1d * (double)a01) * wy0) * tx;
                td = td + ((c0d + c1d * (double)a10) * ty) * wx0;
                td = td + ((c0d + c1d * (double)a11) * ty) * tx;
                t = td < c3d ? 0.f : (float)td;
            } else if (t < c3) {
                t = 0.f;
            }
        }
        dst[p] = fminf(fmaxf(t, 0.f), 1.f);
    }
}

constexpr int SNOW_MAXW = 41;

// line blur of the layer + round(layer * 255) -> uint8 [H][W].  One 1024-thread CTA per image: the prologue (float64 sin / cos /
// hypot of the tap offsets, one warp) costs several microseconds of latency, which four 256-thread CTAs per image each paid.
constexpr int SB_THREADS = 1024;
__global__ void __launch_bounds__(SB_THREADS)
snow_blur_fast_kernel(const float* __restrict__ layer, uint8_t* __restrict__ layer8, const int32_t* __restrict__ idx,
                      const double* __restrict__ param, uint64_t seed, int64_t sample_base, int H, int W, int oh, int ow,
                      const double* __restrict__ kw, int width) {
    __shared__ int s_dy[SNOW_MAXW], s_dx[SNOW_MAXW], s_n;
    __shared__ float s_k[SNOW_MAXW];
    __shared__ int2 s_off[SNOW_MAXW];
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const double angle = param_uniform(param ? param + 4 * i : nullptr, rng, -135.0, -45.0);
    if (threadIdx.x < width) {
        const int t = threadIdx.x;
        s_k[t] = (float)kw[t];
        const double rad = angle * (3.141592653589793 / 180.0);
        const double p0 = (double)width * sin(rad), p1 = (double)width * cos(rad);
        const double hyp = hypot(p0, p1);
        s_dy[t] = -(int)ceil(((double)t * p0) / hyp - 0.5);
        s_dx[t] = -(int)ceil(((double)t * p1) / hyp - 0.5);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = width;
        for (int t = 0; t < width; ++t)
            if (abs(s_dy[t]) >= oh || abs(s_dx[t]) >= ow) { n = t; break; }
        s_n = n;
    }
    __syncthreads();
    const int ntaps = s_n;
    const int owp = ow + 2 * SNOW_PADX;
    if (threadIdx.x < ntaps) s_off[threadIdx.x] = make_int2(s_dy[threadIdx.x] * owp + s_dx[threadIdx.x], __float_as_int(s_k[threadIdx.x]));
    __syncthreads();
    int my0 = 0, my1 = 0, mxa = 0;
    for (int t = 0; t < ntaps; ++t) {
        my0 = min(my0, s_dy[t]); my1 = max(my1, s_dy[t]);
        mxa = max(mxa, abs(s_dx[t]));
    }
    const float* src = layer + (int64_t)i * oh * owp + SNOW_PADX;
    uint8_t* dst = layer8 + (int64_t)i * H * W;
    // 4 vertically adjacent pixels per thread (lanes along x: every load of a warp is one coalesced 128-byte line; four
    // horizontally adjacent pixels per thread cost 4x the L1 wavefronts and the kernel was L1 / latency bound): a tap costs one
    // table read and four loads + FMAs.  The layer carries SNOW_PADX replicated columns on each side, so no tap needs an x
    // clamp (no divergent border path); rows are clamped per tap only in the few rows whose taps leave the layer
    const bool padded_ok = mxa <= SNOW_PADX;
    const int hq = (H + 3) >> 2;
    for (int g = blockIdx.x * SB_THREADS + threadIdx.x; g < hq * W; g += gridDim.x * SB_THREADS) {
        const int yq = g / W, x = g - yq * W, y = yq << 2;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const bool rows_inside = y - my1 >= 0 && y + 3 - my0 < oh;
        const float* c = src + (y * owp + x);
        if (padded_ok && rows_inside) {
#pragma unroll 5
            for (int t = 0; t < ntaps; ++t) {
                const int2 T = s_off[t];
                const float k = __int_as_float(T.y);
                const float* q = c - T.x;
                acc[0] = fmaf(k, __ldg(q), acc[0]); acc[1] = fmaf(k, __ldg(q + owp), acc[1]);
                acc[2] = fmaf(k, __ldg(q + 2 * owp), acc[2]); acc[3] = fmaf(k, __ldg(q + 3 * owp), acc[3]);
            }
        } else {
            for (int t = 0; t < ntaps; ++t) {
                const int xx = padded_ok ? x - s_dx[t] : clampi(x - s_dx[t], 0, ow - 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fmaf(s_k[t], src[clampi(y + j - s_dy[t], 0, oh - 1) * owp + xx], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (y + j < H) dst[(y + j) * W + x] = (uint8_t)__float2int_rn(acc[j] * 255.0f);
    }
}

__global__ void __launch_bounds__(ST_THREADS)
snow_apply_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                       const uint8_t* __restrict__ layer8, int H, int W, float c6, float omc6) {
    __shared__ float f255[256];
    for (int i = threadIdx.x; i < 256; i += ST_THREADS) f255[i] = __fdiv_rn((float)i, 255.0f);
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = out + (int64_t)slot * H * W * 3;
    const uint8_t* L = layer8 + (int64_t)i * H * W;
    const int npix = H * W;
    // 4 pixels (12 bytes) per thread step; H*W is a multiple of 4 and images start 4-byte aligned in a dense batch
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(L)) & 3) == 0;
    for (int g = blockIdx.x * ST_THREADS + threadIdx.x; g < npix / 4; g += gridDim.x * ST_THREADS) {
        const int p0 = 4 * g;
        uint32_t w[3], la, lr;
        if (vec) {
            const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + p0 * 3);
            w[0] = __ldg(s4); w[1] = __ldg(s4 + 1); w[2] = __ldg(s4 + 2);
            la = __ldg(reinterpret_cast<const uint32_t*>(L + p0));
        } else {
            const uint8_t* s1 = src + p0 * 3;
            for (int k = 0; k < 3; ++k) w[k] = s1[4 * k] | (s1[4 * k + 1] << 8) | (s1[4 * k + 2] << 16) | ((uint32_t)s1[4 * k + 3] << 24);
            la = L[p0] | (L[p0 + 1] << 8) | (L[p0 + 2] << 16) | ((uint32_t)L[p0 + 3] << 24);
        }
        // rot90(k=2): pixel p pairs with npix - 1 - p
        const uint8_t* Lr = L + (npix - 4 - p0);
        lr = Lr[3] | (Lr[2] << 8) | (Lr[1] << 16) | ((uint32_t)Lr[0] << 24);
        uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float xs[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { const int e = 3 * k + c; xs[c] = f255[(w[e >> 2] >> (8 * (e & 3))) & 255u]; }
            const float gray = __fadd_rn(__fadd_rn(__fmul_rn(xs[0], 0.299f), __fmul_rn(xs[1], 0.587f)), __fmul_rn(xs[2], 0.114f));
            const float m = __fadd_rn(__fmul_rn(gray, 1.5f), 0.5f);
            const float add = f255[(la >> (8 * k)) & 255u] + f255[(lr >> (8 * k)) & 255u];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xv = __fadd_rn(__fmul_rn(c6, xs[c]), __fmul_rn(omc6, fmaxf(xs[c], m)));
                const int e = 3 * k + c;
                o[e >> 2] |= f32_to_u8_255b(xv + add) << (8 * (e & 3));
            }
        }
        if (vec) {
            uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + p0 * 3);
            d4[0] = o[0]; d4[1] = o[1]; d4[2] = o[2];
        } else {
            for (int k = 0; k < 12; ++k) dst[p0 * 3 + k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
        }
    }
}

void snow_layer_dims(int severity, int H, int W, int* oh, int* ow);

int run_snow_fast(const CorruptArgs& a) {
    struct SP { double c0, c1, c2, c3; int radius; double sigma, c6; };
    const SP p[5] = {{0.1, 0.3, 3, 0.5, 10, 4, 0.8}, {0.2, 0.3, 2, 0.5, 12, 4, 0.7}, {0.55, 0.3, 4, 0.9, 12, 8, 0.7},
                     {0.55, 0.3, 4.5, 0.85, 12, 8, 0.65}, {0.55, 0.3, 2.5, 0.85, 12, 12, 0.55}};
    const SP sp = p[a.severity - 1];
    ZoomLayerF z;
    z.in0 = (int)std::ceil(a.H / sp.c2); z.top0 = (a.H - z.in0) / 2;
    z.in1 = (int)std::ceil(a.W / sp.c2); z.top1 = (a.W - z.in1) / 2;
    z.out0 = (int)std::nearbyint(z.in0 * sp.c2); z.out1 = (int)std::nearbyint(z.in1 * sp.c2);
    z.z0 = z.out0 > 1 ? (double)(z.in0 - 1) / (double)(z.out0 - 1) : 1.0;
    z.z1 = z.out1 > 1 ? (double)(z.in1 - 1) / (double)(z.out1 - 1) : 1.0;
    const int width = 2 * sp.radius + 1;
    const double* d_k = reinterpret_cast<const double*>(cached_table("snowk_" + std::to_string(a.severity), SNOW_K[a.severity - 1], width * sizeof(double)));
    if (!d_k) return ADVMIX_ERR_CUDA;
    // workspace (sized for the float64 path): float layer [n][oh][ow] | uint8 layer [n][H][W] | generated field
    float* layer = reinterpret_cast<float*>(a.ws);
    if (z.out1 < 2 * SNOW_PADX) return -1;                            // the padded float32 layer must fit the float64 path's layer slot
    uint8_t* layer8 = reinterpret_cast<uint8_t*>(reinterpret_cast<double*>(a.ws) + (size_t)a.n * z.out0 * z.out1);
    const float* field = reinterpret_cast<const float*>(a.rand_field);
    size_t fstride = a.field_bytes;
    if (!field) {
        float* gen = reinterpret_cast<float*>(layer8 + (((size_t)a.n * a.H * a.W + 15) & ~(size_t)15));
        int rc = launch_fill_rand(a, gen, nullptr);
        if (rc) return rc;
        field = gen;
    }
    // zoom geometry table (rows, then columns): same float64 expressions as corrupt_stencil.cu:zoom_coord
    std::vector<SnowTap> taps((size_t)z.out0 + z.out1);
    auto entry = [](int o, double zz, int in, int top) {
        const double cc = (double)o * zz;
        if (cc < 0.0 || cc > (double)(in - 1)) return SnowTap{0, 0, 0.f, 0};
        const double fl = std::floor(cc);
        const int si = (int)fl;
        return SnowTap{top + si, top + std::min(si + 1, in - 1), (float)(cc - fl), 1};
    };
    for (int y = 0; y < z.out0; ++y) taps[y] = entry(y, z.z0, z.in0, z.top0);
    for (int x = 0; x < z.out1; ++x) taps[z.out0 + x] = entry(x, z.z1, z.in1, z.top1);
    const SnowTap* d_taps = reinterpret_cast<const SnowTap*>(cached_table(
        "snowtap_" + std::to_string(a.severity) + "_" + std::to_string(a.H) + "x" + std::to_string(a.W), taps.data(), taps.size() * sizeof(SnowTap)));
    if (!d_taps) return ADVMIX_ERR_CUDA;
    {
        snow_layer_fast_kernel<<<st_grid((int64_t)z.out0 * (z.out1 + 2 * SNOW_PADX), a.n), ST_THREADS, 0, a.stream>>>(layer, field, fstride, a.W, z.out0, z.out1, d_taps, d_taps + z.out0,
                                                                                   (float)sp.c0, (float)sp.c1, (float)sp.c3, sp.c0, sp.c1, sp.c3);
    }
    ADVMIX_LAUNCH_OK();
    snow_blur_fast_kernel<<<dim3(a.n >= sm_count() ? 1 : 4, a.n), SB_THREADS, 0, a.stream>>>(
        layer, layer8, a.idx, a.rand_param, a.seed, a.sample_base, a.H, a.W, z.out0, z.out1, d_k, width);
    ADVMIX_LAUNCH_OK();
    snow_apply_fast_kernel<<<st_grid((int64_t)a.H * a.W / 4, a.n), ST_THREADS, 0, a.stream>>>(
        a.in, a.out, a.idx, layer8, a.H, a.W, (float)sp.c6, (float)(1 - sp.c6));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== fog
// One CTA per image runs the whole diamond-square recursion on a float32 map (global scratch, L1 / L2 resident: 256 KB at
// M = 256), reduces min / max of the map and the image maximum, and applies the fog - 1 launch instead of 2 per level + 2.
constexpr int FG_THREADS = 512;       // several CTAs (images) per SM: the recursion is latency-bound, not issue-bound

__device__ __forceinline__ float wibbled_f(float sum4, float wibble, float u) {
    return fmaf(wibble, fmaf(2.0f * wibble, u, -wibble), 0.25f * sum4);
}

__global__ void __launch_bounds__(FG_THREADS, 2)
fog_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                float* __restrict__ maps, const float* __restrict__ field, size_t field_stride, uint64_t seed,
                int64_t sample_base, int n, int H, int W, int M, float decay, float c0) {
    __shared__ float s_mn[32], s_mx[32];
    __shared__ int s_im[32];
    __shared__ float s_stat[3];
    static_assert(FG_THREADS <= 1024, "one partial per warp");
    const int lm = 31 - __clz(M);
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const SampleRng rng(seed, sample_base + slot);
        const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)img * field_stride) : nullptr;
        float* m = maps + (size_t)blockIdx.x * M * M;            // one scratch map per CTA
        const uint8_t* src = in + (int64_t)slot * H * W * 3;
        uint8_t* dst = out + (int64_t)slot * H * W * 3;
        __syncthreads();
        if (threadIdx.x == 0) m[0] = 0.f;
        float wibble = 100.f;
        for (int ls = lm; ls >= 1; --ls) {                       // step = 1 << ls
            const int step = 1 << ls, h = step >> 1, nb = lm - ls, nn = 1 << nb, mask = nn - 1;
            __syncthreads();
            for (int t = threadIdx.x; t < nn * nn; t += FG_THREADS) {
                const int a = t >> nb, b = t & mask, a1 = (a + 1) & mask, b1 = (b + 1) & mask;
                const float c00 = m[((a << ls) << lm) + (b << ls)], c10 = m[((a1 << ls) << lm) + (b << ls)];
                const float c01 = m[((a << ls) << lm) + (b1 << ls)], c11 = m[((a1 << ls) << lm) + (b1 << ls)];
                const int y = (a << ls) + h, x = (b << ls) + h;
                m[(y << lm) + x] = wibbled_f((c00 + c10) + (c01 + c11), wibble, field_uniform1(inj, rng, TAG_FIELD0, ((uint64_t)y << lm) + x));
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nn * nn; t += FG_THREADS) {
                const int a = t >> nb, b = t & mask;
                const int am = (a + nn - 1) & mask, bm = (b + nn - 1) & mask, a1 = (a + 1) & mask, b1 = (b + 1) & mask;
                const float dr = m[(((a << ls) + h) << lm) + (b << ls) + h], ul = m[((a << ls) << lm) + (b << ls)];
                const float lt = (dr + m[(((am << ls) + h) << lm) + (b << ls) + h]) + (ul + m[((a << ls) << lm) + (b1 << ls)]);
                const float tt = (dr + m[(((a << ls) + h) << lm) + (bm << ls) + h]) + (ul + m[((a1 << ls) << lm) + (b << ls)]);
                int y = a << ls, x = (b << ls) + h;
                const float v1 = wibbled_f(lt, wibble, field_uniform1(inj, rng, TAG_FIELD0, ((uint64_t)y << lm) + x));
                m[(y << lm) + x] = v1;
                y = (a << ls) + h; x = b << ls;
                const float v2 = wibbled_f(tt, wibble, field_uniform1(inj, rng, TAG_FIELD0, ((uint64_t)y << lm) + x));
                m[(y << lm) + x] = v2;
            }
            wibble = __fdiv_rn(wibble, decay);
        }
        __syncthreads();
        // min / max of the map, max of the image
        float mn = INFINITY, mx = -INFINITY;
        for (int t = threadIdx.x; t < M * M / 4; t += FG_THREADS) {
            const float4 v = reinterpret_cast<const float4*>(m)[t];
            mn = fminf(fminf(mn, fminf(v.x, v.y)), fminf(v.z, v.w));
            mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
        int im = 0;
        const int nq = H * W * 3 / 4;                            // H*W % 4 == 0
        const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0;
        if (vec) {
            for (int t = threadIdx.x; t < nq; t += FG_THREADS) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src) + t);
                im = max(max(im, (int)(v & 255u)), max(max((int)((v >> 8) & 255u), (int)((v >> 16) & 255u)), (int)(v >> 24)));
            }
        } else {
            for (int t = threadIdx.x; t < 4 * nq; t += FG_THREADS) im = max(im, (int)src[t]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            im = max(im, __shfl_xor_sync(0xffffffffu, im, o));
        }
        if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; s_im[threadIdx.x >> 5] = im; }
        __syncthreads();
        if (threadIdx.x < 32) {
            const bool have = threadIdx.x < FG_THREADS / 32;
            mn = have ? s_mn[threadIdx.x] : INFINITY; mx = have ? s_mx[threadIdx.x] : -INFINITY; im = have ? s_im[threadIdx.x] : 0;
            for (int o = 16; o > 0; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                im = max(im, __shfl_xor_sync(0xffffffffu, im, o));
            }
            if (threadIdx.x == 0) { s_stat[0] = mn; s_stat[1] = mx - mn; s_stat[2] = (float)im; }
        }
        __syncthreads();
        // x = (x/255 + c0 * (map - mn) / (mx - mn)) * max_val / (max_val + c0), clipped, * 255:
        // in units of 255: (byte + 255 * c0 * pf) * scale, scale = max_val / (max_val + c0), max_val = im / 255
        const float mnv = s_stat[0], rng_v = s_stat[1], max_val = s_stat[2] / 255.0f;
        const float scale = max_val / (max_val + c0);
        const float k = rng_v > 0.f ? 255.0f * c0 / rng_v : 0.f;
        const int wq = W >> 2;
        if (vec && (W & 3) == 0) {
            for (int g = threadIdx.x; g < H * wq; g += FG_THREADS) {
                const int y = g / wq, x = (g - y * wq) << 2;
                const float4 pf = *reinterpret_cast<const float4*>(m + (y << lm) + x);
                const float add[4] = {(pf.x - mnv) * k, (pf.y - mnv) * k, (pf.z - mnv) * k, (pf.w - mnv) * k};
                const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + (y * W + x) * 3);
                const uint32_t w[3] = {__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2)};
                uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
                for (int e = 0; e < 12; ++e) {
                    const float b = byte_of_word_f(w[e >> 2], e & 3);
                    const float v = fminf((b + add[e / 3]) * scale, 255.0f);
                    o[e >> 2] |= (uint32_t)__float2int_rz(v) << (8 * (e & 3));
                }
                uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + (y * W + x) * 3);
                d4[0] = o[0]; d4[1] = o[1]; d4[2] = o[2];
            }
        } else {
            for (int p = threadIdx.x; p < H * W; p += FG_THREADS) {
                const int y = p / W, x = p - y * W;
                const float add = (m[(y << lm) + x] - mnv) * k;
#pragma unroll
                for (int c = 0; c < 3; ++c) dst[p * 3 + c] = (uint8_t)__float2int_rz(fminf(((float)src[p * 3 + c] + add) * scale, 255.0f));
            }
        }
    }
}

// Dense-level variant for maps up to 256 x 256, everything in shared memory.  The map after level k is kept as a dense
// n x n array D_k (n = 2^k) instead of a strided sub-lattice of the M x M map, so a level reads D_k and writes D_{k+1}
// with unit strides; D_0 .. D_6 ping-pong between two small buffers, D_7 (128 x 128, 64 KB) and the squares of the last
// level (64 KB) stay resident, and the diamonds of the last level - half of all cells - are never stored: the min / max
// pass and the apply pass evaluate them from their four neighbours and the cell's uniform.  The global-scratch kernel
// above is latency-bound (16 dependent phases of L2 round trips, ~130 us per image whatever the batch); here a phase is
// a shared-memory round trip.
constexpr int FD_THREADS = 1024;

__global__ void __launch_bounds__(FD_THREADS, 1)
fog_dense_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                 const float* __restrict__ field, size_t field_stride, int n, int H, int W, int lm, float decay, float c0,
                 uint8_t* __restrict__ out2 = nullptr, float c0b = 0.f) {      // out2 / c0b: a second severity with the same wibble decay (sweep)
    extern __shared__ __align__(16) float fd_smem[];
    const int M = 1 << lm, nh = M >> 1;                  // nh x nh: D_{lm-1} and the last level's squares
    float* A = fd_smem;                                  // D_{lm-1}
    float* S = A + nh * nh;                              // last-level squares; before that: odd-numbered D_k
    float* T = S + nh * nh;                              // even-numbered D_k up to (nh/2)^2
    __shared__ float s_mn[FD_THREADS / 32], s_mx[FD_THREADS / 32];
    __shared__ int s_im[FD_THREADS / 32];
    __shared__ float s_stat[3];
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const float* U = reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)img * field_stride);
        const uint8_t* src = in + (int64_t)slot * H * W * 3;
        uint8_t* dst = out + (int64_t)slot * H * W * 3;
        __syncthreads();
        // D_0 lives in the buffer of even k
        if (threadIdx.x == 0) (lm - 1 == 0 ? A : T)[0] = 0.f;
        float wibble = 100.f;
        for (int k = 0; k + 1 < lm; ++k) {               // D_k (nn x nn) -> D_{k+1} (2nn x 2nn)
            const int nn = 1 << k, mask = nn - 1, n2 = nn << 1, sh = lm - k - 1;      // full-res coordinate = dense index << sh
            const float* Dk = (k == lm - 1) ? A : ((k & 1) ? S : T);
            float* Dn = (k + 1 == lm - 1) ? A : (((k + 1) & 1) ? S : T);
            __syncthreads();
            for (int t = threadIdx.x; t < nn * nn; t += FD_THREADS) {
                const int a = t >> k, b = t & mask, a1 = (a + 1) & mask, b1 = (b + 1) & mask;
                const float c00 = Dk[a * nn + b], c10 = Dk[a1 * nn + b], c01 = Dk[a * nn + b1], c11 = Dk[a1 * nn + b1];
                Dn[(2 * a) * n2 + 2 * b] = c00;
                const int y = (2 * a + 1) << sh, x = (2 * b + 1) << sh;
                Dn[(2 * a + 1) * n2 + 2 * b + 1] = wibbled_f((c00 + c10) + (c01 + c11), wibble, __ldg(U + ((size_t)y << lm) + x));
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nn * nn; t += FD_THREADS) {
                const int a = t >> k, b = t & mask, am = (a + nn - 1) & mask, bm = (b + nn - 1) & mask, a1 = (a + 1) & mask, b1 = (b + 1) & mask;
                const float dr = Dn[(2 * a + 1) * n2 + 2 * b + 1], ul = Dk[a * nn + b];
                const float lt = (dr + Dn[(2 * am + 1) * n2 + 2 * b + 1]) + (ul + Dk[a * nn + b1]);
                const float tt = (dr + Dn[(2 * a + 1) * n2 + 2 * bm + 1]) + (ul + Dk[a1 * nn + b]);
                int y = (2 * a) << sh, x = (2 * b + 1) << sh;
                Dn[(2 * a) * n2 + 2 * b + 1] = wibbled_f(lt, wibble, __ldg(U + ((size_t)y << lm) + x));
                y = (2 * a + 1) << sh; x = (2 * b) << sh;
                Dn[(2 * a + 1) * n2 + 2 * b] = wibbled_f(tt, wibble, __ldg(U + ((size_t)y << lm) + x));
            }
            wibble = __fdiv_rn(wibble, decay);
        }
        __syncthreads();
        // last level: squares into S, diamonds on the fly
        const int hm = nh - 1, lh = lm - 1;
        for (int t = threadIdx.x; t < nh * nh; t += FD_THREADS) {
            const int a = t >> lh, b = t & hm, a1 = (a + 1) & hm, b1 = (b + 1) & hm;
            const float c00 = A[a * nh + b], c10 = A[a1 * nh + b], c01 = A[a * nh + b1], c11 = A[a1 * nh + b1];
            S[t] = wibbled_f((c00 + c10) + (c01 + c11), wibble, __ldg(U + ((size_t)(2 * a + 1) << lm) + 2 * b + 1));
        }
        __syncthreads();
        auto diamond_lt = [&](int a, int b) {           // map cell (2a, 2b+1)
            const float lt = (S[a * nh + b] + S[((a + nh - 1) & hm) * nh + b]) + (A[a * nh + b] + A[a * nh + ((b + 1) & hm)]);
            return wibbled_f(lt, wibble, __ldg(U + ((size_t)(2 * a) << lm) + 2 * b + 1));
        };
        auto diamond_tt = [&](int a, int b) {           // map cell (2a+1, 2b)
            const float tt = (S[a * nh + b] + S[a * nh + ((b + nh - 1) & hm)]) + (A[a * nh + b] + A[((a + 1) & hm) * nh + b]);
            return wibbled_f(tt, wibble, __ldg(U + ((size_t)(2 * a + 1) << lm) + 2 * b));
        };
        float mn = INFINITY, mx = -INFINITY;
        for (int t = threadIdx.x; t < nh * nh; t += FD_THREADS) {
            const int a = t >> lh, b = t & hm;
            const float v0 = A[t], v1 = S[t], v2 = diamond_lt(a, b), v3 = diamond_tt(a, b);
            mn = fminf(fminf(mn, fminf(v0, v1)), fminf(v2, v3));
            mx = fmaxf(fmaxf(mx, fmaxf(v0, v1)), fmaxf(v2, v3));
        }
        int im = 0;
        const int nq = H * W * 3 / 4;
        const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0 && (W & 3) == 0;
        if (vec) {
            for (int t = threadIdx.x; t < nq; t += FD_THREADS) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src) + t);
                im = max(max(im, (int)(v & 255u)), max(max((int)((v >> 8) & 255u), (int)((v >> 16) & 255u)), (int)(v >> 24)));
            }
        } else {
            for (int t = threadIdx.x; t < 4 * nq; t += FD_THREADS) im = max(im, (int)src[t]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            im = max(im, __shfl_xor_sync(0xffffffffu, im, o));
        }
        if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; s_im[threadIdx.x >> 5] = im; }
        __syncthreads();
        if (threadIdx.x < 32) {
            const bool have = threadIdx.x < FD_THREADS / 32;
            mn = have ? s_mn[threadIdx.x] : INFINITY; mx = have ? s_mx[threadIdx.x] : -INFINITY; im = have ? s_im[threadIdx.x] : 0;
            for (int o = 16; o > 0; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                im = max(im, __shfl_xor_sync(0xffffffffu, im, o));
            }
            if (threadIdx.x == 0) { s_stat[0] = mn; s_stat[1] = mx - mn; s_stat[2] = (float)im; }
        }
        __syncthreads();
        const float mnv = s_stat[0], rng_v = s_stat[1], max_val = s_stat[2] / 255.0f;
        const float scale = max_val / (max_val + c0);
        const float kk = rng_v > 0.f ? 255.0f * c0 / rng_v : 0.f;
        const float scale2 = max_val / (max_val + c0b);
        const float kk2 = rng_v > 0.f ? 255.0f * c0b / rng_v : 0.f;
        uint8_t* dst2 = out2 ? out2 + (int64_t)slot * H * W * 3 : nullptr;
        auto cell = [&](int y, int x) {
            const int a = y >> 1, b = x >> 1;
            if (y & 1) return (x & 1) ? S[a * nh + b] : diamond_tt(a, b);
            return (x & 1) ? diamond_lt(a, b) : A[a * nh + b];
        };
        if (vec) {
            const int wq = W >> 2;
            for (int g = threadIdx.x; g < H * wq; g += FD_THREADS) {
                const int y = g / wq, x = (g - y * wq) << 2;
                float cv[4], add[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { cv[j] = cell(y, x + j) - mnv; add[j] = cv[j] * kk; }
                const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + (y * W + x) * 3);
                const uint32_t w[3] = {__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2)};
                uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
                for (int e = 0; e < 12; ++e) {
                    const float bv = byte_of_word_f(w[e >> 2], e & 3);
                    o[e >> 2] |= (uint32_t)__float2int_rz(fminf((bv + add[e / 3]) * scale, 255.0f)) << (8 * (e & 3));
                }
                uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + (y * W + x) * 3);
                d4[0] = o[0]; d4[1] = o[1]; d4[2] = o[2];
                if (dst2) {
                    uint32_t o2[3] = {0u, 0u, 0u};
#pragma unroll
                    for (int e = 0; e < 12; ++e) {
                        const float bv = byte_of_word_f(w[e >> 2], e & 3);
                        o2[e >> 2] |= (uint32_t)__float2int_rz(fminf((bv + cv[e / 3] * kk2) * scale2, 255.0f)) << (8 * (e & 3));
                    }
                    uint32_t* e4 = reinterpret_cast<uint32_t*>(dst2 + (y * W + x) * 3);
                    e4[0] = o2[0]; e4[1] = o2[1]; e4[2] = o2[2];
                }
            }
        } else {
            for (int p = threadIdx.x; p < H * W; p += FD_THREADS) {
                const int y = p / W, x = p - y * W;
                const float cv = cell(y, x) - mnv, add = cv * kk;
#pragma unroll
                for (int c = 0; c < 3; ++c) dst[p * 3 + c] = (uint8_t)__float2int_rz(fminf(((float)src[p * 3 + c] + add) * scale, 255.0f));
                if (dst2) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) dst2[p * 3 + c] = (uint8_t)__float2int_rz(fminf(((float)src[p * 3 + c] + cv * kk2) * scale2, 255.0f));
                }
            }
        }
    }
}

int run_fog_fast(const CorruptArgs& a) {
    const double c0[5] = {1.5, 2., 2.5, 2.5, 3.}, decay[5] = {2, 2, 1.7, 1.5, 1.4};
    const int M = next_pow2(std::max(a.H, a.W));
    if (M > 1024 || (int64_t)a.H * a.W * 3 >= INT_MAX) return -1;
    // scratch: one float32 map per CTA; the float64 path's workspace holds n float64 maps, so min(n, 4 * #SM) float32 maps fit
    const int ctas = std::min(a.n, 4 * sm_count());
    const float* field = reinterpret_cast<const float*>(a.rand_field);
    if (!field) {
        // perf mode: one Philox block per 4 map cells, written once (the per-cell accessor would redo the block 4 times)
        float* gen = reinterpret_cast<float*>(a.ws) + (size_t)a.n * M * M;
        int rc = launch_fill_rand(a, gen, nullptr);
        if (rc) return rc;
        field = gen;
    }
    if (M <= 256 && M >= 4) {
        int lm = 0;
        while ((1 << lm) < M) ++lm;
        const size_t smem = ((size_t)2 * (M / 2) * (M / 2) + (size_t)(M / 4) * (M / 4)) * sizeof(float);
        ADVMIX_CUDA_OK(ensure_dyn_smem(fog_dense_kernel, 160 * 1024));
        fog_dense_kernel<<<std::min(a.n, sm_count()), FD_THREADS, smem, a.stream>>>(a.in, a.out, a.idx, field, a.field_bytes, a.n, a.H, a.W, lm,
                                                                                  (float)decay[a.severity - 1], (float)c0[a.severity - 1]);
        ADVMIX_LAUNCH_OK();
        return ADVMIX_OK;
    }
    fog_fast_kernel<<<ctas, FG_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, reinterpret_cast<float*>(a.ws),
                                                      field, a.field_bytes, a.seed, a.sample_base,
                                                      a.n, a.H, a.W, M, (float)decay[a.severity - 1], (float)c0[a.severity - 1]);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// sweep: severities 1 and 2 share the wibble decay (2.0), hence the plasma map; one launch writes both.  -1: not this shape / mode.
int run_fog_pair_fast(const CorruptArgs& a, const float* field, uint8_t* out1, uint8_t* out2) {
    const int M = next_pow2(std::max(a.H, a.W));
    if (!a.fast || !field || M > 256 || M < 4 || (int64_t)a.H * a.W * 3 >= INT_MAX) return -1;
    int lm = 0;
    while ((1 << lm) < M) ++lm;
    const size_t smem = ((size_t)2 * (M / 2) * (M / 2) + (size_t)(M / 4) * (M / 4)) * sizeof(float);
    ADVMIX_CUDA_OK(ensure_dyn_smem(fog_dense_kernel, 160 * 1024));
    fog_dense_kernel<<<std::min(a.n, sm_count()), FD_THREADS, smem, a.stream>>>(a.in, out1, a.idx, field, a.field_bytes, a.n, a.H, a.W, lm,
                                                                              2.0f, 1.5f, out2, 2.0f);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== elastic_transform
struct LoadElasticUniformF {
    const float* field; size_t field_stride; int H, W; float maxd;
    __device__ void init() {}
    typedef const float* Row;
    __device__ Row row(int img2, int y) const {
        return reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)(img2 >> 1) * field_stride) +
               (size_t)(img2 & 1) * H * W + (size_t)y * W;
    }
    typedef float Raw;
    __device__ Raw raw(Row r, int xc) const { return __ldg(r + xc); }
    __device__ float cvt(Raw v) const { return fmaf(2.0f * maxd, v, -maxd); }
};
struct StoreF32ScaledF {
    float* base; int64_t stride; int WC; float alpha;
    __device__ void operator()(int img, int y, int xc, float v) const { base[(int64_t)img * stride + (int64_t)y * WC + xc] = v * alpha; }
};

__device__ __forceinline__ double scipy_reflect_d(double in, int len) {
    if (in < 0) {
        if (len <= 1) return 0.0;
        const double sz2 = 2.0 * len;
        if (in < -sz2) in = sz2 * (double)(int)(-in / sz2) + in;
        in = in < -len ? in + sz2 : -in - 1;
    } else if (in > len - 1) {
        if (len <= 1) return 0.0;
        const double sz2 = 2.0 * len;
        in -= sz2 * (double)(int)(in / sz2);
        if (in >= len) in = sz2 - in - 1;
    }
    return in;
}

// map_coordinates(order=1, mode='reflect'): the integer and fractional parts of a coordinate are kept apart (y + floor(d),
// d - floor(d)), so float32 loses nothing to the magnitude of y; coordinates that leave the image take the float64 route.
__device__ __forceinline__ void elastic_gather_px(const uint8_t* __restrict__ src, uint8_t* __restrict__ o, const float* f255, int y, int x,
                                                  float dy, float dx, int H, int W) {
    const float fy = floorf(dy), fx = floorf(dx);
    int y0 = y + (int)fy, x0 = x + (int)fx, y1 = y0 + 1, x1 = x0 + 1;
    float ty = dy - fy, tx = dx - fx;
    if (y0 < 0 || y1 > H - 1 || x0 < 0 || x1 > W - 1) {
        const double cy = scipy_reflect_d((double)y + (double)dy, H), cx = scipy_reflect_d((double)x + (double)dx, W);
        const double gy = floor(cy), gx = floor(cx);
        ty = (float)(cy - gy); tx = (float)(cx - gx);
        y0 = reflect_sym((int)gy, H); y1 = reflect_sym((int)gy + 1, H);
        x0 = reflect_sym((int)gx, W); x1 = reflect_sym((int)gx + 1, W);
    }
    const uint8_t* p00 = src + (y0 * W + x0) * 3;
    const uint8_t* p01 = src + (y0 * W + x1) * 3;
    const uint8_t* p10 = src + (y1 * W + x0) * 3;
    const uint8_t* p11 = src + (y1 * W + x1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = f255[__ldg(p00 + c)], b = f255[__ldg(p01 + c)], cc = f255[__ldg(p10 + c)], d = f255[__ldg(p11 + c)];
        const float top = fmaf(tx, b - a, a), bot = fmaf(tx, d - cc, cc);
        o[c] = (uint8_t)f32_to_u8_255b(fmaf(ty, bot - top, top));
    }
}

__global__ void __launch_bounds__(ST_THREADS)
elastic_gather_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                           const float* __restrict__ disp, int H, int W) {
    __shared__ float f255[256];
    for (int t = threadIdx.x; t < 256; t += ST_THREADS) f255[t] = __fdiv_rn((float)t, 255.0f);
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = out + (int64_t)slot * H * W * 3;
    const int npix = H * W;
    const float* dxf = disp + (int64_t)(2 * i) * npix;
    const float* dyf = disp + (int64_t)(2 * i + 1) * npix;
    for (int p = blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += gridDim.x * ST_THREADS) {
        const int y = p / W, x = p - y * W;
        elastic_gather_px(src, dst + p * 3, f255, y, x, __ldg(dyf + p), __ldg(dxf + p), H, W);
    }
}

// five severities (advmix_corrupt_sweep_u8c3): the smoothed field does not depend on the severity (sigma = 0.01 * size), only
// its scale alpha does - `disp` holds the unscaled field and the product field * alpha_s (the float32 multiply
// StoreF32ScaledF does in the single-severity path) is formed here.
__global__ void __launch_bounds__(ST_THREADS)
elastic_gather_sweep_fast_kernel(const uint8_t* __restrict__ in, Sweep5Out outs, const int32_t* __restrict__ idx,
                                 const float* __restrict__ disp, int H, int W, Sweep5F alpha) {
    __shared__ float f255[256];
    for (int t = threadIdx.x; t < 256; t += ST_THREADS) f255[t] = __fdiv_rn((float)t, 255.0f);
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    const int npix = H * W;
    const float* dxf = disp + (int64_t)(2 * i) * npix;
    const float* dyf = disp + (int64_t)(2 * i + 1) * npix;
    for (int p = blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += gridDim.x * ST_THREADS) {
        const int y = p / W, x = p - y * W;
        const float dy = __ldg(dyf + p), dx = __ldg(dxf + p);
#pragma unroll
        for (int s = 0; s < 5; ++s)
            elastic_gather_px(src, outs.p[s] + (int64_t)slot * H * W * 3 + p * 3, f255, y, x, dy * alpha.v[s], dx * alpha.v[s], H, W);
    }
}

int run_elastic_fast(const CorruptArgs& a) {
    const double alpha[5] = {250 * 0.05, 250 * 0.065, 250 * 0.085, 250 * 0.1, 250 * 0.12};
    const int H = a.H, W = a.W;
    if ((int64_t)H * W * 3 >= INT_MAX) return -1;
    const double sig0 = H * 0.01, sig1 = W * 0.01, maxd = H * 0.005;
    int r0, r1;
    const double* w0 = gauss_table(sig0, 3.0, &r0);
    const double* w1 = gauss_table(sig1, 3.0, &r1);
    if (!w0 || !w1) return ADVMIX_ERR_CUDA;
    if (r0 >= H || r1 >= W) return -1;
    const int64_t plane = (int64_t)H * W;
    float* disp = reinterpret_cast<float*>(a.ws);
    const float* field = reinterpret_cast<const float*>(a.rand_field);
    if (!field) {
        float* gen = disp + (size_t)2 * a.n * plane;
        int rc = launch_fill_rand(a, gen, nullptr);
        if (rc) return rc;
        field = gen;
    }
    int rc = launch_gauss2d_fast(LoadElasticUniformF{field, a.field_bytes, H, W, (float)maxd},
                                 StoreF32ScaledF{disp, plane, W, (float)alpha[a.severity - 1]}, 2 * a.n, H, W, 1, r0, r1, w0, w1, BORDER_REFLECT, a.stream);
    if (rc) return rc;                                           // -1: no float32 variant for these radii
    elastic_gather_fast_kernel<<<st_grid(plane, a.n), ST_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, disp, H, W);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_elastic_sweep_fast(const SweepArgs& sw, const float* shared_field) {
    const CorruptArgs& a = sw.base;
    const double alpha[5] = {250 * 0.05, 250 * 0.065, 250 * 0.085, 250 * 0.1, 250 * 0.12};
    const int H = a.H, W = a.W;
    if (!a.fast || !shared_field || (int64_t)H * W * 3 >= INT_MAX) return -1;
    const double sig0 = H * 0.01, sig1 = W * 0.01, maxd = H * 0.005;
    int r0, r1;
    const double* w0 = gauss_table(sig0, 3.0, &r0);
    const double* w1 = gauss_table(sig1, 3.0, &r1);
    if (!w0 || !w1) return ADVMIX_ERR_CUDA;
    if (r0 >= H || r1 >= W) return -1;
    const int64_t plane = (int64_t)H * W;
    float* disp = reinterpret_cast<float*>(a.ws);
    int rc = launch_gauss2d_fast(LoadElasticUniformF{shared_field, a.field_bytes, H, W, (float)maxd},
                                 StoreF32ScaledF{disp, plane, W, 1.0f}, 2 * a.n, H, W, 1, r0, r1, w0, w1, BORDER_REFLECT, a.stream);
    if (rc) return rc;
    Sweep5Out o;
    Sweep5F al;
    for (int s = 0; s < 5; ++s) { o.p[s] = sw.outs[s]; al.v[s] = (float)alpha[s]; }
    // (a variant with the image resident in shared memory, one 1024-thread CTA per image, was slower: 3.1 instead of 2.1 us per
    // image for the five outputs - the 60 byte loads + 60 LUT reads per pixel then all go through one SM's shared-memory pipe at
    // a quarter of the occupancy; from global memory they hit L1 with 8 CTAs per SM)
    elastic_gather_sweep_fast_kernel<<<st_grid(plane, a.n), ST_THREADS, 0, a.stream>>>(a.in, o, a.idx, disp, H, W, al);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // namespace advmix
