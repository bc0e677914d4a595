// jpeg_compression: the numeric round trip of PIL's save(quality=q) -> open(), i.e.
// libjpeg(-turbo) baseline 4:2:0 with the islow DCT, minus the (lossless) entropy coding:
//   RGB->YCbCr (16-bit fixed tables) -> h2v2 box downsample (bias 1,2,1,2) -> level shift,
//   islow FDCT -> quantise -> dequantise -> islow IDCT + range limit -> h2v2 fancy (triangle)
//   upsample -> YCbCr->RGB.  All integer; bit-exact against PIL is the bar.
#include "corrupt_common.cuh"
#include "jpeg_common.cuh"

#include <cstring>
#include <initializer_list>
#include <string>
#include <vector>

namespace advmix {

constexpr int JP_THREADS = 128;

// n / d for 0 <= n < 2^18, 8 <= d <= 2040 with d's float reciprocal: the float product is within 1 of the quotient, one
// multiply-subtract fixes it up.  (An integer division by a run-time divisor costs ~20 instructions, and the quantiser
// does 64 of them per block.)
__device__ __forceinline__ int div_exact(int n, int d, float rcp) {
    int q = (int)((float)n * rcp);
    const int r = n - q * d;
    q += (r >= d) ? 1 : 0;
    q -= (r < 0) ? 1 : 0;
    return q;
}

// ---- tables --------------------------------------------------------------------------
static const uint8_t STD_LUMA[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57,
                                     69, 56, 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64,
                                     81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t STD_CHROMA[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                       99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                       99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

static void quant_tables(int quality, uint16_t* q /*[2][64]*/) {
    // jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE)
    int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            long temp = ((long)(t ? STD_CHROMA[i] : STD_LUMA[i]) * scale + 50L) / 100L;
            if (temp <= 0L) temp = 1L;
            if (temp > 255L) temp = 255L;
            q[t * 64 + i] = (uint16_t)temp;
        }
}

struct JpegGeom {
    int H, W;       // image
    int Hp, Wp;     // luma plane, multiples of 16
    int ch, cw;     // real chroma size: ceil(H/2), ceil(W/2)
    int Hc, Wc;     // chroma plane: Hp/2, Wp/2
};

static JpegGeom jpeg_geom(int H, int W) {
    JpegGeom g;
    g.H = H; g.W = W;
    g.Hp = (H + 15) / 16 * 16; g.Wp = (W + 15) / 16 * 16;
    g.ch = (H + 1) / 2; g.cw = (W + 1) / 2;
    g.Hc = g.Hp / 2; g.Wc = g.Wp / 2;
    return g;
}

size_t jpeg_ws_bytes(int n, int H, int W) {
    const JpegGeom g = jpeg_geom(H, W);
    return (size_t)n * ((size_t)g.Hp * g.Wp + 2 * (size_t)g.Hc * g.Wc);
}

// ---- J1: colour conversion + chroma downsample -----------------------------------------
__device__ __forceinline__ int ycc_y(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
__device__ __forceinline__ int ycc_cb(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
__device__ __forceinline__ int ycc_cr(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

// one thread per chroma sample (2x2 luma block) of the padded planes
__global__ void __launch_bounds__(256)
jpeg_forward_color_kernel(const uint8_t* __restrict__ in, const int32_t* __restrict__ idx, uint8_t* __restrict__ planes,
                          JpegGeom g, size_t plane_stride) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * g.H * g.W * 3;
    uint8_t* Y = planes + (size_t)i * plane_stride;
    uint8_t* Cb = Y + (size_t)g.Hp * g.Wp;
    uint8_t* Cr = Cb + (size_t)g.Hc * g.Wc;
    const int total = g.Hc * g.Wc;
    for (int t = blockIdx.x * 256 + threadIdx.x; t < total; t += gridDim.x * 256) {
        const int cy = t / g.Wc, cx = t - cy * g.Wc;
        // luma: edge replication == clamped reads
        int sb = 0, sr = 0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int y = 2 * cy + dy, x = 2 * cx + dx;
                const uint8_t* p = src + ((int64_t)min(y, g.H - 1) * g.W + min(x, g.W - 1)) * 3;
                Y[(size_t)y * g.Wp + x] = (uint8_t)ycc_y(p[0], p[1], p[2]);
            }
        // chroma: vertical padding replicates the last REAL downsampled row, horizontal padding
        // replicates the last input column (jcprepct.c / jcsample.c)
        const int ry = min(cy, g.ch - 1);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int y = min(2 * ry + dy, g.H - 1), x = min(2 * cx + dx, g.W - 1);
                const uint8_t* p = src + ((int64_t)y * g.W + x) * 3;
                sb += ycc_cb(p[0], p[1], p[2]);
                sr += ycc_cr(p[0], p[1], p[2]);
            }
        const int bias = (cx & 1) ? 2 : 1;
        Cb[(size_t)cy * g.Wc + cx] = (uint8_t)((sb + bias) >> 2);
        Cr[(size_t)cy * g.Wc + cx] = (uint8_t)((sr + bias) >> 2);
    }
}

// W % 4 == 0: one thread per 2 chroma samples = a 4 x 2 pixel block read as 2 x 3 aligned words (12 RGB bytes per row);
// 4 luma bytes per row and the chroma pair leave as one 32-bit / 16-bit store each.  Blocks that touch the padding
// (edge replication) take the per-sample code of jpeg_forward_color_kernel.
__device__ __forceinline__ void fwd_color_sample(const uint8_t* __restrict__ src, uint8_t* Y, uint8_t* Cb, uint8_t* Cr, const JpegGeom& g,
                                                 int cy, int cx) {
    int sb = 0, sr = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int y = 2 * cy + dy, x = 2 * cx + dx;
            const uint8_t* p = src + ((int64_t)min(y, g.H - 1) * g.W + min(x, g.W - 1)) * 3;
            Y[(size_t)y * g.Wp + x] = (uint8_t)ycc_y(p[0], p[1], p[2]);
        }
    const int ry = min(cy, g.ch - 1);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int y = min(2 * ry + dy, g.H - 1), x = min(2 * cx + dx, g.W - 1);
            const uint8_t* p = src + ((int64_t)y * g.W + x) * 3;
            sb += ycc_cb(p[0], p[1], p[2]);
            sr += ycc_cr(p[0], p[1], p[2]);
        }
    const int bias = (cx & 1) ? 2 : 1;
    Cb[(size_t)cy * g.Wc + cx] = (uint8_t)((sb + bias) >> 2);
    Cr[(size_t)cy * g.Wc + cx] = (uint8_t)((sr + bias) >> 2);
}

__global__ void __launch_bounds__(256)
jpeg_forward_color4_kernel(const uint8_t* __restrict__ in, const int32_t* __restrict__ idx, uint8_t* __restrict__ planes,
                           JpegGeom g, size_t plane_stride) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * g.H * g.W * 3;
    uint8_t* Y = planes + (size_t)i * plane_stride;
    uint8_t* Cb = Y + (size_t)g.Hp * g.Wp;
    uint8_t* Cr = Cb + (size_t)g.Hc * g.Wc;
    const int pw = g.Wc / 2, total = g.Hc * pw;
    for (int t = blockIdx.x * 256 + threadIdx.x; t < total; t += gridDim.x * 256) {
        const int cy = t / pw, q = t - cy * pw, x0 = 4 * q;
        if (x0 + 4 <= g.W && 2 * cy + 2 <= g.H) {
            int sb[2] = {0, 0}, sr[2] = {0, 0};
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int y = 2 * cy + dy;
                const uint32_t* p = reinterpret_cast<const uint32_t*>(src + ((int64_t)y * g.W + x0) * 3);
                const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
                const int r[4] = {(int)(w0 & 255u), (int)(w0 >> 24), (int)((w1 >> 16) & 255u), (int)((w2 >> 8) & 255u)};
                const int gg[4] = {(int)((w0 >> 8) & 255u), (int)(w1 & 255u), (int)(w1 >> 24), (int)((w2 >> 16) & 255u)};
                const int b[4] = {(int)((w0 >> 16) & 255u), (int)((w1 >> 8) & 255u), (int)(w2 & 255u), (int)(w2 >> 24)};
                uint32_t yw = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    yw |= (uint32_t)ycc_y(r[j], gg[j], b[j]) << (8 * j);
                    sb[j >> 1] += ycc_cb(r[j], gg[j], b[j]);
                    sr[j >> 1] += ycc_cr(r[j], gg[j], b[j]);
                }
                *reinterpret_cast<uint32_t*>(Y + (size_t)y * g.Wp + x0) = yw;
            }
            // bias 1 for even, 2 for odd chroma columns (jcsample.c h2v2_downsample)
            *reinterpret_cast<uint16_t*>(Cb + (size_t)cy * g.Wc + 2 * q) = (uint16_t)(((sb[0] + 1) >> 2) | (((sb[1] + 2) >> 2) << 8));
            *reinterpret_cast<uint16_t*>(Cr + (size_t)cy * g.Wc + 2 * q) = (uint16_t)(((sr[0] + 1) >> 2) | (((sr[1] + 2) >> 2) << 8));
        } else {
            fwd_color_sample(src, Y, Cb, Cr, g, cy, 2 * q);
            fwd_color_sample(src, Y, Cb, Cr, g, cy, 2 * q + 1);
        }
    }
}

static void launch_forward_color(const uint8_t* in, const int32_t* idx, uint8_t* planes, const JpegGeom& g, size_t plane_stride, int n,
                                 int cap, cudaStream_t st) {
    if (g.W % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0)
        jpeg_forward_color4_kernel<<<dim3(std::min(ceil_div(g.Hc * g.Wc / 2, 256), cap), n), 256, 0, st>>>(in, idx, planes, g, plane_stride);
    else
        jpeg_forward_color_kernel<<<dim3(std::min(ceil_div(g.Hc * g.Wc, 256), cap), n), 256, 0, st>>>(in, idx, planes, g, plane_stride);
}

// ---- J2: FDCT -> quantise -> dequantise -> IDCT per 8x8 block ------------------------------
// one 1-D forward pass on 8 values; pass2 selects the column-pass scaling
template <bool PASS2>
__device__ __forceinline__ void fdct8(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
    int tmp0 = d0 + d7, tmp7 = d0 - d7, tmp1 = d1 + d6, tmp6 = d1 - d6;
    int tmp2 = d2 + d5, tmp5 = d2 - d5, tmp3 = d3 + d4, tmp4 = d3 - d4;
    int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    if (!PASS2) {
        d0 = (tmp10 + tmp11) << 2;
        d4 = (tmp10 - tmp11) << 2;
    } else {
        d0 = DESCALE(tmp10 + tmp11, 2);
        d4 = DESCALE(tmp10 - tmp11, 2);
    }
    const int sh = PASS2 ? 15 : 11;
    int z1 = (tmp12 + tmp13) * FIX_0_541196100;
    d2 = DESCALE(z1 + tmp13 * FIX_0_765366865, sh);
    d6 = DESCALE(z1 + tmp12 * (-FIX_1_847759065), sh);
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    int z5 = (z3 + z4) * FIX_1_175875602;
    tmp4 *= FIX_0_298631336; tmp5 *= FIX_2_053119869; tmp6 *= FIX_3_072711026; tmp7 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
    z3 += z5; z4 += z5;
    d7 = DESCALE(tmp4 + z1 + z3, sh);
    d5 = DESCALE(tmp5 + z2 + z4, sh);
    d3 = DESCALE(tmp6 + z2 + z3, sh);
    d1 = DESCALE(tmp7 + z1 + z4, sh);
}

// one thread per 8x8 block of any component plane
__global__ void __launch_bounds__(JP_THREADS)
jpeg_block_kernel(uint8_t* __restrict__ planes, JpegGeom g, size_t plane_stride, const uint16_t* __restrict__ qtab) {
    __shared__ uint16_t q[128];
    __shared__ float qr[128];                                 // 1 / (8 q)
    if (threadIdx.x < 128) { q[threadIdx.x] = qtab[threadIdx.x]; qr[threadIdx.x] = 1.0f / (float)(qtab[threadIdx.x] << 3); }
    __syncthreads();
    const int i = blockIdx.y;
    const int yb = (g.Hp / 8) * (g.Wp / 8), cb = (g.Hc / 8) * (g.Wc / 8);
    const int total = yb + 2 * cb;
    for (int t = blockIdx.x * JP_THREADS + threadIdx.x; t < total; t += gridDim.x * JP_THREADS) {
        uint8_t* plane;
        int pitch, bidx;
        const uint16_t* qq;
        if (t < yb) { plane = planes + (size_t)i * plane_stride; pitch = g.Wp; bidx = t; qq = q; }
        else {
            const int c = (t - yb) / cb;
            plane = planes + (size_t)i * plane_stride + (size_t)g.Hp * g.Wp + (size_t)c * g.Hc * g.Wc;
            pitch = g.Wc; bidx = (t - yb) - c * cb; qq = q + 64;
        }
        const int bw = pitch / 8;
        uint8_t* base = plane + (size_t)(bidx / bw) * 8 * pitch + (size_t)(bidx % bw) * 8;
        int d[64];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint2 v = *reinterpret_cast<const uint2*>(base + (size_t)r * pitch);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                d[r * 8 + c] = (int)((v.x >> (8 * c)) & 255) - 128;
                d[r * 8 + 4 + c] = (int)((v.y >> (8 * c)) & 255) - 128;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) ROWS8(fdct8<false>, (d + 8 * r));
#pragma unroll
        for (int c = 0; c < 8; ++c) COL8(fdct8<true>, d, c);
        // quantise (divisor = q*8, round half away from zero) and dequantise
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const int qv = qq[k], div = qv << 3;
            int tcoef = d[k];
            const int neg = tcoef < 0;
            if (neg) tcoef = -tcoef;
            tcoef = div_exact(tcoef + (div >> 1), div, qr[(qq - q) + k]);
            d[k] = (neg ? -tcoef : tcoef) * qv;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) COL8(idct8<false>, d, c);
#pragma unroll
        for (int r = 0; r < 8; ++r) ROWS8(idct8<true>, (d + 8 * r));
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint2 v;
            v.x = range_limit(d[r * 8]) | (range_limit(d[r * 8 + 1]) << 8) | (range_limit(d[r * 8 + 2]) << 16) | (range_limit(d[r * 8 + 3]) << 24);
            v.y = range_limit(d[r * 8 + 4]) | (range_limit(d[r * 8 + 5]) << 8) | (range_limit(d[r * 8 + 6]) << 16) | (range_limit(d[r * 8 + 7]) << 24);
            *reinterpret_cast<uint2*>(base + (size_t)r * pitch) = v;
        }
    }
}

// ---- J3: fancy upsample + YCbCr -> RGB -----------------------------------------------------
__global__ void __launch_bounds__(256)
jpeg_inverse_color_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                          JpegGeom g, size_t plane_stride) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* Y = planes + (size_t)i * plane_stride;
    const uint8_t* Cb = Y + (size_t)g.Hp * g.Wp;
    const uint8_t* Cr = Cb + (size_t)g.Hc * g.Wc;
    uint8_t* dst = out + (int64_t)slot * g.H * g.W * 3;
    const int64_t npix = (int64_t)g.H * g.W;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npix; p += (int64_t)gridDim.x * 256) {
        const int y = (int)((uint32_t)p / (uint32_t)g.W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)g.W);
        const int yy = Y[(size_t)y * g.Wp + x];
        const int cb = up_h2v2(Cb, g.Wc, g.ch, g.cw, y, x) - 128;
        const int cr = up_h2v2(Cr, g.Wc, g.ch, g.cw, y, x) - 128;
        const int r = yy + ((91881 * cr + 32768) >> 16);
        const int gg = yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
        const int b = yy + ((116130 * cb + 32768) >> 16);
        uint8_t* o = dst + p * 3;
        o[0] = clamp255(r); o[1] = clamp255(gg); o[2] = clamp255(b);
    }
}

// W % 4 == 0: one thread per 4 horizontally adjacent pixels; they share chroma columns cx0-1 .. cx0+2 of rows cy and ny
// (6 loads per plane instead of 16), the 12 output bytes leave as three 32-bit stores.  Same integers as up_h2v2.
__global__ void __launch_bounds__(256)
jpeg_inverse_color4_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                           JpegGeom g, size_t plane_stride) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* Y = planes + (size_t)i * plane_stride;
    const uint8_t* Cb = Y + (size_t)g.Hp * g.Wp;
    const uint8_t* Cr = Cb + (size_t)g.Hc * g.Wc;
    uint8_t* dst = out + (int64_t)slot * g.H * g.W * 3;
    const int qw = g.W / 4, ngrp = g.H * qw;
    for (int t = blockIdx.x * 256 + threadIdx.x; t < ngrp; t += gridDim.x * 256) {
        const int y = t / qw, x0 = (t - y * qw) * 4;
        const uint32_t y4 = *reinterpret_cast<const uint32_t*>(Y + (size_t)y * g.Wp + x0);
        const int cy = y >> 1, nyr = min(max((y & 1) ? cy + 1 : cy - 1, 0), g.ch - 1);
        const int cx0 = x0 >> 1, im1 = max(cx0 - 1, 0), i2 = min(cx0 + 2, g.cw - 1);
        int cbv[4], crv[4];
        {
            const uint8_t* a = Cb + (size_t)cy * g.Wc;
            const uint8_t* b = Cb + (size_t)nyr * g.Wc;
            const uint32_t a01 = *reinterpret_cast<const uint16_t*>(a + cx0), b01 = *reinterpret_cast<const uint16_t*>(b + cx0);
            cbv[0] = 3 * a[im1] + b[im1]; cbv[1] = 3 * (int)(a01 & 255u) + (int)(b01 & 255u);
            cbv[2] = 3 * (int)(a01 >> 8) + (int)(b01 >> 8); cbv[3] = 3 * a[i2] + b[i2];
        }
        {
            const uint8_t* a = Cr + (size_t)cy * g.Wc;
            const uint8_t* b = Cr + (size_t)nyr * g.Wc;
            const uint32_t a01 = *reinterpret_cast<const uint16_t*>(a + cx0), b01 = *reinterpret_cast<const uint16_t*>(b + cx0);
            crv[0] = 3 * a[im1] + b[im1]; crv[1] = 3 * (int)(a01 & 255u) + (int)(b01 & 255u);
            crv[2] = 3 * (int)(a01 >> 8) + (int)(b01 >> 8); crv[3] = 3 * a[i2] + b[i2];
        }
        uint32_t px[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cur = 1 + (j >> 1), nbr = (j & 1) ? cur + 1 : cur - 1, bias = (j & 1) ? 7 : 8;
            const int cb = ((3 * cbv[cur] + cbv[nbr] + bias) >> 4) - 128, cr = ((3 * crv[cur] + crv[nbr] + bias) >> 4) - 128;
            const int yy = (y4 >> (8 * j)) & 255;
            px[3 * j] = clamp255(yy + ((91881 * cr + 32768) >> 16));
            px[3 * j + 1] = clamp255(yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
            px[3 * j + 2] = clamp255(yy + ((116130 * cb + 32768) >> 16));
        }
        uint32_t* o = reinterpret_cast<uint32_t*>(dst + ((int64_t)y * g.W + x0) * 3);
        o[0] = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
        o[1] = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
        o[2] = px[8] | (px[9] << 8) | (px[10] << 16) | (px[11] << 24);
    }
}

int run_jpeg(const CorruptArgs& a) {
    const int quality[5] = {25, 18, 15, 10, 7};
    uint16_t q[128];
    quant_tables(quality[a.severity - 1], q);
    const uint16_t* d_q = reinterpret_cast<const uint16_t*>(cached_table("jpegq_" + std::to_string(a.severity), q, sizeof(q)));
    if (!d_q) return ADVMIX_ERR_CUDA;
    const JpegGeom g = jpeg_geom(a.H, a.W);
    const size_t stride = (size_t)g.Hp * g.Wp + 2 * (size_t)g.Hc * g.Wc;
    uint8_t* planes = reinterpret_cast<uint8_t*>(a.ws);
    const int cap = std::max(1, (sm_count() * 16 + a.n - 1) / a.n);
    launch_forward_color(a.in, a.idx, planes, g, stride, a.n, cap, a.stream);
    ADVMIX_LAUNCH_OK();
    const int blocks = (g.Hp / 8) * (g.Wp / 8) + 2 * (g.Hc / 8) * (g.Wc / 8);
    jpeg_block_kernel<<<dim3(std::min(ceil_div(blocks, JP_THREADS), cap), a.n), JP_THREADS, 0, a.stream>>>(planes, g, stride, d_q);
    ADVMIX_LAUNCH_OK();
    if (a.W % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 3) == 0)
        jpeg_inverse_color4_kernel<<<dim3(std::min(ceil_div(a.H * (a.W / 4), 256), cap), a.n), 256, 0, a.stream>>>(planes, a.out, a.idx, g, stride);
    else
        jpeg_inverse_color_kernel<<<dim3(std::min(ceil_div((long long)a.H * a.W, 256), cap), a.n), 256, 0, a.stream>>>(planes, a.out, a.idx, g, stride);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}


// ============================================================================================= JPEG encoder
// The other half of SURVEY 8f rank 1: tools/make_datasets.py:45 writes every corrupted image with PIL's
// Image.save (libjpeg baseline, 4:2:0, standard Huffman tables).  The forward path above already produces
// libjpeg's quantised coefficients; this adds the entropy coder and the file framing, so the COCO-C files can
// be produced on the device, byte-identical to the ones PIL writes.
#include "const_tables.inc"

constexpr int JPEG_HDR_BYTES = 623;          // SOI, APP0, 2 x DQT, SOF0, 4 x DHT, SOS (what libjpeg emits for 3-component 4:2:0)

struct EncTab {                              // derived encoding tables (jpeg_make_c_derived_tbl): 0 DC-Y, 1 AC-Y, 2 DC-C, 3 AC-C
    uint16_t code[4][256];
    uint8_t size[4][256];
};

static void derive_enc(const uint8_t* dht, uint16_t* code, uint8_t* size) {
    const uint8_t* bits = dht + 1;           // 16 counts
    const uint8_t* vals = dht + 17;
    memset(code, 0, 256 * 2); memset(size, 0, 256);
    unsigned c = 0;
    int p = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l - 1]; ++i, ++p) { code[vals[p]] = (uint16_t)c++; size[vals[p]] = (uint8_t)l; }
        c <<= 1;
    }
}

static const uint8_t ZZ_NAT[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static void build_header(int H, int W, const uint16_t* q /*[2][64] natural*/, uint8_t* o) {
    int p = 0;
    auto put = [&](std::initializer_list<int> v) { for (int x : v) o[p++] = (uint8_t)x; };
    put({0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int t = 0; t < 2; ++t) {
        put({0xFF, 0xDB, 0x00, 0x43, t});
        for (int i = 0; i < 64; ++i) o[p++] = (uint8_t)q[t * 64 + ZZ_NAT[i]];
    }
    put({0xFF, 0xC0, 0x00, 0x11, 8, H >> 8, H & 255, W >> 8, W & 255, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
    const uint8_t* dht[4] = {JPEG_STD_DC0, JPEG_STD_AC0, JPEG_STD_DC1, JPEG_STD_AC1};
    const int dlen[4] = {(int)sizeof(JPEG_STD_DC0), (int)sizeof(JPEG_STD_AC0), (int)sizeof(JPEG_STD_DC1), (int)sizeof(JPEG_STD_AC1)};
    for (int t = 0; t < 4; ++t) {
        put({0xFF, 0xC4, (dlen[t] + 2) >> 8, (dlen[t] + 2) & 255});
        for (int i = 0; i < dlen[t]; ++i) o[p++] = dht[t][i];
    }
    put({0xFF, 0xDA, 0x00, 0x0C, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 0x3F, 0});
}

// FDCT + quantise, one thread per 8x8 block; coefficients stored in ZIG-ZAG order (what the entropy coder walks)
__global__ void __launch_bounds__(JP_THREADS)
jpeg_fdct_quant_kernel(const uint8_t* __restrict__ planes, int16_t* __restrict__ coef, JpegGeom g, size_t plane_stride,
                       size_t coef_stride, const uint16_t* __restrict__ qtab) {
    __shared__ uint16_t q[128];
    __shared__ float qr[128];                                 // 1 / (8 q)
    if (threadIdx.x < 128) { q[threadIdx.x] = qtab[threadIdx.x]; qr[threadIdx.x] = 1.0f / (float)(qtab[threadIdx.x] << 3); }
    __syncthreads();
    // natural index of zig-zag position z: compile-time, so the permutation is register naming in the unrolled store loop
    constexpr uint8_t NAT[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    const int i = blockIdx.y;
    const int yb = (g.Hp / 8) * (g.Wp / 8), cb = (g.Hc / 8) * (g.Wc / 8);
    const int total = yb + 2 * cb;
    for (int t = blockIdx.x * JP_THREADS + threadIdx.x; t < total; t += gridDim.x * JP_THREADS) {
        const uint8_t* plane;
        int pitch, bidx, qo;
        if (t < yb) { plane = planes + (size_t)i * plane_stride; pitch = g.Wp; bidx = t; qo = 0; }
        else {
            const int c = (t - yb) / cb;
            plane = planes + (size_t)i * plane_stride + (size_t)g.Hp * g.Wp + (size_t)c * g.Hc * g.Wc;
            pitch = g.Wc; bidx = (t - yb) - c * cb; qo = 64;
        }
        const int bw = pitch / 8;
        const uint8_t* base = plane + (size_t)(bidx / bw) * 8 * pitch + (size_t)(bidx % bw) * 8;
        int d[64];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint2 v = *reinterpret_cast<const uint2*>(base + (size_t)r * pitch);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                d[r * 8 + c] = (int)((v.x >> (8 * c)) & 255) - 128;
                d[r * 8 + 4 + c] = (int)((v.y >> (8 * c)) & 255) - 128;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) ROWS8(fdct8<false>, (d + 8 * r));
#pragma unroll
        for (int c = 0; c < 8; ++c) COL8(fdct8<true>, d, c);
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const int div = (int)q[qo + k] << 3;
            int tcoef = d[k];
            const int neg = tcoef < 0;
            if (neg) tcoef = -tcoef;
            tcoef = div_exact(tcoef + (div >> 1), div, qr[qo + k]);
            d[k] = neg ? -tcoef : tcoef;
        }
        // blocks in plane order (Y, Cb, Cr), coefficients in zig-zag order: eight 16-byte stores
        uint4* o = reinterpret_cast<uint4*>(coef + (size_t)i * coef_stride + (size_t)t * 64);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            uint32_t p[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                p[j] = ((uint32_t)d[NAT[8 * w + 2 * j]] & 0xFFFFu) | ((uint32_t)d[NAT[8 * w + 2 * j + 1]] << 16);
            o[w] = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
}

// scan-order block t of a 4:2:0 image -> index of the coefficient block (plane order Y, Cb, Cr) that supplies its DC.
// *dc_only is set for libjpeg's dummy blocks (jccoefct.c compress_data): luma blocks of the last MCU column / row that
// lie wholly outside ceil(W/8) x ceil(H/8) real blocks are not transformed; they carry zero AC terms and the DC of the
// previous block in the MCU buffer (right edge: the block to the left; bottom edge: whatever block 1 of the MCU holds).
__device__ __forceinline__ int enc_block_index(int t, const JpegGeom& g, int* comp, bool* dc_only) {
    const int mcu = t / 6, mcus_x = g.Wp / 16, my = mcu / mcus_x, mx = mcu - my * mcus_x;
    int j = t - mcu * 6;
    const int yb = (g.Hp / 8) * (g.Wp / 8), cb = (g.Hc / 8) * (g.Wc / 8);
    *dc_only = false;
    if (j >= 4) { *comp = j - 3; return yb + (j - 4) * cb + my * (g.Wc / 8) + mx; }
    *comp = 0;
    const int wib = (g.W + 7) >> 3, hib = (g.H + 7) >> 3;
    if ((j >> 1) && 2 * my + 1 >= hib) { j = 1; *dc_only = true; }             // dummy row: DC of MCU block 1
    if ((j & 1) && 2 * mx + 1 >= wib) { j -= 1; *dc_only = true; }             // dummy column: DC of the block to the left
    return (2 * my + (j >> 1)) * (g.Wp / 8) + 2 * mx + (j & 1);
}

__device__ __forceinline__ int bit_length(int v) { return 32 - __clz(v); }   // v >= 0

// jchuff.c encode_one_block: walks one block, either counting bits (emit == nullptr) or emitting them.
struct BitSink {
    uint32_t* words;     // big-endian bit stream, zero initialised
    uint64_t acc;
    int nacc;            // bits in acc (starts at the bit offset inside the first word)
    int wpos;
    __device__ __forceinline__ void put(uint32_t code, int size) {
        acc = (acc << size) | code;
        nacc += size;
        if (nacc >= 32) {
            atomicOr(words + wpos, (uint32_t)(acc >> (nacc - 32)));
            ++wpos;
            nacc -= 32;
        }
    }
    __device__ __forceinline__ void flush() {
        if (nacc > 0) atomicOr(words + wpos, (uint32_t)(acc << (32 - nacc)));
    }
};

// Stages one coefficient block (64 int16, zig-zag order, 16-byte aligned) in the thread's shared-memory row and returns
// the bit mask of its non-zero coefficients.  Rows are 33 words apart, so equal k of different lanes hit different banks.
constexpr int JE_ROW_WORDS = 33;
__device__ __forceinline__ uint64_t stage_block(const int16_t* __restrict__ blk, uint32_t* row) {
    const uint4* src = reinterpret_cast<const uint4*>(blk);
    uint64_t mask = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint4 v = __ldg(src + w);
        const uint32_t p[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            row[4 * w + j] = p[j];
            mask |= (uint64_t)((p[j] & 0xFFFFu) != 0u) << (8 * w + 2 * j);
            mask |= (uint64_t)((p[j] >> 16) != 0u) << (8 * w + 2 * j + 1);
        }
    }
    return mask;
}

// jchuff.c encode_one_block on a staged block: only the non-zero coefficients are visited (bit scan over the mask); the
// run length is the gap between consecutive set bits.  Returns the number of bits; EMIT also writes them.
template <bool EMIT>
__device__ __forceinline__ int encode_block(const uint32_t* row, uint64_t mask, bool dc_only, int last_dc, int comp,
                                            const EncTab& T, BitSink* sink) {
    const int dct = comp ? 2 : 0, act = comp ? 3 : 1;
    const int16_t* blk = reinterpret_cast<const int16_t*>(row);
    int bits = 0;
    int temp = (int)blk[0] - last_dc, temp2 = temp;
    if (temp < 0) { temp = -temp; --temp2; }
    int nb = bit_length(temp);
    if (EMIT) { sink->put(T.code[dct][nb], T.size[dct][nb]); if (nb) sink->put((uint32_t)temp2 & ((1u << nb) - 1u), nb); }
    bits += T.size[dct][nb] + nb;
    uint64_t m = dc_only ? 0ull : (mask & ~1ull);
    int prevk = 0;
    while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        int r = k - prevk - 1;
        prevk = k;
        while (r > 15) {
            if (EMIT) sink->put(T.code[act][0xF0], T.size[act][0xF0]);
            bits += T.size[act][0xF0];
            r -= 16;
        }
        int v = blk[k], v2 = v;
        if (v < 0) { v = -v; --v2; }
        nb = bit_length(v);
        const int sym = (r << 4) + nb;
        if (EMIT) { sink->put(T.code[act][sym], T.size[act][sym]); sink->put((uint32_t)v2 & ((1u << nb) - 1u), nb); }
        bits += T.size[act][sym] + nb;
    }
    if (prevk != 63) {                                         // trailing zeros: end of block
        if (EMIT) sink->put(T.code[act][0], T.size[act][0]);
        bits += T.size[act][0];
    }
    return bits;
}

// One CTA per image: bit length of every block, exclusive scan -> bit offsets, emission, padding, byte stuffing.
constexpr int JE_THREADS = 256;
__global__ void __launch_bounds__(JE_THREADS)
jpeg_entropy_encode_kernel(const int16_t* __restrict__ coef, size_t coef_stride, JpegGeom g, const EncTab* __restrict__ tabs,
                           const uint8_t* __restrict__ header, uint32_t* __restrict__ words_all, size_t words_stride,
                           int* __restrict__ offs_all, uint8_t* __restrict__ out, size_t out_stride, int32_t* __restrict__ lengths) {
    __shared__ EncTab T;
    __shared__ uint32_t s_rows[JE_THREADS * JE_ROW_WORDS];       // one staged block per thread
    __shared__ int s_scan[JE_THREADS / 32 + 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, img = blockIdx.x;
    uint32_t* row = s_rows + tid * JE_ROW_WORDS;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tabs);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        for (int i = tid; i < (int)(sizeof(EncTab) / 4); i += JE_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const int16_t* cf = coef + (size_t)img * coef_stride;
    uint32_t* words = words_all + (size_t)img * words_stride;
    uint8_t* dst = out + (size_t)img * out_stride;
    const int nblk = (g.Hp / 16) * (g.Wp / 16) * 6;
    auto last_dc_of = [&](int t, int comp) -> int {          // DC of the previous block of the same component in scan order
        const int j = t % 6;
        int pt;
        if (comp == 0) pt = j == 0 ? t - 6 + 3 : t - 1; else pt = t - 6;
        if (pt < 0) return 0;
        int c2;
        bool d2;
        return cf[(size_t)enc_block_index(pt, g, &c2, &d2) * 64];
    };
    // 1. bit length per block
    const int per = (nblk + JE_THREADS - 1) / JE_THREADS;
    const int lo = tid * per, hi = min(lo + per, nblk);
    int sum = 0;
    for (int t = lo; t < hi; ++t) {
        int comp;
        bool dc_only;
        const int16_t* blk = cf + (size_t)enc_block_index(t, g, &comp, &dc_only) * 64;
        const uint64_t mask = stage_block(blk, row);
        sum += encode_block<false>(row, mask, dc_only, last_dc_of(t, comp), comp, T, nullptr);
    }
    // 2. exclusive scan over the blocks (thread-contiguous runs)
    int v = sum;
    for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, off); if ((tid & 31) >= off) v += o; }
    if ((tid & 31) == 31) s_scan[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        int w = tid < JE_THREADS / 32 ? s_scan[tid] : 0;
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, off); if (tid >= off) w += o; }
        s_scan[8 + tid] = w;
    }
    __syncthreads();
    int run = v - sum + ((tid >> 5) ? s_scan[8 + (tid >> 5) - 1] : 0);
    if (tid == JE_THREADS - 1) s_total = run + sum;
    __syncthreads();
    const int total_bits = s_total;
    const int nbytes = (total_bits + 7) >> 3;
    if ((size_t)nbytes + 8 > words_stride * 4) { if (tid == 0) lengths[img] = -1; return; }   // cannot happen: <= 26 bits per sample
    // 3. emission
    for (int t = lo; t < hi; ++t) {
        int comp;
        bool dc_only;
        const int16_t* blk = cf + (size_t)enc_block_index(t, g, &comp, &dc_only) * 64;
        const uint64_t mask = stage_block(blk, row);
        BitSink sink{words, 0, run & 31, run >> 5};
        run += encode_block<true>(row, mask, dc_only, last_dc_of(t, comp), comp, T, &sink);
        sink.flush();
    }
    if (tid == 0 && (total_bits & 7)) {                       // pad the last byte with 1-bits
        const int pad = 8 - (total_bits & 7);
        atomicOr(words + (total_bits >> 5), ((1u << pad) - 1u) << (32 - (total_bits & 31) - pad));
    }
    __syncthreads();
    if (tid == 0) s_total = 0;
    __syncthreads();
    // 4. header, byte stuffing (FF -> FF 00), EOI.  The stuffed size is counted first so that a file that does not fit
    // out_stride is reported (length -1) instead of written.
    {
        int nff = 0;
        for (int w = tid; w < (nbytes + 3) >> 2; w += JE_THREADS) {
            const uint32_t v = words[w];
            const int nb = min(4, nbytes - 4 * w);
            for (int j = 0; j < nb; ++j) nff += ((v >> (24 - 8 * j)) & 255u) == 255u;
        }
        for (int off = 16; off; off >>= 1) nff += __shfl_xor_sync(0xffffffffu, nff, off);
        if ((tid & 31) == 0 && nff) atomicAdd(&s_total, nff);
    }
    __syncthreads();
    if ((size_t)JPEG_HDR_BYTES + (size_t)nbytes + (size_t)s_total + 2 > out_stride) { if (tid == 0) lengths[img] = -1; return; }
    for (int i = tid; i < JPEG_HDR_BYTES; i += JE_THREADS) dst[i] = header[i];
    int base = JPEG_HDR_BYTES;
    for (int b0 = 0; b0 < nbytes; b0 += JE_THREADS * 16) {
        const int lo2 = b0 + tid * 16, hi2 = min(lo2 + 16, nbytes);
        uint8_t by[16];
        int nff = 0, cnt = max(hi2 - lo2, 0);
        for (int j = 0; j < cnt; ++j) {
            const int i = lo2 + j;
            by[j] = (uint8_t)(words[i >> 2] >> (24 - 8 * (i & 3)));
            nff += by[j] == 0xFF;
        }
        int w2 = cnt + nff;
        int incl = w2;
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if ((tid & 31) >= off) incl += o; }
        __syncthreads();
        if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            int w = tid < JE_THREADS / 32 ? s_scan[tid] : 0;
            for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, off); if (tid >= off) w += o; }
            s_scan[8 + tid] = w;
        }
        __syncthreads();
        int o = base + incl - w2 + ((tid >> 5) ? s_scan[8 + (tid >> 5) - 1] : 0);
        for (int j = 0; j < cnt; ++j) {
            dst[o++] = by[j];
            if (by[j] == 0xFF) dst[o++] = 0;
        }
        base += s_scan[8 + JE_THREADS / 32 - 1];
        __syncthreads();
    }
    if (tid == 0) { dst[base] = 0xFF; dst[base + 1] = 0xD9; lengths[img] = base + 2; }
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

struct EncodeWs { size_t planes, coef, words, offs, total, plane_stride, coef_stride, words_stride; };
static EncodeWs encode_ws(int n, int H, int W) {
    const JpegGeom g = jpeg_geom(H, W);
    EncodeWs w;
    w.plane_stride = (size_t)g.Hp * g.Wp + 2 * (size_t)g.Hc * g.Wc;
    w.coef_stride = w.plane_stride;                           // one int16 per sample
    w.words_stride = w.plane_stride + 16;                     // entropy-coded bits: a coefficient costs <= 16 + 10 bits, so 4 bytes per sample always fit
    const size_t nblk = (size_t)(g.Hp / 16) * (g.Wp / 16) * 6 + 1;
    size_t off = 0;
    w.planes = off; off += align256(n * w.plane_stride);
    w.coef = off; off += align256(n * w.coef_stride * 2);
    w.words = off; off += align256(n * w.words_stride * 4);
    w.offs = off; off += align256(n * nblk * 4);
    w.total = off;
    return w;
}

}  // namespace advmix

using namespace advmix;

namespace advmix {

// ---- packing the encoded files: file i (lengths[i] bytes at files + i * stride) -> packed + offsets[i]; offsets[n] = total.
// Only the encoded bytes then cross PCIe / reach host memory (a typical file takes a quarter of a conservative per-file slot).
constexpr int PK_THREADS = 256;
__global__ void __launch_bounds__(1024) pack_offsets_kernel(const int32_t* __restrict__ lengths, int n, int64_t* __restrict__ offsets) {
    __shared__ int64_t s_part[1024];
    // thread t sums its contiguous chunk, a block scan over the chunk sums, then the chunk is written out
    const int per = (n + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, n), hi = min(lo + per, n);
    int64_t sum = 0;
    for (int i = lo; i < hi; ++i) sum += max(lengths[i], 0);
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int64_t v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int64_t run = s_part[threadIdx.x] - sum;            // exclusive prefix of this chunk
    for (int i = lo; i < hi; ++i) { offsets[i] = run; run += max(lengths[i], 0); }
    if (threadIdx.x == 1023) offsets[n] = s_part[1023];
}

__global__ void __launch_bounds__(PK_THREADS) pack_copy_kernel(const uint8_t* __restrict__ files, size_t stride, const int32_t* __restrict__ lengths,
                                                                const int64_t* __restrict__ offsets, uint8_t* __restrict__ packed) {
    const int i = blockIdx.y;
    const int len = max(lengths[i], 0);
    const uint8_t* src = files + (size_t)i * stride;
    uint8_t* dst = packed + offsets[i];
    // the source is 16-byte aligned (stride % 16 == 0), the destination is not: 16 source bytes per thread, byte stores
    for (int q = blockIdx.x * PK_THREADS + threadIdx.x; q * 16 < len; q += gridDim.x * PK_THREADS) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + q);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int b0 = q * 16, nb = min(16, len - b0);
        if ((reinterpret_cast<uintptr_t>(dst + b0) & 3) == 0 && nb == 16) {
            uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + b0);
            d4[0] = w[0]; d4[1] = w[1]; d4[2] = w[2]; d4[3] = w[3];
        } else {
            for (int k = 0; k < nb; ++k) dst[b0 + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
        }
    }
}

}  // namespace advmix

extern "C" {

size_t advmix_jpeg_encode_workspace_bytes(int n, int H, int W) {
    if (n <= 0 || H <= 0 || W <= 0) return 0;
    return encode_ws(n, H, W).total;
}

int advmix_jpeg_encode_u8c3(const uint8_t* images, int n, int H, int W, int quality, uint8_t* out, size_t out_stride,
                            int32_t* lengths, void* workspace, size_t ws_bytes, advmix_stream_t stream) {
    ADVMIX_REQUIRE(n >= 0 && H > 0 && W > 0 && H < 65536 && W < 65536, "jpeg_encode: bad shape");
    ADVMIX_REQUIRE(quality >= 1 && quality <= 100, "jpeg_encode: quality %d outside 1..100", quality);
    if (n == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(images && out && lengths && workspace, "jpeg_encode: null argument");
    ADVMIX_REQUIRE(n <= 65535, "jpeg_encode: n <= 65535 per call");
    const EncodeWs w = encode_ws(n, H, W);
    if (ws_bytes < w.total) return fail(ADVMIX_ERR_WORKSPACE, "jpeg_encode: workspace %zu < %zu bytes", ws_bytes, w.total);
    ADVMIX_REQUIRE(out_stride >= (size_t)JPEG_HDR_BYTES + 64, "jpeg_encode: out_stride too small");
    cudaStream_t st = as_stream(stream);
    uint16_t q[128];
    quant_tables(quality, q);
    const uint16_t* d_q = reinterpret_cast<const uint16_t*>(cached_table("jpegq_enc_" + std::to_string(quality), q, sizeof(q)));
    uint8_t hdr[JPEG_HDR_BYTES + 1];
    build_header(H, W, q, hdr);
    const uint8_t* d_hdr = reinterpret_cast<const uint8_t*>(
        cached_table("jpeghdr_" + std::to_string(H) + "x" + std::to_string(W) + "_" + std::to_string(quality), hdr, JPEG_HDR_BYTES));
    static const EncTab etab = [] {                 // thread-safe one-time initialisation (C++11 magic static)
        EncTab t;
        derive_enc(JPEG_STD_DC0, t.code[0], t.size[0]);
        derive_enc(JPEG_STD_AC0, t.code[1], t.size[1]);
        derive_enc(JPEG_STD_DC1, t.code[2], t.size[2]);
        derive_enc(JPEG_STD_AC1, t.code[3], t.size[3]);
        return t;
    }();
    const EncTab* d_tab = reinterpret_cast<const EncTab*>(cached_table("jpeg_enctab", &etab, sizeof(etab)));
    if (!d_q || !d_hdr || !d_tab) return ADVMIX_ERR_CUDA;
    const JpegGeom g = jpeg_geom(H, W);
    char* ws = reinterpret_cast<char*>(workspace);
    uint8_t* planes = reinterpret_cast<uint8_t*>(ws + w.planes);
    int16_t* coef = reinterpret_cast<int16_t*>(ws + w.coef);
    uint32_t* words = reinterpret_cast<uint32_t*>(ws + w.words);
    int* offs = reinterpret_cast<int*>(ws + w.offs);
    const int cap = std::max(1, (sm_count() * 16 + n - 1) / n);
    launch_forward_color(images, nullptr, planes, g, w.plane_stride, n, cap, st);
    ADVMIX_LAUNCH_OK();
    const int blocks = (g.Hp / 8) * (g.Wp / 8) + 2 * (g.Hc / 8) * (g.Wc / 8);
    jpeg_fdct_quant_kernel<<<dim3(std::min(ceil_div(blocks, JP_THREADS), cap), n), JP_THREADS, 0, st>>>(planes, coef, g, w.plane_stride,
                                                                                                       w.coef_stride, d_q);
    ADVMIX_LAUNCH_OK();
    ADVMIX_CUDA_OK(cudaMemsetAsync(words, 0, (size_t)n * w.words_stride * 4, st));
    jpeg_entropy_encode_kernel<<<n, JE_THREADS, 0, st>>>(coef, w.coef_stride, g, d_tab, d_hdr, words, w.words_stride, offs, out,
                                                        out_stride, lengths);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_pack_files(const uint8_t* files, size_t stride, const int32_t* lengths, int n, uint8_t* packed, int64_t* offsets,
                      advmix_stream_t stream) {
    ADVMIX_REQUIRE(n >= 0, "pack_files: n = %d", n);
    ADVMIX_REQUIRE(offsets != nullptr, "pack_files: null offsets");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        ADVMIX_CUDA_OK(cudaMemsetAsync(offsets, 0, sizeof(int64_t), st));
        return ADVMIX_OK;
    }
    ADVMIX_REQUIRE(files && lengths && packed, "pack_files: null argument");
    ADVMIX_REQUIRE(stride % 16 == 0 && (reinterpret_cast<uintptr_t>(files) & 15) == 0, "pack_files: files / stride must be 16-byte aligned");
    ADVMIX_REQUIRE(n <= 65535, "pack_files: n = %d > 65535 per call", n);
    pack_offsets_kernel<<<1, 1024, 0, st>>>(lengths, n, offsets);
    ADVMIX_LAUNCH_OK();
    const int bx = (int)std::min<size_t>((stride / 16 + PK_THREADS - 1) / PK_THREADS, 8);
    pack_copy_kernel<<<dim3(bx, n), PK_THREADS, 0, st>>>(files, stride, lengths, offsets, packed);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"


namespace advmix {
}  // namespace advmix
