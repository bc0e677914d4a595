// jpeg_compression: the numeric round trip of PIL's save(quality=q) -> open(), i.e.
// libjpeg(-turbo) baseline 4:2:0 with the islow DCT, minus the (lossless) entropy coding:
//   RGB->YCbCr (16-bit fixed tables) -> h2v2 box downsample (bias 1,2,1,2) -> level shift,
//   islow FDCT -> quantise -> dequantise -> islow IDCT + range limit -> h2v2 fancy (triangle)
//   upsample -> YCbCr->RGB.  All integer; bit-exact against PIL is the bar.
#include "corrupt_common.cuh"
#include "jpeg_common.cuh"

#include <vector>

namespace advmix {

constexpr int JP_THREADS = 128;

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
    if (threadIdx.x < 128) q[threadIdx.x] = qtab[threadIdx.x];
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
            tcoef = (tcoef + (div >> 1)) / div;
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
        const int y = (int)(p / g.W), x = (int)(p - (int64_t)y * g.W);
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
    jpeg_forward_color_kernel<<<dim3(std::min(ceil_div(g.Hc * g.Wc, 256), cap), a.n), 256, 0, a.stream>>>(a.in, a.idx, planes, g, stride);
    ADVMIX_LAUNCH_OK();
    const int blocks = (g.Hp / 8) * (g.Wp / 8) + 2 * (g.Hc / 8) * (g.Wc / 8);
    jpeg_block_kernel<<<dim3(std::min(ceil_div(blocks, JP_THREADS), cap), a.n), JP_THREADS, 0, a.stream>>>(planes, g, stride, d_q);
    ADVMIX_LAUNCH_OK();
    jpeg_inverse_color_kernel<<<dim3(std::min(ceil_div((long long)a.H * a.W, 256), cap), a.n), 256, 0, a.stream>>>(planes, a.out, a.idx, g, stride);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // namespace advmix
