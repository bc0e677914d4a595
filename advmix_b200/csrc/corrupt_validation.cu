// The stencil members of imagecorruptions' 'validation' set: gaussian_blur and spatter.
// (speckle_noise and saturate are point kernels and live in corrupt_point.cu.)
//
// spatter, severities 1-3, runs a chain of cv2 primitives on a single-channel liquid map: Canny ->
// distanceTransform(DIST_L2, 5) -> threshold -> blur -> equalizeHist -> filter2D -> blur.  Each stage is
// restated here operation by operation (integer Sobel / fixed-point tangent tests, chamfer costs taken from
// cv2, float64 box sums, cv2's rounding) so the uint8 result is bit-identical to the cv2 chain the oracle calls.
#include "stencil_common.cuh"

namespace advmix {

// ======================================================================== gaussian_blur
static size_t gauss2d_smem(int r0, int r1, int C) {
    const size_t cols = GT_COLS + 2 * (size_t)r1 * C;
    return ((size_t)(GT_ROWS + 2 * r0) * cols + (size_t)GT_ROWS * cols) * sizeof(double);
}

int run_gaussian_blur(const CorruptArgs& a) {
    const double sig[5] = {1, 2, 3, 4, 6};
    int radius;
    std::vector<double> hw;
    const double* d_w = gauss_table(sig[a.severity - 1], 4.0, &radius, &hw);
    if (!d_w) return ADVMIX_ERR_CUDA;
    const int H = a.H, WC = a.W * 3;
    const int64_t img = (int64_t)H * WC;
    if (a.fast) {                                 // ADVMIX_CORRUPT_FAST: float32, all five radii in the fused kernel
        const float top = fast_top(gauss1d_unit_response(hw.data(), radius, gauss1d_unit_response(hw.data(), radius, 1.0)));
        const int rcu = launch_gauss_u8_fast(a.in, a.idx, a.out, a.idx, a.n, H, a.W, radius, d_w, top >= 1.0f ? 255.0f : 254.9999f, a.stream);
        if (rcu != -1) return rcu;
        const int rc = launch_gauss2d_fast(LoadU8Div255F{a.in, a.idx, img, WC, nullptr}, StoreU8Trunc255F{a.out, a.idx, img, WC, 1, top}, a.n, H, WC, 3,
                                           radius, radius, d_w, d_w, BORDER_NEAREST, a.stream);
        if (rc != -1) return rc;
    }
    if (gauss2d_smem(radius, radius, 3) <= 160 * 1024)
        return launch_gauss2d(LoadU8Div255{a.in, a.idx, img, WC, nullptr}, StoreU8Trunc255{a.out, a.idx, img, WC, 1}, a.n, H, WC, 3,
                              radius, radius, d_w, d_w, BORDER_NEAREST, a.stream);
    // sigma 6 (radius 24): the fused tile does not fit in shared memory; two passes through a float64 image
    double* tmp = reinterpret_cast<double*>(a.ws);
    int rc = launch_gauss(LoadU8Div255{a.in, a.idx, img, WC, nullptr}, StoreF64{tmp, img, WC}, a.n, H, WC, 3, 0, radius, d_w,
                          BORDER_NEAREST, a.stream);
    if (rc) return rc;
    return launch_gauss(LoadF64{tmp, img, WC}, StoreU8Trunc255{a.out, a.idx, img, WC, 1}, a.n, H, WC, 3, 1, radius, d_w,
                        BORDER_NEAREST, a.stream);
}

// ======================================================================== spatter
struct SpatterParams { double c0, c1, sigma, c3, c4; int mud; };
static SpatterParams spatter_params(int s) {
    const SpatterParams p[5] = {{0.65, 0.3, 4, 0.69, 0.6, 0}, {0.65, 0.3, 3, 0.68, 0.6, 0}, {0.65, 0.3, 2, 0.68, 0.5, 0},
                                {0.65, 0.3, 1, 0.65, 1.5, 1}, {0.67, 0.4, 1, 0.65, 1.5, 1}};
    return p[s - 1];
}

struct LoadLiquid {             // np.random.normal(loc=c0, scale=c1) from the materialised N(0,1) field [H][W]
    const float* field; size_t field_stride; int W; double c0, c1;
    __device__ void init() {}
    typedef const float* Row;
    __device__ Row row(int img, int y) const {
        return reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)img * field_stride) + (size_t)y * W;
    }
    __device__ double at(Row r, int x) const { return c0 + c1 * (double)r[x]; }
    typedef float Raw;
    __device__ Raw raw(Row r, int x) const { return r[x]; }
    __device__ double cvt(Raw v) const { return c0 + c1 * (double)v; }
};
struct StoreLiquidU8 {          // liquid[liquid < c3] = 0 ; (liquid * 255).astype(np.uint8)
    uint8_t* base; int64_t plane; int W; double c3;
    __device__ void operator()(int img, int y, int x, double v) const {
        if (v < c3) v = 0.0;
        base[(int64_t)img * plane + (int64_t)y * W + x] = (uint8_t)(long long)(v * 255.0);
    }
};
struct StoreLiquidMask {        // np.where(liquid > c3, 1, 0).astype(np.float32)
    float* base; int64_t plane; int W; double c3;
    __device__ void operator()(int img, int y, int x, double v) const {
        base[(int64_t)img * plane + (int64_t)y * W + x] = v > c3 ? 1.0f : 0.0f;
    }
};
struct LoadF32Plane {
    const float* base; int64_t plane; int W;
    __device__ void init() {}
    typedef const float* Row;
    __device__ Row row(int img, int y) const { return base + (int64_t)img * plane + (int64_t)y * W; }
    __device__ double at(Row r, int x) const { return (double)r[x]; }
    typedef float Raw;
    __device__ Raw raw(Row r, int x) const { return r[x]; }
    __device__ double cvt(Raw v) const { return (double)v; }
};
struct StoreMud {               // float32 result of skimage.gaussian on a float32 image ; m[m < 0.8] = 0
    float* base; int64_t plane; int W;
    __device__ void operator()(int img, int y, int x, double v) const {
        float m = (float)v;
        if (m < 0.8f) m = 0.0f;
        base[(int64_t)img * plane + (int64_t)y * W + x] = m;
    }
};
template <> struct MidF32<StoreMud> { static constexpr bool value = true; };

// mud (severities 4-5): x *= (1 - m) in float32 ; clip(x + 63|42|20/255 * m, 0, 1) * 255 in float64
__global__ void __launch_bounds__(ST_THREADS)
spatter_mud_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                   const float* __restrict__ m, int64_t plane) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* s = in + (int64_t)slot * plane * 3;
    uint8_t* d = out + (int64_t)slot * plane * 3;
    const float* mi = m + (int64_t)i * plane;
    const double col[3] = {63 / 255., 42 / 255., 20 / 255.};
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < plane; p += (int64_t)gridDim.x * ST_THREADS) {
        const float mm = mi[p];
        const float keep = __fsub_rn(1.0f, mm);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = __fmul_rn(__fdiv_rn((float)s[3 * p + c], 255.0f), keep);
            const double v = (double)x + col[c] * (double)mm;
            d[3 * p + c] = trunc_u8(clip01(v) * 255.0);
        }
    }
}

// ---- cv2.Canny(img, 50, 150): 3x3 Sobel (BORDER_REPLICATE), L1 magnitude, fixed-point sector test ----------
constexpr int CN_T = 32;        // output tile
__global__ void __launch_bounds__(256)
canny_nms_kernel(const uint8_t* __restrict__ l8, uint8_t* __restrict__ map, int H, int W, int low, int high) {
    __shared__ uint8_t s_px[CN_T + 4][CN_T + 4];
    __shared__ int s_mag[CN_T + 2][CN_T + 2];
    __shared__ short s_dx[CN_T][CN_T], s_dy[CN_T][CN_T];
    const int64_t plane = (int64_t)H * W;
    const uint8_t* src = l8 + (int64_t)blockIdx.z * plane;
    uint8_t* dst = map + (int64_t)blockIdx.z * plane;
    const int x0 = blockIdx.x * CN_T, y0 = blockIdx.y * CN_T;
    for (int t = threadIdx.x; t < (CN_T + 4) * (CN_T + 4); t += 256) {
        const int ty = t / (CN_T + 4), tx = t - ty * (CN_T + 4);
        s_px[ty][tx] = src[(int64_t)clampi(y0 + ty - 2, 0, H - 1) * W + clampi(x0 + tx - 2, 0, W - 1)];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (CN_T + 2) * (CN_T + 2); t += 256) {
        const int ty = t / (CN_T + 2), tx = t - ty * (CN_T + 2);
        const int y = y0 + ty - 1, x = x0 + tx - 1;
        int mag = 0;                                               // cv2's magnitude buffer has a zero frame
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const uint8_t(*p)[CN_T + 4] = s_px;
            const int cy = ty + 1, cx = tx + 1;
            const int dx = (p[cy - 1][cx + 1] + 2 * p[cy][cx + 1] + p[cy + 1][cx + 1]) - (p[cy - 1][cx - 1] + 2 * p[cy][cx - 1] + p[cy + 1][cx - 1]);
            const int dy = (p[cy + 1][cx - 1] + 2 * p[cy + 1][cx] + p[cy + 1][cx + 1]) - (p[cy - 1][cx - 1] + 2 * p[cy - 1][cx] + p[cy - 1][cx + 1]);
            mag = abs(dx) + abs(dy);
            if (ty >= 1 && ty <= CN_T && tx >= 1 && tx <= CN_T) { s_dx[ty - 1][tx - 1] = (short)dx; s_dy[ty - 1][tx - 1] = (short)dy; }
        }
        s_mag[ty][tx] = mag;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < CN_T * CN_T; t += 256) {
        const int ty = t / CN_T, tx = t - ty * CN_T;
        const int y = y0 + ty, x = x0 + tx;
        if (y >= H || x >= W) continue;
        const int m = s_mag[ty + 1][tx + 1];
        uint8_t r = 1;                                             // 1: not an edge, 0: weak candidate, 2: edge
        if (m > low) {
            const int xs = s_dx[ty][tx], ys = s_dy[ty][tx];
            const int ax = abs(xs), ay = abs(ys) << 15;
            const int tg22x = ax * 13573;                          // tan(22.5 deg) * 2^15
            bool keep;
            if (ay < tg22x) keep = m > s_mag[ty + 1][tx] && m >= s_mag[ty + 1][tx + 2];
            else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = m > s_mag[ty][tx + 1] && m >= s_mag[ty + 2][tx + 1];
                else {
                    const int sg = (xs ^ ys) < 0 ? -1 : 1;
                    keep = m > s_mag[ty][tx + 1 - sg] && m > s_mag[ty + 2][tx + 1 + sg];
                }
            }
            if (keep) r = m > high ? 2 : 0;
        }
        dst[(int64_t)y * W + x] = r;
    }
}

// hysteresis: weak candidates 8-connected to an edge become edges.  One CTA per image sweeps until nothing
// changes; updates are monotone (0 -> 2), so the fixed point does not depend on the sweep order.
__global__ void __launch_bounds__(1024)
canny_hysteresis_kernel(uint8_t* __restrict__ map, int H, int W) {
    volatile uint8_t* m = map + (int64_t)blockIdx.x * H * W;
    const int npx = H * W;
    for (;;) {
        int changed = 0;
        for (int p = threadIdx.x; p < npx; p += 1024) {
            if (m[p] != 0) continue;
            const int y = p / W, x = p - y * W;
            bool hit = false;
            for (int dy = -1; dy <= 1 && !hit; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = x + dx;
                    if (xx >= 0 && xx < W && m[yy * W + xx] == 2) { hit = true; break; }
                }
            }
            if (hit) { m[p] = 2; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
}

// per pixel: distance to the nearest edge pixel in its own row, to the left (>= 0, 0 = itself) and to the
// right (>= 1), capped at CHAMFER_R + 1 = "none".  Packed (left | right << 8).
__global__ void __launch_bounds__(ST_THREADS)
row_nearest_kernel(const uint8_t* __restrict__ map, uint16_t* __restrict__ g, int H, int W) {
    const int64_t plane = (int64_t)H * W;
    const uint8_t* m = map + (int64_t)blockIdx.y * plane;
    uint16_t* o = g + (int64_t)blockIdx.y * plane;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < plane; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)W);
        const uint8_t* row = m + (int64_t)y * W;
        int gl = CHAMFER_R + 1, gr = CHAMFER_R + 1;
        for (int k = 0; k <= CHAMFER_R && x - k >= 0; ++k)
            if (row[x - k] == 2) { gl = k; break; }
        for (int k = 1; k <= CHAMFER_R && x + k < W; ++k)
            if (row[x + k] == 2) { gr = k; break; }
        o[p] = (uint16_t)(gl | (gr << 8));
    }
}

// min(cv2.distanceTransform(255 - edges, DIST_L2, 5), 20): minimum over edge pixels q of the chamfer cost of
// (p - q).  Within a source row the cost grows with |dx| on either side, so the nearest edge pixel per side
// is enough: 2 table look-ups per row offset.
__global__ void __launch_bounds__(ST_THREADS)
chamfer_kernel(const uint16_t* __restrict__ g, const float* __restrict__ tab, float* __restrict__ dist, int H, int W) {
    constexpr int R = CHAMFER_R, N = 2 * CHAMFER_R + 1;
    __shared__ float s_t[N * N];
    for (int t = threadIdx.x; t < N * N; t += ST_THREADS) s_t[t] = tab[t];
    __syncthreads();
    const int64_t plane = (int64_t)H * W;
    const uint16_t* gi = g + (int64_t)blockIdx.y * plane;
    float* o = dist + (int64_t)blockIdx.y * plane;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < plane; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)W);
        float best = 20.0f;
        for (int dy = -R; dy <= R; ++dy) {            // source row y - dy
            const int yy = y - dy;
            if (yy < 0 || yy >= H) continue;
            const uint32_t v = gi[(int64_t)yy * W + x];
            const int gl = v & 255, gr = v >> 8;
            const float* row = s_t + (dy + R) * N + R;
            if (gl <= R) best = fminf(best, row[gl]);
            if (gr <= R) best = fminf(best, row[-gr]);
        }
        o[p] = best;
    }
}

// cv2.blur(float32, (3,3)) (float64 sums, BORDER_REFLECT_101) -> astype(uint8), plus the image histogram
__global__ void __launch_bounds__(ST_THREADS)
blur_f32_hist_kernel(const float* __restrict__ dist, uint8_t* __restrict__ q, unsigned int* __restrict__ hist, int H, int W) {
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t plane = (int64_t)H * W;
    const float* d = dist + (int64_t)blockIdx.y * plane;
    uint8_t* o = q + (int64_t)blockIdx.y * plane;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < plane; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)W);
        double s = 0.0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const float* row = d + (int64_t)reflect101(y + dy, H) * W;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) s += (double)row[reflect101(x + dx, W)];
        }
        const float b = (float)(s * (1.0 / 9));
        const uint8_t v = (uint8_t)(int)b;
        o[p] = v;
        atomicAdd(&s_h[v], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&hist[(int64_t)blockIdx.y * 256 + threadIdx.x], s_h[threadIdx.x]);
}

// cv2.equalizeHist (LUT from the histogram) followed by cv2.filter2D(CV_8U, [[-2,-1,0],[-1,1,1],[0,1,2]])
__global__ void __launch_bounds__(ST_THREADS)
equalize_emboss_kernel(const uint8_t* __restrict__ q, const unsigned int* __restrict__ hist, uint8_t* __restrict__ f, int H, int W) {
    __shared__ uint8_t s_lut[256];
    const int64_t plane = (int64_t)H * W;
    if (threadIdx.x == 0) {
        const unsigned int* h = hist + (int64_t)blockIdx.y * 256;
        int i = 0;
        while (i < 255 && !h[i]) ++i;
        const int total = (int)plane;
        for (int k = 0; k < 256; ++k) s_lut[k] = 0;
        if ((int)h[i] == total) s_lut[i] = (uint8_t)i;          // constant image is left unchanged
        else {
            const float scale = 255.0f / (float)(total - (int)h[i]);
            int sum = 0;
            for (int k = i + 1; k < 256; ++k) {
                sum += (int)h[k];
                s_lut[k] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn((float)sum, scale))));
            }
        }
    }
    __syncthreads();
    const uint8_t* s = q + (int64_t)blockIdx.y * plane;
    uint8_t* o = f + (int64_t)blockIdx.y * plane;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < plane; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)W);
        const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H), xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
        const uint8_t* r0 = s + (int64_t)ym * W;
        const uint8_t* r1 = s + (int64_t)y * W;
        const uint8_t* r2 = s + (int64_t)yp * W;
        const int v = -2 * s_lut[r0[xm]] - s_lut[r0[x]] - s_lut[r1[xm]] + s_lut[r1[x]] + s_lut[r1[xp]] + s_lut[r2[x]] + 2 * s_lut[r2[xp]];
        o[p] = (uint8_t)min(255, max(0, v));
    }
}

// cv2.blur(uint8, (3,3)).astype(float32) ; m = liquid_u8 * dist ; per-image max of m
__global__ void __launch_bounds__(ST_THREADS)
blur_u8_mask_kernel(const uint8_t* __restrict__ f, const uint8_t* __restrict__ l8, float* __restrict__ m, unsigned int* __restrict__ mmax,
                    int H, int W) {
    const int64_t plane = (int64_t)H * W;
    const uint8_t* s = f + (int64_t)blockIdx.y * plane;
    const uint8_t* l = l8 + (int64_t)blockIdx.y * plane;
    float* o = m + (int64_t)blockIdx.y * plane;
    float best = 0.0f;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < plane; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)W);
        int sum = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const uint8_t* row = s + (int64_t)reflect101(y + dy, H) * W;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) sum += row[reflect101(x + dx, W)];
        }
        const float d2 = (float)__double2int_rn((double)sum * (1.0 / 9));
        const float v = __fmul_rn((float)l[p], d2);
        o[p] = v;
        best = fmaxf(best, v);
    }
    for (int off = 16; off; off >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0 && best > 0.0f) atomicMax(&mmax[blockIdx.y], __float_as_uint(best));   // non-negative floats order like uints
}

// water (severities 1-3): clip(x + (m / max(m)) * c4 * (175|238|238)/255, 0, 1) * 255, float32
__global__ void __launch_bounds__(ST_THREADS)
spatter_water_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                     const float* __restrict__ m, const unsigned int* __restrict__ mmax, int64_t plane, float c4) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* s = in + (int64_t)slot * plane * 3;
    uint8_t* d = out + (int64_t)slot * plane * 3;
    const float* mi = m + (int64_t)i * plane;
    const float mx = __uint_as_float(mmax[i]);
    const float col[3] = {(float)(175 / 255.), (float)(238 / 255.), (float)(238 / 255.)};
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < plane; p += (int64_t)gridDim.x * ST_THREADS) {
        const float mm = __fmul_rn(__fdiv_rn(mi[p], mx), c4);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = __fdiv_rn((float)s[3 * p + c], 255.0f);
            float v = __fadd_rn(x, __fmul_rn(mm, col[c]));
            // max(m) == 0 (no liquid above the threshold) makes the reference divide 0/0: NaN survives np.clip
            // and np.uint8(NaN) is 0 on x86
            v = v != v ? 0.0f : fminf(fmaxf(v, 0.0f), 1.0f);
            d[3 * p + c] = (uint8_t)(int)__fmul_rn(v, 255.0f);
        }
    }
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

struct SpatterWs {
    size_t field, l8, map, g, dist, q, hist, f, m, total;
};
static SpatterWs spatter_ws(int severity, int n, int H, int W) {
    const size_t P = (size_t)H * W, np_ = (size_t)n * P;
    SpatterWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes); return o; };
    w.field = take(np_ * sizeof(float));            // generated N(0,1) field (perf mode)
    if (spatter_params(severity).mud) {
        w.m = take(np_ * sizeof(float));            // liquid mask
        w.dist = take(np_ * sizeof(float));         // smoothed mask
    } else {
        w.l8 = take(np_);
        w.map = take(np_);
        w.g = take(np_ * sizeof(uint16_t));
        w.dist = take(np_ * sizeof(float));
        w.q = take(np_);
        w.hist = take((size_t)n * 257 * sizeof(unsigned int));   // 256 bins + the per-image max of m
        w.f = take(np_);
        w.m = take(np_ * sizeof(float));
    }
    w.total = off;
    return w;
}

int run_spatter(const CorruptArgs& a) {
    const SpatterParams sp = spatter_params(a.severity);
    const int H = a.H, W = a.W, n = a.n;
    const int64_t plane = (int64_t)H * W;
    ADVMIX_REQUIRE(plane < (1ll << 31), "spatter: image too large");
    const SpatterWs w = spatter_ws(a.severity, n, H, W);
    char* ws = reinterpret_cast<char*>(a.ws);
    const float* field = reinterpret_cast<const float*>(a.rand_field);
    size_t fstride = a.field_bytes;
    int rc;
    if (!field) {   // perf mode: draw the field once (every value is read by (2R+1)^2 filter taps)
        float* gen = reinterpret_cast<float*>(ws + w.field);
        rc = launch_fill_rand(a, gen, nullptr);
        if (rc) return rc;
        field = gen;
    }
    int radius;
    const double* d_w = gauss_table(sp.sigma, 4.0, &radius);
    if (!d_w) return ADVMIX_ERR_CUDA;
    const LoadLiquid ld{field, fstride, W, sp.c0, sp.c1};
    if (sp.mud) {
        float* mask = reinterpret_cast<float*>(ws + w.m);
        float* sm = reinterpret_cast<float*>(ws + w.dist);
        rc = launch_gauss2d(ld, StoreLiquidMask{mask, plane, W, sp.c3}, n, H, W, 1, radius, radius, d_w, d_w, BORDER_NEAREST, a.stream);
        if (rc) return rc;
        int r2;
        const double* d_w2 = gauss_table(sp.c4, 4.0, &r2);
        if (!d_w2) return ADVMIX_ERR_CUDA;
        rc = launch_gauss2d(LoadF32Plane{mask, plane, W}, StoreMud{sm, plane, W}, n, H, W, 1, r2, r2, d_w2, d_w2, BORDER_NEAREST, a.stream);
        if (rc) return rc;
        spatter_mud_kernel<<<st_grid(plane, n), ST_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, sm, plane);
        ADVMIX_LAUNCH_OK();
        return ADVMIX_OK;
    }
    uint8_t* l8 = reinterpret_cast<uint8_t*>(ws + w.l8);
    uint8_t* map = reinterpret_cast<uint8_t*>(ws + w.map);
    uint16_t* g = reinterpret_cast<uint16_t*>(ws + w.g);
    float* dist = reinterpret_cast<float*>(ws + w.dist);
    uint8_t* q = reinterpret_cast<uint8_t*>(ws + w.q);
    unsigned int* hist = reinterpret_cast<unsigned int*>(ws + w.hist);
    unsigned int* mmax = hist + (size_t)n * 256;
    uint8_t* f = reinterpret_cast<uint8_t*>(ws + w.f);
    float* m = reinterpret_cast<float*>(ws + w.m);
    const float* d_tab = reinterpret_cast<const float*>(cached_table("chamfer_l2_5", CHAMFER_L2_5, sizeof(CHAMFER_L2_5)));
    if (!d_tab) return ADVMIX_ERR_CUDA;
    rc = launch_gauss2d(ld, StoreLiquidU8{l8, plane, W, sp.c3}, n, H, W, 1, radius, radius, d_w, d_w, BORDER_NEAREST, a.stream);
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)n * 257 * sizeof(unsigned int), a.stream));
    canny_nms_kernel<<<dim3(ceil_div(W, CN_T), ceil_div(H, CN_T), n), 256, 0, a.stream>>>(l8, map, H, W, 50, 150);
    ADVMIX_LAUNCH_OK();
    canny_hysteresis_kernel<<<n, 1024, 0, a.stream>>>(map, H, W);
    ADVMIX_LAUNCH_OK();
    row_nearest_kernel<<<st_grid(plane, n), ST_THREADS, 0, a.stream>>>(map, g, H, W);
    ADVMIX_LAUNCH_OK();
    chamfer_kernel<<<st_grid(plane, n), ST_THREADS, 0, a.stream>>>(g, d_tab, dist, H, W);
    ADVMIX_LAUNCH_OK();
    blur_f32_hist_kernel<<<st_grid(plane, n), ST_THREADS, 0, a.stream>>>(dist, q, hist, H, W);
    ADVMIX_LAUNCH_OK();
    equalize_emboss_kernel<<<st_grid(plane, n), ST_THREADS, 0, a.stream>>>(q, hist, f, H, W);
    ADVMIX_LAUNCH_OK();
    blur_u8_mask_kernel<<<st_grid(plane, n), ST_THREADS, 0, a.stream>>>(f, l8, m, mmax, H, W);
    ADVMIX_LAUNCH_OK();
    spatter_water_kernel<<<st_grid(plane, n), ST_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, m, mmax, plane, (float)sp.c4);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

size_t validation_ws_bytes(int op, int severity, int n, int H, int W) {
    if (op == C_GAUSSIAN_BLUR) {
        const int radius = (int)(4.0 * (severity == 5 ? 6 : severity) + 0.5);
        return gauss2d_smem(radius, radius, 3) <= 160 * 1024 ? 0 : (size_t)n * H * W * 3 * sizeof(double);
    }
    if (op == C_SPATTER) return spatter_ws(severity, n, H, W).total;
    return 0;
}

}  // namespace advmix
