// Shared pieces of the stencil corruptions: launch helpers, border index maps, the constant tables and the
// scipy-exact separable Gaussian (used by glass_blur, elastic_transform, gaussian_blur, spatter).
#pragma once
#include "corrupt_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

namespace advmix {

constexpr int ST_THREADS = 256;

static inline dim3 st_grid(int64_t work_per_image, int n) {
    int64_t bx = (work_per_image + ST_THREADS - 1) / ST_THREADS;
    int64_t cap = std::max<int64_t>(1, ((int64_t)sm_count() * 16 + n - 1) / n);
    return dim3((unsigned)std::min(bx, cap), (unsigned)n);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(hi, v)); }
__device__ __forceinline__ int reflect101(int i, int n) {   // cv2 BORDER_REFLECT_101
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return clampi(i, 0, n - 1);
}
__device__ __forceinline__ int reflect_sym(int i, int n) {  // scipy 'reflect' (half-sample symmetric)
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    return clampi(i, 0, n - 1);
}

#include "const_tables.inc"

// ======================================================================== separable Gaussian
// scipy.ndimage.correlate1d with symmetric weights: tmp = x[l]*w0; for j = R..1:
// tmp += (x[l-j] + x[l+j]) * w[j]   (float64, outermost pair first).
static std::vector<double> scipy_gauss_weights(double sigma, int radius) {
    std::vector<double> phi(2 * radius + 1);
    double sum = 0;
    for (int x = -radius; x <= radius; ++x) {
        phi[x + radius] = std::exp(-0.5 / (sigma * sigma) * (double)(x * x));
        sum += phi[x + radius];
    }
    std::vector<double> w(radius + 1);
    for (int j = 0; j <= radius; ++j) w[j] = phi[radius + j] / sum;
    return w;   // w[0] centre
}

constexpr int GAUSS_MAXR = 24;
enum { BORDER_NEAREST = 0, BORDER_REFLECT = 1 };

// Tiled: a CTA stages a (GT_ROWS x GT_COLS) block of the [H][W*C] plane plus its halo along the filtered
// axis in shared memory (border mode resolved at load time), then every tap is one LDS.64.
constexpr int GT_ROWS = 32, GT_COLS = 64;

template <class Load, class Store>
__global__ void __launch_bounds__(ST_THREADS)
gauss1d_kernel(Load ld, Store st, int H, int WC, int C, int axis, int radius, const double* __restrict__ wts, int border) {
    extern __shared__ double s_gt[];
    __shared__ double w[GAUSS_MAXR + 1];
    if (threadIdx.x <= radius) w[threadIdx.x] = wts[threadIdx.x];
    ld.init();
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * GT_COLS, y0 = blockIdx.y * GT_ROWS;
    const int halo = axis == 0 ? radius : radius * C;
    const int trows = axis == 0 ? GT_ROWS + 2 * radius : GT_ROWS;
    const int tcols = axis == 0 ? GT_COLS : GT_COLS + 2 * halo;
    const int Wpix = WC / C;
    __syncthreads();   // ld.init()
    for (int e = threadIdx.x; e < trows * tcols; e += ST_THREADS) {
        const int ty = e / tcols, tx = e - ty * tcols;
        double v = 0.0;
        if (axis == 0) {
            int y = y0 + ty - radius;
            y = border == BORDER_NEAREST ? clampi(y, 0, H - 1) : reflect_sym(y, H);
            const int xc = x0 + tx;
            if (xc < WC) v = ld(img, y, xc);
        } else {
            const int y = y0 + ty, xc = x0 + tx - halo;
            if (y < H) {
                int px = xc >= 0 ? xc / C : -((-xc + C - 1) / C);
                const int ch = xc - px * C;
                px = border == BORDER_NEAREST ? clampi(px, 0, Wpix - 1) : reflect_sym(px, Wpix);
                v = ld(img, y, px * C + ch);
            }
        }
        s_gt[e] = v;
    }
    __syncthreads();
    const int stride = axis == 0 ? tcols : C;
    for (int o = threadIdx.x; o < GT_ROWS * GT_COLS; o += ST_THREADS) {
        const int oy = o / GT_COLS, ox = o - oy * GT_COLS;
        const int y = y0 + oy, xc = x0 + ox;
        if (y >= H || xc >= WC) continue;
        const double* c = s_gt + (axis == 0 ? (oy + radius) * tcols + ox : oy * tcols + ox + halo);
        double tmp = c[0] * w[0];
        for (int j = radius; j >= 1; --j) tmp = tmp + (c[-j * stride] + c[j * stride]) * w[j];
        st(img, y, xc, tmp);
    }
}

// loaders / storers
struct LoadU8Div255 {       // x/255. from a uint8 HWC image (slot-indexed or dense)
    const uint8_t* base; const int32_t* idx; int64_t stride; int WC;
    double* tab;
    __device__ void init() {
        __shared__ double d255[256];
        fill_div255(d255);
        tab = d255;
    }
    __device__ double operator()(int img, int y, int xc) const {
        const int s = idx ? idx[img] : img;
        return tab[base[(int64_t)s * stride + (int64_t)y * WC + xc]];
    }
    // row-wise access for the fused kernel: 64-bit address math once per row
    typedef const uint8_t* Row;
    __device__ Row row(int img, int y) const { return base + (int64_t)(idx ? idx[img] : img) * stride + (int64_t)y * WC; }
    __device__ double at(Row r, int xc) const { return tab[r[xc]]; }
    typedef uint8_t Raw;            // load / convert split so a thread can have several loads in flight
    __device__ Raw raw(Row r, int xc) const { return __ldg(r + xc); }
    __device__ double cvt(Raw v) const { return tab[v]; }
};
struct LoadF64 {
    const double* base; int64_t stride; int WC;
    __device__ void init() {}
    __device__ double operator()(int img, int y, int xc) const { return base[(int64_t)img * stride + (int64_t)y * WC + xc]; }
};
struct StoreF64 {
    double* base; int64_t stride; int WC;
    __device__ void operator()(int img, int y, int xc, double v) const { base[(int64_t)img * stride + (int64_t)y * WC + xc] = v; }
};
struct StoreU8Trunc255 {    // np.uint8(v*255)  (optionally clipped to [0,1] first)
    uint8_t* base; const int32_t* idx; int64_t stride; int WC; int clip;
    __device__ void operator()(int img, int y, int xc, double v) const {
        const int s = idx ? idx[img] : img;
        if (clip) v = clip01(v);
        // un-clipped values can only exceed 1 by rounding; uint8 wrap like numpy
        base[(int64_t)s * stride + (int64_t)y * WC + xc] = (uint8_t)__double2int_rz(v * 255.0);
    }
};

template <class Load, class Store>
static int launch_gauss(Load ld, Store st, int n, int H, int WC, int C, int axis, int radius, const double* d_w, int border, cudaStream_t s) {
    ADVMIX_REQUIRE(n <= 65535, "gaussian filter: n<=65535 images per call");
    const int halo = axis == 0 ? radius : radius * C;
    const size_t smem = (axis == 0 ? (size_t)(GT_ROWS + 2 * radius) * GT_COLS : (size_t)GT_ROWS * (GT_COLS + 2 * halo)) * sizeof(double);
    ADVMIX_CUDA_OK(ensure_dyn_smem(gauss1d_kernel<Load, Store>, 96 * 1024));
    ADVMIX_REQUIRE(smem <= 96 * 1024, "gaussian filter: radius %d too large", radius);
    dim3 grid(ceil_div(WC, GT_COLS), ceil_div(H, GT_ROWS), n);
    gauss1d_kernel<Load, Store><<<grid, ST_THREADS, smem, s>>>(ld, st, H, WC, C, axis, radius, d_w, border);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// Fused 2-D variant (scipy filters axis 0 then axis 1): the input block with both halos is staged once,
// the axis-0 result stays in shared memory, so no float64 intermediate image touches HBM.  Border handling
// is an index map per axis, hence resolving both maps at load time reproduces scipy's two passes exactly.
// scipy keeps the array dtype between the two axis passes: a float32 image is rounded to float32 after
// the axis-0 pass.  Store functors of float32 pipelines specialise this trait.
template <class Store> struct MidF32 { static constexpr bool value = false; };

template <class Load, class Store, int TR0, int TR1>
__global__ void __launch_bounds__(ST_THREADS)
gauss2d_kernel(Load ld, Store st, int H, int WC, int C, int r0_rt, int r1_rt, const double* __restrict__ w0g,
               const double* __restrict__ w1g, int border) {
    // TR0/TR1 > 0: compile-time radii (unrolled tap loops); 0: runtime radii (any size, generic fallback)
    constexpr int M0 = TR0 > 0 ? TR0 : GAUSS_MAXR, M1 = TR1 > 0 ? TR1 : GAUSS_MAXR;
    const int R0 = TR0 > 0 ? TR0 : r0_rt, R1 = TR1 > 0 ? TR1 : r1_rt;
    extern __shared__ double s_g2[];
    __shared__ double w0[M0 + 1], w1[M1 + 1];
    __shared__ int s_ymap[GT_ROWS + 2 * M0], s_xmap[GT_COLS + 2 * M1 * 4];
    if (threadIdx.x <= R0) w0[threadIdx.x] = w0g[threadIdx.x];
    if (threadIdx.x <= R1) w1[threadIdx.x] = w1g[threadIdx.x];
    ld.init();
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * GT_COLS, y0 = blockIdx.y * GT_ROWS;
    const int halo1 = R1 * C;
    const int arows = GT_ROWS + 2 * R0, cols = GT_COLS + 2 * halo1;
    double* A = s_g2;                     // [arows][cols]  input
    double* Bm = s_g2 + arows * cols;     // [GT_ROWS][cols] after the axis-0 pass
    const int Wpix = WC / C;
    // border-resolved source row / element of every tile row / column, once per CTA
    for (int t = threadIdx.x; t < arows; t += ST_THREADS) {
        const int y = y0 + t - R0;
        s_ymap[t] = border == BORDER_NEAREST ? clampi(y, 0, H - 1) : reflect_sym(y, H);
    }
    for (int t = threadIdx.x; t < cols; t += ST_THREADS) {
        const int xc = x0 + t - halo1;
        int px = xc >= 0 ? xc / C : -((-xc + C - 1) / C);
        const int ch = xc - px * C;
        px = border == BORDER_NEAREST ? clampi(px, 0, Wpix - 1) : reflect_sym(px, Wpix);
        s_xmap[t] = px * C + ch;
    }
    __syncthreads();    // (also publishes w0 / w1)
    // 64 threads per row group, 4 row groups: no per-element divisions, 64-bit address math once per row
    const int lx = threadIdx.x & 63, ly = threadIdx.x >> 6;
    for (int ty = ly; ty < arows; ty += ST_THREADS / 64) {
        const typename Load::Row r = ld.row(img, s_ymap[ty]);
        double* a = A + ty * cols;
        for (int tx = lx; tx < cols; tx += 64) a[tx] = ld.at(r, s_xmap[tx]);
    }
    __syncthreads();
    if (TR0 > 0) {
        // register tiling along the filtered axis: a thread produces 4 consecutive rows of one column from a
        // window of 2*R0+4 values held in registers (one LDS.64 per 4 outputs and tap instead of one per tap);
        // every output still sums its taps in scipy's order
        constexpr int RQ = 4, WIN = 2 * (TR0 > 0 ? TR0 : 1) + RQ;
        for (int q = ly; q < GT_ROWS / RQ; q += ST_THREADS / 64)
            for (int tx = lx; tx < cols; tx += 64) {
                const double* c = A + (q * RQ) * cols + tx;          // window row 0 = tile row q*RQ - R0 (+R0 halo offset)
                double win[WIN];
#pragma unroll
                for (int m = 0; m < WIN; ++m) win[m] = c[m * cols];
#pragma unroll
                for (int o = 0; o < RQ; ++o) {
                    double tmp = win[o + R0] * w0[0];
#pragma unroll
                    for (int j = (TR0 > 0 ? TR0 : 1); j >= 1; --j) tmp = tmp + (win[o + R0 - j] + win[o + R0 + j]) * w0[j];
                    Bm[(q * RQ + o) * cols + tx] = MidF32<Store>::value ? (double)(float)tmp : tmp;
                }
            }
    } else {
        for (int ty = ly; ty < GT_ROWS; ty += ST_THREADS / 64)
            for (int tx = lx; tx < cols; tx += 64) {
                const double* c = A + (ty + R0) * cols + tx;
                double tmp = c[0] * w0[0];
                for (int j = R0; j >= 1; --j) tmp = tmp + (c[-j * cols] + c[j * cols]) * w0[j];
                Bm[ty * cols + tx] = MidF32<Store>::value ? (double)(float)tmp : tmp;
            }
    }
    __syncthreads();
    const int xc = x0 + lx;
    if (xc < WC)
        for (int oy = ly; oy < GT_ROWS; oy += ST_THREADS / 64) {
            const int y = y0 + oy;
            if (y >= H) break;
            const double* c = Bm + oy * cols + lx + halo1;
            double tmp = c[0] * w1[0];
#pragma unroll
            for (int j = R1; j >= 1; --j) tmp = tmp + (c[-j * C] + c[j * C]) * w1[j];
            st(img, y, xc, tmp);
        }
}

template <class Load, class Store, int R0, int R1>
static int launch_gauss2d_r(Load ld, Store st, int n, int H, int WC, int C, int r0, int r1, const double* d_w0, const double* d_w1,
                            int border, cudaStream_t s) {
    const int cols = GT_COLS + 2 * r1 * C;
    const size_t smem = ((size_t)(GT_ROWS + 2 * r0) * cols + (size_t)GT_ROWS * cols) * sizeof(double);
    ADVMIX_REQUIRE(smem <= 160 * 1024 && r0 <= GAUSS_MAXR && r1 <= GAUSS_MAXR, "gaussian filter: radii (%d,%d) too large", r0, r1);
    ADVMIX_CUDA_OK(ensure_dyn_smem(gauss2d_kernel<Load, Store, R0, R1>, 160 * 1024));
    dim3 grid(ceil_div(WC, GT_COLS), ceil_div(H, GT_ROWS), n);
    gauss2d_kernel<Load, Store, R0, R1><<<grid, ST_THREADS, smem, s>>>(ld, st, H, WC, C, r0, r1, d_w0, d_w1, border);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ---- register-tiled variant for the configured radii -------------------------------------------------------------
// 32 rows x 96 elements per CTA, 12 warps.  Axis-0 pass: lanes along x, a thread produces 4 consecutive rows of one
// column from a register window.  Axis-1 pass: lanes along ROWS (odd row pitch -> conflict-free 64-bit reads), a thread
// produces 8 outputs of one row spaced C elements apart (same channel) from a register window of 2*R1+8 values - one
// shared-memory read per 8 outputs and tap instead of two per output and tap, and 8 (4) independent accumulation chains
// per thread instead of one (the float64 adds of one output are a dependent chain in scipy's order).  Results go through
// a staging tile so the global stores keep lanes along x.  Every output sums its taps in exactly the same order as
// gauss2d_kernel / scipy.
constexpr int G2_ROWS = 32, G2_Q = 8, G2_TASKS = 12, G2_COLS = G2_Q * G2_TASKS, G2_THREADS = G2_TASKS * 32, G2_RQ = 4;

// Real = double: the reference's float64 arithmetic (bit-exact mode); Real = float: ADVMIX_CORRUPT_FAST (same tap order in
// float32, <= 1 LSB after the final truncation) - half the shared memory, FP32 pipe instead of the FP64 one.
template <class Load, class Store, int R0, int R1, int C, typename Real = double>
__global__ void __launch_bounds__(G2_THREADS)
gauss2d_rt_kernel(Load ld, Store st, int H, int WC, const double* __restrict__ w0g, const double* __restrict__ w1g, int border) {
    static_assert(G2_TASKS % C == 0, "tasks per row must split evenly over the channels");
    constexpr int HALO1 = R1 * C, COLS = G2_COLS + 2 * HALO1, AROWS = G2_ROWS + 2 * R0;
    constexpr int PB = COLS | 1, PO = G2_COLS + 1;
    extern __shared__ __align__(16) unsigned char s_g2_raw[];
    Real* s_g2 = reinterpret_cast<Real*>(s_g2_raw);
    Real* A = s_g2;                          // [AROWS][COLS] input; later the output staging tile [G2_ROWS][PO]
    Real* Bm = s_g2 + AROWS * COLS;          // [G2_ROWS][PB] after the axis-0 pass
    __shared__ Real w0[R0 + 1], w1[R1 + 1];
    __shared__ int s_ymap[AROWS], s_xmap[COLS];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    if (tid <= R0) w0[tid] = (Real)w0g[tid];
    if (tid <= R1) w1[tid] = (Real)w1g[tid];
    ld.init();
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * G2_COLS, y0 = blockIdx.y * G2_ROWS;
    const int Wpix = WC / C;
    for (int t = tid; t < AROWS; t += G2_THREADS) {
        const int y = y0 + t - R0;
        s_ymap[t] = border == BORDER_NEAREST ? clampi(y, 0, H - 1) : reflect_sym(y, H);
    }
    for (int t = tid; t < COLS; t += G2_THREADS) {
        const int xc = x0 + t - HALO1;
        int px = xc >= 0 ? xc / C : -((-xc + C - 1) / C);
        const int ch = xc - px * C;
        px = border == BORDER_NEAREST ? clampi(px, 0, Wpix - 1) : reflect_sym(px, Wpix);
        s_xmap[t] = px * C + ch;
    }
    __syncthreads();
    {
        // one warp per tile row; the lane's source elements are the same for every row, and a row's loads are all issued
        // before the first one is used (latency, not bandwidth, bounds this phase)
        constexpr int NX = (COLS + 31) / 32;
        int xm[NX];
#pragma unroll
        for (int k = 0; k < NX; ++k) xm[k] = s_xmap[min(lane + 32 * k, COLS - 1)];
        for (int ty = wrp; ty < AROWS; ty += G2_TASKS) {
            const typename Load::Row r = ld.row(img, s_ymap[ty]);
            Real* a = A + ty * COLS;
            typename Load::Raw raw[NX];
#pragma unroll
            for (int k = 0; k < NX; ++k) raw[k] = ld.raw(r, xm[k]);
#pragma unroll
            for (int k = 0; k < NX; ++k)
                if (lane + 32 * k < COLS) a[lane + 32 * k] = (Real)ld.cvt(raw[k]);
        }
    }
    __syncthreads();
    // axis 0
    for (int item = tid; item < (G2_ROWS / G2_RQ) * COLS; item += G2_THREADS) {
        const int q = item / COLS, tx = item - q * COLS;
        const Real* c = A + (q * G2_RQ) * COLS + tx;
        Real win[2 * R0 + G2_RQ], acc[G2_RQ];
#pragma unroll
        for (int m = 0; m < 2 * R0 + G2_RQ; ++m) win[m] = c[m * COLS];
        const Real wc = w0[0];
#pragma unroll
        for (int o = 0; o < G2_RQ; ++o) acc[o] = win[o + R0] * wc;
#pragma unroll
        for (int j = R0; j >= 1; --j) {
            const Real wj = w0[j];
#pragma unroll
            for (int o = 0; o < G2_RQ; ++o) acc[o] = acc[o] + (win[o + R0 - j] + win[o + R0 + j]) * wj;
        }
#pragma unroll
        for (int o = 0; o < G2_RQ; ++o) Bm[(q * G2_RQ + o) * PB + tx] = MidF32<Store>::value ? (Real)(float)acc[o] : acc[o];
    }
    __syncthreads();
    // axis 1: warp = (pixel group, channel), lane = row
    {
        const int base = (wrp / C) * (G2_Q * C) + (wrp % C);
        const Real* c = Bm + lane * PB + base;
        Real win[2 * R1 + G2_Q], acc[G2_Q];
#pragma unroll
        for (int m = 0; m < 2 * R1 + G2_Q; ++m) win[m] = c[m * C];
        const Real wc = w1[0];
#pragma unroll
        for (int o = 0; o < G2_Q; ++o) acc[o] = win[o + R1] * wc;
#pragma unroll
        for (int j = R1; j >= 1; --j) {
            const Real wj = w1[j];
#pragma unroll
            for (int o = 0; o < G2_Q; ++o) acc[o] = acc[o] + (win[o + R1 - j] + win[o + R1 + j]) * wj;
        }
        Real* o_ = A + lane * PO + base;           // A is dead: every thread passed the barrier after the axis-0 pass
#pragma unroll
        for (int o = 0; o < G2_Q; ++o) o_[o * C] = acc[o];
    }
    __syncthreads();
    for (int e = tid; e < G2_ROWS * G2_COLS; e += G2_THREADS) {
        const int row = e / G2_COLS, x = e - row * G2_COLS;
        const int y = y0 + row, xc = x0 + x;
        if (y < H && xc < WC) st(img, y, xc, A[row * PO + x]);
    }
}

template <class Load, class Store, int R0, int R1, int C, typename Real = double>
static int launch_gauss2d_rt(Load ld, Store st, int n, int H, int WC, const double* d_w0, const double* d_w1, int border, cudaStream_t s) {
    constexpr int COLS = G2_COLS + 2 * R1 * C, AROWS = G2_ROWS + 2 * R0;
    constexpr size_t smem = ((size_t)AROWS * COLS + (size_t)G2_ROWS * (COLS | 1)) * sizeof(Real);
    static_assert(smem <= 200 * 1024, "tile does not fit shared memory");
    ADVMIX_CUDA_OK(ensure_dyn_smem(gauss2d_rt_kernel<Load, Store, R0, R1, C, Real>, (int)smem));
    dim3 grid(ceil_div(WC, G2_COLS), ceil_div(H, G2_ROWS), n);
    gauss2d_rt_kernel<Load, Store, R0, R1, C, Real><<<grid, G2_THREADS, smem, s>>>(ld, st, H, WC, d_w0, d_w1, border);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// radii (and the channel count) are compile-time for the configured sizes so the tap loops unroll into register
// windows; any other size uses the runtime-radius kernel
template <class Load, class Store>
static int launch_gauss2d(Load ld, Store st, int n, int H, int WC, int C, int r0, int r1, const double* d_w0, const double* d_w1,
                          int border, cudaStream_t s) {
    ADVMIX_REQUIRE(n <= 65535 && C <= 4, "gaussian filter: n<=65535 images, C<=4");
#define G2D_CASE(a, b, c) if (r0 == a && r1 == b && C == c) return launch_gauss2d_rt<Load, Store, a, b, c>(ld, st, n, H, WC, d_w0, d_w1, border, s);
    G2D_CASE(3, 3, 3) G2D_CASE(4, 4, 3) G2D_CASE(6, 6, 3)                   // glass_blur sigmas (also gaussian_blur sigma 1)
    G2D_CASE(8, 8, 3) G2D_CASE(12, 12, 3) G2D_CASE(16, 16, 3)               // gaussian_blur sigmas 2, 3, 4
    G2D_CASE(8, 6, 1) G2D_CASE(8, 8, 1) G2D_CASE(15, 15, 1)                 // elastic_transform at 256x192, 256x256, 512x512; spatter
    G2D_CASE(4, 4, 1) G2D_CASE(6, 6, 1)                                     // spatter
#undef G2D_CASE
    return launch_gauss2d_r<Load, Store, 0, 0>(ld, st, n, H, WC, C, r0, r1, d_w0, d_w1, border, s);   // any other size
}

// ADVMIX_CORRUPT_FAST: the same register-tiled kernel in float32 for the configured radii; returns -1 when (r0, r1, C) has no
// compiled variant (the caller then takes the float64 path).
template <class Load, class Store>
static int launch_gauss2d_fast(Load ld, Store st, int n, int H, int WC, int C, int r0, int r1, const double* d_w0, const double* d_w1,
                               int border, cudaStream_t s) {
    if (n > 65535) return -1;
#define G2D_CASE(a, b, c) if (r0 == a && r1 == b && C == c) return launch_gauss2d_rt<Load, Store, a, b, c, float>(ld, st, n, H, WC, d_w0, d_w1, border, s);
    G2D_CASE(3, 3, 3) G2D_CASE(4, 4, 3) G2D_CASE(6, 6, 3)
    G2D_CASE(8, 8, 3) G2D_CASE(12, 12, 3) G2D_CASE(16, 16, 3) G2D_CASE(24, 24, 3)
    G2D_CASE(8, 6, 1) G2D_CASE(8, 8, 1) G2D_CASE(15, 15, 1)
#undef G2D_CASE
    return -1;
}

// The exact path's response to a constant-1.0 input (scipy's tap order, float64): decides whether a saturated region
// truncates to 255 or 254, which the float32 path reproduces by snapping values within 2e-5 of 1 to `top`.
static inline double gauss1d_unit_response(const double* w, int radius, double x) {
    double t = x * w[0];
    for (int j = radius; j >= 1; --j) t = t + (x + x) * w[j];
    return t;
}
static inline float fast_top(double unit_response) { return unit_response >= 1.0 ? 1.0f : 0.99999994f; }

struct LoadU8Div255F {      // float32 twin of LoadU8Div255
    const uint8_t* base; const int32_t* idx; int64_t stride; int WC;
    float* tab;
    __device__ void init() {
        __shared__ float f255[256];
        for (int i = threadIdx.x; i < 256; i += blockDim.x) f255[i] = (float)__ddiv_rn((double)i, 255.0);
        tab = f255;
    }
    typedef const uint8_t* Row;
    __device__ Row row(int img, int y) const { return base + (int64_t)(idx ? idx[img] : img) * stride + (int64_t)y * WC; }
    typedef uint8_t Raw;
    __device__ Raw raw(Row r, int xc) const { return __ldg(r + xc); }
    __device__ float cvt(Raw v) const { return tab[v]; }
};
struct StoreU8Trunc255F {   // float32 twin of StoreU8Trunc255 with the saturation snap
    uint8_t* base; const int32_t* idx; int64_t stride; int WC; int clip; float top;
    __device__ void operator()(int img, int y, int xc, float v) const {
        const int s = idx ? idx[img] : img;
        if (v > 0.999999f) v = top;         // saturated region (float32 accumulation error ~1e-7): what the float64 tap order gives there
        v = fmaxf(v, 0.f);
        base[(int64_t)s * stride + (int64_t)y * WC + xc] = (uint8_t)__float2int_rz(__fmul_rn(v, 255.0f));
    }
};

static const double* gauss_table(double sigma, double truncate, int* radius, std::vector<double>* host_w = nullptr) {
    *radius = (int)(truncate * sigma + 0.5);
    std::vector<double> w;
    for (int i = 0; i < GAUSS_NTABS; ++i)   // bit-exact scipy weights for the sigmas the configs use
        if (GAUSS_TABS[i].sigma == sigma && GAUSS_TABS[i].truncate == truncate) {
            *radius = GAUSS_TABS[i].radius;
            w.assign(GAUSS_TABS[i].w, GAUSS_TABS[i].w + GAUSS_TABS[i].radius + 1);
        }
    if (w.empty()) w = scipy_gauss_weights(sigma, *radius);   // other sizes: libm exp, <= 1 ulp off numpy
    if (host_w) *host_w = w;
    char key[96];
    snprintf(key, sizeof(key), "gauss_%.17g_%d", sigma, *radius);
    return reinterpret_cast<const double*>(cached_table(key, w.data(), w.size() * sizeof(double)));
}

}  // namespace advmix
