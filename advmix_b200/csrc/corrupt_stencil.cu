// Stencil / gather corruptions: defocus_blur, glass_blur, motion_blur, zoom_blur, snow, fog,
// elastic_transform.  Arithmetic follows the scipy / cv2 / skimage primitives the package
// calls operation by operation, so results are bit-exact wherever the
// reference's own summation order is reproducible and within 1 LSB elsewhere.
#include "stencil_common.cuh"

#include <climits>

namespace advmix {

// ======================================================================== defocus_blur
// cv2.filter2D(x/255 (float64), -1, disk kernel (float32)), BORDER_REFLECT_101.
struct TapList {
    std::vector<int8_t> off;     // dy, dx pairs
    std::vector<double> w;
};


// The five kernels are constant data of the op: imagecorruptions' disk() output, taken
// bit-for-bit from cv2.GaussianBlur (oracle/gen_disk_kernels.py) because cv2's separable
// float filter accumulates with platform-dependent FMA and saturated regions are sensitive
// to the last ulp of sum(kernel).
static TapList disk_taps(int severity) {
    TapList t;
    const DiskTap* taps = DISK_TAPS[severity - 1];
    for (int i = 0; i < DISK_NTAPS[severity - 1]; ++i) {
        t.off.push_back(taps[i].dy);
        t.off.push_back(taps[i].dx);
        t.w.push_back((double)taps[i].w);
    }
    return t;
}

// Register-tiled: a CTA computes 64x32 outputs, each thread a 4 (horizontally adjacent) x 2 (vertically adjacent) block.
// The source tile (+halo, reflect-101 resolved at load time) sits in shared memory as x/255 float64, planar per channel;
// the kernel is a dense (2h+1) x pad4(2h+1) float64 weight array (zeros where the disk has no tap - adding 0*v leaves
// the sum bit-identical) with one extra all-zero row above and below.  Per tile row a thread slides a 4-value window
// along x and feeds it to BOTH of its output rows (kernel row r for the upper one, r-1 for the lower one), so one new
// LDS.64 serves 8 DFMA and the weights come as uniform LDS.128: the loop is bound by the FP64 pipe, not by shared-memory
// bandwidth (the 4x1 version needed 12 wavefronts per 16 DFMA).  Every output still sums its taps in row-major kernel
// order.
constexpr int DEF_BW = 64;

template <int DEF_BH>                                         // 32 (256 threads) or 16 (128 threads: twice the CTAs for small batches)
__global__ void __launch_bounds__(DEF_BH * 8)
defocus_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
               int H, int W, const double* __restrict__ wdense, int h, int ncols) {
    extern __shared__ __align__(16) double smem_d[];
    const int nrows = 2 * h + 1;
    const int th = DEF_BH + 2 * h, tw = DEF_BW + ncols;       // >= DEF_BW + 2h + 3
    double* d255 = smem_d;                                    // 256
    double* wts = d255 + 256;                                 // (nrows + 2) * ncols, rows 0 and nrows + 1 are zero
    double* tile = wts + (nrows + 2) * ncols;                 // 3 * th * tw
    fill_div255(d255);
    for (int i = threadIdx.x; i < (nrows + 2) * ncols; i += (DEF_BH * 8)) wts[i] = wdense[i];
    __syncthreads();
    const int slot = slot_of(idx, blockIdx.z);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    const int x0 = blockIdx.x * DEF_BW, y0 = blockIdx.y * DEF_BH;
    // tile layout: column c of a row lives at sub-plane (c & 3), index (c >> 2): the 16 threads of a row
    // (4 adjacent pixels each) then read 16 consecutive doubles per LDS.64 - no bank conflicts
    const int wq = tw >> 2;                                   // tw is a multiple of 4
    for (int e = threadIdx.x; e < th * tw; e += (DEF_BH * 8)) {
        const int ty = e / tw, tx = e - ty * tw;
        const int gy = reflect101(y0 + ty - h, H), gx = reflect101(x0 + tx - h, W);
        const uint8_t* p = src + ((int64_t)gy * W + gx) * 3;
        const int o = (ty * 4 + (tx & 3)) * wq + (tx >> 2);
        tile[o] = d255[p[0]];
        tile[th * tw + o] = d255[p[1]];
        tile[2 * th * tw + o] = d255[p[2]];
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = 2 * (threadIdx.x >> 4);
    const int X = 4 * tx;
    const int x = x0 + X, y = y0 + ty;
    uint8_t res[2][4][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;        // upper output row
        double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;        // lower output row
        for (int r = 0; r <= nrows; ++r) {                    // tile row ty + r: kernel row r (upper), r - 1 (lower)
            const double* p0 = tile + ((c * th + ty + r) * 4) * wq + tx;   // sub-plane 0 of this row, at this thread's column group
            const double *p1 = p0 + wq, *p2 = p1 + wq, *p3 = p2 + wq;
            const double2* wa = reinterpret_cast<const double2*>(wts + (r + 1) * ncols);
            const double2* wb = reinterpret_cast<const double2*>(wts + r * ncols);
            // (visiting only the column groups that hold a tap was measured: the dynamic loop bounds cost what the skipped
            // zero products save)
            double v0 = p0[0], v1 = p1[0], v2 = p2[0];
            for (int d = 0, g = 0; d < ncols; d += 4, ++g) {
                const double2 wa01 = wa[2 * g], wa23 = wa[2 * g + 1], wb01 = wb[2 * g], wb23 = wb[2 * g + 1];
                const double v3 = p3[g];
                a0 = fma(wa01.x, v0, a0); a1 = fma(wa01.x, v1, a1); a2 = fma(wa01.x, v2, a2); a3 = fma(wa01.x, v3, a3);
                b0 = fma(wb01.x, v0, b0); b1 = fma(wb01.x, v1, b1); b2 = fma(wb01.x, v2, b2); b3 = fma(wb01.x, v3, b3);
                const double v4 = p0[g + 1];
                a0 = fma(wa01.y, v1, a0); a1 = fma(wa01.y, v2, a1); a2 = fma(wa01.y, v3, a2); a3 = fma(wa01.y, v4, a3);
                b0 = fma(wb01.y, v1, b0); b1 = fma(wb01.y, v2, b1); b2 = fma(wb01.y, v3, b2); b3 = fma(wb01.y, v4, b3);
                const double v5 = p1[g + 1];
                a0 = fma(wa23.x, v2, a0); a1 = fma(wa23.x, v3, a1); a2 = fma(wa23.x, v4, a2); a3 = fma(wa23.x, v5, a3);
                b0 = fma(wb23.x, v2, b0); b1 = fma(wb23.x, v3, b1); b2 = fma(wb23.x, v4, b2); b3 = fma(wb23.x, v5, b3);
                const double v6 = p2[g + 1];
                a0 = fma(wa23.y, v3, a0); a1 = fma(wa23.y, v4, a1); a2 = fma(wa23.y, v5, a2); a3 = fma(wa23.y, v6, a3);
                b0 = fma(wb23.y, v3, b0); b1 = fma(wb23.y, v4, b1); b2 = fma(wb23.y, v5, b2); b3 = fma(wb23.y, v6, b3);
                v0 = v4; v1 = v5; v2 = v6;
            }
        }
        res[0][0][c] = trunc_u8(clip01(a0) * 255.0);
        res[0][1][c] = trunc_u8(clip01(a1) * 255.0);
        res[0][2][c] = trunc_u8(clip01(a2) * 255.0);
        res[0][3][c] = trunc_u8(clip01(a3) * 255.0);
        res[1][0][c] = trunc_u8(clip01(b0) * 255.0);
        res[1][1][c] = trunc_u8(clip01(b1) * 255.0);
        res[1][2][c] = trunc_u8(clip01(b2) * 255.0);
        res[1][3][c] = trunc_u8(clip01(b3) * 255.0);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (y + j < H) {
            uint8_t* o = out + (int64_t)slot * H * W * 3 + ((int64_t)(y + j) * W + x) * 3;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x + i < W) { o[3 * i] = res[j][i][0]; o[3 * i + 1] = res[j][i][1]; o[3 * i + 2] = res[j][i][2]; }
        }
}

int run_defocus_blur(const CorruptArgs& a) {
    if (a.fast) { const int frc = run_defocus_blur_fast(a); if (frc != -1) return frc; }      // ADVMIX_CORRUPT_FAST
    TapList t = disk_taps(a.severity);
    int h = 0;
    for (auto v : t.off) h = std::max(h, (int)std::abs((int)v));
    const int nrows = 2 * h + 1, ncols = (nrows + 3) / 4 * 4;
    std::vector<double> dense((size_t)(nrows + 2) * ncols, 0.0);      // zero row above and below (see the kernel)
    for (size_t k = 0; k < t.w.size(); ++k) dense[(size_t)(t.off[2 * k] + h + 1) * ncols + (t.off[2 * k + 1] + h)] = t.w[k];
    const double* d_w = reinterpret_cast<const double*>(cached_table("diskdense2_" + std::to_string(a.severity), dense.data(), dense.size() * sizeof(double)));
    if (!d_w) return ADVMIX_ERR_CUDA;
    ADVMIX_REQUIRE(a.n <= 65535, "defocus: n<=65535 per call");
    // a 64x32 tile per CTA when the batch fills the GPU, 64x16 (twice the CTAs, half the work each) for small batches
    const bool small = (int64_t)a.n * ceil_div(a.W, DEF_BW) * ceil_div(a.H, 32) < 2 * sm_count();
    const int BH = small ? 16 : 32;
    const size_t smem = (256 + (size_t)(nrows + 2) * ncols + (size_t)3 * (BH + 2 * h) * (DEF_BW + ncols)) * sizeof(double);
    ADVMIX_REQUIRE(smem <= 160 * 1024, "defocus: kernel too large");
    ADVMIX_CUDA_OK(ensure_dyn_smem(defocus_kernel<32>, 160 * 1024));
    ADVMIX_CUDA_OK(ensure_dyn_smem(defocus_kernel<16>, 160 * 1024));
    dim3 grid(ceil_div(a.W, DEF_BW), ceil_div(a.H, BH), a.n);
    if (small) defocus_kernel<16><<<grid, 128, smem, a.stream>>>(a.in, a.out, a.idx, a.H, a.W, d_w, h, ncols);
    else defocus_kernel<32><<<grid, 256, smem, a.stream>>>(a.in, a.out, a.idx, a.H, a.W, d_w, h, ncols);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== motion blur taps
// _motion_blur(): blurred = sum_i k_i * shift(x, dx_i, dy_i) in float64, i ascending;
// shift = roll + replicate fill = clamp-to-edge sampling at (y - dy, x - dx).
constexpr int MOTION_MAXW = 41;

// kernel weights: bit-for-bit numpy values (const_tables.inc); constant regions make the
// truncated output depend on the last ulp of sum(k).
static std::vector<double> table_weights(const double* k, int radius) {
    return std::vector<double>(k, k + 2 * radius + 1);
}

// fills s_dy/s_dx[0..ntaps) for this image; returns ntaps (the package `break`s when an
// offset leaves the image).  Call from all threads, followed by __syncthreads().
__device__ __forceinline__ void motion_offsets(int width, double angle_deg, int H, int W, int* s_dy, int* s_dx, int* s_n) {
    if (threadIdx.x < width) {
        const int i = threadIdx.x;
        const double rad = angle_deg * (3.141592653589793 / 180.0);      // np.deg2rad
        const double p0 = (double)width * sin(rad), p1 = (double)width * cos(rad);
        const double hyp = hypot(p0, p1);
        s_dy[i] = -(int)ceil(((double)i * p0) / hyp - 0.5);
        s_dx[i] = -(int)ceil(((double)i * p1) / hyp - 0.5);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = width;
        for (int i = 0; i < width; ++i)
            if (abs(s_dy[i]) >= H || abs(s_dx[i]) >= W) { n = i; break; }
        *s_n = n;
    }
    __syncthreads();
}

struct MotionTap { int dy, dx; double k; };     // one 16-byte shared-memory read per tap (uniform over the warp)
struct MotionTapOff { int off, pad; double k; };   // interior pixels: byte offset of the tap relative to the pixel

__global__ void __launch_bounds__(ST_THREADS)
motion_blur_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                   const double* __restrict__ param, uint64_t seed, int64_t sample_base, int H, int W,
                   const double* __restrict__ kw, int width) {
    __shared__ int s_dy[MOTION_MAXW], s_dx[MOTION_MAXW], s_n;
    __shared__ __align__(16) MotionTap s_tap[MOTION_MAXW];
    __shared__ double s_d[256];                 // (double)byte without a conversion instruction per tap and channel
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const double angle = param_uniform(param ? param + 4 * i : nullptr, rng, -45.0, 45.0);
    for (int t = threadIdx.x; t < 256; t += ST_THREADS) s_d[t] = (double)t;
    motion_offsets(width, angle, H, W, s_dy, s_dx, &s_n);
    const int ntaps = s_n;
    if (threadIdx.x < ntaps) s_tap[threadIdx.x] = MotionTap{s_dy[threadIdx.x], s_dx[threadIdx.x], kw[threadIdx.x]};
    __syncthreads();
    // taps never leave [y - my1, y - my0] x [x - mx1, x - mx0]: pixels whose whole line is inside the image skip the clamps
    int my0 = 0, my1 = 0, mx0 = 0, mx1 = 0;
    for (int t = 0; t < ntaps; ++t) {
        my0 = min(my0, s_dy[t]); my1 = max(my1, s_dy[t]);
        mx0 = min(mx0, s_dx[t]); mx1 = max(mx1, s_dx[t]);
    }
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = out + (int64_t)slot * H * W * 3;
    const int npix = H * W, W3 = 3 * W;
    for (int p = blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += gridDim.x * ST_THREADS) {
        const int y = p / W, x = p - y * W;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        if (y - my1 >= 0 && y - my0 < H && x - mx1 >= 0 && x - mx0 < W) {
            const uint8_t* c = src + (int64_t)p * 3;
#pragma unroll 4
            for (int t = 0; t < ntaps; ++t) {
                const MotionTap T = s_tap[t];
                const uint8_t* q = c - (T.dy * W3 + T.dx * 3);
                a0 = a0 + T.k * s_d[__ldg(q)];
                a1 = a1 + T.k * s_d[__ldg(q + 1)];
                a2 = a2 + T.k * s_d[__ldg(q + 2)];
            }
        } else {
            for (int t = 0; t < ntaps; ++t) {
                const MotionTap T = s_tap[t];
                const int yy = clampi(y - T.dy, 0, H - 1), xx = clampi(x - T.dx, 0, W - 1);
                const uint8_t* q = src + (yy * W + xx) * 3;
                a0 = a0 + T.k * s_d[__ldg(q)];
                a1 = a1 + T.k * s_d[__ldg(q + 1)];
                a2 = a2 + T.k * s_d[__ldg(q + 2)];
            }
        }
        uint8_t* o = dst + (int64_t)p * 3;
        o[0] = trunc_u8(fmin(fmax(a0, 0.0), 255.0));
        o[1] = trunc_u8(fmin(fmax(a1, 0.0), 255.0));
        o[2] = trunc_u8(fmin(fmax(a2, 0.0), 255.0));
    }
}

// Image-resident variant (same idea as zoom_blur_smem_kernel): the taps are shared-memory byte reads; bytes become
// float64 by building 2^52 + b in the mantissa and subtracting 2^52 (exact, no conversion instruction).
constexpr int MS_THREADS = 1024;
__device__ __forceinline__ double byte_to_f64(uint32_t b) { return __hiloint2double(0x43300000, (int)b) - 4503599627370496.0; }

__global__ void __launch_bounds__(MS_THREADS, 1)
motion_blur_smem_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                        const double* __restrict__ param, uint64_t seed, int64_t sample_base, int n, int H, int W,
                        const double* __restrict__ kw, int width) {
    extern __shared__ __align__(16) uint8_t s_img[];
    __shared__ int s_dy[MOTION_MAXW], s_dx[MOTION_MAXW], s_n;
    __shared__ __align__(16) MotionTap s_tap[MOTION_MAXW];
    __shared__ __align__(16) MotionTapOff s_off[MOTION_MAXW];
    const int nbytes = H * W * 3, npix = H * W, W3 = 3 * W;
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const SampleRng rng(seed, sample_base + slot);
        const double angle = param_uniform(param ? param + 4 * img : nullptr, rng, -45.0, 45.0);
        const uint8_t* src = in + (int64_t)slot * nbytes;
        uint8_t* dst = out + (int64_t)slot * nbytes;
        __syncthreads();                                    // previous image fully consumed
        if ((nbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(s_img);
            for (int i = threadIdx.x; i < nbytes / 16; i += MS_THREADS) d4[i] = ld_stream_u4(s4 + i);
        } else {
            for (int i = threadIdx.x; i < nbytes; i += MS_THREADS) s_img[i] = src[i];
        }
        motion_offsets(width, angle, H, W, s_dy, s_dx, &s_n);
        const int ntaps = s_n;
        if (threadIdx.x < ntaps) {
            s_tap[threadIdx.x] = MotionTap{s_dy[threadIdx.x], s_dx[threadIdx.x], kw[threadIdx.x]};
            s_off[threadIdx.x] = MotionTapOff{s_dy[threadIdx.x] * W3 + s_dx[threadIdx.x] * 3, 0, kw[threadIdx.x]};
        }
        __syncthreads();
        int my0 = 0, my1 = 0, mx0 = 0, mx1 = 0;
        for (int t = 0; t < ntaps; ++t) {
            my0 = min(my0, s_dy[t]); my1 = max(my1, s_dy[t]);
            mx0 = min(mx0, s_dx[t]); mx1 = max(mx1, s_dx[t]);
        }
        for (int p = threadIdx.x; p < npix; p += MS_THREADS) {
            const int y = p / W, x = p - y * W;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            if (y - my1 >= 0 && y - my0 < H && x - mx1 >= 0 && x - mx0 < W) {
                const uint8_t* c = s_img + p * 3;
#pragma unroll 4
                for (int t = 0; t < ntaps; ++t) {
                    const MotionTapOff T = s_off[t];
                    const uint8_t* q = c - T.off;
                    a0 = a0 + T.k * byte_to_f64(q[0]);
                    a1 = a1 + T.k * byte_to_f64(q[1]);
                    a2 = a2 + T.k * byte_to_f64(q[2]);
                }
            } else {
                for (int t = 0; t < ntaps; ++t) {
                    const MotionTap T = s_tap[t];
                    const int yy = clampi(y - T.dy, 0, H - 1), xx = clampi(x - T.dx, 0, W - 1);
                    const uint8_t* q = s_img + (yy * W + xx) * 3;
                    a0 = a0 + T.k * byte_to_f64(q[0]);
                    a1 = a1 + T.k * byte_to_f64(q[1]);
                    a2 = a2 + T.k * byte_to_f64(q[2]);
                }
            }
            uint8_t* o = dst + p * 3;
            o[0] = trunc_u8(fmin(fmax(a0, 0.0), 255.0));
            o[1] = trunc_u8(fmin(fmax(a1, 0.0), 255.0));
            o[2] = trunc_u8(fmin(fmax(a2, 0.0), 255.0));
        }
    }
}

int run_motion_blur(const CorruptArgs& a) {
    if (a.fast) { const int frc = run_motion_blur_fast(a); if (frc != -1) return frc; }      // ADVMIX_CORRUPT_FAST
    const int r = MOTION_RADIUS[a.severity - 1];
    std::vector<double> k = table_weights(MOTION_K[a.severity - 1], r);
    const double* d_k = reinterpret_cast<const double*>(cached_table("motion_" + std::to_string(a.severity), k.data(), k.size() * sizeof(double)));
    if (!d_k) return ADVMIX_ERR_CUDA;
    const size_t img_bytes = (size_t)a.H * a.W * 3;
    if (img_bytes <= 200 * 1024 && a.n >= sm_count() / 2) {       // enough images to fill the GPU with one CTA each
        ADVMIX_CUDA_OK(ensure_dyn_smem(motion_blur_smem_kernel, 200 * 1024));
        motion_blur_smem_kernel<<<std::min(a.n, sm_count()), MS_THREADS, (img_bytes + 15) & ~(size_t)15, a.stream>>>(
            a.in, a.out, a.idx, a.rand_param, a.seed, a.sample_base, a.n, a.H, a.W, d_k, 2 * r + 1);
    } else {
        motion_blur_kernel<<<st_grid((int64_t)a.H * a.W, a.n), ST_THREADS, 0, a.stream>>>(
            a.in, a.out, a.idx, a.rand_param, a.seed, a.sample_base, a.H, a.W, d_k, 2 * r + 1);
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== zoom_blur
// scipy.ndimage.zoom(order=1, mode='constant') on the centre crop, per zoom factor.
struct ZoomLayer {
    int top0, in0, out0, top1, in1, out1;
    double z0, z1;   // (in-1)/(out-1)
};

static int py_round(double v) { return (int)std::nearbyint(v); }   // round-half-even like Python

static ZoomLayer zoom_layer(int H, int W, double zf) {
    ZoomLayer L;
    L.in0 = (int)std::ceil(H / zf);
    L.top0 = (H - L.in0) / 2;
    L.in1 = (int)std::ceil(W / zf);
    L.top1 = (W - L.in1) / 2;
    L.out0 = py_round(L.in0 * zf);
    L.out1 = py_round(L.in1 * zf);
    L.z0 = L.out0 > 1 ? (double)(L.in0 - 1) / (double)(L.out0 - 1) : 1.0;
    L.z1 = L.out1 > 1 ? (double)(L.in1 - 1) / (double)(L.out1 - 1) : 1.0;
    return L;
}

static std::vector<double> zoom_factors(int severity) {
    // np.arange(start, stop, step): length ceil((stop-start)/step), values start + i*step
    const double stop[5] = {1.11, 1.16, 1.21, 1.26, 1.33}, step[5] = {0.01, 0.01, 0.02, 0.02, 0.03};
    const int n = (int)std::ceil((stop[severity - 1] - 1.0) / step[severity - 1]);
    std::vector<double> f(n);
    for (int i = 0; i < n; ++i) f[i] = 1.0 + i * step[severity - 1];
    return f;
}

// one interpolated sample along scipy's order-1 path; returns false if the coordinate falls
// outside [0, in-1] ('constant' mode -> cval 0).  Host side: the coordinates depend only on (layer, row) and
// (layer, column), so they are tabulated once per (H, W, severity) instead of being recomputed per pixel.
__host__ __device__ __forceinline__ bool zoom_coord(int o, double z, int in, int* s, double* t) {
    const double cc = (double)o * z;
    if (cc < 0.0 || cc > (double)(in - 1)) return false;
    const double f = floor(cc);
    *s = (int)f;
    *t = cc - f;
    return true;
}

constexpr int ZOOM_MAXL = 16;
struct ZoomTap { int o0, o1; double t; };     // byte offsets of the two source rows (columns); o0 < 0: outside -> layer value 0

__device__ __forceinline__ ZoomTap zoom_tap(const ZoomTap* p) {      // one 16-byte load
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    return ZoomTap{v.x, v.y, __hiloint2double(v.w, v.z)};
}

// One thread per pixel, all layers.  Per layer: one row entry (uniform over a warp), one column entry (coalesced),
// 12 source bytes, and per channel the reference's float64 sum ((v*wy)*wx, four terms in order) rounded to float32;
// v comes from a float64 table of float32(i/255) (no conversion instructions in the loop).
__global__ void __launch_bounds__(ST_THREADS)
zoom_blur_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                 int H, int W, const ZoomTap* __restrict__ taps, int nl) {
    __shared__ double lut[256];
    for (int i = threadIdx.x; i < 256; i += ST_THREADS) lut[i] = (double)(float)__ddiv_rn((double)i, 255.0);
    __syncthreads();
    const int slot = slot_of(idx, blockIdx.y);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = out + (int64_t)slot * H * W * 3;
    const int npix = H * W;
    const float denom = (float)(nl + 1);
    for (int p = blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += gridDim.x * ST_THREADS) {
        const int y = p / W, x = p - y * W;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
        const ZoomTap* tr = taps + y;
        const ZoomTap* tc = taps + H + x;
#pragma unroll 2
        for (int l = 0; l < nl; ++l, tr += H + W, tc += H + W) {
            const ZoomTap R = zoom_tap(tr);
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (R.o0 >= 0) {
                const ZoomTap Cc = zoom_tap(tc);
                if (Cc.o0 >= 0) {
                    const double ty = R.t, tx = Cc.t, wy0 = 1.0 - ty, wx0 = 1.0 - tx;
                    const uint8_t* p00 = src + R.o0 + Cc.o0;
                    const uint8_t* p01 = src + R.o0 + Cc.o1;
                    const uint8_t* p10 = src + R.o1 + Cc.o0;
                    const uint8_t* p11 = src + R.o1 + Cc.o1;
                    double t;
#define ZB_CH(c, dstv)                                      \
    t = (lut[__ldg(p00 + c)] * wy0) * wx0;                  \
    t = t + (lut[__ldg(p01 + c)] * wy0) * tx;               \
    t = t + (lut[__ldg(p10 + c)] * ty) * wx0;               \
    t = t + (lut[__ldg(p11 + c)] * ty) * tx;                \
    dstv = (float)t;
                    ZB_CH(0, v0)
                    ZB_CH(1, v1)
                    ZB_CH(2, v2)
#undef ZB_CH
                }
            }
            acc0 = __fadd_rn(acc0, v0);
            acc1 = __fadd_rn(acc1, v1);
            acc2 = __fadd_rn(acc2, v2);
        }
        const uint8_t* q = src + (int64_t)p * 3;
        const float r0 = __fdiv_rn(__fadd_rn((float)lut[q[0]], acc0), denom);
        const float r1 = __fdiv_rn(__fadd_rn((float)lut[q[1]], acc1), denom);
        const float r2 = __fdiv_rn(__fadd_rn((float)lut[q[2]], acc2), denom);
        uint8_t* o = dst + (int64_t)p * 3;
        o[0] = (uint8_t)(int)__fmul_rn(fminf(fmaxf(r0, 0.f), 1.f), 255.f);
        o[1] = (uint8_t)(int)__fmul_rn(fminf(fmaxf(r1, 0.f), 1.f), 255.f);
        o[2] = (uint8_t)(int)__fmul_rn(fminf(fmaxf(r2, 0.f), 1.f), 255.f);
    }
}

// Image-resident variant: one CTA per image keeps the whole uint8 image in shared memory (147 KB at 256x192), so the
// 12 source bytes per pixel and layer are shared-memory reads instead of L1 requests (the global-memory version is bound
// by its byte loads).  A thread owns one column of 4 rows: the column entry of a layer is loaded once per 4 pixels, the
// row entries are uniform over the warp.  Arithmetic identical to zoom_blur_kernel.
constexpr int ZS_THREADS = 1024, ZS_ROWS = 4;
__global__ void __launch_bounds__(ZS_THREADS, 1)
zoom_blur_smem_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                      int n, int H, int W, const ZoomTap* __restrict__ taps, int nl) {
    extern __shared__ __align__(16) uint8_t s_img[];
    // float64 table of float32(i/255), 16 copies interleaved ([value][lane & 15]): the 16 lanes of a 64-bit shared-memory
    // wavefront read 16 different banks whatever their byte values are (a single table costs ~2x in bank conflicts)
    double* lut = reinterpret_cast<double*>(s_img + (((size_t)H * W * 3 + 15) & ~(size_t)15)) + (threadIdx.x & 15);
    for (int i = threadIdx.x; i < 256 * 16; i += ZS_THREADS) lut[i - (threadIdx.x & 15)] = (double)(float)__ddiv_rn((double)(i >> 4), 255.0);
    const int nbytes = H * W * 3;
    const float denom = (float)(nl + 1);
    const int hq = (H + ZS_ROWS - 1) / ZS_ROWS;
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const uint8_t* src = in + (int64_t)slot * nbytes;
        uint8_t* dst = out + (int64_t)slot * nbytes;
        __syncthreads();                                    // previous image fully consumed (and lut written)
        if ((nbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(s_img);
            for (int i = threadIdx.x; i < nbytes / 16; i += ZS_THREADS) d4[i] = ld_stream_u4(s4 + i);
        } else {
            for (int i = threadIdx.x; i < nbytes; i += ZS_THREADS) s_img[i] = src[i];
        }
        __syncthreads();
        for (int item = threadIdx.x; item < hq * W; item += ZS_THREADS) {
            const int yq = item / W, x = item - yq * W;
            const int y0 = yq * ZS_ROWS;
            float acc[ZS_ROWS][3];
#pragma unroll
            for (int i = 0; i < ZS_ROWS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
            const ZoomTap* tl = taps;
            for (int l = 0; l < nl; ++l, tl += H + W) {
                const ZoomTap Cc = zoom_tap(tl + H + x);
                const double tx = Cc.t, wx0 = 1.0 - tx;
#pragma unroll
                for (int i = 0; i < ZS_ROWS; ++i) {
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                    if (y0 + i < H) {
                        const ZoomTap R = zoom_tap(tl + y0 + i);
                        if (R.o0 >= 0 && Cc.o0 >= 0) {
                            const double ty = R.t, wy0 = 1.0 - ty;
                            const uint8_t* p00 = s_img + R.o0 + Cc.o0;
                            const uint8_t* p01 = s_img + R.o0 + Cc.o1;
                            const uint8_t* p10 = s_img + R.o1 + Cc.o0;
                            const uint8_t* p11 = s_img + R.o1 + Cc.o1;
                            double t;
#define ZB_CH(c, dstv)                                   \
    t = (lut[16 * p00[c]] * wy0) * wx0;                  \
    t = t + (lut[16 * p01[c]] * wy0) * tx;               \
    t = t + (lut[16 * p10[c]] * ty) * wx0;               \
    t = t + (lut[16 * p11[c]] * ty) * tx;                \
    dstv = (float)t;
                            ZB_CH(0, v0)
                            ZB_CH(1, v1)
                            ZB_CH(2, v2)
#undef ZB_CH
                        }
                    }
                    acc[i][0] = __fadd_rn(acc[i][0], v0);
                    acc[i][1] = __fadd_rn(acc[i][1], v1);
                    acc[i][2] = __fadd_rn(acc[i][2], v2);
                }
            }
#pragma unroll
            for (int i = 0; i < ZS_ROWS; ++i) {
                if (y0 + i >= H) break;
                const int pb = ((y0 + i) * W + x) * 3;
                uint8_t* o = dst + pb;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float r = __fdiv_rn(__fadd_rn((float)lut[16 * s_img[pb + c]], acc[i][c]), denom);
                    o[c] = (uint8_t)(int)__fmul_rn(fminf(fmaxf(r, 0.f), 1.f), 255.f);
                }
            }
        }
    }
}

int run_zoom_blur(const CorruptArgs& a) {
    if (a.fast) { const int frc = run_zoom_blur_fast(a); if (frc != -1) return frc; }      // ADVMIX_CORRUPT_FAST
    std::vector<double> f = zoom_factors(a.severity);
    ADVMIX_REQUIRE((int)f.size() <= ZOOM_MAXL, "zoom_blur: too many layers");
    ADVMIX_REQUIRE((int64_t)a.H * a.W * 3 < INT_MAX, "zoom_blur: image too large");
    const int H = a.H, W = a.W, nl = (int)f.size();
    const std::string key = "zoomtaps_" + std::to_string(H) + "x" + std::to_string(W) + "_" + std::to_string(a.severity);
    // layout: [layer][H row entries, W column entries]
    std::vector<ZoomTap> T((size_t)nl * (H + W));
    for (int l = 0; l < nl; ++l) {
        const ZoomLayer z = zoom_layer(H, W, f[l]);
        ZoomTap* tr = T.data() + (size_t)l * (H + W);
        for (int y = 0; y < H; ++y) {
            int s; double t;
            if (y < z.out0 && zoom_coord(y, z.z0, z.in0, &s, &t))
                tr[y] = ZoomTap{(z.top0 + s) * W * 3, (z.top0 + std::min(s + 1, z.in0 - 1)) * W * 3, t};
            else
                tr[y] = ZoomTap{-1, -1, 0.0};
        }
        for (int x = 0; x < W; ++x) {
            int s; double t;
            if (x < z.out1 && zoom_coord(x, z.z1, z.in1, &s, &t))
                tr[H + x] = ZoomTap{(z.top1 + s) * 3, (z.top1 + std::min(s + 1, z.in1 - 1)) * 3, t};
            else
                tr[H + x] = ZoomTap{-1, -1, 0.0};
        }
    }
    const ZoomTap* d_T = reinterpret_cast<const ZoomTap*>(cached_table(key, T.data(), T.size() * sizeof(ZoomTap)));
    if (!d_T) return ADVMIX_ERR_CUDA;
    const size_t img_bytes = (size_t)H * W * 3;
    // one CTA per image only pays when there are enough images to fill the GPU (small per-op groups of the AdvMix
    // chains: many CTAs per image instead)
    if (img_bytes <= 192 * 1024 && a.n >= sm_count() / 2) {
        ADVMIX_CUDA_OK(ensure_dyn_smem(zoom_blur_smem_kernel, 225 * 1024));
        zoom_blur_smem_kernel<<<std::min(a.n, sm_count()), ZS_THREADS, ((img_bytes + 15) & ~(size_t)15) + 256 * 16 * sizeof(double), a.stream>>>(
            a.in, a.out, a.idx, a.n, H, W, d_T, nl);
    } else {
        zoom_blur_kernel<<<st_grid((int64_t)H * W, a.n), ST_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, H, W, d_T, nl);
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== glass_blur
// One iteration of the package's in-place scan: out[p] = in[root(p)], where root follows the
// (dx,dy) references through cells the scan has already rewritten (SURVEY hard part 3).
__global__ void __launch_bounds__(ST_THREADS)
glass_shuffle_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int32_t* __restrict__ idx,
                     const int8_t* __restrict__ field, size_t field_stride, uint64_t seed, int64_t sample_base,
                     int H, int W, int delta, int iter) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const int8_t* inj = field ? field + (size_t)i * field_stride : nullptr;
    const uint8_t* s = src + (int64_t)i * H * W * 3;
    uint8_t* d = dst + (int64_t)i * H * W * 3;
    const int64_t npix = (int64_t)H * W;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += (int64_t)gridDim.x * ST_THREADS) {
        int h = (int)((uint32_t)p / (uint32_t)W), w = (int)((uint32_t)p - (uint32_t)h * (uint32_t)W);
        const bool visited = h > delta && h <= H - delta && w > delta && w <= W - delta;
        if (visited) {
            for (int hop = 0; hop < 4096; ++hop) {
                const int2 o = field_glass(inj, rng, ((uint64_t)iter * H + h) * W + w, delta);
                const int nh = h + o.y, nw = w + o.x;
                const bool earlier = nh > delta && nh <= H - delta && nw > delta && nw <= W - delta &&
                                     (nh > h || (nh == h && nw > w));
                h = nh; w = nw;
                if (!earlier) break;
            }
        }
        const uint8_t* q = s + ((int64_t)h * W + w) * 3;
        uint8_t* o = d + p * 3;
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
    }
}

// Table form of the same scan, one CTA per image.  In glass_shuffle_kernel every pixel walks its own chain of references, one
// Philox block per hop, and a warp waits for its longest chain (about 6 hops where the mean is 2).  Here the offsets are
// drawn once per cell (one Philox block per four cells), a chain node stores its successor in a uint16 table in shared
// memory (G[p] = p for the node whose reference leaves the already-rewritten region; that node also keeps its (dy, dx) in one
// byte), and every pixel follows G to its root: out[p] = in[root + offset(root)].  A thread follows the chains of four
// neighbouring pixels in lock step (independent shared-memory loads in flight; reading G[root] again is harmless).  The first
// version resolved the roots with rounds of G[p] = G[G[p]] over all pixels (5-6 rounds with a CTA barrier each, ~40 % of the
// kernel); the chains are short, so following them costs about half the loads and no barriers.  Bit-identical output.
constexpr int GJ_THREADS = 1024;
__global__ void __launch_bounds__(GJ_THREADS, 1)
glass_shuffle_jump_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int32_t* __restrict__ idx,
                          const int8_t* __restrict__ field, size_t field_stride, uint64_t seed, int64_t sample_base,
                          int n, int H, int W, int delta, int iter) {
    extern __shared__ __align__(16) uint16_t gj_G[];
    const int npix = H * W;
    uint8_t* S = reinterpret_cast<uint8_t*>(gj_G + npix);
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const SampleRng rng(seed, sample_base + slot);
        const int8_t* inj = field ? field + (size_t)img * field_stride : nullptr;
        const uint8_t* s = src + (int64_t)img * npix * 3;
        uint8_t* d = dst + (int64_t)img * npix * 3;
        const uint64_t ebase = (uint64_t)iter * npix;               // npix % 4 == 0: cell quads never straddle two iterations
        __syncthreads();
        for (int j = threadIdx.x; j < npix / 4; j += GJ_THREADS) {
            int ox[4], oy[4];
            if (inj) {
                const int8_t* f = inj + 2 * (ebase + 4 * j);
#pragma unroll
                for (int k = 0; k < 4; ++k) { ox[k] = f[2 * k]; oy[k] = f[2 * k + 1]; }
            } else {
                const uint4 u = rng.quad(TAG_GLASS, (ebase >> 2) + j);
                const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ox[k] = -delta + (int)(((uw[k] & 0xFFFFu) * (2u * delta)) >> 16);
                    oy[k] = -delta + (int)(((uw[k] >> 16) * (2u * delta)) >> 16);
                }
            }
            int h = (int)((uint32_t)(4 * j) / (uint32_t)W), w = 4 * j - h * W;
            uint32_t gq[4], codes = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = 4 * j + k;
                int g = p, code = 0x88;                              // (dy + 8) << 4 | (dx + 8); 0x88 = no displacement
                if (h > delta && h <= H - delta && w > delta && w <= W - delta) {
                    const int nh = h + oy[k], nw = w + ox[k];
                    const bool earlier = nh > delta && nh <= H - delta && nw > delta && nw <= W - delta && (nh > h || (nh == h && nw > w));
                    if (earlier) g = nh * W + nw; else code = ((oy[k] + 8) << 4) | (ox[k] + 8);
                }
                gq[k] = (uint32_t)g;
                codes |= (uint32_t)code << (8 * k);
                if (++w == W) { w = 0; ++h; }
            }
            *reinterpret_cast<uint2*>(gj_G + 4 * j) = make_uint2(gq[0] | (gq[1] << 16), gq[2] | (gq[3] << 16));
            *reinterpret_cast<uint32_t*>(S + 4 * j) = codes;
        }
        __syncthreads();
        const bool vec = ((reinterpret_cast<uintptr_t>(d)) & 3) == 0;
        for (int q = threadIdx.x; q < npix / 4; q += GJ_THREADS) {
            const uint2 g0 = *reinterpret_cast<const uint2*>(gj_G + 4 * q);
            uint32_t r[4] = {g0.x & 0xFFFFu, g0.x >> 16, g0.y & 0xFFFFu, g0.y >> 16};
            for (;;) {
                const uint32_t a0 = gj_G[r[0]], a1 = gj_G[r[1]], a2 = gj_G[r[2]], a3 = gj_G[r[3]];
                const bool moved = (a0 != r[0]) | (a1 != r[1]) | (a2 != r[2]) | (a3 != r[3]);
                r[0] = a0; r[1] = a1; r[2] = a2; r[3] = a3;
                if (!moved) break;
            }
            uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int code = S[r[k]];
                const uint8_t* t = s + ((int)r[k] + ((code >> 4) - 8) * W + ((code & 15) - 8)) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) { const int e = 3 * k + c; o[e >> 2] |= (uint32_t)__ldg(t + c) << (8 * (e & 3)); }
            }
            if (vec) {
                uint32_t* d4 = reinterpret_cast<uint32_t*>(d + 12 * q);
                d4[0] = o[0]; d4[1] = o[1]; d4[2] = o[2];
            } else {
                for (int e = 0; e < 12; ++e) d[12 * q + e] = (uint8_t)(o[e >> 2] >> (8 * (e & 3)));
            }
        }
    }
}

int run_glass_blur(const CorruptArgs& a) {
    const double sigma = glass_sigma(a.severity);
    const int delta = glass_delta(a.severity), iters = glass_iters(a.severity);
    int radius;
    std::vector<double> hw;
    const double* d_w = gauss_table(sigma, 4.0, &radius, &hw);
    if (!d_w) return ADVMIX_ERR_CUDA;
    const int H = a.H, W = a.W, WC = W * 3;
    const int64_t img = (int64_t)H * WC;
    uint8_t* bufA = reinterpret_cast<uint8_t*>(a.ws);                    // [n][H][W][3] u8 (ping)
    uint8_t* bufB = bufA + (size_t)a.n * img;
    int rc;
    // ADVMIX_CORRUPT_FAST: both Gaussians in float32 (the shuffle is integer work either way)
    const float top = fast_top(gauss1d_unit_response(hw.data(), radius, gauss1d_unit_response(hw.data(), radius, 1.0)));
    bool fast = a.fast;
    // x = uint8(gaussian(img/255, sigma) * 255)
    const float top255 = top >= 1.0f ? 255.0f : 254.9999f;
    bool fast_u8 = false;                                            // the uint8-specialised kernel (radius 3 / 4 / 6, W % 4 == 0)
    rc = fast ? launch_gauss_u8_fast(a.in, a.idx, bufA, nullptr, a.n, H, W, radius, d_w, top255, a.stream) : -1;
    if (fast && rc != -1) fast_u8 = true;
    else if (fast)
        rc = launch_gauss2d_fast(LoadU8Div255F{a.in, a.idx, img, WC, nullptr}, StoreU8Trunc255F{bufA, nullptr, img, WC, 0, top}, a.n, H, WC, 3,
                                 radius, radius, d_w, d_w, BORDER_NEAREST, a.stream);
    if (rc == -1) {
        fast = false;
        rc = launch_gauss2d(LoadU8Div255{a.in, a.idx, img, WC, nullptr}, StoreU8Trunc255{bufA, nullptr, img, WC, 0}, a.n, H, WC, 3, radius,
                            radius, d_w, d_w, BORDER_NEAREST, a.stream);
    }
    if (rc) return rc;
    uint8_t *cur = bufA, *nxt = bufB;
    const size_t fstride = a.field_bytes;
    const bool jump = (int64_t)H * W <= 65536;                      // uint16 successor table + 1 byte per pixel in shared memory (<= 192 KB)
    if (jump) ADVMIX_CUDA_OK(ensure_dyn_smem(glass_shuffle_jump_kernel, 3 * 65536));
    for (int it = 0; it < iters; ++it) {
        if (jump)
            glass_shuffle_jump_kernel<<<std::min(a.n, sm_count()), GJ_THREADS, (size_t)3 * H * W, a.stream>>>(
                cur, nxt, a.idx, reinterpret_cast<const int8_t*>(a.rand_field), fstride, a.seed, a.sample_base, a.n, H, W, delta, it);
        else
            glass_shuffle_kernel<<<st_grid((int64_t)H * W, a.n), ST_THREADS, 0, a.stream>>>(
                cur, nxt, a.idx, reinterpret_cast<const int8_t*>(a.rand_field), fstride, a.seed, a.sample_base, H, W, delta, it);
        ADVMIX_LAUNCH_OK();
        std::swap(cur, nxt);
    }
    // clip(gaussian(x/255, sigma), 0, 1) * 255
    if (fast_u8) return launch_gauss_u8_fast(cur, nullptr, a.out, a.idx, a.n, H, W, radius, d_w, top255, a.stream);
    if (fast)
        return launch_gauss2d_fast(LoadU8Div255F{cur, nullptr, img, WC, nullptr}, StoreU8Trunc255F{a.out, a.idx, img, WC, 1, top}, a.n, H, WC, 3,
                                   radius, radius, d_w, d_w, BORDER_NEAREST, a.stream);
    return launch_gauss2d(LoadU8Div255{cur, nullptr, img, WC, nullptr}, StoreU8Trunc255{a.out, a.idx, img, WC, 1}, a.n, H, WC, 3, radius,
                          radius, d_w, d_w, BORDER_NEAREST, a.stream);
}

// ======================================================================== snow
struct SnowParams { double c0, c1, c2, c3; int radius; double sigma, c6; };
static SnowParams snow_params(int s) {
    const SnowParams p[5] = {{0.1, 0.3, 3, 0.5, 10, 4, 0.8}, {0.2, 0.3, 2, 0.5, 12, 4, 0.7}, {0.55, 0.3, 4, 0.9, 12, 8, 0.7},
                             {0.55, 0.3, 4.5, 0.85, 12, 8, 0.65}, {0.55, 0.3, 2.5, 0.85, 12, 12, 0.55}};
    return p[s - 1];
}

// S1: zoomed (clipped_zoom), thresholded, clipped snow layer  [oh][ow] float64
__global__ void __launch_bounds__(ST_THREADS)
snow_layer_kernel(double* __restrict__ layer, const int32_t* __restrict__ idx, const float* __restrict__ field,
                  size_t field_stride, uint64_t seed, int64_t sample_base, int W, ZoomLayer z, double c0, double c1, double c3) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride) : nullptr;
    double* dst = layer + (int64_t)i * z.out0 * z.out1;
    const int64_t total = (int64_t)z.out0 * z.out1;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < total; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)z.out1), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)z.out1);
        int sy, sx;
        double ty, tx, t = 0.0;
        if (zoom_coord(y, z.z0, z.in0, &sy, &ty) && zoom_coord(x, z.z1, z.in1, &sx, &tx)) {
            const int r0 = z.top0 + sy, r1 = z.top0 + min(sy + 1, z.in0 - 1);
            const int q0 = z.top1 + sx, q1 = z.top1 + min(sx + 1, z.in1 - 1);
            const double wy0 = 1.0 - ty, wx0 = 1.0 - tx;
            const double v00 = c0 + c1 * (double)field_normal1(inj, rng, TAG_FIELD0, (uint64_t)r0 * W + q0);
            const double v01 = c0 + c1 * (double)field_normal1(inj, rng, TAG_FIELD0, (uint64_t)r0 * W + q1);
            const double v10 = c0 + c1 * (double)field_normal1(inj, rng, TAG_FIELD0, (uint64_t)r1 * W + q0);
            const double v11 = c0 + c1 * (double)field_normal1(inj, rng, TAG_FIELD0, (uint64_t)r1 * W + q1);
            t = t + (v00 * wy0) * wx0;
            t = t + (v01 * wy0) * tx;
            t = t + (v10 * ty) * wx0;
            t = t + (v11 * ty) * tx;
        }
        if (t < c3) t = 0.0;
        dst[p] = clip01(t);
    }
}

// S2: motion-blur the layer, round(layer*255) -> uint8, cropped to [H][W]
__global__ void __launch_bounds__(ST_THREADS)
snow_blur_kernel(const double* __restrict__ layer, uint8_t* __restrict__ layer8, const int32_t* __restrict__ idx,
                 const double* __restrict__ param, uint64_t seed, int64_t sample_base, int H, int W, int oh, int ow,
                 const double* __restrict__ kw, int width) {
    __shared__ int s_dy[MOTION_MAXW], s_dx[MOTION_MAXW], s_n;
    __shared__ double s_k[MOTION_MAXW];
    __shared__ __align__(16) MotionTapOff s_off[MOTION_MAXW];
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const double angle = param_uniform(param ? param + 4 * i : nullptr, rng, -135.0, -45.0);
    if (threadIdx.x < width) s_k[threadIdx.x] = kw[threadIdx.x];
    motion_offsets(width, angle, oh, ow, s_dy, s_dx, &s_n);
    const int ntaps = s_n;
    if (threadIdx.x < ntaps) s_off[threadIdx.x] = MotionTapOff{s_dy[threadIdx.x] * ow + s_dx[threadIdx.x], 0, s_k[threadIdx.x]};
    __syncthreads();
    int my0 = 0, my1 = 0, mx0 = 0, mx1 = 0;
    for (int t = 0; t < ntaps; ++t) {
        my0 = min(my0, s_dy[t]); my1 = max(my1, s_dy[t]);
        mx0 = min(mx0, s_dx[t]); mx1 = max(mx1, s_dx[t]);
    }
    const double* src = layer + (int64_t)i * oh * ow;
    uint8_t* dst = layer8 + (int64_t)i * H * W;
    const int npix = H * W;
    for (int p = blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += gridDim.x * ST_THREADS) {
        const int y = p / W, x = p - y * W;
        double acc = 0.0;
        if (y - my1 >= 0 && y - my0 < oh && x - mx1 >= 0 && x - mx0 < ow) {
            // the whole line of taps lies inside the layer: no clamps, one 16-byte tap entry per step
            const double* c = src + (y * ow + x);
#pragma unroll 5
            for (int t = 0; t < ntaps; ++t) {
                const MotionTapOff T = s_off[t];
                acc = acc + T.k * __ldg(c - T.off);
            }
        } else {
            for (int t = 0; t < ntaps; ++t) {
                const int yy = clampi(y - s_dy[t], 0, oh - 1), xx = clampi(x - s_dx[t], 0, ow - 1);
                acc = acc + s_k[t] * src[yy * ow + xx];
            }
        }
        dst[p] = (uint8_t)__double2int_rn(acc * 255.0);   // np.round = half-to-even
    }
}

// S3: whiten + composite with the layer and its 180-degree rotation
__global__ void __launch_bounds__(ST_THREADS)
snow_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                  const uint8_t* __restrict__ layer8, int H, int W, float c6, float omc6) {
    __shared__ double d255[256];
    __shared__ float f255[256];
    fill_div255(d255);
    for (int i = threadIdx.x; i < 256; i += ST_THREADS) f255[i] = __fdiv_rn((float)i, 255.0f);
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = out + (int64_t)slot * H * W * 3;
    const uint8_t* L = layer8 + (int64_t)i * H * W;
    const int64_t npix = (int64_t)H * W;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += (int64_t)gridDim.x * ST_THREADS) {
        const uint8_t* q = src + p * 3;
        const float r = f255[q[0]], g = f255[q[1]], b = f255[q[2]];
        // cv2.cvtColor(RGB2GRAY) on float32
        const float gray = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(b, 0.114f));
        const float m = __fadd_rn(__fmul_rn(gray, 1.5f), 0.5f);
        const double add = d255[L[p]] ;
        const double rot = d255[L[npix - 1 - p]];
        const float xs[3] = {r, g, b};
        uint8_t* o = dst + p * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float xv = __fadd_rn(__fmul_rn(c6, xs[c]), __fmul_rn(omc6, fmaxf(xs[c], m)));
            o[c] = trunc_u8(clip01(((double)xv + add) + rot) * 255.0);
        }
    }
}

void snow_layer_dims(int severity, int H, int W, int* oh, int* ow) {
    const ZoomLayer z = zoom_layer(H, W, snow_params(severity).c2);
    *oh = z.out0;
    *ow = z.out1;
}

int run_snow(const CorruptArgs& a) {
    if (a.fast) { const int frc = run_snow_fast(a); if (frc != -1) return frc; }      // ADVMIX_CORRUPT_FAST
    const SnowParams sp = snow_params(a.severity);
    const ZoomLayer z = zoom_layer(a.H, a.W, sp.c2);
    std::vector<double> k = table_weights(SNOW_K[a.severity - 1], sp.radius);
    const double* d_k = reinterpret_cast<const double*>(cached_table("snowk_" + std::to_string(a.severity), k.data(), k.size() * sizeof(double)));
    if (!d_k) return ADVMIX_ERR_CUDA;
    double* layer = reinterpret_cast<double*>(a.ws);
    uint8_t* layer8 = reinterpret_cast<uint8_t*>(layer + (size_t)a.n * z.out0 * z.out1);
    const float* field = reinterpret_cast<const float*>(a.rand_field);
    if (!field) {
        // perf mode: draw the normal field once (each value feeds up to 4 bilinear taps x zoom^2 outputs)
        float* gen = reinterpret_cast<float*>(layer8 + (((size_t)a.n * a.H * a.W + 15) & ~(size_t)15));
        int rc = launch_fill_rand(a, gen, nullptr);
        if (rc) return rc;
        field = gen;
    }
    snow_layer_kernel<<<st_grid((int64_t)z.out0 * z.out1, a.n), ST_THREADS, 0, a.stream>>>(
        layer, a.idx, field, a.field_bytes, a.seed, a.sample_base, a.W, z, sp.c0, sp.c1, sp.c3);
    ADVMIX_LAUNCH_OK();
    snow_blur_kernel<<<st_grid((int64_t)a.H * a.W, a.n), ST_THREADS, 0, a.stream>>>(
        layer, layer8, a.idx, a.rand_param, a.seed, a.sample_base, a.H, a.W, z.out0, z.out1, d_k, 2 * sp.radius + 1);
    ADVMIX_LAUNCH_OK();
    snow_apply_kernel<<<st_grid((int64_t)a.H * a.W, a.n), ST_THREADS, 0, a.stream>>>(
        a.in, a.out, a.idx, layer8, a.H, a.W, (float)sp.c6, (float)(1 - sp.c6));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== fog
// Diamond-square plasma fractal on a torus, one launch per half-level, float64 map in the
// workspace.  U(i,j) is the uniform consumed by map cell (i,j).
__device__ __forceinline__ double wibbled(double sum4, double wibble, float u) {
    // array/4 + wibble * uniform(-wibble, wibble);  uniform = lo + (hi-lo)*u
    return sum4 / 4.0 + wibble * (-wibble + (wibble - (-wibble)) * (double)u);
}

__global__ void __launch_bounds__(ST_THREADS)
fog_squares_kernel(double* __restrict__ maps, const int32_t* __restrict__ idx, const float* __restrict__ field,
                   size_t field_stride, uint64_t seed, int64_t sample_base, int M, int step, double wibble) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride) : nullptr;
    double* m = maps + (int64_t)i * M * M;
    const int n = M / step, h = step / 2;
    if (step == M && blockIdx.x == 0 && threadIdx.x == 0) m[0] = 0.0;   // maparray[0,0] = 0 (only first level reads it)
    for (int t = blockIdx.x * ST_THREADS + threadIdx.x; t < n * n; t += gridDim.x * ST_THREADS) {
        const int a = t / n, b = t - a * n;
        const int a1 = (a + 1) % n, b1 = (b + 1) % n;
        const double c00 = (step == M) ? 0.0 : m[(int64_t)a * step * M + b * step];
        const double c10 = (step == M) ? 0.0 : m[(int64_t)a1 * step * M + b * step];
        const double c01 = (step == M) ? 0.0 : m[(int64_t)a * step * M + b1 * step];
        const double c11 = (step == M) ? 0.0 : m[(int64_t)a1 * step * M + b1 * step];
        // squareaccum = c + roll(c,-1,0);  squareaccum += roll(squareaccum,-1,1)
        const double s = (c00 + c10) + (c01 + c11);
        const int y = a * step + h, x = b * step + h;
        m[(int64_t)y * M + x] = wibbled(s, wibble, field_uniform1(inj, rng, TAG_FIELD0, (uint64_t)y * M + x));
    }
}

__global__ void __launch_bounds__(ST_THREADS)
fog_diamonds_kernel(double* __restrict__ maps, const int32_t* __restrict__ idx, const float* __restrict__ field,
                    size_t field_stride, uint64_t seed, int64_t sample_base, int M, int step, double wibble) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride) : nullptr;
    double* m = maps + (int64_t)i * M * M;
    const int n = M / step, h = step / 2;
    for (int t = blockIdx.x * ST_THREADS + threadIdx.x; t < n * n; t += gridDim.x * ST_THREADS) {
        const int a = t / n, b = t - a * n;
        const int am = (a + n - 1) % n, bm = (b + n - 1) % n, a1 = (a + 1) % n, b1 = (b + 1) % n;
        const double dr = m[(int64_t)(a * step + h) * M + b * step + h];
        const double ul = m[(int64_t)a * step * M + b * step];
        // ltsum = (dr + roll(dr,1,0)) + (ul + roll(ul,-1,1))
        const double lt = (dr + m[(int64_t)(am * step + h) * M + b * step + h]) + (ul + m[(int64_t)a * step * M + b1 * step]);
        // ttsum = (dr + roll(dr,1,1)) + (ul + roll(ul,-1,0))
        const double tt = (dr + m[(int64_t)(a * step + h) * M + bm * step + h]) + (ul + m[(int64_t)a1 * step * M + b * step]);
        int y = a * step, x = b * step + h;
        const double v1 = wibbled(lt, wibble, field_uniform1(inj, rng, TAG_FIELD0, (uint64_t)y * M + x));
        y = a * step + h; x = b * step;
        const double v2 = wibbled(tt, wibble, field_uniform1(inj, rng, TAG_FIELD0, (uint64_t)y * M + x));
        m[(int64_t)(a * step) * M + b * step + h] = v1;
        m[(int64_t)(a * step + h) * M + b * step] = v2;
    }
}

// per image: min / max of the map, max of the image.  One 1024-thread CTA per image.
__global__ void __launch_bounds__(1024)
fog_reduce_kernel(const double* __restrict__ maps, const uint8_t* __restrict__ in, const int32_t* __restrict__ idx,
                  int M, int64_t img_bytes, double* __restrict__ stats) {
    __shared__ double s_mn[32], s_mx[32];
    __shared__ int s_im[32];
    const int i = blockIdx.x, slot = slot_of(idx, i);
    const double* m = maps + (int64_t)i * M * M;
    double mn = INFINITY, mx = -INFINITY;
    for (int t = threadIdx.x; t < M * M; t += 1024) { const double v = m[t]; mn = fmin(mn, v); mx = fmax(mx, v); }
    int im = 0;
    const uint8_t* src = in + (int64_t)slot * img_bytes;
    for (int64_t t = threadIdx.x; t < img_bytes; t += 1024) im = max(im, (int)src[t]);
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        im = max(im, __shfl_xor_sync(0xffffffffu, im, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; s_im[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x < 32) {
        mn = s_mn[threadIdx.x]; mx = s_mx[threadIdx.x]; im = s_im[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            im = max(im, __shfl_xor_sync(0xffffffffu, im, o));
        }
        if (threadIdx.x == 0) {
            stats[4 * i] = mn;
            stats[4 * i + 1] = mx - mn;                    // (maparray - min).max()
            stats[4 * i + 2] = __ddiv_rn((double)im, 255.0);  // x.max()
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS)
fog_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                 const double* __restrict__ maps, const double* __restrict__ stats, int H, int W, int M, double c0) {
    __shared__ double d255[256];
    fill_div255(d255);
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const double mn = stats[4 * i], mx = stats[4 * i + 1], max_val = stats[4 * i + 2];
    const double denom = max_val + c0;
    // Both divisors are per-image constants.  With y = RN(1/b): q0 = RN(a*y), r = a - q0*b (exact, one FMA),
    // q = RN(q0 + r*y) is the correctly rounded a/b (Markstein) - 3 float64 instructions instead of a ~12-instruction
    // division, four times per pixel.  Degenerate divisors (0, inf, nan) keep the plain division.
    const double r_mx = 1.0 / mx, r_den = 1.0 / denom;
    const bool fast_mx = isfinite(r_mx) && r_mx != 0.0, fast_den = isfinite(r_den) && r_den != 0.0;
    auto div_const = [](double a, double b, double rb, bool fast) {
        if (!fast) return a / b;
        const double q0 = a * rb;
        return fma(fma(-q0, b, a), rb, q0);
    };
    const double* m = maps + (int64_t)i * M * M;
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = out + (int64_t)slot * H * W * 3;
    const int64_t npix = (int64_t)H * W;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)W);
        const double add = c0 * div_const(m[(int64_t)y * M + x] - mn, mx, r_mx, fast_mx);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = d255[src[p * 3 + c]] + add;
            dst[p * 3 + c] = trunc_u8(clip01(div_const(v * max_val, denom, r_den, fast_den)) * 255.0);
        }
    }
}

int run_fog(const CorruptArgs& a) {
    if (a.fast) { const int frc = run_fog_fast(a); if (frc != -1) return frc; }      // ADVMIX_CORRUPT_FAST
    const double c0[5] = {1.5, 2., 2.5, 2.5, 3.}, decay[5] = {2, 2, 1.7, 1.5, 1.4};
    const int M = next_pow2(std::max(a.H, a.W));
    double* maps = reinterpret_cast<double*>(a.ws);
    double* stats = maps + (size_t)a.n * M * M;
    const float* f = reinterpret_cast<const float*>(a.rand_field);
    double wibble = 100.0;
    for (int step = M; step >= 2; step /= 2) {
        const int n = M / step;
        fog_squares_kernel<<<st_grid((int64_t)n * n, a.n), ST_THREADS, 0, a.stream>>>(maps, a.idx, f, a.field_bytes, a.seed, a.sample_base, M, step, wibble);
        ADVMIX_LAUNCH_OK();
        fog_diamonds_kernel<<<st_grid((int64_t)n * n, a.n), ST_THREADS, 0, a.stream>>>(maps, a.idx, f, a.field_bytes, a.seed, a.sample_base, M, step, wibble);
        ADVMIX_LAUNCH_OK();
        wibble /= decay[a.severity - 1];
    }
    fog_reduce_kernel<<<a.n, 1024, 0, a.stream>>>(maps, a.in, a.idx, M, (int64_t)a.H * a.W * 3, stats);
    ADVMIX_LAUNCH_OK();
    fog_apply_kernel<<<st_grid((int64_t)a.H * a.W, a.n), ST_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, maps, stats, a.H, a.W, M, c0[a.severity - 1]);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== elastic_transform
struct LoadElasticUniform {     // -max + 2max*u  for field f in {0,1}: layout [2][H][W]
    const float* field; size_t field_stride; const int32_t* idx; uint64_t seed; int64_t sample_base; int H, W; double maxd;
    __device__ void init() {}
    __device__ double operator()(int img2, int y, int x) const {
        const int img = img2 >> 1, f = img2 & 1;
        const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)img * field_stride) + (size_t)f * H * W : nullptr;
        const SampleRng rng(seed, sample_base + (idx ? idx[img] : img));
        const float u = field_uniform1(inj, rng, f ? TAG_FIELD1 : TAG_FIELD0, (uint64_t)y * W + x);
        return -maxd + (maxd - (-maxd)) * (double)u;
    }
    // row-wise access (fused kernel; the field is always materialised there)
    typedef const float* Row;
    __device__ Row row(int img2, int y) const {
        return reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)(img2 >> 1) * field_stride) +
               (size_t)(img2 & 1) * H * W + (size_t)y * W;
    }
    __device__ double at(Row r, int xc) const { return -maxd + (maxd - (-maxd)) * (double)r[xc]; }
    typedef float Raw;
    __device__ Raw raw(Row r, int xc) const { return __ldg(r + xc); }
    __device__ double cvt(Raw v) const { return -maxd + (maxd - (-maxd)) * (double)v; }
};
struct StoreF32Scaled {
    float* base; int64_t stride; int WC; double alpha;
    __device__ void operator()(int img, int y, int xc, double v) const { base[(int64_t)img * stride + (int64_t)y * WC + xc] = (float)(v * alpha); }
};

// scipy map_coordinate() for mode='reflect'
__device__ __forceinline__ double scipy_reflect(double in, int len) {
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

__global__ void __launch_bounds__(ST_THREADS)
elastic_gather_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                      const float* __restrict__ disp, int H, int W) {
    __shared__ float f255[256];
    for (int i = threadIdx.x; i < 256; i += ST_THREADS) f255[i] = __fdiv_rn((float)i, 255.0f);
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    uint8_t* dst = out + (int64_t)slot * H * W * 3;
    const int64_t npix = (int64_t)H * W;
    const float* dxf = disp + (int64_t)(2 * i) * npix;
    const float* dyf = disp + (int64_t)(2 * i + 1) * npix;
    for (int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; p < npix; p += (int64_t)gridDim.x * ST_THREADS) {
        const int y = (int)((uint32_t)p / (uint32_t)W), x = (int)((uint32_t)p - (uint32_t)y * (uint32_t)W);
        const double cy = scipy_reflect((double)y + (double)dyf[p], H);
        const double cx = scipy_reflect((double)x + (double)dxf[p], W);
        const double fy = floor(cy), fx = floor(cx);
        const double ty = cy - fy, tx = cx - fx;
        const int y0 = reflect_sym((int)fy, H), y1 = reflect_sym((int)fy + 1, H);
        const int x0 = reflect_sym((int)fx, W), x1 = reflect_sym((int)fx + 1, W);
        const double wy0 = 1.0 - ty, wx0 = 1.0 - tx;
        const uint8_t* p00 = src + ((int64_t)y0 * W + x0) * 3;
        const uint8_t* p01 = src + ((int64_t)y0 * W + x1) * 3;
        const uint8_t* p10 = src + ((int64_t)y1 * W + x0) * 3;
        const uint8_t* p11 = src + ((int64_t)y1 * W + x1) * 3;
        uint8_t* o = dst + p * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double t = 0.0;
            t = t + ((double)f255[__ldg(p00 + c)] * wy0) * wx0;
            t = t + ((double)f255[__ldg(p01 + c)] * wy0) * tx;
            t = t + ((double)f255[__ldg(p10 + c)] * ty) * wx0;
            t = t + ((double)f255[__ldg(p11 + c)] * ty) * tx;
            const float r = fminf(fmaxf((float)t, 0.f), 1.f);
            o[c] = (uint8_t)(int)__fmul_rn(r, 255.f);
        }
    }
}

int run_elastic(const CorruptArgs& a) {
    if (a.fast) { const int frc = run_elastic_fast(a); if (frc != -1) return frc; }      // ADVMIX_CORRUPT_FAST
    const double alpha[5] = {250 * 0.05, 250 * 0.065, 250 * 0.085, 250 * 0.1, 250 * 0.12};
    const int H = a.H, W = a.W;
    const double sig0 = H * 0.01, sig1 = W * 0.01, maxd = H * 0.005;
    int r0, r1;
    const double* w0 = gauss_table(sig0, 3.0, &r0);
    const double* w1 = gauss_table(sig1, 3.0, &r1);
    if (!w0 || !w1) return ADVMIX_ERR_CUDA;
    ADVMIX_REQUIRE(r0 <= GAUSS_MAXR && r1 <= GAUSS_MAXR, "elastic: image too large for the Gaussian radius cap (%d)", GAUSS_MAXR);
    ADVMIX_REQUIRE(r0 < H && r1 < W, "elastic: image too small");
    const int64_t plane = (int64_t)H * W;
    float* disp = reinterpret_cast<float*>(a.ws);                     // [2n][H][W]  (dx, dy)
    const float* field = reinterpret_cast<const float*>(a.rand_field);
    int rc;
    if (!field) {
        // perf mode: draw the two uniform fields once (each value is read by 2R+1 filter taps)
        float* gen = disp + (size_t)2 * a.n * plane;
        rc = launch_fill_rand(a, gen, nullptr);
        if (rc) return rc;
        field = gen;
    }
    rc = launch_gauss2d(LoadElasticUniform{field, a.field_bytes, a.idx, a.seed, a.sample_base, H, W, maxd},
                        StoreF32Scaled{disp, plane, W, alpha[a.severity - 1]}, 2 * a.n, H, W, 1, r0, r1, w0, w1, BORDER_REFLECT, a.stream);
    if (rc) return rc;
    elastic_gather_kernel<<<st_grid(plane, a.n), ST_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, disp, H, W);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ------------------------------------------------------------------------ workspace sizes
size_t stencil_ws_bytes(int op, int severity, int n, int H, int W) {
    const size_t img = (size_t)H * W * 3;
    switch (op) {
        case C_GLASS_BLUR: return 2 * (size_t)n * img;
        case C_SNOW: {
            int oh, ow;
            snow_layer_dims(severity, H, W, &oh, &ow);
            return (size_t)n * oh * ow * sizeof(double) + (size_t)n * H * W + 16 + (size_t)n * H * W * sizeof(float);
        }
        case C_FOG: {
            const size_t M = next_pow2(std::max(H, W));
            return (size_t)n * M * M * sizeof(double) + (size_t)n * 4 * sizeof(double);
        }
        case C_ELASTIC: return (size_t)2 * n * H * W * 2 * sizeof(float);   // disp, generated fields
        default: return 0;
    }
}

}  // namespace advmix
