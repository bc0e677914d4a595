// ADVMIX_CORRUPT_FAST variants of the stencil / gather corruptions: the float64 operation order of
// corrupt_stencil.cu costs the FP64 pipe (64 lanes per SM) and 8-byte shared-memory traffic; north_star's bar for
// floating-point ops is "max abs <= 1 LSB after the final truncation", which float32 / 24-bit fixed-point arithmetic
// meets with < 0.2 % of values moving by one LSB.  Same algorithms, same constant tables, same random draws:
//   defocus_blur  sparse disk taps as FFMA immediates (compile-time tap table per severity, zero taps elided)
//   motion_blur   24-bit fixed-point weights, IMAD on bytes, image resident in shared memory
//   zoom_blur     12-bit x 12-bit fixed-point bilinear weights (8 + 24 bits per product), image resident in shared memory
//   snow          float32 layer + line blur
//   fog           the whole diamond-square recursion, the reduction and the apply pass in one CTA per image
//   elastic       float32 Gaussian of the displacement fields (stencil_common.cuh) + float32 gather
// Saturated regions: whether sum(weights) * 1.0 lands on or just below 1.0 in the reference's float64 order decides
// 255 vs 254 there; each launcher evaluates that order on the host and passes the result (`top`) to the kernel.
#include "stencil_common.cuh"

#include <climits>

namespace advmix {

// value in units of 1/255 -> uint8 (truncation); `top255` is what a saturated region gives in the reference's float64 order
__device__ __forceinline__ uint32_t v255_to_u8(float v, float top255) {
    v = fminf(v, 255.0f);
    if (v > 254.9997f) v = top255;                       // ~1e-6 relative: the float32 accumulation error on a saturated region
    return (uint32_t)__float2int_rz(fmaxf(v, 0.f));
}
static inline float fast_top255(double unit_response) { return unit_response >= 1.0 ? 255.0f : 254.9999f; }

// ======================================================================== defocus_blur
template <int SEV> struct DiskFast;
#define ADVMIX_DISK_FAST(S, HH)                                                                  \
    template <> struct DiskFast<S> {                                                             \
        static constexpr int h = HH;                                                             \
        static constexpr int n = (int)(sizeof(DISK_TAPS_##S) / sizeof(DiskTap));                 \
        static constexpr DiskTap tap(int i) { return DISK_TAPS_##S[i]; }                         \
    };
ADVMIX_DISK_FAST(1, 3) ADVMIX_DISK_FAST(2, 5) ADVMIX_DISK_FAST(3, 7) ADVMIX_DISK_FAST(4, 8) ADVMIX_DISK_FAST(5, 10)
#undef ADVMIX_DISK_FAST

constexpr float DISK_EPS = 1e-7f;      // the alias blur leaves taps of 2^-77 .. 4e-10 around the disk (1.1e-6 LSB in total): dropped here

template <int SEV> struct DiskDense { float w[2 * DiskFast<SEV>::h + 1][2 * DiskFast<SEV>::h + 1]; float cmax; };
template <int SEV> constexpr DiskDense<SEV> disk_dense() {
    DiskDense<SEV> d{};
    constexpr int h = DiskFast<SEV>::h;
    for (int i = 0; i < DiskFast<SEV>::n; ++i) {
        const DiskTap t = DiskFast<SEV>::tap(i);
        if (t.w > DISK_EPS && t.dy >= -h && t.dy <= h && t.dx >= -h && t.dx <= h) {
            d.w[t.dy + h][t.dx + h] = t.w;
            if (t.w > d.cmax) d.cmax = t.w;           // the weight 1/N of the disk's interior (the alias blur only touches the rim)
        }
    }
    return d;
}

// the interior taps (== cmax) of every kernel row must form one contiguous run: the kernel sums them with a sliding window
template <int SEV> constexpr bool disk_runs_contiguous() {
    constexpr DiskDense<SEV> D = disk_dense<SEV>();
    constexpr int n = 2 * DiskFast<SEV>::h + 1;
    for (int r = 0; r < n; ++r) {
        int state = 0;                       // 0: before the run, 1: inside, 2: after
        for (int d = 0; d < n; ++d) {
            const bool in = D.w[r][d] == D.cmax;
            if (in && state == 2) return false;
            if (in) state = 1; else if (state == 1) state = 2;
        }
    }
    return true;
}

constexpr int DF_BW = 64, DF_BH = 32, DF_THREADS = 256;

// 64x32 outputs per CTA, a 4 (x) x 2 (y) block per thread, float32 tile of the BYTE values (+halo, reflect-101 resolved at
// load time) in shared memory, planar per channel.  The tap loops are fully unrolled over a constexpr dense weight array:
// every rim tap is one FFMA with the weight as an immediate operand, the zero taps vanish, and the interior taps (all equal
// to 1/N) are summed as exact integers and scaled once - with N equal weights the reference's result sits a few 1e-8 below
// an integer whenever the byte sum is a multiple of N (3.4 % of all values at severity 1), so the decision `k or k - 1`
// needs the sum exact and a single rounding.
template <int SEV>
__global__ void __launch_bounds__(DF_THREADS)
defocus_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx, int H, int W, float top255) {
    constexpr int h = DiskFast<SEV>::h, WN = (2 * h + 4 + 3) / 4 * 4, TW = DF_BW - 4 + WN, TH = DF_BH + 2 * h;
    constexpr DiskDense<SEV> D = disk_dense<SEV>();
    static_assert(disk_runs_contiguous<SEV>(), "interior taps of a kernel row must be contiguous");
    extern __shared__ __align__(16) float df_tile[];          // [3][TH][TW]
    const int slot = slot_of(idx, blockIdx.z);
    const uint8_t* src = in + (int64_t)slot * H * W * 3;
    const int x0 = blockIdx.x * DF_BW, y0 = blockIdx.y * DF_BH;
    for (int ty = threadIdx.x >> 5; ty < TH; ty += DF_THREADS / 32) {
        const int gy = reflect101(y0 + ty - h, H);
        const uint8_t* row = src + (int64_t)gy * W * 3;
        float* t0 = df_tile + ty * TW;
        for (int tx = threadIdx.x & 31; tx < TW; tx += 32) {
            const int gx = reflect101(x0 + tx - h, W);
            const uint8_t* p = row + gx * 3;
            t0[tx] = u16_to_float(p[0]);
            t0[TH * TW + tx] = u16_to_float(p[1]);
            t0[2 * TH * TW + tx] = u16_to_float(p[2]);
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = 2 * (threadIdx.x >> 4);
    const int x = x0 + 4 * tx, y = y0 + ty;
    uint32_t res[2][4][3];
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        float acc[2][4], isum[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = isum[j][i] = 0.f;
        const float* base = df_tile + (c * TH + ty) * TW + 4 * tx;
#pragma unroll
        for (int t = 0; t < 2 * h + 2; ++t) {                 // tile row ty + t: kernel row t (upper output row), t - 1 (lower)
            float win[WN];
#pragma unroll
            for (int k = 0; k < WN / 4; ++k) {
                const float4 v = *reinterpret_cast<const float4*>(base + t * TW + 4 * k);
                win[4 * k] = v.x; win[4 * k + 1] = v.y; win[4 * k + 2] = v.z; win[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = t - j;
                if (r >= 0 && r <= 2 * h) {
                    // interior run [ra, rb] of this kernel row (taps == cmax; folds to constants once the loops are unrolled):
                    // the first output sums the run, the next three slide it by one (+1 in, -1 out) - len + 6 exact integer
                    // adds for four outputs instead of 4 * len
                    int ra = -1, rb = -1;
#pragma unroll
                    for (int d = 0; d <= 2 * h; ++d)
                        if (D.w[r][d] == D.cmax) { if (ra < 0) ra = d; rb = d; }
                    if (ra >= 0) {
                        float s0 = 0.f;
#pragma unroll
                        for (int d = 0; d <= 2 * h; ++d)
                            if (d >= ra && d <= rb) s0 = __fadd_rn(s0, win[d]);
                        float sl[4];
                        sl[0] = s0;
#pragma unroll
                        for (int i = 1; i < 4; ++i) {
                            float add_v = 0.f, sub_v = 0.f;
#pragma unroll
                            for (int d = 0; d < WN; ++d) {               // constant-index picks after unrolling
                                if (d == rb + i) add_v = win[d];
                                if (d == ra + i - 1) sub_v = win[d];
                            }
                            sl[i] = __fsub_rn(__fadd_rn(sl[i - 1], add_v), sub_v);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) isum[j][i] = __fadd_rn(isum[j][i], sl[i]);               // exact: integers < 2^24
                    }
#pragma unroll
                    for (int d = 0; d <= 2 * h; ++d) {
                        const float w = D.w[r][d];
                        if (w != D.cmax && w != 0.f) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(w, win[d + i], acc[j][i]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // v = cmax * isum + rim.  When v lands exactly on an integer the truncation hangs on the sign of the rounding
                // error, which one more FMA recovers (cmax * N is a few 1e-9 below 1: byte sums that are multiples of N give
                // k - 4e-9 k in the reference, i.e. k - 1 after truncation).
                float v = fmaf(D.cmax, isum[j][i], acc[j][i]);
                const float resid = __fadd_rn(fmaf(D.cmax, isum[j][i], -v), acc[j][i]);
                if (v == truncf(v) && resid < 0.f) v = __fadd_rn(v, -0.5f);
                const uint32_t v8 = v255_to_u8(v, top255);
                if (c == 0) res[j][i][0] = v8; else if (c == 1) res[j][i][1] = v8; else res[j][i][2] = v8;
            }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (y + j < H) {
            uint8_t* o = out + (int64_t)slot * H * W * 3 + ((int64_t)(y + j) * W + x) * 3;
            if (x + 3 < W && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
                uint32_t* o4 = reinterpret_cast<uint32_t*>(o);
                o4[0] = res[j][0][0] | (res[j][0][1] << 8) | (res[j][0][2] << 16) | (res[j][1][0] << 24);
                o4[1] = res[j][1][1] | (res[j][1][2] << 8) | (res[j][2][0] << 16) | (res[j][2][1] << 24);
                o4[2] = res[j][2][2] | (res[j][3][0] << 8) | (res[j][3][1] << 16) | (res[j][3][2] << 24);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (x + i < W) { o[3 * i] = (uint8_t)res[j][i][0]; o[3 * i + 1] = (uint8_t)res[j][i][1]; o[3 * i + 2] = (uint8_t)res[j][i][2]; }
            }
        }
}

template <int SEV>
static int launch_defocus_fast(const CorruptArgs& a) {
    constexpr int h = DiskFast<SEV>::h, WN = (2 * h + 4 + 3) / 4 * 4, TW = DF_BW - 4 + WN, TH = DF_BH + 2 * h;
    constexpr size_t smem = (size_t)3 * TH * TW * sizeof(float);
    // the reference's float64 sum of the taps in row-major order on a constant 1.0 image
    double s = 0.0;
    for (int i = 0; i < DiskFast<SEV>::n; ++i) s = std::fma((double)DiskFast<SEV>::tap(i).w, 1.0, s);
    ADVMIX_CUDA_OK(ensure_dyn_smem(defocus_fast_kernel<SEV>, (int)smem));
    dim3 grid(ceil_div(a.W, DF_BW), ceil_div(a.H, DF_BH), a.n);
    defocus_fast_kernel<SEV><<<grid, DF_THREADS, smem, a.stream>>>(a.in, a.out, a.idx, a.H, a.W, fast_top255(s));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_defocus_blur_fast(const CorruptArgs& a) {
    switch (a.severity) {
        case 1: return launch_defocus_fast<1>(a);
        case 2: return launch_defocus_fast<2>(a);
        case 3: return launch_defocus_fast<3>(a);
        case 4: return launch_defocus_fast<4>(a);
        default: return launch_defocus_fast<5>(a);
    }
}

// ======================================================================== motion_blur
// blurred = sum_i k_i * shift(x, dx_i, dy_i); k_i as 24-bit fixed point K_i (sum K_i = 2^24, or 2^24 - 1 when the reference's
// float64 sum on a saturated region stays below 255), so a channel value is sum(byte * K) >> 24: one IMAD per tap and
// channel.  The image sits in shared memory as packed RGB bytes, a thread owns 4 horizontally adjacent pixels = three
// aligned words; a tap's 12 bytes are three words funnel-shifted out of four.
constexpr int MF_THREADS = 1024, MF_MAXW = 41;

__device__ __forceinline__ void motion_offsets_f(int width, double angle_deg, int H, int W, int* s_dy, int* s_dx, int* s_n) {
    if (threadIdx.x < width) {
        const int i = threadIdx.x;
        const double rad = angle_deg * (3.141592653589793 / 180.0);
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

// Shared-memory image with x padding: for |angle| <= 45 degrees every tap samples at or to the RIGHT of its pixel (dx <= 0), up to
// 2 * radius pixels away, so each staged row carries `pad` replicated copies of its last pixel (= clamp-to-edge).  No tap of
// any pixel then needs an x clamp, which removes the border path - with rows of 48 four-pixel groups nearly every warp holds a
// border group, so the divergent slow path used to run in (almost) all warps.  Rows are clamped per tap (warp-uniform) only in
// the few rows whose taps leave the image.
__global__ void __launch_bounds__(MF_THREADS, 1)
motion_blur_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                        const double* __restrict__ param, uint64_t seed, int64_t sample_base, int n, int H, int W, int pad,
                        const uint32_t* __restrict__ kfix, int width) {
    extern __shared__ __align__(16) uint8_t mf_img[];
    __shared__ int s_dy[MF_MAXW], s_dx[MF_MAXW], s_n;
    __shared__ int2 s_tap[MF_MAXW];                 // {byte offset of the tap relative to the pixel (padded pitch), K}
    const int W3 = 3 * W, pitch = 3 * (W + pad), nwords = H * W3 / 4, rw = W3 / 4;
    const int wq = W >> 2;                          // W % 4 == 0 (checked by the launcher)
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const SampleRng rng(seed, sample_base + slot);
        const double angle = param_uniform(param ? param + 4 * img : nullptr, rng, -45.0, 45.0);
        const uint8_t* src = in + (int64_t)slot * H * W3;
        uint8_t* dst = out + (int64_t)slot * H * W3;
        __syncthreads();
        {
            const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
            for (int i = threadIdx.x; i < nwords; i += MF_THREADS) {
                const int y = i / rw, xw = i - y * rw;
                *reinterpret_cast<uint32_t*>(mf_img + y * pitch + 4 * xw) = __ldg(s4 + i);
            }
            for (int i = threadIdx.x; i < H * pad; i += MF_THREADS) {      // replicate the last pixel of every row
                const int y = i / pad, k = i - y * pad;
                const uint8_t* last = src + y * W3 + W3 - 3;
                uint8_t* d = mf_img + y * pitch + W3 + 3 * k;
                d[0] = last[0]; d[1] = last[1]; d[2] = last[2];
            }
        }
        motion_offsets_f(width, angle, H, W, s_dy, s_dx, &s_n);
        const int ntaps = s_n;                      // == width (the launcher sends smaller images to the float64 path)
        if (threadIdx.x < ntaps) s_tap[threadIdx.x] = make_int2(s_dy[threadIdx.x] * pitch + s_dx[threadIdx.x] * 3, (int)kfix[threadIdx.x]);
        __syncthreads();
        int my0 = 0, my1 = 0;
        for (int t = 0; t < ntaps; ++t) { my0 = min(my0, s_dy[t]); my1 = max(my1, s_dy[t]); }
        for (int g = threadIdx.x; g < H * wq; g += MF_THREADS) {
            const int y = g / wq, x = (g - y * wq) << 2;
            uint32_t acc[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) acc[k] = 0u;
            const bool rows_inside = y - my1 >= 0 && y - my0 < H;
            const int c = y * pitch + x * 3;
#pragma unroll 2
            for (int t = 0; t < ntaps; ++t) {
                const int2 T = s_tap[t];
                int b = c - T.x;                                         // byte address of the tap's first pixel
                if (!rows_inside) b = clampi(y - s_dy[t], 0, H - 1) * pitch + (x - s_dx[t]) * 3;
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(mf_img + (b & ~3));
                const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(b & 3);      // byte-granular funnel shift = one PRMT
                const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];   // w3 may read 3 bytes past the group: inside the padded buffer
                const uint32_t q[3] = {__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel)};
                const uint32_t K = (uint32_t)T.y;
#pragma unroll
                for (int e = 0; e < 12; ++e) acc[e] += __byte_perm(q[e >> 2], 0u, 0x4440u | (uint32_t)(e & 3)) * K;
            }
            uint32_t* o = reinterpret_cast<uint32_t*>(dst + (y * W + x) * 3);
            o[0] = (acc[0] >> 24) | ((acc[1] >> 24) << 8) | ((acc[2] >> 24) << 16) | (acc[3] & 0xFF000000u);
            o[1] = (acc[4] >> 24) | ((acc[5] >> 24) << 8) | ((acc[6] >> 24) << 16) | (acc[7] & 0xFF000000u);
            o[2] = (acc[8] >> 24) | ((acc[9] >> 24) << 8) | ((acc[10] >> 24) << 16) | (acc[11] & 0xFF000000u);
        }
    }
}

// 24-bit fixed-point weights whose sum is exactly 2^24 (2^24 - 1 when `below_one`): largest-remainder rounding
static std::vector<uint32_t> fixed_weights(const double* k, int n, bool below_one) {
    std::vector<uint32_t> K(n);
    std::vector<std::pair<double, int>> frac(n);
    long long sum = 0;
    double ks = 0;
    for (int i = 0; i < n; ++i) ks += k[i];
    for (int i = 0; i < n; ++i) {
        const double v = k[i] / ks * 16777216.0;
        K[i] = (uint32_t)std::floor(v);
        frac[i] = {v - std::floor(v), i};
        sum += K[i];
    }
    std::sort(frac.begin(), frac.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; });
    long long target = 16777216LL - (below_one ? 1 : 0);
    for (int i = 0; sum < target && i < n; ++i, ++sum) K[frac[i].second] += 1;
    return K;
}

int run_motion_blur_fast(const CorruptArgs& a) {
    const int r = MOTION_RADIUS[a.severity - 1], width = 2 * r + 1;
    // image-resident kernel: W % 4 == 0, 4-byte aligned images whose x-padded copy fits shared memory, no truncated tap list
    const int pad = (2 * r + 3) / 4 * 4;                             // |dx| <= 2 * radius, taps only to the right (|angle| <= 45)
    const size_t smem = (size_t)a.H * (a.W + pad) * 3 + 16;
    if (a.W % 4 != 0 || smem > 227 * 1024 - 2048 || a.H <= width || a.W <= width ||
        ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out)) & 3) != 0)
        return -1;
    const double* k = MOTION_K[a.severity - 1];
    // the float64 path on a saturated (255) region: blurred = blurred + k_i * 255, then clip to [0, 255]
    double s = 0.0;
    for (int i = 0; i < width; ++i) s = s + k[i] * 255.0;
    std::vector<uint32_t> K = fixed_weights(k, width, s < 255.0);
    const uint32_t* d_K = reinterpret_cast<const uint32_t*>(cached_table("motionfix_" + std::to_string(a.severity), K.data(), K.size() * 4));
    if (!d_K) return ADVMIX_ERR_CUDA;
    ADVMIX_CUDA_OK(ensure_dyn_smem(motion_blur_fast_kernel, 227 * 1024 - 2048));
    motion_blur_fast_kernel<<<std::min(a.n, sm_count()), MF_THREADS, smem, a.stream>>>(
        a.in, a.out, a.idx, a.rand_param, a.seed, a.sample_base, a.n, a.H, a.W, pad, d_K, width);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// ======================================================================== zoom_blur
// Layer value = bilinear sample with 12-bit fixed-point weights per axis: sum of 4 byte * (WY * WX) products, WY * WX < 2^24,
// so a sample fits 32 bits exactly.  Each layer's sample is reduced to 24 bits (>> 8) and the layers are summed in a
// 32-bit integer; one float32 division at the end.  Weight quantisation moves a layer by <= 2^-13 of the local gradient.
struct ZoomTapF { int o0, o1; uint32_t w1; int pad; };      // byte offsets of the two rows (columns); w1 = round(t * 4096); o0 < 0: outside

constexpr int ZF_THREADS = 1024;

// Tap entries in the kernel: 8 bytes {o0 (-1: outside), w1 | step << 16} where o1 = o0 + (step ? pitch : 0).  When the image
// AND the tables (nl * (H + W) * 8 bytes) fit shared memory the tables are staged there: ncu showed the first version
// (entries read from global memory every layer) issue-active 74 % but halving its instruction count did not change its
// time - the per-layer chain {column entry -> row entry -> byte reads -> multiplies} was bound by the entries' L1 / L2
// latency with only 8 warps per scheduler.
__device__ __forceinline__ uint2 zoom_pack(const ZoomTapF& t) {     // outside: offset 0 and bit 31 (both weights are then taken as 0)
    return t.o0 < 0 ? make_uint2(0u, 0x80000000u) : make_uint2((uint32_t)t.o0, t.w1 | ((t.o1 != t.o0 ? 1u : 0u) << 16));
}

template <bool TAB_SMEM>
__global__ void __launch_bounds__(ZF_THREADS, 1)
zoom_blur_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                      int n, int H, int W, const ZoomTapF* __restrict__ taps, int nl, float denom) {
    extern __shared__ __align__(16) uint8_t zf_img[];
    const int nbytes = H * W * 3, W3 = W * 3;
    uint2* s_tab = reinterpret_cast<uint2*>(zf_img + ((nbytes + 15) & ~15));
    if (TAB_SMEM)
        for (int i = threadIdx.x; i < nl * (H + W); i += ZF_THREADS) s_tab[i] = zoom_pack(taps[i]);
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const uint8_t* src = in + (int64_t)slot * nbytes;
        uint8_t* dst = out + (int64_t)slot * nbytes;
        __syncthreads();
        {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(zf_img);
            for (int i = threadIdx.x; i < nbytes / 16; i += ZF_THREADS) d4[i] = ld_stream_u4(s4 + i);
        }
        __syncthreads();
        // a thread owns one column x 4 rows: the column entry of a layer is loaded once per 4 pixels.  Straight-line code: two row
        // evaluations per output row and no tests (an earlier version reused a source row's horizontal result between the thread's
        // output rows; its tests and moves cost more than the 0.8 row evaluations they saved: ~33 -> ~21 instructions per value and
        // layer), 'outside' is folded into zero weights.  Same integers as the 4-product form.
        const int hq = (H + 3) >> 2;
        for (int item = threadIdx.x; item < hq * W; item += ZF_THREADS) {
            const int yq = item / W, x = item - yq * W;
            const int y0 = yq << 2;
            uint32_t acc[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0u;
#pragma unroll 2
            for (int l = 0; l < nl; ++l) {
                const int base = l * (H + W);
                const uint2 cc = TAB_SMEM ? s_tab[base + H + x] : zoom_pack(taps[base + H + x]);
                const uint32_t wx1 = cc.y & 0xFFFFu, wx0 = (cc.y >> 31) ? 0u : 4096u - wx1;
                const uint8_t* c0p = zf_img + cc.x;
                const uint8_t* c1p = c0p + ((cc.y >> 16) & 1u) * 3u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int yi = min(y0 + i, H - 1);                   // rows past the image are evaluated on row H-1 and not stored
                    const uint2 rr = TAB_SMEM ? s_tab[base + yi] : zoom_pack(taps[base + yi]);
                    const uint32_t r0 = rr.x, r1 = r0 + ((rr.y >> 16) & 1u) * (uint32_t)W3;
                    const uint32_t wy1 = rr.y & 0xFFFFu, wy0 = (rr.y >> 31) ? 0u : 4096u - wy1;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const uint32_t hT = c0p[r0 + c] * wx0 + c1p[r0 + c] * wx1;
                        const uint32_t hB = c0p[r1 + c] * wx0 + c1p[r1 + c] * wx1;
                        acc[i][c] += (hT * wy0 + hB * wy1 + 128u) >> 8;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (y0 + i >= H) break;
                const int pb = ((y0 + i) * W + x) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // (x/255 + sum_l layer_l) / (nl + 1) * 255, layers in units of 255 * 2^16
                    const float s = fmaf((float)acc[i][c], 1.0f / 65536.0f, (float)zf_img[pb + c]);
                    dst[pb + c] = (uint8_t)__float2int_rz(fminf(__fdiv_rn(s, denom), 255.0f));     // a true division: (nl+1)*255 / (nl+1) must give 255
                }
            }
        }
    }
}

static int py_round_f(double v) { return (int)std::nearbyint(v); }

int run_zoom_blur_fast(const CorruptArgs& a) {
    const int H = a.H, W = a.W;
    const size_t img_bytes = (size_t)H * W * 3;
    if (img_bytes % 16 != 0 || img_bytes > 220 * 1024 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0) return -1;
    const double stop[5] = {1.11, 1.16, 1.21, 1.26, 1.33}, step[5] = {0.01, 0.01, 0.02, 0.02, 0.03};
    const int nl = (int)std::ceil((stop[a.severity - 1] - 1.0) / step[a.severity - 1]);
    const std::string key = "zoomfix_" + std::to_string(H) + "x" + std::to_string(W) + "_" + std::to_string(a.severity);
    std::vector<ZoomTapF> T((size_t)nl * (H + W));
    for (int l = 0; l < nl; ++l) {
        const double zf = 1.0 + l * step[a.severity - 1];
        // clipped_zoom + scipy.ndimage.zoom(order=1) geometry (same as corrupt_stencil.cu:zoom_layer)
        const int in0 = (int)std::ceil(H / zf), top0 = (H - in0) / 2, in1 = (int)std::ceil(W / zf), top1 = (W - in1) / 2;
        const int out0 = py_round_f(in0 * zf), out1 = py_round_f(in1 * zf);
        const double z0 = out0 > 1 ? (double)(in0 - 1) / (double)(out0 - 1) : 1.0, z1 = out1 > 1 ? (double)(in1 - 1) / (double)(out1 - 1) : 1.0;
        ZoomTapF* tr = T.data() + (size_t)l * (H + W);
        auto entry = [](int o, int outn, double z, int inn, int top, int pitch) {
            const double cc = (double)o * z;
            if (o >= outn || cc < 0.0 || cc > (double)(inn - 1)) return ZoomTapF{-1, -1, 0u, 0};
            const double f = std::floor(cc);
            const int s = (int)f;
            uint32_t w1 = (uint32_t)std::nearbyint((cc - f) * 4096.0);
            return ZoomTapF{(top + s) * pitch, (top + std::min(s + 1, inn - 1)) * pitch, w1, 0};
        };
        for (int y = 0; y < H; ++y) tr[y] = entry(y, out0, z0, in0, top0, W * 3);
        for (int x = 0; x < W; ++x) tr[H + x] = entry(x, out1, z1, in1, top1, 3);
    }
    const ZoomTapF* d_T = reinterpret_cast<const ZoomTapF*>(cached_table(key, T.data(), T.size() * sizeof(ZoomTapF)));
    if (!d_T) return ADVMIX_ERR_CUDA;
    const size_t img_al = (img_bytes + 15) & ~(size_t)15, tab_bytes = (size_t)nl * (H + W) * 8;
    const size_t smem_cap = 227 * 1024 - 2048;
    if (img_al + tab_bytes <= smem_cap) {
        ADVMIX_CUDA_OK(ensure_dyn_smem(zoom_blur_fast_kernel<true>, (int)smem_cap));
        zoom_blur_fast_kernel<true><<<std::min(a.n, sm_count()), ZF_THREADS, img_al + tab_bytes, a.stream>>>(
            a.in, a.out, a.idx, a.n, H, W, d_T, nl, (float)(nl + 1));
    } else {
        ADVMIX_CUDA_OK(ensure_dyn_smem(zoom_blur_fast_kernel<false>, (int)smem_cap));
        zoom_blur_fast_kernel<false><<<std::min(a.n, sm_count()), ZF_THREADS, img_al, a.stream>>>(
            a.in, a.out, a.idx, a.n, H, W, d_T, nl, (float)(nl + 1));
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // namespace advmix
