// Row a1 / a7: affine crop (cv2.warpAffine fixed-point semantics), batched
// get_affine_transform, joint flip+affine, ToTensor+Normalize.
//
// Reference: lib/dataset/JointsDataset.py:167-199, :324-332; lib/utils/transforms.py:44-122.
#include "common.cuh"

namespace advmix {

constexpr int AB_BITS = 10;
constexpr int INTER_BITS = 5;
constexpr int ROUND_DELTA = (1 << AB_BITS) / (1 << INTER_BITS) / 2;  // 16
constexpr int WARP_THREADS = 256;

// cv::warpAffine's inversion of the forward matrix, float64, same operation order.
__device__ __forceinline__ void invert_affine(const double* __restrict__ Min, double* M) {
    double m0 = Min[0], m1 = Min[1], m2 = Min[2], m3 = Min[3], m4 = Min[4], m5 = Min[5];
    double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
    D = (D != 0.0) ? __ddiv_rn(1.0, D) : 0.0;
    double A11 = __dmul_rn(m4, D), A22 = __dmul_rn(m0, D);
    m0 = A11;
    m1 = __dmul_rn(m1, -D);
    m3 = __dmul_rn(m3, -D);
    m4 = A22;
    double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
    double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
    M[0] = m0; M[1] = m1; M[2] = b1; M[3] = m3; M[4] = m4; M[5] = b2;
}

struct WarpArgs {
    const uint8_t* src_base;
    const int64_t* src_off;
    const int32_t* src_h;
    const int32_t* src_w;
    const int64_t* src_pitch;
    const uint8_t* flip;
    const double* M;
    uint8_t* dst_u8;
    void* dst_norm;
    const float* lut;
    int dw, dh, norm_dtype;
};

// One CTA = WARP_THREADS groups of PX consecutive destination pixels of one sample.
// smem: adelta[dw], bdelta[dw] (OpenCV precomputes the same per-column tables),
// 3x256 normalisation LUT.
template <int PX>
__global__ void __launch_bounds__(WARP_THREADS) warp_affine_kernel(WarpArgs a) {
    extern __shared__ int smem_i[];
    int* adelta = smem_i;
    int* bdelta = smem_i + a.dw;
    float* lut = reinterpret_cast<float*>(smem_i + 2 * a.dw);
    __shared__ double Minv[6];

    const int b = blockIdx.y;
    if (threadIdx.x == 0) invert_affine(a.M + 6 * b, Minv);
    if (a.dst_norm)
        for (int i = threadIdx.x; i < 768; i += WARP_THREADS) lut[i] = a.lut[i];
    __syncthreads();
    {
        const double m0 = Minv[0], m3 = Minv[3];
        for (int x = threadIdx.x; x < a.dw; x += WARP_THREADS) {
            // saturate_cast<int>(M[0]*x*AB_SCALE): ((M0*x)*1024), rint
            adelta[x] = __double2int_rn(__dmul_rn(__dmul_rn(m0, (double)x), 1024.0));
            bdelta[x] = __double2int_rn(__dmul_rn(__dmul_rn(m3, (double)x), 1024.0));
        }
    }
    __syncthreads();

    const int groups_per_row = a.dw / PX;
    const int g = blockIdx.x * WARP_THREADS + threadIdx.x;
    if (g >= groups_per_row * a.dh) return;
    const int y = g / groups_per_row;
    const int x0 = (g - y * groups_per_row) * PX;

    const int H = a.src_h[b], W = a.src_w[b];
    const int64_t pitch = a.src_pitch[b];
    const uint8_t* __restrict__ src = a.src_base + a.src_off[b];
    const bool flip = a.flip && a.flip[b];

    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Minv[1], (double)y), Minv[2]), 1024.0)) + ROUND_DELTA;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Minv[4], (double)y), Minv[5]), 1024.0)) + ROUND_DELTA;

    uint8_t px[PX][3];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        const int X = (X0 + adelta[x0 + i]) >> (AB_BITS - INTER_BITS);
        const int Y = (Y0 + bdelta[x0 + i]) >> (AB_BITS - INTER_BITS);
        int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
        sx = max(-32768, min(32767, sx));
        sy = max(-32768, min(32767, sy));
        const int fx = X & 31, fy = Y & 31;
        const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
        const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
        int acc0 = 0, acc1 = 0, acc2 = 0;
        const bool y0ok = (unsigned)sy < (unsigned)H, y1ok = (unsigned)(sy + 1) < (unsigned)H;
        const bool x0ok = (unsigned)sx < (unsigned)W, x1ok = (unsigned)(sx + 1) < (unsigned)W;
        const int cx0 = flip ? (W - 1 - sx) : sx;
        const int cx1 = flip ? (W - 2 - sx) : (sx + 1);
        const uint8_t* r0 = src + (int64_t)sy * pitch;
        const uint8_t* r1 = r0 + pitch;
        if (y0ok && x0ok) {
            const uint8_t* p = r0 + 3 * cx0;
            acc0 += __ldg(p) * w00; acc1 += __ldg(p + 1) * w00; acc2 += __ldg(p + 2) * w00;
        }
        if (y0ok && x1ok) {
            const uint8_t* p = r0 + 3 * cx1;
            acc0 += __ldg(p) * w01; acc1 += __ldg(p + 1) * w01; acc2 += __ldg(p + 2) * w01;
        }
        if (y1ok && x0ok) {
            const uint8_t* p = r1 + 3 * cx0;
            acc0 += __ldg(p) * w10; acc1 += __ldg(p + 1) * w10; acc2 += __ldg(p + 2) * w10;
        }
        if (y1ok && x1ok) {
            const uint8_t* p = r1 + 3 * cx1;
            acc0 += __ldg(p) * w11; acc1 += __ldg(p + 1) * w11; acc2 += __ldg(p + 2) * w11;
        }
        px[i][0] = (uint8_t)((acc0 + (1 << 14)) >> 15);
        px[i][1] = (uint8_t)((acc1 + (1 << 14)) >> 15);
        px[i][2] = (uint8_t)((acc2 + (1 << 14)) >> 15);
    }

    const int64_t pix = ((int64_t)b * a.dh + y) * a.dw + x0;
    if (a.dst_u8) {
        if (PX == 4) {
            uint32_t w0 = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[1][0] << 24);
            uint32_t w1 = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | (px[2][1] << 24);
            uint32_t w2 = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | (px[3][2] << 24);
            uint32_t* o = reinterpret_cast<uint32_t*>(a.dst_u8 + pix * 3);
            o[0] = w0; o[1] = w1; o[2] = w2;
        } else {
#pragma unroll
            for (int i = 0; i < PX; ++i)
                for (int c = 0; c < 3; ++c) a.dst_u8[(pix + i) * 3 + c] = px[i][c];
        }
    }
    if (a.dst_norm) {
        const int64_t plane = (int64_t)a.dh * a.dw;
        const int64_t o = (int64_t)b * 3 * plane + (int64_t)y * a.dw + x0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[PX];
#pragma unroll
            for (int i = 0; i < PX; ++i) v[i] = lut[c * 256 + px[i][c]];
            if (a.norm_dtype == ADVMIX_F32) {
                float* d = reinterpret_cast<float*>(a.dst_norm) + o + c * plane;
                if (PX == 4) *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
                else
                    for (int i = 0; i < PX; ++i) d[i] = v[i];
            } else {
                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(a.dst_norm) + o + c * plane;
                if (PX == 4) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
                    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                    *reinterpret_cast<uint2*>(d) = pk;
                } else
                    for (int i = 0; i < PX; ++i) d[i] = __float2bfloat16_rn(v[i]);
            }
        }
    }
}

// ---- get_affine_transform (transforms.py:69-101), batched ---------------------------
__global__ void affine_matrices_kernel(const float* __restrict__ center, const double* __restrict__ scale,
                                       const double* __restrict__ rot, double* __restrict__ M, int B,
                                       int out_w, int out_h, int scale_f32) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // float32 / float64 promotions follow numpy on the reference's expressions.
    const float cx = center[2 * b], cy = center[2 * b + 1];
    // scale_tmp = scale * 200.0 ; src_w * -0.5 : in the dtype numpy gives `scale`
    // (float32 under numpy<2 value-based casting, float64 under NEP 50 after `s * np.clip(...)`)
    double half;
    if (scale_f32) half = (double)__fmul_rn(__fmul_rn((float)scale[2 * b], 200.0f), -0.5f);
    else half = __dmul_rn(__dmul_rn(scale[2 * b], 200.0), -0.5);
    const double rot_rad = __ddiv_rn(__dmul_rn(3.141592653589793, rot[b]), 180.0);
    const double sn = sin(rot_rad), cs = cos(rot_rad);
    const double dirx = __dsub_rn(__dmul_rn(0.0, cs), __dmul_rn(half, sn));
    const double diry = __dadd_rn(__dmul_rn(0.0, sn), __dmul_rn(half, cs));
    float s[3][2], d[3][2];
    s[0][0] = cx; s[0][1] = cy;                                    // center + scale_tmp*shift(0)
    s[1][0] = (float)__dadd_rn(__dadd_rn((double)cx, dirx), 0.0);
    s[1][1] = (float)__dadd_rn(__dadd_rn((double)cy, diry), 0.0);
    d[0][0] = (float)(out_w * 0.5); d[0][1] = (float)(out_h * 0.5);
    const float ddy = (float)(out_w * -0.5);
    d[1][0] = (float)__dadd_rn(out_w * 0.5, 0.0);
    d[1][1] = (float)__dadd_rn(out_h * 0.5, (double)ddy);
    // get_3rd_point(a,b) = b + (-(a-b).y, (a-b).x), float32
    {
        float dx = __fsub_rn(s[0][0], s[1][0]), dy = __fsub_rn(s[0][1], s[1][1]);
        s[2][0] = __fadd_rn(s[1][0], -dy); s[2][1] = __fadd_rn(s[1][1], dx);
        dx = __fsub_rn(d[0][0], d[1][0]); dy = __fsub_rn(d[0][1], d[1][1]);
        d[2][0] = __fadd_rn(d[1][0], -dy); d[2][1] = __fadd_rn(d[1][1], dx);
    }
    // closed-form solve M*[p,1] = q  (float64)
    const double p0x = s[0][0], p0y = s[0][1];
    const double ax = (double)s[1][0] - p0x, ay = (double)s[1][1] - p0y;
    const double bx = (double)s[2][0] - p0x, by = (double)s[2][1] - p0y;
    const double det = ax * by - ay * bx;
    const double inv = det != 0.0 ? 1.0 / det : 0.0;
    double* m = M + 6 * b;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double q0 = d[0][r], u = (double)d[1][r] - q0, v = (double)d[2][r] - q0;
        const double m0 = (u * by - v * ay) * inv;
        const double m1 = (v * ax - u * bx) * inv;
        m[3 * r + 0] = m0;
        m[3 * r + 1] = m1;
        m[3 * r + 2] = q0 - m0 * p0x - m1 * p0y;
    }
}

// ---- fliplr_joints + affine_transform ----------------------------------------------
__global__ void joints_kernel(const double* __restrict__ jin, const double* __restrict__ vin,
                              const uint8_t* __restrict__ flip, const int32_t* __restrict__ src_w,
                              const int32_t* __restrict__ perm, const double* __restrict__ M,
                              double* __restrict__ jout, double* __restrict__ vout, int B, int J) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * J) return;
    const int b = t / J, j = t - b * J;
    double x, y, z, v0, v1, v2;
    if (flip && flip[b]) {
        const int s = perm ? perm[j] : j;
        const double* p = jin + ((int64_t)b * J + s) * 3;
        const double* q = vin + ((int64_t)b * J + s) * 3;
        v0 = q[0]; v1 = q[1]; v2 = q[2];
        // joints[:,0] = width - joints[:,0] - 1 ; then joints*joints_vis
        x = __dmul_rn(__dsub_rn(__dsub_rn((double)src_w[b], p[0]), 1.0), v0);
        y = __dmul_rn(p[1], v1);
        z = __dmul_rn(p[2], v2);
    } else {
        const double* p = jin + (int64_t)t * 3;
        const double* q = vin + (int64_t)t * 3;
        x = p[0]; y = p[1]; z = p[2];
        v0 = q[0]; v1 = q[1]; v2 = q[2];
    }
    if (v0 > 0.0) {
        const double* m = M + 6 * b;
        const double nx = fma(m[0], x, fma(m[1], y, m[2]));
        const double ny = fma(m[3], x, fma(m[4], y, m[5]));
        x = nx; y = ny;
    }
    double* o = jout + (int64_t)t * 3;
    o[0] = x; o[1] = y; o[2] = z;
    double* w = vout + (int64_t)t * 3;
    w[0] = v0; w[1] = v1; w[2] = v2;
}

// ---- ToTensor + Normalize ------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_kernel(const uint8_t* __restrict__ in, void* __restrict__ out,
                                                        const float* __restrict__ lut_g, int64_t groups,
                                                        int64_t plane, int dtype) {
    __shared__ float lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = g * 4;
        const int64_t b = pix / plane, r = pix - b * plane;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(in + pix * 3);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        uint8_t v[4][3] = {{(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16)},
                           {(uint8_t)(w0 >> 24), (uint8_t)w1, (uint8_t)(w1 >> 8)},
                           {(uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24), (uint8_t)w2},
                           {(uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)}};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f0 = lut[c * 256 + v[0][c]], f1 = lut[c * 256 + v[1][c]];
            const float f2 = lut[c * 256 + v[2][c]], f3 = lut[c * 256 + v[3][c]];
            const int64_t o = (b * 3 + c) * plane + r;
            if (dtype == ADVMIX_F32) {
                st_stream_f4(reinterpret_cast<float*>(out) + o, make_float4(f0, f1, f2, f3));
            } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(f0, f1), hi = __floats2bfloat162_rn(f2, f3);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + o) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
        }
    }
}

__global__ void normalize_kernel_scalar(const uint8_t* __restrict__ in, void* __restrict__ out,
                                        const float* __restrict__ lut, int64_t npix, int64_t plane, int dtype) {
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
         pix += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = pix / plane, r = pix - b * plane;
        for (int c = 0; c < 3; ++c) {
            const float f = lut[c * 256 + in[pix * 3 + c]];
            const int64_t o = (b * 3 + c) * plane + r;
            if (dtype == ADVMIX_F32) reinterpret_cast<float*>(out)[o] = f;
            else reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(f);
        }
    }
}

}  // namespace advmix

using namespace advmix;

extern "C" {

int advmix_warp_affine_u8c3(const uint8_t* src_base, const int64_t* src_off, const int32_t* src_h,
                            const int32_t* src_w, const int64_t* src_pitch, const uint8_t* flip_lr,
                            const double* M_fwd, uint8_t* dst_u8, void* dst_norm, const float* norm_lut,
                            int B, int dw, int dh, int norm_dtype, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && dw > 0 && dh > 0, "warp_affine: bad shape B=%d dw=%d dh=%d", B, dw, dh);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(src_off && src_h && src_w && src_pitch && M_fwd, "warp_affine: null argument");
    ADVMIX_REQUIRE(dst_u8 || dst_norm, "warp_affine: no output requested");
    ADVMIX_REQUIRE(!dst_norm || norm_lut, "warp_affine: dst_norm needs norm_lut");
    ADVMIX_REQUIRE(norm_dtype == ADVMIX_F32 || norm_dtype == ADVMIX_BF16, "warp_affine: bad dtype %d", norm_dtype);
    ADVMIX_REQUIRE(dw <= 8192 && B <= 65535, "warp_affine: dw<=8192, B<=65535 per call");
    WarpArgs a{src_base, src_off, src_h, src_w, src_pitch, flip_lr, M_fwd, dst_u8, dst_norm, norm_lut, dw, dh, norm_dtype};
    const size_t smem = (size_t)2 * dw * sizeof(int) + 768 * sizeof(float);
    if (dw % 4 == 0) {
        dim3 grid(ceil_div((long long)(dw / 4) * dh, WARP_THREADS), B);
        warp_affine_kernel<4><<<grid, WARP_THREADS, smem, as_stream(stream)>>>(a);
    } else {
        dim3 grid(ceil_div((long long)dw * dh, WARP_THREADS), B);
        warp_affine_kernel<1><<<grid, WARP_THREADS, smem, as_stream(stream)>>>(a);
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_affine_matrices(const float* center, const double* scale, int scale_is_f32, const double* rot_deg,
                           double* M_fwd, int B, int out_w, int out_h, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && out_w > 0 && out_h > 0, "affine_matrices: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(center && scale && rot_deg && M_fwd, "affine_matrices: null argument");
    affine_matrices_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(center, scale, rot_deg, M_fwd, B, out_w, out_h, scale_is_f32);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_joints_flip_affine(const double* joints_in, const double* vis_in, const uint8_t* flip_lr,
                              const int32_t* src_w, const int32_t* flip_perm, const double* M_fwd,
                              double* joints_out, double* vis_out, int B, int J, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0, "joints_flip_affine: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(joints_in && vis_in && M_fwd && joints_out && vis_out, "joints_flip_affine: null argument");
    ADVMIX_REQUIRE(!flip_lr || src_w, "joints_flip_affine: flip needs src_w");
    joints_kernel<<<ceil_div((long long)B * J, 128), 128, 0, as_stream(stream)>>>(joints_in, vis_in, flip_lr, src_w, flip_perm,
                                                                                  M_fwd, joints_out, vis_out, B, J);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_normalize_u8c3(const uint8_t* in, void* out, const float* norm_lut, int B, int H, int W,
                          int norm_dtype, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && H > 0 && W > 0, "normalize: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(in && out && norm_lut, "normalize: null argument");
    ADVMIX_REQUIRE(norm_dtype == ADVMIX_F32 || norm_dtype == ADVMIX_BF16, "normalize: bad dtype %d", norm_dtype);
    const int64_t plane = (int64_t)H * W, npix = plane * B;
    if (plane % 4 == 0) {
        const int64_t groups = npix / 4;
        int blocks = (int)std::min<int64_t>((groups + 255) / 256, (int64_t)sm_count() * 16);
        normalize_kernel<<<blocks, 256, 0, as_stream(stream)>>>(in, out, norm_lut, groups, plane, norm_dtype);
    } else {
        int blocks = (int)std::min<int64_t>((npix + 255) / 256, (int64_t)sm_count() * 16);
        normalize_kernel_scalar<<<blocks, 256, 0, as_stream(stream)>>>(in, out, norm_lut, npix, plane, norm_dtype);
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"
