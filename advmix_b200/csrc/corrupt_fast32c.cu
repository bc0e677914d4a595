// ADVMIX_CORRUPT_FAST, third part: a separable float32 Gaussian specialised for uint8 HWC images (glass_blur's two passes,
// gaussian_blur).  The generic register-tiled kernel of stencil_common.cuh spends ~105 instructions per output byte on
// per-element index maps, LUT conversions and single-byte loads / stores (ncu: 76 % issue-active, 3.6 % of DRAM peak);
// here a thread always works on 4 consecutive bytes of a row (one 32-bit word: four independent channel values), the
// vertical pass slides a register window down a column of words, the horizontal pass reads its 4 + 6R floats as 128-bit
// shared-memory loads, and global traffic is whole words: ~26 instructions per output byte.
#include "stencil_common.cuh"

namespace advmix {

constexpr int GU_THREADS = 256, GU_ROWS = 32, GU_TCB = 192, GU_SEG = 8;     // tile: 32 rows x 192 bytes (64 pixels)

// float(byte k of w): ONE PRMT drops the byte into the mantissa of 2^23 (0x4B000000), one FADD removes the 2^23
__device__ __forceinline__ float byte_to_float(uint32_t w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)k)) - 8388608.0f;
}

template <int R>
__global__ void __launch_bounds__(GU_THREADS)
gauss_u8c3_fast_kernel(const uint8_t* __restrict__ in, const int32_t* __restrict__ in_idx, uint8_t* __restrict__ out,
                       const int32_t* __restrict__ out_idx, int H, int W, const double* __restrict__ wts, float top255) {
    constexpr int HB = (3 * R + 3) / 4 * 4;              // halo bytes each side, rounded to whole words
    constexpr int CB = GU_TCB + 2 * HB, CQ = CB / 4;     // staged bytes / words per row
    constexpr int AR = GU_ROWS + 2 * R;
    __shared__ __align__(16) uint32_t raw[AR * CQ];
    __shared__ __align__(16) float mid[GU_ROWS * CB];
    __shared__ float w[R + 1];
    if (threadIdx.x <= R) w[threadIdx.x] = (float)wts[threadIdx.x];
    const int WC = W * 3;
    const int img = blockIdx.z;
    const uint8_t* src = in + (int64_t)(in_idx ? in_idx[img] : img) * H * WC;
    uint8_t* dst = out + (int64_t)(out_idx ? out_idx[img] : img) * H * WC;
    const int x0 = blockIdx.x * GU_TCB, y0 = blockIdx.y * GU_ROWS;
    // ---- stage the raw bytes (rows clamped = scipy 'nearest'; columns clamped per pixel).  All of a thread's loads are issued
    // before the first store: with one dependent load -> store per loop iteration this loop held 59 % of the kernel's stall
    // samples (ncu source view) although it executes 28 % of its instructions.
    {
        constexpr int NI = (AR * CQ + GU_THREADS - 1) / GU_THREADS;
        uint32_t v[NI];
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            const int e = threadIdx.x + k * GU_THREADS;
            v[k] = 0u;
            if (e < AR * CQ) {
                const int ty = e / CQ, q = e - ty * CQ;
                const int gy = clampi(y0 + ty - R, 0, H - 1);
                const int bc = x0 - HB + 4 * q;                  // first byte of this word in the row
                if (bc >= 0 && bc + 3 < WC) v[k] = __ldg(reinterpret_cast<const uint32_t*>(src + (int64_t)gy * WC + bc));
            }
        }
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            const int e = threadIdx.x + k * GU_THREADS;
            if (e < AR * CQ) {
                const int ty = e / CQ, q = e - ty * CQ;
                const int bc = x0 - HB + 4 * q;
                uint32_t w = v[k];
                if (!(bc >= 0 && bc + 3 < WC)) {                 // a word that straddles the left / right image border
                    const uint8_t* row = src + (int64_t)clampi(y0 + ty - R, 0, H - 1) * WC;
                    w = 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int b = bc + j;
                        int px = b >= 0 ? b / 3 : -((-b + 2) / 3);
                        const int ch = b - px * 3;
                        px = clampi(px, 0, W - 1);
                        w |= (uint32_t)row[px * 3 + ch] << (8 * j);
                    }
                }
                raw[e] = w;
            }
        }
    }
    __syncthreads();
    // ---- vertical pass: item = (column word, segment of GU_SEG output rows); window of 2R+1 rows x 4 channel values
    for (int item = threadIdx.x; item < CQ * (GU_ROWS / GU_SEG); item += GU_THREADS) {
        const int seg = item / CQ, q = item - seg * CQ;
        const uint32_t* col = raw + (seg * GU_SEG) * CQ + q;             // tile row of output row seg*SEG at offset R
        float win[2 * R + 1][4];
#pragma unroll
        for (int m = 0; m < 2 * R; ++m) {
            const uint32_t v = col[m * CQ];
#pragma unroll
            for (int k = 0; k < 4; ++k) win[m + 1][k] = byte_to_float(v, k);
        }
#pragma unroll
        for (int o = 0; o < GU_SEG; ++o) {
#pragma unroll
            for (int m = 0; m < 2 * R; ++m)
#pragma unroll
                for (int k = 0; k < 4; ++k) win[m][k] = win[m + 1][k];
            const uint32_t v = col[(o + 2 * R) * CQ];
#pragma unroll
            for (int k = 0; k < 4; ++k) win[2 * R][k] = byte_to_float(v, k);
            float4 r;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = win[R][k] * w[0];
#pragma unroll
                for (int j = R; j >= 1; --j) t = fmaf(win[R - j][k] + win[R + j][k], w[j], t);
                (&r.x)[k] = t;
            }
            *reinterpret_cast<float4*>(mid + (seg * GU_SEG + o) * CB + 4 * q) = r;
        }
    }
    __syncthreads();
    // ---- horizontal pass + store: item = (row, output word); channel stride 3 bytes
    constexpr int OQ = GU_TCB / 4, NV = (4 + 6 * R + 3) / 4 + 1;
    for (int item = threadIdx.x; item < GU_ROWS * OQ; item += GU_THREADS) {
        const int ty = item / OQ, q = item - ty * OQ;
        const int y = y0 + ty, bx = x0 + 4 * q;
        if (y >= H || bx >= WC) continue;
        // outputs at tile byte HB + 4q + k need mid[HB + 4q + k - 3R .. + 3R]; first needed float f0 = HB - 3R + 4q (>= 0)
        constexpr int OFF = HB - 3 * R;                   // 0..3: misalignment of the window start inside its first word
        const float4* mp = reinterpret_cast<const float4*>(mid + ty * CB + 4 * q);
        float v[4 * NV];
#pragma unroll
        for (int m = 0; m < NV; ++m) {
            if (4 * (q + m) < CB) {
                const float4 t = mp[m];
                v[4 * m] = t.x; v[4 * m + 1] = t.y; v[4 * m + 2] = t.z; v[4 * m + 3] = t.w;
            } else {
                v[4 * m] = v[4 * m + 1] = v[4 * m + 2] = v[4 * m + 3] = 0.f;
            }
        }
        uint32_t o = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = OFF + 3 * R + k;                // index of the centre value in v[]
            float t = v[c] * w[0];
#pragma unroll
            for (int j = R; j >= 1; --j) t = fmaf(v[c - 3 * j] + v[c + 3 * j], w[j], t);
            // values are in units of 1/255 (the byte domain).  Saturated region: what the float64 tap order gives there
            // (255 or 254); the un-clipped first pass of glass_blur can only exceed 255 by rounding, which uint8() maps to 255
            t = fminf(t, 255.0f);
            if (t > 254.9997f) t = top255;
            o |= (uint32_t)__float2int_rz(fmaxf(t, 0.f)) << (8 * k);
        }
        if (bx + 3 < WC) {
            *reinterpret_cast<uint32_t*>(dst + (int64_t)y * WC + bx) = o;
        } else {
            for (int k = 0; bx + k < WC; ++k) dst[(int64_t)y * WC + bx + k] = (uint8_t)(o >> (8 * k));
        }
    }
}

template <int R>
static int launch_gauss_u8_r(const uint8_t* in, const int32_t* in_idx, uint8_t* out, const int32_t* out_idx, int n, int H, int W,
                             const double* d_w, float top255, cudaStream_t s) {
    dim3 grid(ceil_div(W * 3, GU_TCB), ceil_div(H, GU_ROWS), n);
    gauss_u8c3_fast_kernel<R><<<grid, GU_THREADS, 0, s>>>(in, in_idx, out, out_idx, H, W, d_w, top255);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

// returns -1 when the shape / radius has no specialised kernel
int launch_gauss_u8_fast(const uint8_t* in, const int32_t* in_idx, uint8_t* out, const int32_t* out_idx, int n, int H, int W,
                         int radius, const double* d_w, float top255, cudaStream_t s) {
    if (n > 65535 || (W & 3) != 0 || ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3) != 0) return -1;
    switch (radius) {
        case 3: return launch_gauss_u8_r<3>(in, in_idx, out, out_idx, n, H, W, d_w, top255, s);
        case 4: return launch_gauss_u8_r<4>(in, in_idx, out, out_idx, n, H, W, d_w, top255, s);
        case 6: return launch_gauss_u8_r<6>(in, in_idx, out, out_idx, n, H, W, d_w, top255, s);
        case 8: return launch_gauss_u8_r<8>(in, in_idx, out, out_idx, n, H, W, d_w, top255, s);
        default: return -1;
    }
}

}  // namespace advmix
