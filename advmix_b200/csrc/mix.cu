// Row a3: AdvMix per-pixel convex mix, forward + backward (lib/core/function.py:137-146,
// :158-164).  Pure streaming: (K*C + K) reads + C writes per pixel, 128-bit accesses.
#include "chains.cuh"

namespace advmix {

constexpr int MIX_MAXK = 8;
constexpr int MIX_THREADS = 256;

struct MixPtrs {
    const void* x[MIX_MAXK];
};

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    __device__ static float4 load(const float* p) { return ld_stream_f4(p); }
    __device__ static void store(float* p, float4 v) { st_stream_f4(p, v); }
};
template <> struct Vec4<__nv_bfloat16> {
    __device__ static float4 load(const __nv_bfloat16* p) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&r.x), hi = *reinterpret_cast<__nv_bfloat162*>(&r.y);
        float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    __device__ static void store(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 r = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(r.x), "r"(r.y) : "memory");
    }
};

__device__ __forceinline__ float& f4(float4& v, int i) { return (&v.x)[i]; }

// softmax over K for four pixels at once: softmax4_sfu (chains.cuh) - the same function the fused chain + mix kernels use
// (chainmix.cu), so both paths give identical bits
template <int K>
__device__ __forceinline__ void softmax4(float4 (&w)[K]) { softmax4_sfu<K>(w); }

// grid-stride over float4 groups of the [B][HW] pixel space.
template <typename T, int K>
__global__ void __launch_bounds__(MIX_THREADS)
mix_fwd_kernel(MixPtrs ptrs, const float* __restrict__ wl, int apply_softmax, T* __restrict__ out,
               float* __restrict__ w_out, int64_t groups, int64_t hw4, int C) {
    for (int64_t g = (int64_t)blockIdx.x * MIX_THREADS + threadIdx.x; g < groups;
         g += (int64_t)gridDim.x * MIX_THREADS) {
        const int64_t b = g / hw4, r = (g - b * hw4) << 2;
        const int64_t hw = hw4 << 2;
        float4 w[K];
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = ld_stream_f4(wl + (b * K + k) * hw + r);
        if (apply_softmax) {
            softmax4<K>(w);
            if (w_out) {
#pragma unroll
                for (int k = 0; k < K; ++k) st_stream_f4(w_out + (b * K + k) * hw + r, w[k]);
            }
        }
        for (int c = 0; c < C; ++c) {
            const int64_t o = (b * C + c) * hw + r;
            float4 xv[K];
#pragma unroll
            for (int k = 0; k < K; ++k) xv[k] = Vec4<T>::load(reinterpret_cast<const T*>(ptrs.x[k]) + o);
            float4 acc;
            // tmp = x0*w0 ; tmp += xk*wk   (separate mul and add, as eager torch does)
            acc.x = __fmul_rn(xv[0].x, w[0].x); acc.y = __fmul_rn(xv[0].y, w[0].y);
            acc.z = __fmul_rn(xv[0].z, w[0].z); acc.w = __fmul_rn(xv[0].w, w[0].w);
#pragma unroll
            for (int k = 1; k < K; ++k) {
                acc.x = __fadd_rn(acc.x, __fmul_rn(xv[k].x, w[k].x));
                acc.y = __fadd_rn(acc.y, __fmul_rn(xv[k].y, w[k].y));
                acc.z = __fadd_rn(acc.z, __fmul_rn(xv[k].z, w[k].z));
                acc.w = __fadd_rn(acc.w, __fmul_rn(xv[k].w, w[k].w));
            }
            Vec4<T>::store(out + o, acc);
        }
    }
}

// grad_w[k] = sum_c g_c * x_{k,c};  through softmax: gl_k = w_k * (gw_k - sum_j w_j gw_j)
template <typename T, int K>
__global__ void __launch_bounds__(MIX_THREADS)
mix_bwd_kernel(MixPtrs ptrs, const float* __restrict__ w, const T* __restrict__ go,
               float* __restrict__ gw_out, int through_softmax, int64_t groups, int64_t hw4, int C) {
    for (int64_t g = (int64_t)blockIdx.x * MIX_THREADS + threadIdx.x; g < groups;
         g += (int64_t)gridDim.x * MIX_THREADS) {
        const int64_t b = g / hw4, r = (g - b * hw4) << 2;
        const int64_t hw = hw4 << 2;
        float4 gw[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < C; ++c) {
            const int64_t o = (b * C + c) * hw + r;
            const float4 gv = Vec4<T>::load(go + o);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4 xv = Vec4<T>::load(reinterpret_cast<const T*>(ptrs.x[k]) + o);
                gw[k].x = fmaf(gv.x, xv.x, gw[k].x); gw[k].y = fmaf(gv.y, xv.y, gw[k].y);
                gw[k].z = fmaf(gv.z, xv.z, gw[k].z); gw[k].w = fmaf(gv.w, xv.w, gw[k].w);
            }
        }
        if (through_softmax) {
            float4 wv[K];
            float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                wv[k] = ld_stream_f4(w + (b * K + k) * hw + r);
                dot.x = fmaf(wv[k].x, gw[k].x, dot.x); dot.y = fmaf(wv[k].y, gw[k].y, dot.y);
                dot.z = fmaf(wv[k].z, gw[k].z, dot.z); dot.w = fmaf(wv[k].w, gw[k].w, dot.w);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                gw[k].x = wv[k].x * (gw[k].x - dot.x); gw[k].y = wv[k].y * (gw[k].y - dot.y);
                gw[k].z = wv[k].z * (gw[k].z - dot.z); gw[k].w = wv[k].w * (gw[k].w - dot.w);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) st_stream_f4(gw_out + (b * K + k) * hw + r, gw[k]);
    }
}

template <typename T, int K>
static int launch_fwd(const MixPtrs& p, const float* wl, int sm, void* out, float* w_out, int64_t groups,
                      int64_t hw4, int C, cudaStream_t s) {
    // one 4-pixel group per thread up to 32 CTAs per SM, grid-stride beyond (keeps small batches balanced)
    int blocks = (int)std::min<int64_t>((groups + MIX_THREADS - 1) / MIX_THREADS, (int64_t)sm_count() * 32);
    mix_fwd_kernel<T, K><<<blocks, MIX_THREADS, 0, s>>>(p, wl, sm, reinterpret_cast<T*>(out), w_out, groups, hw4, C);
    return 0;
}
template <typename T, int K>
static int launch_bwd(const MixPtrs& p, const float* w, const void* go, float* gw, int ts, int64_t groups,
                      int64_t hw4, int C, cudaStream_t s) {
    int blocks = (int)std::min<int64_t>((groups + MIX_THREADS - 1) / MIX_THREADS, (int64_t)sm_count() * 32);
    mix_bwd_kernel<T, K><<<blocks, MIX_THREADS, 0, s>>>(p, w, reinterpret_cast<const T*>(go), gw, ts, groups, hw4, C);
    return 0;
}

#define MIX_DISPATCH_K(FN, T, ...)                         \
    switch (K) {                                           \
        case 1: FN<T, 1>(__VA_ARGS__); break;              \
        case 2: FN<T, 2>(__VA_ARGS__); break;              \
        case 3: FN<T, 3>(__VA_ARGS__); break;              \
        case 4: FN<T, 4>(__VA_ARGS__); break;              \
        case 5: FN<T, 5>(__VA_ARGS__); break;              \
        case 6: FN<T, 6>(__VA_ARGS__); break;              \
        case 7: FN<T, 7>(__VA_ARGS__); break;              \
        default: FN<T, 8>(__VA_ARGS__); break;             \
    }

}  // namespace advmix

using namespace advmix;

static int mix_check(const void* const* x_h, int B, int K, int C, int H, int W, int dtype) {
    ADVMIX_REQUIRE(B >= 0 && K >= 1 && K <= MIX_MAXK && C >= 1 && H > 0 && W > 0, "mix: bad shape B=%d K=%d C=%d H=%d W=%d", B, K, C, H, W);
    ADVMIX_REQUIRE(dtype == ADVMIX_F32 || dtype == ADVMIX_BF16, "mix: bad dtype %d", dtype);
    ADVMIX_REQUIRE(((int64_t)H * W) % 4 == 0, "mix: H*W must be a multiple of 4 (got %dx%d)", H, W);
    ADVMIX_REQUIRE(x_h != nullptr, "mix: null x_h");
    for (int k = 0; k < K; ++k) ADVMIX_REQUIRE(x_h[k] != nullptr, "mix: x_h[%d] is null", k);
    return ADVMIX_OK;
}

extern "C" {

int advmix_mix_fwd(const void* const* x_h, const float* w_or_logits, int apply_softmax, void* out, float* w_out,
                   int B, int K, int C, int H, int W, int dtype, advmix_stream_t stream) {
    int rc = mix_check(x_h, B, K, C, H, W, dtype);
    if (rc) return rc;
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(w_or_logits && out, "mix_fwd: null argument");
    MixPtrs p{};
    for (int k = 0; k < K; ++k) p.x[k] = x_h[k];
    const int64_t hw4 = (int64_t)H * W / 4, groups = hw4 * B;
    cudaStream_t s = as_stream(stream);
    if (dtype == ADVMIX_F32) {
        MIX_DISPATCH_K(launch_fwd, float, p, w_or_logits, apply_softmax, out, w_out, groups, hw4, C, s)
    } else {
        MIX_DISPATCH_K(launch_fwd, __nv_bfloat16, p, w_or_logits, apply_softmax, out, w_out, groups, hw4, C, s)
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_mix_bwd(const void* const* x_h, const float* w, const void* grad_out, float* grad_w, int through_softmax,
                   int B, int K, int C, int H, int W, int dtype, advmix_stream_t stream) {
    int rc = mix_check(x_h, B, K, C, H, W, dtype);
    if (rc) return rc;
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(grad_out && grad_w, "mix_bwd: null argument");
    ADVMIX_REQUIRE(!through_softmax || w, "mix_bwd: through_softmax needs the weights");
    MixPtrs p{};
    for (int k = 0; k < K; ++k) p.x[k] = x_h[k];
    const int64_t hw4 = (int64_t)H * W / 4, groups = hw4 * B;
    cudaStream_t s = as_stream(stream);
    if (dtype == ADVMIX_F32) {
        MIX_DISPATCH_K(launch_bwd, float, p, w, grad_out, grad_w, through_softmax, groups, hw4, C, s)
    } else {
        MIX_DISPATCH_K(launch_bwd, __nv_bfloat16, p, w, grad_out, grad_w, through_softmax, groups, hw4, C, s)
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"
