// Row N1 / SURVEY 7 step 7: the fused chain + mix kernels.
//
// The reference materialises K = 3 normalised float32 chain tensors per sample in CPU workers
// (lib/dataset/JointsDataset.py:124-131), ships them to the GPU and mixes them with five eager kernels
// (lib/core/function.py:137-146): 2 949 120 B of HBM traffic per 256x192 sample for the mix alone.  Here the mix
// reads the uint8 crop (147 456 B) and the generator's logits, RECOMPUTES the three chains
//     clean    = Normalize(ToTensor(crop))                                  LUT            (tools/train.py:116-126)
//     autoaug  = Normalize(ToTensor(post[sharpen?(pre[crop])]))             per-image plan (advaug.py:10-107)
//     gridmask = clean * mask(d, st_h, st_w)                                closed form    (advaug.py:111-170)
// in registers and writes `tmp` once: 147 456 + 589 824 + 589 824 B with float32 logits / output, 737 280 B with
// bfloat16 ones.  The backward pass (function.py:158-164: the G step back-propagates through `tmp` into the mixing
// weights) recomputes the chains and the softmax the same way from the crop, the logits and grad_out.
// Arithmetic: the same float32 mul-then-add order as the reference expression, so `tmp` is bit-identical to
// advmix_mix_fwd on the materialised chains (and to the reference for given weights).
//
// advmix_mix_u8_fwd / _bwd are the general form for the target workload (chains drawn from the 15x5 corruption set,
// BASELINE configs[2]): K uint8 HWC chain images, normalised in registers.
#include "chains.cuh"

#include <cstdlib>

namespace advmix {

constexpr int CM_THREADS = 256;
constexpr int CM_K = 3;

struct ChainMixArgs {
    const uint8_t* crop;          // [B][H][W][3]
    const AutoPlan* plans;        // [B] (nullable: chain 1 == clean)
    const int32_t* gm;            // [B][4] apply, d, st_h, st_w (nullable: chain 2 == clean)
    const float* lut;             // [3][256]
    const void* logits;           // [B][3][H][W] float32 | bfloat16 (weights when !softmax)
    const void* grad_out;         // backward only: [B][3][H][W] in the output dtype
    void* out;                    // forward: tmp [B][3][H][W]; backward: grad_logits (float32)
    float* w_out;                 // forward, nullable
    int H, W, softmax;
    uint32_t inv_wq;              // ceil(2^32 / (W / 4)): g / wq == __umulhi(g, inv_wq) for g * wq < 2^32
};

template <typename T> __device__ __forceinline__ float4 cm_load4(const T* p);
template <> __device__ __forceinline__ float4 cm_load4<float>(const float* p) { return ld_stream_f4(p); }
template <> __device__ __forceinline__ float4 cm_load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void cm_store4(T* p, float4 v);
template <> __device__ __forceinline__ void cm_store4<float>(float* p, float4 v) { st_stream_f4(p, v); }
template <> __device__ __forceinline__ void cm_store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(*reinterpret_cast<uint32_t*>(&lo)),
                 "r"(*reinterpret_cast<uint32_t*>(&hi)) : "memory");
}

__device__ __forceinline__ float& cf4(float4& v, int i) { return (&v.x)[i]; }

// softmax over K chains for four pixels (softmax4_sfu, chains.cuh): the same function mix.cu uses, so the fused and the
// materialised paths give identical bits
template <int K>
__device__ __forceinline__ void softmax4_k(float4 (&w)[K]) { softmax4_sfu<K>(w); }
__device__ __forceinline__ void cm_softmax(float4 (&w)[CM_K]) { softmax4_k<CM_K>(w); }

// Per-CTA tables of one image: tab[c][v] = {Normalize(v), Normalize(autoaug(v))} (one 8-byte read per value and channel;
// autoaug folded in when the sub-policy has no sharpness stage), the gridmask row / column line flags, and for the
// rare sharpness sub-policy (p = 1/12 * 0.4) the raw pre / post byte tables.
struct ChainTables {
    float2 tab[768];
    uint8_t pre[768], post[768];
    uint8_t rowline[1024], colline[1024];
    float factor;
    int stencil, gm_on;
};

__device__ __forceinline__ void chain_tables_fill(ChainTables& T, const ChainMixArgs& a, int b) {
    // phase 1: everything that comes from global memory, all loads independent (one round trip)
    const AutoPlan* plan = a.plans ? a.plans + b : nullptr;
    const bool stencil = plan && plan->stencil != 0;
    for (int i = threadIdx.x; i < 768; i += CM_THREADS) {
        T.tab[i].x = __ldg(a.lut + i);
        T.pre[i] = plan ? plan->pre[i] : (uint8_t)i;
        if (stencil) T.post[i] = plan->post[i];
    }
    int on = 0, d = 2, st_h = 0, st_w = 0;
    if (a.gm) { on = a.gm[4 * b]; d = max(a.gm[4 * b + 1], 2); st_h = a.gm[4 * b + 2]; st_w = a.gm[4 * b + 3]; }
    if (on) {
        // grid_mask_at(y, x) == rowline[y] | colline[x]  (mode = 1: keep the grid lines, zero the cells)
        const GridGeom g = grid_geom(a.H, a.W, d);
        for (int y = threadIdx.x; y < a.H; y += CM_THREADS) {
            const int t = y + g.oy - st_h;
            bool line = false;
            if (t >= 0) { const int i = t / d; line = (i < g.hh / d) && (t - i * d < g.l); }
            T.rowline[y] = line;
        }
        for (int x = threadIdx.x; x < a.W; x += CM_THREADS) {
            const int t = x + g.ox - st_w;
            bool line = false;
            if (t >= 0) { const int i = t / d; line = (i < g.ww / d) && (t - i * d < g.l); }
            T.colline[x] = line;
        }
    }
    if (threadIdx.x == 0) { T.factor = stencil ? plan->factor : 1.f; T.stencil = stencil; T.gm_on = on; }
    __syncthreads();
    // phase 2: fold the autoaug byte map into the normalisation table (no sharpness stage: out = pre[in])
    for (int i = threadIdx.x; i < 768; i += CM_THREADS)
        T.tab[i].y = stencil ? T.tab[i].x : T.tab[(i & ~255) + T.pre[i]].x;
    __syncthreads();
}

// the three chain values of 4 horizontally adjacent pixels, channel c: x[k] for k = clean, autoaug, gridmask
template <bool STENCIL>
__device__ __forceinline__ void chain_values(const ChainTables& T, const ChainMixArgs& a, const uint8_t* __restrict__ img,
                                             const uint32_t (&wds)[3], int c, int y, int x, bool gm_on, const float (&m)[4],
                                             float4 (&xv)[CM_K]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = 3 * i + c;                                   // byte index within the 12-byte group
        const uint32_t v = (wds[e >> 2] >> (8 * (e & 3))) & 255u;
        const float2 t = T.tab[c * 256 + v];
        float x1 = t.y;
        if (STENCIL) {                                             // rare (p = 1/30): PIL SMOOTH needs the 3x3 neighbourhood
            const bool interior = y > 0 && y < a.H - 1 && (x + i) > 0 && (x + i) < a.W - 1;
            const uint8_t s = sharpen_px(img + ((int64_t)y * a.W + x + i) * 3 + c, (int64_t)a.W * 3, T.pre + c * 256, T.factor, interior);
            x1 = T.tab[c * 256 + T.post[c * 256 + s]].x;
        }
        cf4(xv[0], i) = t.x;
        cf4(xv[1], i) = x1;
        cf4(xv[2], i) = gm_on ? __fmul_rn(t.x, m[i]) : t.x;        // img *= mask on the normalised tensor (advaug.py:166)
    }
}

enum { CM_FWD = 0, CM_BWD = 1, CM_EMIT = 2 };

// One loop for the three kernels.  FWD: tmp = sum_k x_k * w_k.  BWD: grad_w[k] = sum_c g_c * x_{k,c}, through softmax
// gl_k = w_k * (gw_k - sum_j w_j gw_j) (same math as mix_bwd_kernel).  EMIT: G_input = cat(inputs, dim=1).
template <int MODE, typename LT, typename OT, bool STENCIL>
__device__ __forceinline__ void chain_loop(const ChainMixArgs& a, const ChainTables& T, int b) {
    const int64_t hw = (int64_t)a.H * a.W;
    const int wq = a.W >> 2, groups = a.H * wq;
    const bool gm_on = T.gm_on != 0;
    const uint8_t* img = a.crop + (int64_t)b * hw * 3;
    const LT* lg = reinterpret_cast<const LT*>(a.logits) + (int64_t)b * CM_K * hw;
    for (int g = blockIdx.x * CM_THREADS + threadIdx.x; g < groups; g += gridDim.x * CM_THREADS) {
        const int y = (int)__umulhi((uint32_t)g, a.inv_wq), x = (g - y * wq) << 2;
        const int64_t r = (int64_t)g << 2;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(img + r * 3);
        const uint32_t wds[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
        float4 w[CM_K], gv[3];
        if (MODE == CM_FWD || (MODE == CM_BWD && a.softmax)) {
#pragma unroll
            for (int k = 0; k < CM_K; ++k) w[k] = cm_load4<LT>(lg + k * hw + r);
        }
        if (MODE == CM_BWD) {
            const OT* go = reinterpret_cast<const OT*>(a.grad_out) + (int64_t)b * 3 * hw;
#pragma unroll
            for (int c = 0; c < 3; ++c) gv[c] = cm_load4<OT>(go + c * hw + r);
        }
        float m[4] = {1.f, 1.f, 1.f, 1.f};
        if (gm_on) {
            const bool rl = T.rowline[y] != 0;
            const uint32_t cl = *reinterpret_cast<const uint32_t*>(T.colline + x);      // x is a multiple of 4
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = (rl || ((cl >> (8 * i)) & 255u)) ? 1.f : 0.f;
        }
        if (MODE != CM_EMIT && a.softmax) cm_softmax(w);
        if (MODE == CM_FWD && a.softmax && a.w_out) {
#pragma unroll
            for (int k = 0; k < CM_K; ++k) st_stream_f4(a.w_out + ((int64_t)b * CM_K + k) * hw + r, w[k]);
        }
        float4 gw[CM_K];
        if (MODE == CM_BWD) {
#pragma unroll
            for (int k = 0; k < CM_K; ++k) gw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4 xv[CM_K];
            chain_values<STENCIL>(T, a, img, wds, c, y, x, gm_on, m, xv);
            if (MODE == CM_FWD) {
                float4 acc;
                // tmp = x0*w0 ; tmp += x1*w1 ; tmp += x2*w2   (function.py:142-144: separate mul and add)
                acc.x = __fmul_rn(xv[0].x, w[0].x); acc.y = __fmul_rn(xv[0].y, w[0].y);
                acc.z = __fmul_rn(xv[0].z, w[0].z); acc.w = __fmul_rn(xv[0].w, w[0].w);
#pragma unroll
                for (int k = 1; k < CM_K; ++k) {
                    acc.x = __fadd_rn(acc.x, __fmul_rn(xv[k].x, w[k].x)); acc.y = __fadd_rn(acc.y, __fmul_rn(xv[k].y, w[k].y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(xv[k].z, w[k].z)); acc.w = __fadd_rn(acc.w, __fmul_rn(xv[k].w, w[k].w));
                }
                cm_store4<OT>(reinterpret_cast<OT*>(a.out) + ((int64_t)b * 3 + c) * hw + r, acc);
            } else if (MODE == CM_BWD) {
#pragma unroll
                for (int k = 0; k < CM_K; ++k) {
                    gw[k].x = fmaf(gv[c].x, xv[k].x, gw[k].x); gw[k].y = fmaf(gv[c].y, xv[k].y, gw[k].y);
                    gw[k].z = fmaf(gv[c].z, xv[k].z, gw[k].z); gw[k].w = fmaf(gv[c].w, xv[k].w, gw[k].w);
                }
            } else {
#pragma unroll
                for (int k = 0; k < CM_K; ++k) cm_store4<OT>(reinterpret_cast<OT*>(a.out) + ((int64_t)b * 3 * CM_K + 3 * k + c) * hw + r, xv[k]);
            }
        }
        if (MODE == CM_BWD) {
            if (a.softmax) {                                     // w holds the recomputed softmax: no saved weights tensor
                float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < CM_K; ++k) {
                    dot.x = fmaf(w[k].x, gw[k].x, dot.x); dot.y = fmaf(w[k].y, gw[k].y, dot.y);
                    dot.z = fmaf(w[k].z, gw[k].z, dot.z); dot.w = fmaf(w[k].w, gw[k].w, dot.w);
                }
#pragma unroll
                for (int k = 0; k < CM_K; ++k) {
                    gw[k].x = w[k].x * (gw[k].x - dot.x); gw[k].y = w[k].y * (gw[k].y - dot.y);
                    gw[k].z = w[k].z * (gw[k].z - dot.z); gw[k].w = w[k].w * (gw[k].w - dot.w);
                }
            }
#pragma unroll
            for (int k = 0; k < CM_K; ++k) st_stream_f4(reinterpret_cast<float*>(a.out) + ((int64_t)b * CM_K + k) * hw + r, gw[k]);
        }
    }
}

template <int MODE, typename LT, typename OT>
__global__ void __launch_bounds__(CM_THREADS, 4)
chain_kernel(ChainMixArgs a) {
    __shared__ ChainTables T;
    const int b = blockIdx.y;
    chain_tables_fill(T, a, b);
    if (T.stencil) chain_loop<MODE, LT, OT, true>(a, T, b);
    else chain_loop<MODE, LT, OT, false>(a, T, b);
}

// ---- general form: K uint8 HWC chain images ----------------------------------------------------------------
constexpr int MU_MAXK = 4;
struct MixU8Args {
    const uint8_t* x[MU_MAXK];    // each [B][H][W][3]
    const float* lut;
    const void* logits;
    const void* grad_out;
    void* out;
    float* w_out;
    int64_t hw;
    int softmax;
};

template <typename LT, typename OT, int K, bool BWD>
__global__ void __launch_bounds__(CM_THREADS)
mix_u8_kernel(MixU8Args a) {
    __shared__ float nl[768];
    for (int i = threadIdx.x; i < 768; i += CM_THREADS) nl[i] = a.lut[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int64_t hw = a.hw;
    const int groups = (int)(hw >> 2);
    const LT* lg = reinterpret_cast<const LT*>(a.logits) + (int64_t)b * K * hw;
    for (int g = blockIdx.x * CM_THREADS + threadIdx.x; g < groups; g += gridDim.x * CM_THREADS) {
        const int64_t r = (int64_t)g << 2;
        uint32_t wds[K][3];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(a.x[k] + ((int64_t)b * hw + r) * 3);
            wds[k][0] = __ldg(p); wds[k][1] = __ldg(p + 1); wds[k][2] = __ldg(p + 2);
        }
        float4 w[K];
        if (!BWD || a.softmax) {
#pragma unroll
            for (int k = 0; k < K; ++k) w[k] = cm_load4<LT>(lg + k * hw + r);
            if (a.softmax) softmax4_k<K>(w);
        }
        if (!BWD && a.softmax && a.w_out) {
#pragma unroll
            for (int k = 0; k < K; ++k) st_stream_f4(a.w_out + ((int64_t)b * K + k) * hw + r, w[k]);
        }
        float4 gw[K];
        if (BWD) {
#pragma unroll
            for (int k = 0; k < K; ++k) gw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4 xv[K];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = 3 * i + c;
                    cf4(xv[k], i) = nl[c * 256 + ((wds[k][e >> 2] >> (8 * (e & 3))) & 255u)];
                }
            if (!BWD) {
                float4 acc;
                acc.x = __fmul_rn(xv[0].x, w[0].x); acc.y = __fmul_rn(xv[0].y, w[0].y);
                acc.z = __fmul_rn(xv[0].z, w[0].z); acc.w = __fmul_rn(xv[0].w, w[0].w);
#pragma unroll
                for (int k = 1; k < K; ++k) {
                    acc.x = __fadd_rn(acc.x, __fmul_rn(xv[k].x, w[k].x)); acc.y = __fadd_rn(acc.y, __fmul_rn(xv[k].y, w[k].y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(xv[k].z, w[k].z)); acc.w = __fadd_rn(acc.w, __fmul_rn(xv[k].w, w[k].w));
                }
                cm_store4<OT>(reinterpret_cast<OT*>(a.out) + ((int64_t)b * 3 + c) * hw + r, acc);
            } else {
                const float4 gv = cm_load4<OT>(reinterpret_cast<const OT*>(a.grad_out) + ((int64_t)b * 3 + c) * hw + r);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    gw[k].x = fmaf(gv.x, xv[k].x, gw[k].x); gw[k].y = fmaf(gv.y, xv[k].y, gw[k].y);
                    gw[k].z = fmaf(gv.z, xv[k].z, gw[k].z); gw[k].w = fmaf(gv.w, xv[k].w, gw[k].w);
                }
            }
        }
        if (BWD) {
            if (a.softmax) {
                float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    dot.x = fmaf(w[k].x, gw[k].x, dot.x); dot.y = fmaf(w[k].y, gw[k].y, dot.y);
                    dot.z = fmaf(w[k].z, gw[k].z, dot.z); dot.w = fmaf(w[k].w, gw[k].w, dot.w);
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    gw[k].x = w[k].x * (gw[k].x - dot.x); gw[k].y = w[k].y * (gw[k].y - dot.y);
                    gw[k].z = w[k].z * (gw[k].z - dot.z); gw[k].w = w[k].w * (gw[k].w - dot.w);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) st_stream_f4(reinterpret_cast<float*>(a.out) + ((int64_t)b * K + k) * hw + r, gw[k]);
        }
    }
}

static uint32_t cm_inv(int wq) { return (uint32_t)((0x100000000ull + (uint64_t)wq - 1) / (uint64_t)wq); }

static dim3 cm_grid(int groups, int B) {
    // Every CTA serves one image (it builds that image's tables once).  Measured on B200 at B = 256, 256x192 (us, fwd / emit /
    // bwd): 2 CTAs per image (one wave of 512 CTAs, 24 groups per thread) 154 / 161 / 173; 4: 118 / 130 / 138; 8: 89 / 101 / 107;
    // 16 (3 groups per thread, 7 waves) 77 / 89 / 92; 48 (1 group per thread) 100 / 96 / 112.  A thread has one iteration of
    // loads in flight, so few fat CTAs starve the memory system and end in a long tail; ~3 groups per thread amortise the
    // table fill without that.  Small batches still get at least enough CTAs to fill the GPU.
    const int max_ctas = (groups + CM_THREADS - 1) / CM_THREADS;
    int per_img = std::max((max_ctas + 2) / 3, (4 * sm_count()) / std::max(B, 1));
    // small batches (the reference's 32 per GPU): the launch is latency-bound, one group per thread so that every load of the
    // launch is in flight at once - B = 32 (us, emit / fwd / bwd): 18 CTAs per image 19.3 / 20.0 / 22.0, 24: 16.5 / 18.8 / 20.5,
    // 48 (one group per thread): 15.5 / 16.3 / 17.5
    if ((int64_t)B * max_ctas <= 12 * (int64_t)sm_count()) per_img = max_ctas;
    if (const char* e = getenv("ADVMIX_CM_PER_IMG")) per_img = atoi(e);                  // experiment knob
    per_img = std::max(1, std::min(per_img, max_ctas));
    return dim3((unsigned)per_img, (unsigned)B);
}

}  // namespace advmix

using namespace advmix;

static int cm_check(const char* what, int B, int H, int W, int ldt, int odt) {
    ADVMIX_REQUIRE(B >= 0 && B <= 65535 && H > 0 && W > 0, "%s: bad shape B=%d H=%d W=%d", what, B, H, W);
    ADVMIX_REQUIRE(W % 4 == 0 && H <= 1024 && W <= 1024, "%s: W must be a multiple of 4, H and W at most 1024 (got %dx%d)", what, H, W);
    ADVMIX_REQUIRE((ldt == ADVMIX_F32 || ldt == ADVMIX_BF16) && (odt == ADVMIX_F32 || odt == ADVMIX_BF16), "%s: bad dtype", what);
    return ADVMIX_OK;
}

#define CM_DISPATCH(MODE, ldt, odt, grid, args, st)                                                               \
    do {                                                                                                           \
        if (ldt == ADVMIX_F32 && odt == ADVMIX_F32) chain_kernel<MODE, float, float><<<grid, CM_THREADS, 0, st>>>(args);          \
        else if (ldt == ADVMIX_F32) chain_kernel<MODE, float, __nv_bfloat16><<<grid, CM_THREADS, 0, st>>>(args);                 \
        else if (odt == ADVMIX_F32) chain_kernel<MODE, __nv_bfloat16, float><<<grid, CM_THREADS, 0, st>>>(args);                 \
        else chain_kernel<MODE, __nv_bfloat16, __nv_bfloat16><<<grid, CM_THREADS, 0, st>>>(args);                                \
    } while (0)

extern "C" {

size_t advmix_autoaug_plan_bytes(int B) { return B > 0 ? (size_t)B * sizeof(AutoPlan) : 0; }

int advmix_chains_emit_u8c3(const uint8_t* crop, const void* plans, const int32_t* gridmask_params, const float* norm_lut,
                            void* g_input, int B, int H, int W, int dtype, advmix_stream_t stream) {
    int rc = cm_check("chains_emit", B, H, W, ADVMIX_F32, dtype);
    if (rc) return rc;
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(crop && norm_lut && g_input, "chains_emit: null argument");
    ChainMixArgs a{crop, reinterpret_cast<const AutoPlan*>(plans), gridmask_params, norm_lut, nullptr, nullptr, g_input, nullptr, H, W, 0, cm_inv(W / 4)};
    const dim3 grid = cm_grid(H * (W / 4), B);
    CM_DISPATCH(CM_EMIT, ADVMIX_F32, dtype, grid, a, as_stream(stream));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_chainmix_fwd(const uint8_t* crop, const void* plans, const int32_t* gridmask_params, const float* norm_lut,
                        const void* w_or_logits, int w_dtype, int apply_softmax, void* out, int out_dtype, float* w_out,
                        int B, int H, int W, advmix_stream_t stream) {
    int rc = cm_check("chainmix_fwd", B, H, W, w_dtype, out_dtype);
    if (rc) return rc;
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(crop && norm_lut && w_or_logits && out, "chainmix_fwd: null argument");
    ChainMixArgs a{crop, reinterpret_cast<const AutoPlan*>(plans), gridmask_params, norm_lut, w_or_logits, nullptr, out, w_out, H, W, apply_softmax, cm_inv(W / 4)};
    const dim3 grid = cm_grid(H * (W / 4), B);
    CM_DISPATCH(CM_FWD, w_dtype, out_dtype, grid, a, as_stream(stream));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_chainmix_bwd(const uint8_t* crop, const void* plans, const int32_t* gridmask_params, const float* norm_lut,
                        const void* w_or_logits, int w_dtype, int through_softmax, const void* grad_out, int out_dtype,
                        float* grad_w, int B, int H, int W, advmix_stream_t stream) {
    int rc = cm_check("chainmix_bwd", B, H, W, w_dtype, out_dtype);
    if (rc) return rc;
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(crop && norm_lut && grad_out && grad_w, "chainmix_bwd: null argument");
    ADVMIX_REQUIRE(!through_softmax || w_or_logits, "chainmix_bwd: through_softmax needs the logits");
    ChainMixArgs a{crop, reinterpret_cast<const AutoPlan*>(plans), gridmask_params, norm_lut, w_or_logits, grad_out, grad_w, nullptr, H, W, through_softmax, cm_inv(W / 4)};
    const dim3 grid = cm_grid(H * (W / 4), B);
    CM_DISPATCH(CM_BWD, w_dtype, out_dtype, grid, a, as_stream(stream));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"

// ---- general uint8-chain mix ----------------------------------------------------------------------------------
template <typename LT, typename OT, bool BWD>
static void mu_launch_k(int K, dim3 grid, const MixU8Args& a, cudaStream_t st) {
    switch (K) {
        case 1: mix_u8_kernel<LT, OT, 1, BWD><<<grid, CM_THREADS, 0, st>>>(a); break;
        case 2: mix_u8_kernel<LT, OT, 2, BWD><<<grid, CM_THREADS, 0, st>>>(a); break;
        case 3: mix_u8_kernel<LT, OT, 3, BWD><<<grid, CM_THREADS, 0, st>>>(a); break;
        default: mix_u8_kernel<LT, OT, 4, BWD><<<grid, CM_THREADS, 0, st>>>(a); break;
    }
}
template <bool BWD>
static void mu_launch(int K, int ldt, int odt, dim3 grid, const MixU8Args& a, cudaStream_t st) {
    if (ldt == ADVMIX_F32 && odt == ADVMIX_F32) mu_launch_k<float, float, BWD>(K, grid, a, st);
    else if (ldt == ADVMIX_F32) mu_launch_k<float, __nv_bfloat16, BWD>(K, grid, a, st);
    else if (odt == ADVMIX_F32) mu_launch_k<__nv_bfloat16, float, BWD>(K, grid, a, st);
    else mu_launch_k<__nv_bfloat16, __nv_bfloat16, BWD>(K, grid, a, st);
}

static int mu_check(const char* what, const uint8_t* const* x_h, int B, int K, int H, int W, int ldt, int odt) {
    ADVMIX_REQUIRE(B >= 0 && B <= 65535 && K >= 1 && K <= MU_MAXK && H > 0 && W > 0, "%s: bad shape B=%d K=%d H=%d W=%d", what, B, K, H, W);
    ADVMIX_REQUIRE(((int64_t)H * W) % 4 == 0, "%s: H*W must be a multiple of 4", what);
    ADVMIX_REQUIRE((ldt == ADVMIX_F32 || ldt == ADVMIX_BF16) && (odt == ADVMIX_F32 || odt == ADVMIX_BF16), "%s: bad dtype", what);
    ADVMIX_REQUIRE(x_h != nullptr, "%s: null x_h", what);
    for (int k = 0; k < K; ++k) ADVMIX_REQUIRE(x_h[k] != nullptr, "%s: x_h[%d] is null", what, k);
    return ADVMIX_OK;
}

extern "C" {

int advmix_mix_u8_fwd(const uint8_t* const* x_h, const float* norm_lut, const void* w_or_logits, int w_dtype, int apply_softmax,
                      void* out, int out_dtype, float* w_out, int B, int K, int H, int W, advmix_stream_t stream) {
    int rc = mu_check("mix_u8_fwd", x_h, B, K, H, W, w_dtype, out_dtype);
    if (rc) return rc;
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(norm_lut && w_or_logits && out, "mix_u8_fwd: null argument");
    MixU8Args a{};
    for (int k = 0; k < K; ++k) a.x[k] = x_h[k];
    a.lut = norm_lut; a.logits = w_or_logits; a.out = out; a.w_out = w_out; a.hw = (int64_t)H * W; a.softmax = apply_softmax;
    mu_launch<false>(K, w_dtype, out_dtype, cm_grid((int)(a.hw / 4), B), a, as_stream(stream));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_mix_u8_bwd(const uint8_t* const* x_h, const float* norm_lut, const void* w_or_logits, int w_dtype, int through_softmax,
                      const void* grad_out, int out_dtype, float* grad_w, int B, int K, int H, int W, advmix_stream_t stream) {
    int rc = mu_check("mix_u8_bwd", x_h, B, K, H, W, w_dtype, out_dtype);
    if (rc) return rc;
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(norm_lut && grad_out && grad_w, "mix_u8_bwd: null argument");
    ADVMIX_REQUIRE(!through_softmax || w_or_logits, "mix_u8_bwd: through_softmax needs the logits");
    MixU8Args a{};
    for (int k = 0; k < K; ++k) a.x[k] = x_h[k];
    a.lut = norm_lut; a.logits = w_or_logits; a.grad_out = grad_out; a.out = grad_w; a.hw = (int64_t)H * W; a.softmax = through_softmax;
    mu_launch<true>(K, w_dtype, out_dtype, cm_grid((int)(a.hw / 4), B), a, as_stream(stream));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"
