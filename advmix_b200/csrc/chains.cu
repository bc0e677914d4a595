// Rows a5 / a6: the reference-actual AdvMix chains (lib/dataset/advaug.py).
//
// autoaug - every reachable PIL op except sharpness is a per-channel 256-entry LUT, and
// equalize's LUT depends only on the histogram of its input, which for a LUT predecessor
// is the push-forward of the source histogram.  So a sub-policy collapses to
//     out = post_lut[ sharpen?( pre_lut[in] ) ]
// planned per image from ONE histogram pass; the image is then read and written once.
#include "chains.cuh"

namespace advmix {

__device__ __forceinline__ uint8_t pointwise_op(int op, float mag, int v) {
    if (op == OP_POSTERIZE) {
        const int bits = (int)mag;
        return (uint8_t)(v & (~((1 << (8 - bits)) - 1) & 0xFF));
    }
    if (op == OP_SOLARIZE) return (uint8_t)(((float)v < mag) ? v : 255 - v);
    if (op == OP_INVERT) return (uint8_t)(255 - v);
    return (uint8_t)v;
}

// Histogram + plan of one image in ONE launch: a thread-block cluster of AA_CLUSTER CTAs per image.  Every CTA histograms its
// interleaved share of the image into its own shared memory; after a cluster barrier the CTAs of rank 0 / 1 / 2 each add up
// one colour channel's eight partial histograms through distributed shared memory and build that channel's tables (256
// threads = 256 grey levels).  The histogram never goes through global memory: no workspace, no memset node, no second
// and third launch (memset + histogram kernel + one-CTA-per-image plan kernel took 16.9 us for 32 images, three dependent
// graph nodes of mostly launch latency).
// The histogram is that of the stage-1 output whenever stage 2 is equalize after a sharpness stage, otherwise of the raw input.
constexpr int AA_CLUSTER = 8;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ld_dsmem(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t remote, v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
    return v;
}

__global__ void __cluster_dims__(AA_CLUSTER, 1, 1) __launch_bounds__(256)
autoaug_hist_plan_kernel(const uint8_t* __restrict__ in, const int32_t* __restrict__ ops, const float* __restrict__ mags,
                         AutoPlan* __restrict__ plans, int H, int W) {
    __shared__ uint32_t sh[768];
    __shared__ uint32_t h[256], h2[256], scan[256], wtot[8];
    __shared__ uint8_t cur[256], ident[256];
    __shared__ int s_nonzero, s_last;
    const int b = blockIdx.y, t = threadIdx.x;
    const uint32_t rank = cluster_ctarank();              // == blockIdx.x: gridDim.x is the cluster size
    const int op[2] = {ops[2 * b], ops[2 * b + 1]};
    const float mag[2] = {mags[2 * b], mags[2 * b + 1]};
    const bool need_hist = op[0] == OP_EQUALIZE || op[1] == OP_EQUALIZE;     // the same for the whole cluster
    for (int i = t; i < 768; i += 256) sh[i] = 0;
    ident[t] = (uint8_t)t;
    __syncthreads();
    if (need_hist) {
        const bool sharp_first = (op[0] == OP_SHARPNESS);
        const float factor = mag[0];
        const uint32_t n3 = (uint32_t)H * (uint32_t)W * 3u;
        const uint8_t* img = in + (int64_t)b * n3;
        const int64_t pitch = (int64_t)W * 3;
        if (!sharp_first && (n3 & 3u) == 0 && (reinterpret_cast<uintptr_t>(in) & 3u) == 0) {
            // whole words; the channel of byte k of word w is (4w + k) mod 3
            const uint32_t nw = n3 >> 2;
            const uint32_t* img4 = reinterpret_cast<const uint32_t*>(img);
            for (uint32_t w = rank * 256u + t; w < nw; w += AA_CLUSTER * 256u) {
                const uint32_t v = __ldg(img4 + w);
                const uint32_t c0 = (4u * w) % 3u;
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                    uint32_t c = c0 + k;
                    c -= c >= 3u ? 3u : 0u;                        // c0 <= 2, k <= 3
                    atomicAdd(&sh[c * 256u + ((v >> (8u * k)) & 0xFFu)], 1u);
                }
            }
        } else {
            for (uint32_t e = rank * 256u + t; e < n3; e += AA_CLUSTER * 256u) {
                uint8_t v;
                if (sharp_first) {
                    const uint32_t pix = e / 3u;
                    const int y = (int)(pix / (uint32_t)W), x = (int)(pix - (uint32_t)y * (uint32_t)W);
                    const bool interior = y > 0 && y < H - 1 && x > 0 && x < W - 1;
                    v = sharpen_px(img + e, pitch, ident, factor, interior);
                } else {
                    v = img[e];
                }
                atomicAdd(&sh[(e % 3u) * 256u + v], 1u);
            }
        }
    }
    cluster_barrier();                                    // all partial histograms are complete and visible cluster-wide
    if (rank < 3) {
        const int c = (int)rank;
        AutoPlan* plan = plans + b;
        // where does the stencil sit?  (S,L): pre = id, post = L.  (L,S): pre = L, post = id.
        const int sidx = op[0] == OP_SHARPNESS ? 0 : (op[1] == OP_SHARPNESS ? 1 : -1);
        if (c == 0 && t == 0) {
            plan->stencil = sidx >= 0;
            plan->factor = sidx >= 0 ? mag[sidx] : 1.0f;
        }
        uint32_t hsum = 0;
        if (need_hist) {
            const uint32_t mine = (uint32_t)__cvta_generic_to_shared(&sh[c * 256 + t]);
#pragma unroll
            for (uint32_t r = 0; r < AA_CLUSTER; ++r) hsum += ld_dsmem(mine, r);
        }
        h[t] = hsum;                          // histogram of the input of the first equalize
        cur[t] = (uint8_t)t;
        __syncthreads();
        bool hist_is_current = true;          // h describes the image `cur` maps to
        for (int s = 0; s < 2; ++s) {
            if (op[s] == OP_SHARPNESS) {
                plan->pre[c * 256 + t] = cur[t];
                __syncthreads();
                cur[t] = (uint8_t)t;
                // after a stencil the histogram is only valid if the hist pass recomputed it
                hist_is_current = (s == 0);
                __syncthreads();
                continue;
            }
            uint8_t f;  // this stage's map applied to level t
            if (op[s] == OP_EQUALIZE) {
                // PIL ImageOps.equalize on histogram h
                if (t == 0) { s_nonzero = 0; s_last = 0; }
                __syncthreads();
                const uint32_t ht = h[t];
                if (ht) { atomicAdd(&s_nonzero, 1); atomicMax(&s_last, t); }
                // inclusive scan: shuffles inside a warp, then the eight warp totals
                uint32_t inc = ht;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if ((t & 31) >= o) inc += n;
                }
                if ((t & 31) == 31) wtot[t >> 5] = inc;
                __syncthreads();
                uint32_t total = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < (t >> 5)) inc += wtot[i];
                    total += wtot[i];
                }
                const uint32_t excl = inc - ht;
                f = (uint8_t)t;
                if (s_nonzero > 1 && hist_is_current) {
                    const uint32_t step = (total - h[s_last]) / 255u;
                    if (step) f = (uint8_t)min((step / 2 + excl) / step, 255u);
                }
            } else {
                f = pointwise_op(op[s], mag[s], t);
            }
            // push the histogram forward through f, compose cur = f o cur
            h2[t] = 0;
            __syncthreads();
            if (h[t]) atomicAdd(&h2[f], h[t]);
            scan[t] = f;      // the stage table
            __syncthreads();
            cur[t] = (uint8_t)scan[cur[t]];
            h[t] = h2[t];
            __syncthreads();
        }
        if (sidx < 0) {
            // no stencil: fold everything into pre so the apply kernel does one lookup
            plan->pre[c * 256 + t] = cur[t];
            plan->post[c * 256 + t] = (uint8_t)t;
        } else {
            plan->post[c * 256 + t] = cur[t];
        }
    }
    cluster_barrier();                                    // the partial histograms stay alive until ranks 0..2 have read them
}

// out = post[ sharpen?( pre[in] ) ], optionally also ToTensor+Normalize of the result.
// grid (chunks, B); each thread handles 4 pixels (12 bytes).
__global__ void __launch_bounds__(256)
autoaug_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, void* __restrict__ out_norm,
                     const float* __restrict__ norm_lut, const AutoPlan* __restrict__ plans, int H, int W,
                     int norm_dtype) {
    __shared__ uint8_t pre[768], post[768];
    __shared__ float nl[768];
    const int b = blockIdx.y;
    const AutoPlan* plan = plans + b;
    for (int i = threadIdx.x; i < 768; i += 256) {
        pre[i] = plan->pre[i];
        post[i] = plan->post[i];
        if (out_norm) nl[i] = norm_lut[i];
    }
    __syncthreads();
    const bool stencil = plan->stencil != 0;
    const float factor = plan->factor;
    const int64_t npix = (int64_t)H * W;
    const uint8_t* img = in + (int64_t)b * npix * 3;
    const int64_t pitch = (int64_t)W * 3;
    for (int64_t pix = (int64_t)blockIdx.x * 256 + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * 256) {
        const int y = (int)((uint32_t)pix / (uint32_t)W), x = (int)((uint32_t)pix - (uint32_t)y * (uint32_t)W);
        const bool interior = y > 0 && y < H - 1 && x > 0 && x < W - 1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint8_t* p = img + pix * 3 + c;
            uint8_t v = stencil ? post[c * 256 + sharpen_px(p, pitch, pre + c * 256, factor, interior)]
                                : pre[c * 256 + *p];
            if (out) out[((int64_t)b * npix + pix) * 3 + c] = v;
            if (out_norm) {
                const float f = nl[c * 256 + v];
                const int64_t o = ((int64_t)b * 3 + c) * npix + pix;
                if (norm_dtype == ADVMIX_F32) reinterpret_cast<float*>(out_norm)[o] = f;
                else reinterpret_cast<__nv_bfloat16*>(out_norm)[o] = __float2bfloat16_rn(f);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
gridmask_kernel(const T* __restrict__ in, T* __restrict__ out, const int32_t* __restrict__ params, int H, int W) {
    const int b = blockIdx.y;
    const int apply = params[4 * b], d = params[4 * b + 1], st_h = params[4 * b + 2], st_w = params[4 * b + 3];
    const int64_t plane = (int64_t)H * W;
    const GridGeom g = grid_geom(H, W, max(d, 2));
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < plane * 3; e += (int64_t)gridDim.x * 256) {
        const uint32_t pix = (uint32_t)e % (uint32_t)plane;
        const int y = (int)(pix / (uint32_t)W), x = (int)(pix - (uint32_t)y * (uint32_t)W);
        const int64_t o = (int64_t)b * 3 * plane + e;
        float v = (float)in[o];
        if (apply) v = __fmul_rn(v, grid_mask_at(y, x, d, st_h, st_w, g));
        out[o] = (T)v;
    }
}

__global__ void gridmask_vis_kernel(const int32_t* __restrict__ params, const double* __restrict__ joints,
                                    const double* __restrict__ vin, double* __restrict__ vout, int B, int J, int H, int W) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * J) return;
    const int b = t / J;
    const int apply = params[4 * b], d = params[4 * b + 1], st_h = params[4 * b + 2], st_w = params[4 * b + 3];
    double v0 = vin[(int64_t)t * 3], v1 = vin[(int64_t)t * 3 + 1];
    const double v2 = vin[(int64_t)t * 3 + 2];
    if (apply) {
        const GridGeom g = grid_geom(H, W, max(d, 2));
        int tx = min(__double2int_rz(joints[(int64_t)t * 3]), W - 1);
        tx = max(tx, 0);
        int ty = min(__double2int_rz(joints[(int64_t)t * 3 + 1]), H - 1);
        ty = max(ty, 0);
        if (grid_mask_at(ty, tx, d, st_h, st_w, g) == 0.0f) { v0 = 0.0; v1 = 0.0; }
    }
    vout[(int64_t)t * 3] = v0;
    vout[(int64_t)t * 3 + 1] = v1;
    vout[(int64_t)t * 3 + 2] = v2;
}

}  // namespace advmix

using namespace advmix;

extern "C" {

size_t advmix_autoaug_workspace_bytes(int B, int H, int W) {
    (void)H; (void)W;
    if (B <= 0) return 0;
    return (size_t)B * 768 * sizeof(uint32_t) + (size_t)B * sizeof(AutoPlan);
}

int advmix_autoaug_u8c3(const uint8_t* in, uint8_t* out, void* out_norm, const float* norm_lut, const int32_t* ops,
                        const float* mags, int B, int H, int W, int norm_dtype, void* workspace, size_t ws_bytes,
                        advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && H >= 3 && W >= 3 && (int64_t)H * W * 3 < (int64_t)1 << 31, "autoaug: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(in && ops && mags && (out || out_norm), "autoaug: null argument");
    ADVMIX_REQUIRE(in != out, "autoaug: in-place not supported (sharpness reads neighbours)");
    ADVMIX_REQUIRE(!out_norm || norm_lut, "autoaug: out_norm needs norm_lut");
    ADVMIX_REQUIRE(norm_dtype == ADVMIX_F32 || norm_dtype == ADVMIX_BF16, "autoaug: bad dtype");
    ADVMIX_REQUIRE(B <= 65535, "autoaug: B<=65535 per call");
    if (!workspace || ws_bytes < advmix_autoaug_workspace_bytes(B, H, W))
        return fail(ADVMIX_ERR_WORKSPACE, "autoaug: workspace %zu < %zu", ws_bytes, advmix_autoaug_workspace_bytes(B, H, W));
    cudaStream_t s = as_stream(stream);
    // the plans live behind the (no longer used) histogram area of the workspace: the layout callers sized it for
    AutoPlan* plans = reinterpret_cast<AutoPlan*>(reinterpret_cast<uint32_t*>(workspace) + (size_t)B * 768);
    const int64_t npix = (int64_t)H * W;
    autoaug_hist_plan_kernel<<<dim3(AA_CLUSTER, B), 256, 0, s>>>(in, ops, mags, plans, H, W);
    ADVMIX_LAUNCH_OK();
    const int achunks = (int)std::min<int64_t>((npix + 255) / 256, 64);
    autoaug_apply_kernel<<<dim3(achunks, B), 256, 0, s>>>(in, out, out_norm, norm_lut, plans, H, W, norm_dtype);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_autoaug_plan_u8c3(const uint8_t* in, const int32_t* ops, const float* mags, void* plans_out, int B, int H, int W,
                             void* workspace, size_t ws_bytes, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && H >= 3 && W >= 3 && (int64_t)H * W * 3 < (int64_t)1 << 31, "autoaug_plan: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(in && ops && mags && plans_out, "autoaug_plan: null argument");
    ADVMIX_REQUIRE(B <= 65535, "autoaug_plan: B<=65535 per call");
    const size_t need = (size_t)B * 768 * sizeof(uint32_t);
    if (!workspace || ws_bytes < need) return fail(ADVMIX_ERR_WORKSPACE, "autoaug_plan: workspace %zu < %zu", ws_bytes, need);
    // (the workspace was the global histogram of the two-kernel version; the cluster kernel keeps it in shared memory.  The
    // argument and its size check stay: same C ABI and error behaviour)
    autoaug_hist_plan_kernel<<<dim3(AA_CLUSTER, B), 256, 0, as_stream(stream)>>>(in, ops, mags, reinterpret_cast<AutoPlan*>(plans_out), H, W);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_gridmask(const void* img_in, void* img_out, const int32_t* params, const double* joints, const double* vis_in,
                    double* vis_out, int B, int H, int W, int J, int dtype, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && H > 0 && W > 0 && J >= 0, "gridmask: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(params && ((img_in && img_out) || (!img_in && !img_out)), "gridmask: null argument");
    ADVMIX_REQUIRE(dtype == ADVMIX_F32 || dtype == ADVMIX_BF16, "gridmask: bad dtype");
    ADVMIX_REQUIRE(B <= 65535, "gridmask: B<=65535 per call");
    cudaStream_t s = as_stream(stream);
    if (img_in) {                                   // img_in == img_out == NULL: only the visibility update (fused chain + mix path)
        const int64_t n = (int64_t)H * W * 3;
        const int chunks = (int)std::min<int64_t>((n + 255) / 256, 96);
        if (dtype == ADVMIX_F32)
            gridmask_kernel<float><<<dim3(chunks, B), 256, 0, s>>>(reinterpret_cast<const float*>(img_in), reinterpret_cast<float*>(img_out), params, H, W);
        else
            gridmask_kernel<__nv_bfloat16><<<dim3(chunks, B), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(img_in), reinterpret_cast<__nv_bfloat16*>(img_out), params, H, W);
        ADVMIX_LAUNCH_OK();
    }
    if (J > 0 && joints && vis_in && vis_out) {
        gridmask_vis_kernel<<<ceil_div((long long)B * J, 128), 128, 0, s>>>(params, joints, vis_in, vis_out, B, J, H, W);
        ADVMIX_LAUNCH_OK();
    }
    return ADVMIX_OK;
}

}  // extern "C"
