// Row f3: the consumers of the network's heat maps (lib/core/inference.py:22-95, lib/core/function.py:241-261).
//   advmix_heatmap_decode : get_max_preds (+ the POST_PROCESS quarter-pixel step + transform_preds)
//   advmix_flip_merge     : flip_back + SHIFT_HEATMAP + (output + output_flipped) * 0.5
// Both read each heat map once: HBM-read bound, 4*J*Hh*Wh bytes per sample and map.
#include "affine.cuh"
#include "common.cuh"

#include <algorithm>

namespace advmix {

constexpr int DEC_WARPS = 8;

// One warp per (b, j) plane.  np.argmax returns the FIRST maximum, so ties keep the smaller index.
__global__ void __launch_bounds__(DEC_WARPS * 32)
heatmap_decode_kernel(const float* __restrict__ hm, const float* __restrict__ center, const double* __restrict__ scale,
                      int scale_f32, int post_process, float* __restrict__ preds, float* __restrict__ maxvals,
                      float* __restrict__ coords, int planes, int J, int H, int W) {
    const int lane = threadIdx.x & 31;
    const int plane = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
    if (plane >= planes) return;
    const int n = H * W;
    const float* p = hm + (int64_t)plane * n;
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    if ((n & 3) == 0) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        for (int q = lane; q < n / 4; q += 32) {
            const float4 v = ld_stream_f4(p4 + q);
            const int i0 = 4 * q;
            if (v.x > best) { best = v.x; bidx = i0; }
            if (v.y > best) { best = v.y; bidx = i0 + 1; }
            if (v.z > best) { best = v.z; bidx = i0 + 2; }
            if (v.w > best) { best = v.w; bidx = i0 + 3; }
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            const float v = p[i];
            if (v > best) { best = v; bidx = i; }
        }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane != 0) return;
    if (bidx == 0x7fffffff) bidx = 0;                       // all -inf / NaN plane: numpy's argmax gives 0
    // preds = (idx % width, floor(idx / width)) in float32, zeroed where maxval <= 0
    float x = (float)(bidx % W), y = (float)(bidx / W);
    if (!(best > 0.0f)) { x = 0.0f; y = 0.0f; }
    if (post_process) {
        const int px = (int)floorf(x + 0.5f), py = (int)floorf(y + 0.5f);
        if (1 < px && px < W - 1 && 1 < py && py < H - 1) {
            const float dx = __fsub_rn(p[py * W + px + 1], p[py * W + px - 1]);
            const float dy = __fsub_rn(p[(py + 1) * W + px], p[(py - 1) * W + px]);
            x += (dx > 0.0f ? 0.25f : dx < 0.0f ? -0.25f : 0.0f);
            y += (dy > 0.0f ? 0.25f : dy < 0.0f ? -0.25f : 0.0f);
        }
    }
    maxvals[plane] = best;
    if (coords) { coords[2 * plane] = x; coords[2 * plane + 1] = y; }
    if (preds) {
        // transform_preds: get_affine_transform(center, scale, 0, [W, H], inv=1) applied in float64, stored as float32
        const int b = plane / J;
        double m[6];
        affine_from_csr(center[2 * b], center[2 * b + 1], scale[2 * b], scale_f32, 0.0, W, H, m, 1);
        const double xd = (double)x, yd = (double)y;
        preds[2 * plane] = (float)fma(m[0], xd, fma(m[1], yd, m[2]));          // same dot order as the joints kernel
        preds[2 * plane + 1] = (float)fma(m[3], xd, fma(m[4], yd, m[5]));
    }
}

// out[b,j,y,x] = (a[b,j,y,x] + f[b,perm[j],y, W-1-max(x-shift,0)]) * 0.5   (x = 0 keeps the un-shifted column)
__global__ void __launch_bounds__(256)
flip_merge_kernel(const float* __restrict__ a, const float* __restrict__ f, const int32_t* __restrict__ perm, int shift,
                  float* __restrict__ out, int J, int H, int W, int64_t rows) {
    // one warp per heat-map row: coalesced loads of both rows, reversed through the index
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const int y = (int)(r % H);
        const int64_t bj = r / H;
        const int j = (int)(bj % J);
        const int64_t b = bj / J;
        const int pj = perm ? perm[j] : j;
        const float* ar = a ? a + r * W : nullptr;
        const float* fr = f + ((b * J + pj) * H + y) * (int64_t)W;
        float* o = out + r * W;
        for (int x = lane; x < W; x += 32) {
            const int xs = shift ? max(x - 1, 0) : x;
            const float g = fr[W - 1 - xs];
            o[x] = ar ? __fmul_rn(__fadd_rn(ar[x], g), 0.5f) : g;          // a == nullptr: flip_back alone
        }
    }
}

}  // namespace advmix

using namespace advmix;

extern "C" {

int advmix_heatmap_decode(const float* heatmaps, const float* center, const double* scale, int scale_is_f32, int post_process,
                          float* preds, float* maxvals, float* coords_hm, int B, int J, int Hh, int Wh, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0 && Hh > 0 && Wh > 0, "heatmap_decode: bad shape B=%d J=%d %dx%d", B, J, Hh, Wh);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(heatmaps && maxvals, "heatmap_decode: null argument");
    ADVMIX_REQUIRE(!preds || (center && scale), "heatmap_decode: preds need center and scale");
    ADVMIX_REQUIRE((int64_t)Hh * Wh < (1ll << 24), "heatmap_decode: map too large for float32 indices");
    const int planes = B * J;
    heatmap_decode_kernel<<<ceil_div(planes, DEC_WARPS), DEC_WARPS * 32, 0, as_stream(stream)>>>(
        heatmaps, center, scale, scale_is_f32, post_process, preds, maxvals, coords_hm, planes, J, Hh, Wh);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_flip_merge(const float* output, const float* output_flipped, const int32_t* flip_perm, int shift_heatmap,
                      float* merged, int B, int J, int Hh, int Wh, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0 && Hh > 0 && Wh > 0, "flip_merge: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(output_flipped && merged, "flip_merge: null argument");
    ADVMIX_REQUIRE(merged != output_flipped, "flip_merge: merged may alias output but not output_flipped");
    const int64_t rows = (int64_t)B * J * Hh;
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 16);
    flip_merge_kernel<<<grid, 256, 0, as_stream(stream)>>>(output, output_flipped, flip_perm, shift_heatmap, merged, J, Hh, Wh, rows);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"
