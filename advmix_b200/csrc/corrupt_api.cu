// C-ABI dispatcher for the imagecorruptions operators: argument checks,
// workspace / random-buffer sizing, and the fill kernel that materialises exactly the draws
// the fused perf-mode kernels generate in-register.
#include "corrupt_common.cuh"

#include <algorithm>

namespace advmix {

size_t stencil_ws_bytes(int op, int severity, int n, int H, int W);
void pixelate_dims(int severity, int H, int W, int* h2, int* w2);
size_t jpeg_ws_bytes(int n, int H, int W);
size_t validation_ws_bytes(int op, int severity, int n, int H, int W);

size_t ws_bytes_for(int op, int severity, int n, int H, int W) {
    if (n <= 0) return 0;
    size_t b = 0;
    switch (op) {
        case C_CONTRAST: b = (size_t)n * 3 * sizeof(unsigned long long); break;
        case C_PIXELATE: {
            int h2, w2;
            pixelate_dims(severity, H, W, &h2, &w2);
            b = (size_t)n * H * w2 * 3 + (size_t)n * h2 * w2 * 3;
            break;
        }
        case C_JPEG: b = jpeg_ws_bytes(n, H, W); break;
        case C_GAUSSIAN_BLUR: case C_SPATTER: b = validation_ws_bytes(op, severity, n, H, W); break;
        default: b = stencil_ws_bytes(op, severity, n, H, W); break;
    }
    return (b + 255) & ~(size_t)255;
}

size_t field_bytes_for(int op, int severity, int H, int W) {
    const size_t hw = (size_t)H * W;
    switch (op) {
        case C_GAUSSIAN_NOISE: case C_SHOT_NOISE: case C_SPECKLE_NOISE: return hw * 3 * sizeof(float);
        case C_IMPULSE_NOISE: return 2 * hw * 3 * sizeof(float);
        case C_GLASS_BLUR: return (size_t)glass_iters(severity) * hw * 2;
        case C_SNOW: case C_SPATTER: return hw * sizeof(float);
        case C_FOG: { const size_t M = next_pow2(std::max(H, W)); return M * M * sizeof(float); }
        case C_ELASTIC: return 2 * hw * sizeof(float);
        default: return 0;
    }
}

// Writes the injected-layout buffers from the same device functions the fused kernels use.
__global__ void __launch_bounds__(256)
fill_rand_kernel(int op, int severity, const int32_t* __restrict__ idx, int H, int W, uint64_t seed, int64_t sample_base,
                 char* __restrict__ field, size_t field_stride, double* __restrict__ param, int fn, int fh, int fw) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const int64_t hw = (int64_t)H * W;
    const int64_t t0 = (int64_t)blockIdx.x * 256 + threadIdx.x, ts = (int64_t)gridDim.x * 256;
    float* f = reinterpret_cast<float*>(field + (size_t)i * field_stride);
    if (param && t0 == 0) {
        double* p = param + 4 * i;
        p[0] = p[1] = p[2] = p[3] = 0.0;
        if (op == C_MOTION_BLUR) p[0] = param_uniform(nullptr, rng, -45.0, 45.0);
        if (op == C_SNOW) p[0] = param_uniform(nullptr, rng, -135.0, -45.0);
        if (op == C_FROST) {
            const uint4 u = rng.quad(TAG_PARAM, 0);
            p[0] = (double)__umulhi(u.x, (uint32_t)min(5, fn));
            p[1] = fh > H ? (double)__umulhi(u.y, (uint32_t)(fh - H)) : 0.0;
            p[2] = fw > W ? (double)__umulhi(u.z, (uint32_t)(fw - W)) : 0.0;
        }
    }
    if (!field) return;
    switch (op) {
        case C_GAUSSIAN_NOISE: case C_SPECKLE_NOISE:
            for (int64_t b = t0; b < (hw * 3 + 7) / 8; b += ts) {
                float n[8];
                noise_normal8(rng, TAG_FIELD0, b, n);
                for (int k = 0; k < 8; ++k)
                    if (8 * b + k < hw * 3) f[8 * b + k] = n[k];
            }
            break;
        case C_SHOT_NOISE:
            for (int64_t b = t0; b < (hw * 3 + 7) / 8; b += ts) {
                uint32_t kk[8];
                noise_bits8(rng, TAG_FIELD0, b, kk);
                for (int k = 0; k < 8; ++k)
                    if (8 * b + k < hw * 3) f[8 * b + k] = (float)kk[k] * (1.0f / 65536.0f);
            }
            break;
        case C_IMPULSE_NOISE:
            for (int64_t b = t0; b < (hw * 3 + 7) / 8; b += ts) {
                uint32_t kk[8];
                noise_bits8(rng, TAG_FIELD0, b, kk);
                for (int k = 0; k < 8; ++k)
                    if (8 * b + k < hw * 3) {
                        f[8 * b + k] = (float)(kk[k] & 0x7FFFu) * (1.0f / 32768.0f);          // flip draw
                        f[hw * 3 + 8 * b + k] = (kk[k] & 0x8000u) ? 0.25f : 0.75f;             // salt (<0.5) / pepper
                    }
            }
            break;
        case C_GLASS_BLUR: {
            int8_t* g = reinterpret_cast<int8_t*>(f);
            const int delta = glass_delta(severity);
            for (int64_t e = t0; e < (int64_t)glass_iters(severity) * hw; e += ts) {
                const int2 o = field_glass(nullptr, rng, e, delta);
                g[2 * e] = (int8_t)o.x;
                g[2 * e + 1] = (int8_t)o.y;
            }
            break;
        }
        // the per-element accessors (field_normal1 / field_uniform1) evaluate one Philox block for 4 consecutive elements;
        // here a thread produces all 4 at once (same values, a quarter of the Philox rounds) as one 16-byte store.
        // H*W is a multiple of 4 (check_common), so element quads never straddle two fields.
        case C_SNOW: case C_SPATTER:
            for (int64_t q = t0; q < hw / 4; q += ts) {
                const uint4 u = rng.quad(TAG_FIELD0, (uint64_t)q);
                const float2 a = box_muller(u.x, u.y), b = box_muller(u.z, u.w);
                reinterpret_cast<float4*>(f)[q] = make_float4(a.x, a.y, b.x, b.y);
            }
            break;
        case C_FOG: {
            int M = 1;
            while (M < max(H, W)) M <<= 1;
            for (int64_t q = t0; q < (int64_t)M * M / 4; q += ts) {
                const uint4 u = rng.quad(TAG_FIELD0, (uint64_t)q);
                reinterpret_cast<float4*>(f)[q] = make_float4(u01(u.x), u01(u.y), u01(u.z), u01(u.w));
            }
            break;
        }
        case C_ELASTIC:
            for (int64_t q = t0; q < hw / 4; q += ts) {
                const uint4 u = rng.quad(TAG_FIELD0, (uint64_t)q), v = rng.quad(TAG_FIELD1, (uint64_t)q);
                reinterpret_cast<float4*>(f)[q] = make_float4(u01(u.x), u01(u.y), u01(u.z), u01(u.w));
                reinterpret_cast<float4*>(f + hw)[q] = make_float4(u01(v.x), u01(v.y), u01(v.z), u01(v.w));
            }
            break;
        default: break;
    }
}

int launch_fill_rand(const CorruptArgs& a, void* field, double* param) {
    const size_t fb = field_bytes_for(a.op, a.severity, a.H, a.W);
    if (!param && (!field || fb == 0)) return ADVMIX_OK;
    const int64_t work = std::max<int64_t>(1, (int64_t)(fb / 8));
    const int bx = (int)std::min<int64_t>((work + 255) / 256, 256);
    fill_rand_kernel<<<dim3(bx, a.n), 256, 0, a.stream>>>(a.op, a.severity, a.idx, a.H, a.W, a.seed, a.sample_base,
                                                        fb ? reinterpret_cast<char*>(field) : nullptr, fb, param,
                                                        a.frost_n, a.frost_h, a.frost_w);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

static int check_common(int op, int severity, int n, int H, int W) {
    ADVMIX_REQUIRE(op >= 0 && op < C_NUM_OPS, "corrupt: op %d outside 0..18", op);
    ADVMIX_REQUIRE(severity >= 1 && severity <= 5, "corrupt: severity %d outside 1..5", severity);
    ADVMIX_REQUIRE(n >= 0 && n <= 65535, "corrupt: n=%d outside 0..65535 per call", n);
    ADVMIX_REQUIRE(H >= 32 && W >= 32, "corrupt: image width and height must be at least 32 pixels (got %dx%d)", H, W);
    if (((int64_t)H * W) % 4 != 0)
        return fail(ADVMIX_ERR_UNSUPPORTED, "corrupt: H*W must be a multiple of 4 (got %dx%d)", H, W);
    return ADVMIX_OK;
}

static int run_one(const CorruptArgs& a) {
    const int op = a.op;
    switch (op) {
        case C_GAUSSIAN_NOISE: return run_gaussian_noise(a);
        case C_SHOT_NOISE: return run_shot_noise(a);
        case C_IMPULSE_NOISE: return run_impulse_noise(a);
        case C_DEFOCUS_BLUR: return run_defocus_blur(a);
        case C_GLASS_BLUR: return run_glass_blur(a);
        case C_MOTION_BLUR: return run_motion_blur(a);
        case C_ZOOM_BLUR: return run_zoom_blur(a);
        case C_SNOW: return run_snow(a);
        case C_FROST: return run_frost(a);
        case C_FOG: return run_fog(a);
        case C_BRIGHTNESS: return run_brightness(a);
        case C_CONTRAST: return run_contrast(a);
        case C_ELASTIC: return run_elastic(a);
        case C_PIXELATE: return run_pixelate(a);
        case C_JPEG: return run_jpeg(a);
        case C_SPECKLE_NOISE: return run_speckle_noise(a);
        case C_GAUSSIAN_BLUR: return run_gaussian_blur(a);
        case C_SPATTER: return run_spatter(a);
        case C_SATURATE: return run_saturate(a);
    }
    return fail(ADVMIX_ERR_INVALID, "corrupt: unknown op %d", op);
}

// snow, fog, elastic_transform: the random field has the same size and the same draws at every severity; the sweep
// materialises it once behind the largest per-severity workspace and injects it into the five runs
static bool sweep_shares_field(int op) { return op == C_SNOW || op == C_FOG || op == C_ELASTIC; }
static size_t sweep_ws_main(int op, int n, int H, int W) {
    size_t m = 0;
    for (int s = 1; s <= 5; ++s) m = std::max(m, ws_bytes_for(op, s, n, H, W));
    return (m + 255) & ~(size_t)255;
}
static size_t sweep_ws_bytes(int op, int n, int H, int W) {
    return sweep_ws_main(op, n, H, W) + (sweep_shares_field(op) ? (size_t)n * field_bytes_for(op, 1, H, W) : 0);
}

}  // namespace advmix

using namespace advmix;

extern "C" {

size_t advmix_corrupt_workspace_bytes(int op, int severity, int n, int H, int W) {
    if (op < 0 || op >= C_NUM_OPS || severity < 1 || severity > 5 || H < 1 || W < 1) return 0;
    return ws_bytes_for(op, severity, n, H, W);
}

size_t advmix_corrupt_rand_field_bytes(int op, int severity, int H, int W) {
    if (op < 0 || op >= C_NUM_OPS || severity < 1 || severity > 5 || H < 1 || W < 1) return 0;
    return field_bytes_for(op, severity, H, W);
}

int advmix_corrupt_fill_rand(int op, int severity, int n, int H, int W, uint64_t seed, int64_t sample_base,
                             const int32_t* idx, void* rand_field, double* rand_param, int frost_n, int frost_h,
                             int frost_w, advmix_stream_t stream) {
    int rc = check_common(op, severity, n, H, W);
    if (rc) return rc;
    if (n == 0) return ADVMIX_OK;
    const size_t fb = field_bytes_for(op, severity, H, W);
    if (!rand_param && (!rand_field || fb == 0)) return ADVMIX_OK;
    const int64_t work = std::max<int64_t>(1, (int64_t)(fb / 8));
    const int bx = (int)std::min<int64_t>((work + 255) / 256, 256);
    fill_rand_kernel<<<dim3(bx, n), 256, 0, as_stream(stream)>>>(op, severity, idx, H, W, seed, sample_base,
                                                               fb ? reinterpret_cast<char*>(rand_field) : nullptr, fb,
                                                               rand_param, frost_n, frost_h, frost_w);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_corrupt_u8c3(int op, int severity, const uint8_t* in, uint8_t* out, int n, const int32_t* idx, int H, int W,
                        const void* rand_field, const double* rand_param, uint64_t seed, int64_t sample_base,
                        const uint8_t* frost_bank, int frost_n, int frost_h, int frost_w, void* workspace,
                        size_t ws_bytes, advmix_stream_t stream) {
    const bool fast = (op & ADVMIX_CORRUPT_FAST) != 0;
    op &= ~ADVMIX_CORRUPT_FAST;
    int rc = check_common(op, severity, n, H, W);
    if (rc) return rc;
    if (n == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(in && out, "corrupt: null image pointer");
    ADVMIX_REQUIRE(in != out, "corrupt: in-place operation is not supported");
    const size_t need = ws_bytes_for(op, severity, n, H, W);
    if (need && (!workspace || ws_bytes < need))
        return fail(ADVMIX_ERR_WORKSPACE, "corrupt(op=%d): workspace %zu < %zu bytes", op, ws_bytes, need);
    CorruptArgs a{op, severity, in, out, n, idx, H, W, rand_field, rand_param, seed, sample_base,
                  frost_bank, frost_n, frost_h, frost_w, workspace, ws_bytes, as_stream(stream),
                  field_bytes_for(op, severity, H, W), fast};
    return run_one(a);
}

size_t advmix_corrupt_sweep_workspace_bytes(int op, int n, int H, int W) {
    op &= ~ADVMIX_CORRUPT_FAST;
    if (op < 0 || op >= C_NUM_OPS || H < 1 || W < 1) return 0;
    return sweep_ws_bytes(op, n, H, W);
}

int advmix_corrupt_sweep_u8c3(int op, const uint8_t* in, uint8_t* const* outs, int n, const int32_t* idx, int H, int W,
                              uint64_t seed, int64_t sample_base, const uint8_t* frost_bank, int frost_n, int frost_h,
                              int frost_w, void* workspace, size_t ws_bytes, advmix_stream_t stream) {
    const bool fast = (op & ADVMIX_CORRUPT_FAST) != 0;
    op &= ~ADVMIX_CORRUPT_FAST;
    int rc = check_common(op, 1, n, H, W);
    if (rc) return rc;
    if (n == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(in && outs, "corrupt_sweep: null pointer");
    for (int s = 0; s < 5; ++s) {
        ADVMIX_REQUIRE(outs[s] != nullptr && outs[s] != in, "corrupt_sweep: output %d is null or aliases the input", s);
        for (int t = 0; t < s; ++t) ADVMIX_REQUIRE(outs[s] != outs[t], "corrupt_sweep: outputs %d and %d alias", t, s);
    }
    const size_t need = sweep_ws_bytes(op, n, H, W);
    if (need && (!workspace || ws_bytes < need))
        return fail(ADVMIX_ERR_WORKSPACE, "corrupt_sweep(op=%d): workspace %zu < %zu bytes", op, ws_bytes, need);
    SweepArgs sw{CorruptArgs{op, 1, in, nullptr, n, idx, H, W, nullptr, nullptr, seed, sample_base, frost_bank, frost_n, frost_h,
                             frost_w, workspace, ws_bytes, as_stream(stream), 0, fast},
                 {outs[0], outs[1], outs[2], outs[3], outs[4]}};
    int frc = -1;
    switch (op) {
        case C_GAUSSIAN_NOISE: frc = run_gaussian_noise_sweep(sw); break;
        case C_IMPULSE_NOISE: frc = run_impulse_noise_sweep(sw); break;
        case C_ZOOM_BLUR: frc = run_zoom_blur_sweep_fast(sw); break;
        case C_FROST: frc = run_frost_sweep(sw); break;
        case C_BRIGHTNESS: frc = run_brightness_sweep(sw); break;
        case C_CONTRAST: frc = run_contrast_sweep(sw); break;
        default: break;
    }
    if (frc != -1) return frc;
    const void* shared = nullptr;
    if (sweep_shares_field(op)) {
        void* gen = reinterpret_cast<char*>(workspace) + sweep_ws_main(op, n, H, W);
        rc = launch_fill_rand(sw.base, gen, nullptr);
        if (rc) return rc;
        shared = gen;
        if (op == C_ELASTIC) {
            sw.base.field_bytes = field_bytes_for(op, 1, H, W);
            frc = run_elastic_sweep_fast(sw, reinterpret_cast<const float*>(gen));
            if (frc != -1) return frc;
        }
    }
    int first = 1;
    if (op == C_FOG && shared) {
        CorruptArgs a = sw.base;
        a.field_bytes = field_bytes_for(op, 1, H, W);
        if (run_fog_pair_fast(a, reinterpret_cast<const float*>(shared), outs[0], outs[1]) == ADVMIX_OK) first = 3;
    }
    for (int s = first; s <= 5; ++s) {      // no fused kernel: the five per-severity launches
        CorruptArgs a = sw.base;
        a.severity = s;
        a.out = outs[s - 1];
        a.field_bytes = field_bytes_for(op, s, H, W);
        a.rand_field = shared;              // same values the in-register path draws (tests: perf == injected dump)
        rc = run_one(a);
        if (rc) return rc;
    }
    return ADVMIX_OK;
}

}  // extern "C"
