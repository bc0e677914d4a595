// advmix_corrupt_sweep_u8c3: the five severities of one corruption from ONE read of the crops.
// tools/make_datasets.py:38-45 runs `for severity in range(5)` innermost over the same image, so per (image, corruption) the
// crop is read once and five outputs are written; work that does not depend on the severity (random draws, colour-space
// conversion, zoom layers that several severities share) is done once.  Every output is bit-identical to what
// advmix_corrupt_u8c3(op, severity) writes for the same seed (tests/test_gpu_chains_corruptions.py::test_sweep_*): the
// Philox key is (seed, sample, op) - the severity is not part of it - and the fused kernels evaluate the same expressions.
// Ops without a fused kernel (or shapes a fused kernel does not take) run their five per-severity launches.
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "corrupt_common.cuh"

namespace advmix {

// ======================================================================== zoom_blur (ADVMIX_CORRUPT_FAST arithmetic)
// Severity s averages the crop with nl_s zoomed copies (factors 1 + l*step_s).  The integer layer samples of
// zoom_blur_fast_kernel do not depend on the order they are summed in, and the 62 (severity, layer) pairs hold only 31
// different layers: 1.00 belongs to all five severities, 1.02 .. 1.10 to four, and so on.  The kernel evaluates each unique
// layer once, sums the layers of one membership class, and adds the class sum to the severities that contain it.
constexpr int ZS_THREADS = 1024, ZS_MAXCLS = 16;

struct ZoomSweepPlan {
    int ncls;
    int first[ZS_MAXCLS], count[ZS_MAXCLS];
    uint32_t mask[ZS_MAXCLS];
    float denom[5], rdenom[5];     // nl_s + 1 and its correctly rounded reciprocal
};
struct SweepOuts { uint8_t* p[5]; };

// table entry: index (row or column, 10 bits) | step << 10 | w1 << 11 (w1 = round(t * 4096) <= 4096) ; bit 31: outside (index 0, w1 0:
// both weights are then taken as 0 and the layer contributes nothing - no branch)
__device__ __forceinline__ uint32_t zs_entry(const uint32_t* __restrict__ g, const uint32_t* s, bool smem, int i) {
    return smem ? s[i] : __ldg(g + i);
}

// The first version carried zoom_blur_fast_kernel's reuse of horizontally interpolated source rows between a thread's output rows;
// its tests and moves compiled to ~33 instructions per value and layer (ncu: 90 % issue-active).  Straight-line code - two row
// evaluations per output row, no reuse, no branches - needs ~17, and the outside test folds into zero weights.  The final division is
// the FMA-refined product with the correctly rounded reciprocal (q = v*y, r = fma(-q, d, v), q' = fma(r, y, q)): checked exhaustively
// against float division for every float32 in [1e-37, 4400] and d = 12, 13, 14, 17 (scratch check, numpy), so the outputs stay those of zoom_blur_fast_kernel.
template <bool TAB_SMEM>
__global__ void __launch_bounds__(ZS_THREADS, 1)
zoom_sweep_fast_kernel(const uint8_t* __restrict__ in, SweepOuts outs, const int32_t* __restrict__ idx, int n, int H, int W,
                       const uint32_t* __restrict__ tab, int nu, ZoomSweepPlan plan) {
    extern __shared__ __align__(16) uint8_t zs_img[];
    const int nbytes = H * W * 3, W3 = W * 3, HW = H + W;
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(zs_img + ((nbytes + 15) & ~15));
    if (TAB_SMEM)
        for (int i = threadIdx.x; i < nu * HW; i += ZS_THREADS) s_tab[i] = tab[i];
    for (int img = blockIdx.x; img < n; img += gridDim.x) {
        const int slot = slot_of(idx, img);
        const uint8_t* src = in + (int64_t)slot * nbytes;
        __syncthreads();
        {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(zs_img);
            for (int i = threadIdx.x; i < nbytes / 16; i += ZS_THREADS) d4[i] = ld_stream_u4(s4 + i);
        }
        __syncthreads();
        // a thread owns one column x 2 rows x 3 channels; per severity 6 accumulators
        const int hq = (H + 1) >> 1;
        for (int item = threadIdx.x; item < hq * W; item += ZS_THREADS) {
            const int yq = item / W, x = item - yq * W;
            const int y0 = yq << 1, y1 = min(y0 + 1, H - 1);           // odd H: the last item evaluates row H-1 twice and stores it once
            uint32_t acc[5][2][3];
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
                for (int i = 0; i < 2; ++i) acc[s][i][0] = acc[s][i][1] = acc[s][i][2] = 0u;
#pragma unroll 1
            for (int k = 0; k < plan.ncls; ++k) {
                uint32_t cs[2][3] = {{0u, 0u, 0u}, {0u, 0u, 0u}};
                const int l1 = plan.first[k] + plan.count[k];
#pragma unroll 2
                for (int l = plan.first[k]; l < l1; ++l) {
                    const int base = l * HW;
                    const uint32_t cc = zs_entry(tab, s_tab, TAB_SMEM, base + H + x);
                    const uint32_t wx1 = (cc >> 11) & 0x1FFFu, wx0 = (cc >> 31) ? 0u : 4096u - wx1;
                    const uint8_t* c0p = zs_img + (cc & 1023u) * 3u;
                    const uint8_t* c1p = c0p + ((cc >> 10) & 1u) * 3u;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint32_t rr = zs_entry(tab, s_tab, TAB_SMEM, base + (i ? y1 : y0));
                        const uint32_t r0 = (rr & 1023u) * (uint32_t)W3, r1 = r0 + ((rr >> 10) & 1u) * (uint32_t)W3;
                        const uint32_t wy1 = (rr >> 11) & 0x1FFFu, wy0 = (rr >> 31) ? 0u : 4096u - wy1;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const uint32_t hT = c0p[r0 + c] * wx0 + c1p[r0 + c] * wx1;
                            const uint32_t hB = c0p[r1 + c] * wx0 + c1p[r1 + c] * wx1;
                            cs[i][c] += (hT * wy0 + hB * wy1 + 128u) >> 8;
                        }
                    }
                }
                const uint32_t m = plan.mask[k];
#pragma unroll
                for (int s = 0; s < 5; ++s)
                    if (m & (1u << s)) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int c = 0; c < 3; ++c) acc[s][i][c] += cs[i][c];
                    }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i == 1 && y0 + 1 >= H) break;
                const int pb = ((y0 + i) * W + x) * 3;
                const float px[3] = {(float)zs_img[pb], (float)zs_img[pb + 1], (float)zs_img[pb + 2]};
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    uint8_t* dst = outs.p[s] + (int64_t)slot * nbytes + pb;
                    const float d = plan.denom[s], y = plan.rdenom[s];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float v = fmaf((float)acc[s][i][c], 1.0f / 65536.0f, px[c]);      // same expression as zoom_blur_fast_kernel
                        const float q0 = v * y;
                        const float q = fmaf(fmaf(-q0, d, v), y, q0);                          // == __fdiv_rn(v, d) (see above)
                        dst[c] = (uint8_t)__float2int_rz(fminf(q, 255.0f));
                    }
                }
            }
        }
    }
}

static int py_round_i(double v) { return (int)std::nearbyint(v); }

struct ZoomSweepTable { const uint32_t* d_tab; int nu; ZoomSweepPlan plan; bool ok; };

// the (severity, layer) tap tables of run_zoom_blur_fast, de-duplicated by content and grouped by membership mask
static const ZoomSweepTable& zoom_sweep_table(int H, int W) {
    static std::mutex mu;
    static std::map<std::string, ZoomSweepTable> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::string key = "zoomsweep_" + std::to_string(dev) + "_" + std::to_string(H) + "x" + std::to_string(W);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    ZoomSweepTable t{};
    const int HW = H + W;
    const double stop[5] = {1.11, 1.16, 1.21, 1.26, 1.33}, step[5] = {0.01, 0.01, 0.02, 0.02, 0.03};
    std::vector<std::vector<uint32_t>> uniq;
    std::vector<uint32_t> umask;
    bool ok = H <= 1023 && W <= 1023;
    for (int s = 0; s < 5 && ok; ++s) {
        const int nl = (int)std::ceil((stop[s] - 1.0) / step[s]);
        t.plan.denom[s] = (float)(nl + 1);
        t.plan.rdenom[s] = (float)(1.0 / (double)(nl + 1));
        if (nl + 1 != 12 && nl + 1 != 13 && nl + 1 != 14 && nl + 1 != 17) ok = false;      // the divisors the refined division was checked for
        for (int l = 0; l < nl; ++l) {
            const double zf = 1.0 + l * step[s];
            const int in0 = (int)std::ceil(H / zf), top0 = (H - in0) / 2, in1 = (int)std::ceil(W / zf), top1 = (W - in1) / 2;
            const int out0 = py_round_i(in0 * zf), out1 = py_round_i(in1 * zf);
            const double z0 = out0 > 1 ? (double)(in0 - 1) / (double)(out0 - 1) : 1.0, z1 = out1 > 1 ? (double)(in1 - 1) / (double)(out1 - 1) : 1.0;
            std::vector<uint32_t> e(HW);
            auto entry = [](int o, int outn, double z, int inn, int top) -> uint32_t {
                const double cc = (double)o * z;
                if (o >= outn || cc < 0.0 || cc > (double)(inn - 1)) return 0x80000000u;
                const double f = std::floor(cc);
                const int sidx = (int)f;
                const uint32_t w1 = (uint32_t)std::nearbyint((cc - f) * 4096.0);
                const uint32_t stp = std::min(sidx + 1, inn - 1) != sidx ? 1u : 0u;
                return (uint32_t)(top + sidx) | (stp << 10) | (w1 << 11);
            };
            for (int y = 0; y < H; ++y) e[y] = entry(y, out0, z0, in0, top0);
            for (int x = 0; x < W; ++x) e[H + x] = entry(x, out1, z1, in1, top1);
            size_t u = 0;
            for (; u < uniq.size(); ++u)
                if (std::memcmp(uniq[u].data(), e.data(), HW * 4) == 0) break;
            if (u == uniq.size()) { uniq.push_back(std::move(e)); umask.push_back(0u); }
            if (umask[u] & (1u << s)) ok = false;        // a severity holding the same layer twice: sums would differ
            umask[u] |= 1u << s;
        }
    }
    // group by mask
    std::vector<uint32_t> flat;
    int ncls = 0;
    std::vector<bool> used(uniq.size(), false);
    for (size_t u = 0; u < uniq.size() && ok; ++u) {
        if (used[u]) continue;
        if (ncls == ZS_MAXCLS) { ok = false; break; }
        t.plan.first[ncls] = (int)(flat.size() / HW);
        t.plan.mask[ncls] = umask[u];
        int cnt = 0;
        for (size_t v = u; v < uniq.size(); ++v)
            if (!used[v] && umask[v] == umask[u]) { used[v] = true; flat.insert(flat.end(), uniq[v].begin(), uniq[v].end()); ++cnt; }
        t.plan.count[ncls++] = cnt;
    }
    t.plan.ncls = ncls;
    t.nu = (int)uniq.size();
    t.ok = ok;
    if (ok) {
        t.d_tab = reinterpret_cast<const uint32_t*>(cached_table(key, flat.data(), flat.size() * 4));
        if (!t.d_tab) t.ok = false;
    }
    return cache.emplace(key, t).first->second;
}

int run_zoom_blur_sweep_fast(const SweepArgs& sw) {
    const CorruptArgs& a = sw.base;
    const int H = a.H, W = a.W;
    const size_t img_bytes = (size_t)H * W * 3;
    if (!a.fast || img_bytes % 16 != 0 || img_bytes > 200 * 1024 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0) return -1;
    const ZoomSweepTable& t = zoom_sweep_table(H, W);
    if (!t.ok) return -1;
    SweepOuts outs;
    for (int s = 0; s < 5; ++s) outs.p[s] = sw.outs[s];
    const size_t img_al = (img_bytes + 15) & ~(size_t)15, tab_bytes = (size_t)t.nu * (H + W) * 4;
    const size_t smem_cap = 227 * 1024 - 2048;
    const int grid = std::min(a.n, sm_count());
    if (img_al + tab_bytes <= smem_cap) {
        ADVMIX_CUDA_OK(ensure_dyn_smem(zoom_sweep_fast_kernel<true>, (int)smem_cap));
        zoom_sweep_fast_kernel<true><<<grid, ZS_THREADS, img_al + tab_bytes, a.stream>>>(a.in, outs, a.idx, a.n, H, W, t.d_tab, t.nu, t.plan);
    } else {
        ADVMIX_CUDA_OK(ensure_dyn_smem(zoom_sweep_fast_kernel<false>, (int)smem_cap));
        zoom_sweep_fast_kernel<false><<<grid, ZS_THREADS, img_al, a.stream>>>(a.in, outs, a.idx, a.n, H, W, t.d_tab, t.nu, t.plan);
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // namespace advmix
