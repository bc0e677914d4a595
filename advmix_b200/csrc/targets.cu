// Row a4: generate_target Gaussian heatmaps (lib/dataset/JointsDataset.py:412-491).
// Write-only, HBM-store bound: 4*J*Hh*Wh bytes per sample.
#include "common.cuh"

namespace advmix {

constexpr int HM_THREADS = 256;
constexpr int HM_PLANES = 8;   // planes per CTA work item
constexpr int HM_MAXTAB = 37 * 37;  // sigma <= 6

struct PlaneInfo {
    int ul_x, ul_y;   // window origin (mu - 3*sigma)
    int paste;        // 1 if the Gaussian is written
};

// One work item = HM_PLANES consecutive (b,j) planes.  Lanes 0..7 derive mu / window /
// weight for their plane in float64 exactly as the reference does, then the whole CTA
// streams the planes out with 128-bit stores (zeros outside the 13x13 window).
__global__ void __launch_bounds__(HM_THREADS)
heatmap_kernel(const double* __restrict__ joints, const double* __restrict__ vis,
               const float* __restrict__ gtab, const float* __restrict__ jw, float* __restrict__ hm,
               float* __restrict__ mu, float* __restrict__ tw, int planes, int J, int Hh, int Wh,
               double stride_x, double stride_y, int tmp_size, int chunks, int rows_chunk) {
    __shared__ PlaneInfo info[HM_PLANES];
    __shared__ float tab[HM_MAXTAB];
    const int size = 2 * tmp_size + 1;
    for (int i = threadIdx.x; i < size * size; i += HM_THREADS) tab[i] = gtab[i];
    const int plane_elems = Hh * Wh;
    // large planes (configs[3]: 128x128, 256x256) are cut into row chunks so that small batches still fill the GPU
    const int items = ((planes + HM_PLANES - 1) / HM_PLANES) * chunks;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int p0 = (item / chunks) * HM_PLANES, chunk = item % chunks;
        const int y_lo = chunk * rows_chunk, y_hi = min(Hh, y_lo + rows_chunk);
        __syncthreads();
        if (threadIdx.x < HM_PLANES) {
            const int p = p0 + threadIdx.x;
            PlaneInfo pi{0, 0, 0};
            if (p < planes) {
                const double jx = joints[(int64_t)p * 3], jy = joints[(int64_t)p * 3 + 1];
                float w = (float)vis[(int64_t)p * 3];
                // int(x / feat_stride + 0.5): float64, truncation toward zero
                const int mu_x = __double2int_rz(__dadd_rn(__ddiv_rn(jx, stride_x), 0.5));
                const int mu_y = __double2int_rz(__dadd_rn(__ddiv_rn(jy, stride_y), 0.5));
                const int ulx = mu_x - tmp_size, uly = mu_y - tmp_size;
                const int brx = mu_x + tmp_size + 1, bry = mu_y + tmp_size + 1;
                float mx = 0.f, my = 0.f;
                if (ulx >= Wh || uly >= Hh || brx < 0 || bry < 0) {
                    w = 0.f;
                } else if (w > 0.5f) {
                    pi.paste = 1;
                    mx = (float)mu_x;
                    my = (float)mu_y;
                }
                pi.ul_x = ulx;
                pi.ul_y = uly;
                if (jw) w = __fmul_rn(w, jw[p % J]);
                if (chunk == 0) {
                    tw[p] = w;
                    if (mu) { mu[2 * p] = mx; mu[2 * p + 1] = my; }
                }
            }
            info[threadIdx.x] = pi;
        }
        __syncthreads();
        const int nplanes = min(HM_PLANES, planes - p0);
        if ((Wh & 3) == 0 && (Wh >> 2) <= HM_THREADS) {
            // thread -> (plane slot, row group, float4 column); the column and its 4 window offsets are loop
            // invariant, rows advance by a fixed step, so the inner loop is: 4 table reads, one 128-bit store
            const int wv = Wh >> 2;                                // float4 per row
            const int rows_per_pass = HM_THREADS / wv;             // rows covered by the CTA per pass
            const int tcol = threadIdx.x % wv, trow = threadIdx.x / wv;
            if (trow < rows_per_pass) {
                const int x = tcol << 2;
                for (int lp = 0; lp < nplanes; ++lp) {
                    const PlaneInfo pi = info[lp];
                    const int gx = x - pi.ul_x;
                    const bool c0 = (unsigned)gx < (unsigned)size, c1 = (unsigned)(gx + 1) < (unsigned)size;
                    const bool c2 = (unsigned)(gx + 2) < (unsigned)size, c3 = (unsigned)(gx + 3) < (unsigned)size;
                    const bool any = pi.paste && (c0 || c1 || c2 || c3);
                    float* dst = hm + (int64_t)(p0 + lp) * plane_elems + x;
                    for (int y = y_lo + trow; y < y_hi; y += rows_per_pass) {
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                        const int gy = y - pi.ul_y;
                        if (any && (unsigned)gy < (unsigned)size) {
                            const float* row = tab + gy * size + gx;
                            if (c0) o.x = row[0];
                            if (c1) o.y = row[1];
                            if (c2) o.z = row[2];
                            if (c3) o.w = row[3];
                        }
                        st_stream_f4(dst + (int64_t)y * Wh, o);
                    }
                }
            }
        } else {
            const int chunk_elems = (y_hi - y_lo) * Wh;
            for (int v = threadIdx.x; v < nplanes * chunk_elems; v += HM_THREADS) {
                const int lp = v / chunk_elems, r = v - lp * chunk_elems + y_lo * Wh;
                const int y = r / Wh, x = r - y * Wh;
                const PlaneInfo pi = info[lp];
                const int gy = y - pi.ul_y, gx = x - pi.ul_x;
                float o = 0.f;
                if (pi.paste && (unsigned)gy < (unsigned)size && (unsigned)gx < (unsigned)size) o = tab[gy * size + gx];
                hm[(int64_t)(p0 + lp) * plane_elems + r] = o;
            }
        }
    }
}

}  // namespace advmix

using namespace advmix;

extern "C" int advmix_heatmap_targets(const double* joints, const double* vis, const float* gauss_tab,
                                      const float* joints_weight, float* hm, float* mu, float* tw, int B,
                                      int J, int Hh, int Wh, int img_w, int img_h, int sigma,
                                      advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0 && Hh > 0 && Wh > 0 && img_w > 0 && img_h > 0, "heatmap_targets: bad shape");
    ADVMIX_REQUIRE(sigma >= 1 && sigma <= 6, "heatmap_targets: sigma %d outside 1..6", sigma);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(joints && vis && gauss_tab && hm && tw, "heatmap_targets: null argument");
    const int planes = B * J;
    const int items = (planes + HM_PLANES - 1) / HM_PLANES;
    // row chunks: enough work items for ~4 CTAs per SM, at least 8 rows each
    int chunks = 1;
    while (items * chunks < 4 * sm_count() && Hh / (2 * chunks) >= 8) chunks *= 2;
    const int rows_chunk = (Hh + chunks - 1) / chunks;
    const int blocks = std::min(items * chunks, sm_count() * 8);
    // feat_stride = image_size / heatmap_size (numpy float64 true division)
    const double sx = (double)img_w / (double)Wh, sy = (double)img_h / (double)Hh;
    heatmap_kernel<<<blocks, HM_THREADS, 0, as_stream(stream)>>>(joints, vis, gauss_tab, joints_weight, hm, mu, tw,
                                                              planes, J, Hh, Wh, sx, sy, 3 * sigma, chunks, rows_chunk);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}
