// One C call per training step of the K = 1 crop + target path: JointsDataset.get_clean
// (lib/dataset/JointsDataset.py:258-364) for a whole batch - get_affine_transform, cv2.warpAffine (+flip) with the
// ToTensor / Normalize epilogue, fliplr_joints + affine_transform, generate_target.  The host side used to make ~6 ctypes calls
// per step (~1 ms of Python per 256 samples, more than ten times the device time); this entry point takes every per-step
// array out of ONE packed device buffer (the host fills one pinned buffer and issues one copy) and launches
//     matrices -> { crop (caller's stream) || joints + heat maps (library side stream) }
// with the same fork / join the bench's CUDA graph uses, so it can also be captured into a graph.
#include "common.cuh"

#include <map>
#include <mutex>

namespace advmix {

struct StepStreams { cudaStream_t side; cudaEvent_t fork, join; };
static std::mutex g_step_mu;
static std::map<int, StepStreams> g_step_streams;

static int step_streams(StepStreams* out) {
    int dev = 0;
    ADVMIX_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_step_mu);
    auto it = g_step_streams.find(dev);
    if (it == g_step_streams.end()) {
        StepStreams s;
        ADVMIX_CUDA_OK(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
        ADVMIX_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        ADVMIX_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
        it = g_step_streams.insert(std::make_pair(dev, s)).first;
    }
    *out = it->second;
    return ADVMIX_OK;
}

}  // namespace advmix

using namespace advmix;

extern "C" {

size_t advmix_step_params_bytes(int B, int J) {
    // layout of the packed per-step buffer (every section 16-byte aligned, see advmix_crop_targets_step)
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t o = 0;
    o += al((size_t)B * 8);          // src_off   int64 [B]
    o += al((size_t)B * 8);          // src_pitch int64 [B]
    o += al((size_t)B * 4);          // src_h     int32 [B]
    o += al((size_t)B * 4);          // src_w     int32 [B]
    o += al((size_t)B * 2 * 8);      // scale     f64   [B][2]
    o += al((size_t)B * 8);          // rot       f64   [B]
    o += al((size_t)B * 2 * 4);      // center    f32   [B][2]
    o += al((size_t)B);              // flip      u8    [B]
    o += al((size_t)B * J * 3 * 8);  // joints    f64   [B][J][3]
    o += al((size_t)B * J * 3 * 8);  // vis       f64   [B][J][3]
    return o;
}

int advmix_crop_targets_step(const uint8_t* src_base, const void* params, const int32_t* flip_perm, const float* norm_lut,
                             const float* gauss_tab, const float* joints_weight, double* M_fwd, void* inp_norm, int norm_dtype,
                             double* joints_out, double* vis_out, float* hm, float* mu, float* tw, int B, int J, int out_w, int out_h,
                             int Hh, int Wh, int sigma, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0, "crop_targets_step: bad shape B=%d J=%d", B, J);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(src_base && params && norm_lut && gauss_tab && M_fwd && inp_norm && joints_out && vis_out && hm && tw,
                   "crop_targets_step: null argument");
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const char* p = reinterpret_cast<const char*>(params);
    const int64_t* src_off = reinterpret_cast<const int64_t*>(p); p += al((size_t)B * 8);
    const int64_t* src_pitch = reinterpret_cast<const int64_t*>(p); p += al((size_t)B * 8);
    const int32_t* src_h = reinterpret_cast<const int32_t*>(p); p += al((size_t)B * 4);
    const int32_t* src_w = reinterpret_cast<const int32_t*>(p); p += al((size_t)B * 4);
    const double* scale = reinterpret_cast<const double*>(p); p += al((size_t)B * 16);
    const double* rot = reinterpret_cast<const double*>(p); p += al((size_t)B * 8);
    const float* center = reinterpret_cast<const float*>(p); p += al((size_t)B * 8);
    const uint8_t* flip = reinterpret_cast<const uint8_t*>(p); p += al((size_t)B);
    const double* joints = reinterpret_cast<const double*>(p); p += al((size_t)B * J * 24);
    const double* vis = reinterpret_cast<const double*>(p);
    StepStreams ss;
    int rc = step_streams(&ss);
    if (rc) return rc;
    cudaStream_t main = as_stream(stream);
    rc = advmix_affine_matrices(center, scale, 0, rot, M_fwd, B, out_w, out_h, stream);
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaEventRecord(ss.fork, main));
    ADVMIX_CUDA_OK(cudaStreamWaitEvent(ss.side, ss.fork, 0));
    rc = advmix_joints_flip_affine(joints, vis, flip, src_w, flip_perm, M_fwd, joints_out, vis_out, B, J, reinterpret_cast<advmix_stream_t>(ss.side));
    if (rc) return rc;
    rc = advmix_heatmap_targets(joints_out, vis_out, gauss_tab, joints_weight, hm, mu, tw, B, J, Hh, Wh, out_w, out_h, sigma,
                                reinterpret_cast<advmix_stream_t>(ss.side));
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaEventRecord(ss.join, ss.side));
    rc = advmix_warp_affine_u8c3(src_base, src_off, src_h, src_w, src_pitch, flip, M_fwd, nullptr, inp_norm, norm_lut, B, out_w, out_h,
                                 norm_dtype, stream);
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaStreamWaitEvent(main, ss.join, 0));
    return ADVMIX_OK;
}

/* Record-table form: the joints / joints_vis of the whole dataset shard stay on the device (rec_joints, rec_vis: float64
 * [N][J][3]) and the per-step buffer carries the row indices instead of the rows - 14 KB instead of 223 KB per 256-sample
 * step cross PCIe and the host does not gather 2 x B x J x 3 doubles.  `params` = the advmix_crop_targets_step layout with the
 * joints and vis sections replaced by ONE section  rec_idx int32[B]. */
size_t advmix_step_rec_params_bytes(int B, int J) {
    (void)J;
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    return al((size_t)B * 8) * 2 + al((size_t)B * 4) * 2 + al((size_t)B * 16) + al((size_t)B * 8) * 2 + al((size_t)B) + al((size_t)B * 4);
}

int advmix_crop_targets_step_rec(const uint8_t* src_base, const void* params, const double* rec_joints, const double* rec_vis,
                                 const int32_t* flip_perm, const float* norm_lut, const float* gauss_tab, const float* joints_weight,
                                 double* M_fwd, void* inp_norm, int norm_dtype, double* joints_out, double* vis_out, float* hm, float* mu,
                                 float* tw, int B, int J, int out_w, int out_h, int Hh, int Wh, int sigma, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0, "crop_targets_step_rec: bad shape B=%d J=%d", B, J);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(src_base && params && rec_joints && rec_vis && norm_lut && gauss_tab && M_fwd && inp_norm && joints_out && vis_out && hm && tw,
                   "crop_targets_step_rec: null argument");
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const char* p = reinterpret_cast<const char*>(params);
    const int64_t* src_off = reinterpret_cast<const int64_t*>(p); p += al((size_t)B * 8);
    const int64_t* src_pitch = reinterpret_cast<const int64_t*>(p); p += al((size_t)B * 8);
    const int32_t* src_h = reinterpret_cast<const int32_t*>(p); p += al((size_t)B * 4);
    const int32_t* src_w = reinterpret_cast<const int32_t*>(p); p += al((size_t)B * 4);
    const double* scale = reinterpret_cast<const double*>(p); p += al((size_t)B * 16);
    const double* rot = reinterpret_cast<const double*>(p); p += al((size_t)B * 8);
    const float* center = reinterpret_cast<const float*>(p); p += al((size_t)B * 8);
    const uint8_t* flip = reinterpret_cast<const uint8_t*>(p); p += al((size_t)B);
    const int32_t* rec_idx = reinterpret_cast<const int32_t*>(p);
    StepStreams ss;
    int rc = step_streams(&ss);
    if (rc) return rc;
    cudaStream_t main = as_stream(stream);
    rc = advmix_affine_matrices(center, scale, 0, rot, M_fwd, B, out_w, out_h, stream);
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaEventRecord(ss.fork, main));
    ADVMIX_CUDA_OK(cudaStreamWaitEvent(ss.side, ss.fork, 0));
    rc = advmix_joints_flip_affine_rec(rec_joints, rec_vis, rec_idx, flip, src_w, flip_perm, M_fwd, joints_out, vis_out, B, J,
                                       reinterpret_cast<advmix_stream_t>(ss.side));
    if (rc) return rc;
    rc = advmix_heatmap_targets(joints_out, vis_out, gauss_tab, joints_weight, hm, mu, tw, B, J, Hh, Wh, out_w, out_h, sigma,
                                reinterpret_cast<advmix_stream_t>(ss.side));
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaEventRecord(ss.join, ss.side));
    rc = advmix_warp_affine_u8c3(src_base, src_off, src_h, src_w, src_pitch, flip, M_fwd, nullptr, inp_norm, norm_lut, B, out_w, out_h,
                                 norm_dtype, stream);
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaStreamWaitEvent(main, ss.join, 0));
    return ADVMIX_OK;
}

/* K = 3 (sample_times = 3) form for the fused chain + mix path: the step produces what the AdvMix inner loop needs WITHOUT
 * materialising the three chains - the uint8 crop, the per-image autoaug plans, the gridmask parameters, the joints and the
 * heat-map targets of the clean / autoaug chains (shared, JointsDataset.py:236-256) and, optionally, those of the gridmask chain
 * (visibility of masked joints dropped, advaug.py:153-165).  `params` = the K = 1 layout followed by
 *   autoaug ops int32[B][2] | autoaug mags f32[B][2] | gridmask params int32[B][4]     (each section 16-byte aligned). */
size_t advmix_chains_step_params_bytes(int B, int J) {
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    return advmix_step_params_bytes(B, J) + al((size_t)B * 8) + al((size_t)B * 8) + al((size_t)B * 16);
}

int advmix_crop_chains_step(const uint8_t* src_base, const void* params, const int32_t* flip_perm, const float* gauss_tab,
                            const float* joints_weight, double* M_fwd, uint8_t* crop_u8, void* plans, void* plan_ws, size_t plan_ws_bytes,
                            double* joints_out, double* vis_out, double* vis_gm_out, float* hm, float* mu, float* tw, float* hm_gm,
                            float* tw_gm, int B, int J, int out_w, int out_h, int Hh, int Wh, int sigma, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0, "crop_chains_step: bad shape B=%d J=%d", B, J);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(src_base && params && gauss_tab && M_fwd && crop_u8 && plans && plan_ws && joints_out && vis_out && hm && tw,
                   "crop_chains_step: null argument");
    ADVMIX_REQUIRE(!hm_gm || (vis_gm_out && tw_gm), "crop_chains_step: hm_gm needs vis_gm_out and tw_gm");
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const char* p = reinterpret_cast<const char*>(params);
    const int64_t* src_off = reinterpret_cast<const int64_t*>(p); p += al((size_t)B * 8);
    const int64_t* src_pitch = reinterpret_cast<const int64_t*>(p); p += al((size_t)B * 8);
    const int32_t* src_h = reinterpret_cast<const int32_t*>(p); p += al((size_t)B * 4);
    const int32_t* src_w = reinterpret_cast<const int32_t*>(p); p += al((size_t)B * 4);
    const double* scale = reinterpret_cast<const double*>(p); p += al((size_t)B * 16);
    const double* rot = reinterpret_cast<const double*>(p); p += al((size_t)B * 8);
    const float* center = reinterpret_cast<const float*>(p); p += al((size_t)B * 8);
    const uint8_t* flip = reinterpret_cast<const uint8_t*>(p); p += al((size_t)B);
    const double* joints = reinterpret_cast<const double*>(p); p += al((size_t)B * J * 24);
    const double* vis = reinterpret_cast<const double*>(p); p += al((size_t)B * J * 24);
    const int32_t* aa_ops = reinterpret_cast<const int32_t*>(p); p += al((size_t)B * 8);
    const float* aa_mags = reinterpret_cast<const float*>(p); p += al((size_t)B * 8);
    const int32_t* gm = reinterpret_cast<const int32_t*>(p);
    StepStreams ss;
    int rc = step_streams(&ss);
    if (rc) return rc;
    cudaStream_t main = as_stream(stream);
    advmix_stream_t side = reinterpret_cast<advmix_stream_t>(ss.side);
    rc = advmix_affine_matrices(center, scale, 0, rot, M_fwd, B, out_w, out_h, stream);
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaEventRecord(ss.fork, main));
    ADVMIX_CUDA_OK(cudaStreamWaitEvent(ss.side, ss.fork, 0));
    rc = advmix_joints_flip_affine(joints, vis, flip, src_w, flip_perm, M_fwd, joints_out, vis_out, B, J, side);
    if (rc) return rc;
    rc = advmix_heatmap_targets(joints_out, vis_out, gauss_tab, joints_weight, hm, mu, tw, B, J, Hh, Wh, out_w, out_h, sigma, side);
    if (rc) return rc;
    if (vis_gm_out) {
        rc = advmix_gridmask(nullptr, nullptr, gm, joints_out, vis_out, vis_gm_out, B, out_h, out_w, J, ADVMIX_F32, side);
        if (rc) return rc;
        if (hm_gm) {
            rc = advmix_heatmap_targets(joints_out, vis_gm_out, gauss_tab, joints_weight, hm_gm, nullptr, tw_gm, B, J, Hh, Wh, out_w, out_h, sigma, side);
            if (rc) return rc;
        }
    }
    ADVMIX_CUDA_OK(cudaEventRecord(ss.join, ss.side));
    rc = advmix_warp_affine_u8c3(src_base, src_off, src_h, src_w, src_pitch, flip, M_fwd, crop_u8, nullptr, nullptr, B, out_w, out_h,
                                 ADVMIX_F32, stream);
    if (rc) return rc;
    rc = advmix_autoaug_plan_u8c3(crop_u8, aa_ops, aa_mags, plans, B, out_h, out_w, plan_ws, plan_ws_bytes, stream);
    if (rc) return rc;
    ADVMIX_CUDA_OK(cudaStreamWaitEvent(main, ss.join, 0));
    return ADVMIX_OK;
}

}  // extern "C"
